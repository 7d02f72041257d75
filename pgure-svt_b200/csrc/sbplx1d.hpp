// Host-side 1-D subplex driver for the lambda search of PGURE::Optimize (src/pgure.hpp:196-237).
//
// The reference calls NLopt 2.6.2's LN_SBPLX with n = 1, bounds [0, ub], ftol_rel = tol, xtol_abs = 1e-12,
// maxeval and an initial step sqrt(start) (pgure.hpp:206-216).  NLopt is an external dependency that is not
// vendored by the reference, so this is a from-the-published-algorithm implementation of Rowan's subplex as
// NLopt structures it: for one dimension it is a loop of two-point Nelder–Mead searches (reflect / expand /
// contract / shrink with alpha=1, beta=1/2, gamma=2, delta=1/2, points pinned to the bounds), each stopped
// when the simplex diameter has shrunk by psi = 1/4, followed by the subplex step update
// step <- copysign(psi*step, dx) (or -psi*step if the best point did not move).
//
// The objective is one fused GPU pass (threshold + reconstruct + aggregate + risk reduction) per call.
#pragma once
#include <cmath>

namespace pgs
{

enum
{
    SBPLX_FAILURE = -1,
    SBPLX_INVALID_ARGS = -2,
    SBPLX_SUCCESS = 1,
    SBPLX_FTOL_REACHED = 3,
    SBPLX_XTOL_REACHED = 4,
    SBPLX_MAXEVAL_REACHED = 5
};

struct Sbplx1D
{
    double lb, ub, ftol_rel, xtol_abs;
    int maxeval;
    // state
    int nevals = 0;
    double xbest = 0, fbest = 0;

    static bool approx_equal(double a, double b) { return std::fabs(a - b) <= 1e-13 * (std::fabs(a) + std::fabs(b)); }

    static bool rel_stop(double vold, double vnew, double reltol, double abstol)
    {
        if (std::isinf(vold))
            return false;
        const double diff = std::fabs(vnew - vold);
        return diff < abstol || diff < reltol * (std::fabs(vnew) + std::fabs(vold)) * 0.5 || (reltol > 0 && vnew == vold);
    }

    // xnew = c + scale*(c - xold), clipped to the bounds; false when it coincides with c or xold
    bool reflect(double &xnew, double c, double scale, double xold) const
    {
        double v = c + scale * (c - xold);
        if (v < lb)
            v = lb;
        if (v > ub)
            v = ub;
        xnew = v;
        return !(approx_equal(v, c) || approx_equal(v, xold));
    }

    template <typename F>
    bool eval(F &f, double x, double &fx, int &ret)
    {
        fx = f(x);
        nevals++;
        if (fx <= fbest)
        {
            fbest = fx;
            xbest = x;
        }
        if (maxeval > 0 && nevals >= maxeval)
        {
            ret = SBPLX_MAXEVAL_REACHED;
            return false;
        }
        return true;
    }

    // two-point Nelder–Mead from (xbest, fbest) with initial step `step`; stops when |xl - xh| < psi * initial
    template <typename F>
    int nelder_mead(F &f, double step, double psi, double &fdiff)
    {
        int ret = SBPLX_SUCCESS;
        double px[2], pf[2];
        fdiff = HUGE_VAL;
        const double x0 = xbest;
        px[0] = x0;
        pf[0] = fbest;
        px[1] = x0 + step;
        if (px[1] > ub)
            px[1] = (ub - x0 > std::fabs(step) * 0.1) ? ub : x0 - std::fabs(step);
        if (px[1] < lb)
        {
            if (x0 - lb > std::fabs(step) * 0.1)
                px[1] = lb;
            else
            {
                px[1] = x0 + std::fabs(step);
                if (px[1] > ub)
                    px[1] = 0.5 * ((ub - x0 > x0 - lb ? ub : lb) + x0);
            }
        }
        if (approx_equal(px[1], x0))
            return SBPLX_FAILURE;
        if (!eval(f, px[1], pf[1], ret))
            return ret;
        double init_diam = 0;
        for (;;)
        {
            const int lo = (pf[1] < pf[0]) ? 1 : 0, hi = 1 - lo; // ties: the first-stored point is "low"
            const double fl = pf[lo], xl = px[lo];
            double fh = pf[hi], xh = px[hi];
            fdiff = fh - fl;
            if (init_diam == 0)
                init_diam = std::fabs(xl - xh);
            if (std::fabs(xl - xh) < psi * init_diam)
                return SBPLX_XTOL_REACHED;
            const double c = xl; // centroid of the simplex without its worst point
            double xr, fr;
            if (!reflect(xr, c, 1.0, xh))
                return SBPLX_XTOL_REACHED;
            if (!eval(f, xr, fr, ret))
                return ret;
            if (fr < fl)
            { // expansion
                double xe, fe;
                if (!reflect(xe, c, 2.0, xh))
                    return SBPLX_XTOL_REACHED;
                if (!eval(f, xe, fe, ret))
                    return ret;
                if (fe >= fr)
                {
                    xh = xr;
                    fh = fr;
                }
                else
                {
                    xh = xe;
                    fh = fe;
                }
            }
            else
            { // contraction (inside if the worst point is still better than the reflection)
                double xc, fc;
                if (!reflect(xc, c, fh <= fr ? -0.5 : 0.5, xh))
                    return SBPLX_XTOL_REACHED;
                if (!eval(f, xc, fc, ret))
                    return ret;
                if (fc < fr && fc < fh)
                {
                    xh = xc;
                    fh = fc;
                }
                else
                { // shrink towards the best point
                    double xs;
                    if (!reflect(xs, xl, -0.5, xh))
                        return SBPLX_XTOL_REACHED;
                    xh = xs;
                    if (!eval(f, xh, fh, ret))
                        return ret;
                }
            }
            px[hi] = xh;
            pf[hi] = fh;
        }
    }

    template <typename F>
    int minimize(F &f, double x0, double step0)
    {
        const double psi = 0.25;
        nevals = 0;
        if (step0 == 0.0 || !(x0 >= lb && x0 <= ub) || !(lb < ub))
            return SBPLX_INVALID_ARGS; // NLopt rejects a zero initial step
        xbest = x0;
        fbest = f(x0);
        nevals = 1;
        if (maxeval > 0 && nevals >= maxeval)
            return SBPLX_MAXEVAL_REACHED;
        double step = step0;
        for (;;)
        {
            const double xprev = xbest;
            double fdiff;
            int ret = nelder_mead(f, step, psi, fdiff);
            if (ret == SBPLX_FAILURE)
                return SBPLX_XTOL_REACHED;
            if (ret != SBPLX_XTOL_REACHED)
                return ret;
            const double fdiff_max = fdiff > 0 ? fdiff : 0;
            if (rel_stop(fbest + fdiff_max, fbest, ftol_rel, 0.0))
                return SBPLX_FTOL_REACHED;
            if (rel_stop(xprev, xbest, 0.0, xtol_abs) && !(std::fabs(step) * psi > xtol_abs))
                return SBPLX_XTOL_REACHED;
            const double dx = xbest - xprev;
            step = (dx == 0) ? -(step * psi) : std::copysign(step * psi, dx);
        }
    }
};

} // namespace pgs
