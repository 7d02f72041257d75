// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of the PGURE-SVT denoising hot path of tjof2/pgure-svt v0.6.4, written from
// the behaviour of the reference sources (cited per function as file:line relative to
// /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (pgure-svt_b200/) never links,
// imports or calls it.
//
// Pinning status (see DESIGN.md "Oracle"):
//   * pcg64 + Bernoulli perturbations : pinned against the reference's own vendored
//     src/pcg/pcg_random.hpp compiled into oracle/_ref (tests/test_oracle_ref.py) and the
//     golden vectors in SURVEY.md §8c.
//   * median prefilter                : pinned against the reference's src/medfilter.hpp
//     (ConstantTimeMedianFilter) compiled into oracle/_ref.
//   * per-patch SVD                   : LAPACK dgesdd (the routine arma::svd_econ calls),
//     taken from the OpenBLAS bundled with scipy at run time; an in-file one-sided Jacobi is
//     the fallback and cross-check.
//   * NLopt LN_SBPLX (n=1), Armadillo reduction orders, noise.hpp/arps.hpp/svt.hpp/pgure.hpp
//     as a whole: PARITY UNPINNED — the reference cannot be built here (Armadillo, NLopt,
//     libtiff absent) and its own tests pin no hot-path numerics.
//
// FP policy: compiled with -ffp-contract=off; Armadillo's accu() order (two accumulators over
// even/odd linear indices, summed at the end) is followed where a result is order-sensitive.
//
// Layout: everything is column-major like Armadillo: cube(r,c,s) at r + n_rows*(c + n_cols*s).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <mutex>
#include <numeric>
#include <random>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// Armadillo-order reductions (arma arrayops::accumulate / accu_proxy_linear)
// ---------------------------------------------------------------------------------------------
template <typename F>
static inline double accu2(size_t n, F f)
{
    double v1 = 0.0, v2 = 0.0;
    size_t i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2)
    {
        v1 += f(i);
        v2 += f(j);
    }
    if (i < n)
        v1 += f(i);
    return v1 + v2;
}

static double arma_median(std::vector<double> v) // arma::median → op_median::direct_median
{
    const size_t n = v.size();
    if (n == 0)
        return NAN;
    const size_t half = n / 2;
    std::nth_element(v.begin(), v.begin() + half, v.end());
    const double val1 = v[half];
    if (n % 2 == 0)
    {
        const double val2 = *std::max_element(v.begin(), v.begin() + half);
        return val1 + (val2 - val1) / 2.0; // op_mean::robust_mean(A,B)
    }
    return val1;
}

// ---------------------------------------------------------------------------------------------
// LAPACK dgesdd through dlopen (arma::svd_econ → dgesdd JOBZ='S', svt.hpp:111)
// ---------------------------------------------------------------------------------------------
typedef void (*dgesdd_fn)(const char *, const int *, const int *, double *, const int *, double *, double *,
                          const int *, double *, const int *, double *, const int *, int *, int *, size_t);
static dgesdd_fn g_dgesdd = nullptr;
static int g_svd_backend = 0; // 0 = in-file Jacobi, 1 = LAPACK dgesdd

extern "C" int orc_set_lapack(const char *path)
{
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h)
        return -1;
    const char *names[] = {"scipy_dgesdd_", "dgesdd_", "dgesdd_64_", nullptr};
    for (int i = 0; names[i]; i++)
    {
        void *s = dlsym(h, names[i]);
        if (s)
        {
            g_dgesdd = (dgesdd_fn)s;
            break;
        }
    }
    if (!g_dgesdd)
        return -2;
    typedef void (*setthr_fn)(int);
    const char *tn[] = {"scipy_openblas_set_num_threads", "openblas_set_num_threads", nullptr};
    for (int i = 0; tn[i]; i++)
    {
        void *s = dlsym(h, tn[i]);
        if (s)
        {
            ((setthr_fn)s)(1);
            break;
        }
    }
    g_svd_backend = 1;
    return 0;
}
extern "C" void orc_set_svd_backend(int b) { g_svd_backend = (b == 1 && g_dgesdd) ? 1 : 0; }
extern "C" int orc_get_svd_backend() { return g_svd_backend; }

// One-sided (Hestenes) Jacobi thin SVD, m >= n not required.  A is m x n column-major (destroyed).
// Outputs U (m x n), S (n) descending, V (n x n).  Fallback + cross-check for dgesdd.
static void jacobi_svd(int m, int n, double *A, double *U, double *S, double *V)
{
    std::vector<double> W(A, A + (size_t)m * n), Vv((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
        Vv[i + (size_t)i * n] = 1.0;
    const double tol = 1e-15;
    for (int sweep = 0; sweep < 60; sweep++)
    {
        bool rotated = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++)
            {
                double a = 0, b = 0, g = 0;
                double *wp = &W[(size_t)p * m], *wq = &W[(size_t)q * m];
                for (int i = 0; i < m; i++)
                {
                    a += wp[i] * wp[i];
                    b += wq[i] * wq[i];
                    g += wp[i] * wq[i];
                }
                if (g == 0.0 || std::fabs(g) <= tol * std::sqrt(a * b))
                    continue;
                rotated = true;
                const double zeta = (b - a) / (2.0 * g);
                const double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; i++)
                {
                    const double x = wp[i], y = wq[i];
                    wp[i] = c * x - s * y;
                    wq[i] = s * x + c * y;
                }
                double *vp = &Vv[(size_t)p * n], *vq = &Vv[(size_t)q * n];
                for (int i = 0; i < n; i++)
                {
                    const double x = vp[i], y = vq[i];
                    vp[i] = c * x - s * y;
                    vq[i] = s * x + c * y;
                }
            }
        if (!rotated)
            break;
    }
    std::vector<double> sig(n);
    std::vector<int> idx(n);
    for (int j = 0; j < n; j++)
    {
        double a = 0;
        for (int i = 0; i < m; i++)
            a += W[i + (size_t)j * m] * W[i + (size_t)j * m];
        sig[j] = std::sqrt(a);
        idx[j] = j;
    }
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return sig[x] > sig[y]; });
    for (int r = 0; r < n; r++)
    {
        const int j = idx[r];
        S[r] = sig[j];
        for (int i = 0; i < m; i++)
            U[i + (size_t)r * m] = (sig[j] > 0) ? W[i + (size_t)j * m] / sig[j] : 0.0;
        for (int i = 0; i < n; i++)
            V[i + (size_t)r * n] = Vv[i + (size_t)j * n];
    }
}

// thin SVD of an m x n block (m >= n assumed by the reference, pguresvt.hpp:57)
static void svd_econ(int m, int n, const double *block, double *U, double *S, double *V)
{
    std::vector<double> A(block, block + (size_t)m * n);
    if (g_svd_backend == 1 && g_dgesdd)
    {
        const int mn = std::min(m, n);
        std::vector<double> VT((size_t)mn * n), Uu((size_t)m * mn), Ss(mn);
        std::vector<int> iwork(8 * mn);
        int info = 0, lwork = -1;
        double wq = 0;
        const char jobz = 'S';
        g_dgesdd(&jobz, &m, &n, A.data(), &m, Ss.data(), Uu.data(), &m, VT.data(), &mn, &wq, &lwork, iwork.data(),
                 &info, 1);
        lwork = (int)wq + 16;
        std::vector<double> work(lwork);
        g_dgesdd(&jobz, &m, &n, A.data(), &m, Ss.data(), Uu.data(), &m, VT.data(), &mn, work.data(), &lwork,
                 iwork.data(), &info, 1);
        // arma: U m x mn, S mn, V = VT' (n x mn)
        for (int k = 0; k < mn; k++)
        {
            S[k] = Ss[k];
            for (int i = 0; i < m; i++)
                U[i + (size_t)k * m] = Uu[i + (size_t)k * m];
            for (int i = 0; i < n; i++)
                V[i + (size_t)k * n] = VT[k + (size_t)i * mn];
        }
        return;
    }
    jacobi_svd(m, n, A.data(), U, S, V);
}

extern "C" void orc_svd(int m, int n, const double *A, double *U, double *S, double *V) { svd_econ(m, n, A, U, S, V); }

// ---------------------------------------------------------------------------------------------
// pcg64 (setseq_xsl_rr_128_64, default stream) + libstdc++ bernoulli_distribution
// src/pcg/pcg_random.hpp:166-169 (constants), :501-506,539-543 (seeding), :427-451 (step),
// :1085-1113 (XSL-RR output).  SURVEY §11 closed form.
// ---------------------------------------------------------------------------------------------
struct Pcg64
{
    u128 state;
    static u128 MULT() { return ((u128)2549297995355413924ULL << 64) | 4865540595714422341ULL; }
    static u128 INC() { return ((u128)6364136223846793005ULL << 64) | 1442695040888963407ULL; }
    void seed(uint64_t s) { state = ((u128)s + INC()) * MULT() + INC(); }
    uint64_t next()
    {
        state = state * MULT() + INC();
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const unsigned rot = (unsigned)(state >> 122);
        const uint64_t x = hi ^ lo;
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
};
// libstdc++ 13: bernoulli_distribution(p)(g) == (generate_canonical<double,53>(g) < p); with a 64-bit
// engine one draw, u = double(raw) * 2^-64, clamped below 1.
static inline bool bernoulli(Pcg64 &g, double p)
{
    double u = (double)g.next() * 5.42101086242752217e-20;
    if (u >= 1.0)
        u = std::nextafter(1.0, 0.0);
    return u < p;
}

extern "C" void orc_pcg64_raw(int64_t seed, uint64_t *out, int n)
{
    Pcg64 g;
    g.seed((uint64_t)seed);
    for (int i = 0; i < n; i++)
        out[i] = g.next();
}

// pgure.hpp:167-186: all of delta1 first (column-major imbue), then delta2.
extern "C" void orc_perturbations(int64_t seed, int64_t n, int64_t *delta1, double *delta2)
{
    Pcg64 g;
    g.seed((uint64_t)seed);
    const double kappa = 1.;
    const double vP = 0.5 + 0.5 * kappa / std::sqrt(kappa * kappa + 4);
    const double vQ = 1 - vP;
    const double vQvP = std::sqrt(vQ / vP);
    const double vPvQ = std::sqrt(vP / vQ);
    for (int64_t i = 0; i < n; i++)
        delta1[i] = bernoulli(g, 0.5) ? -1 : 1;
    for (int64_t i = 0; i < n; i++)
        delta2[i] = bernoulli(g, vP) ? -1 * vQvP : vPvQ;
}

// ---------------------------------------------------------------------------------------------
// Median prefilter: (2r+1)^2 clamp-to-edge median of uint16 — what ConstantTimeMedianFilter
// computes (medfilter.hpp:247-438,478-539); verified against oracle/_ref.
// ---------------------------------------------------------------------------------------------
extern "C" void orc_median_u16(const uint16_t *src, uint16_t *dst, int n_rows, int n_cols, int r)
{
    std::vector<uint16_t> win((size_t)(2 * r + 1) * (2 * r + 1));
    for (int c = 0; c < n_cols; c++)
        for (int rr = 0; rr < n_rows; rr++)
        {
            size_t k = 0;
            for (int dc = -r; dc <= r; dc++)
                for (int dr = -r; dr <= r; dr++)
                {
                    const int y = std::min(std::max(rr + dr, 0), n_rows - 1);
                    const int x = std::min(std::max(c + dc, 0), n_cols - 1);
                    win[k++] = src[y + (size_t)n_rows * x];
                }
            std::nth_element(win.begin(), win.begin() + k / 2, win.begin() + k);
            dst[rr + (size_t)n_rows * c] = win[k / 2];
        }
}

// ---------------------------------------------------------------------------------------------
// ARPS motion estimation — arps.hpp:21-375
// ---------------------------------------------------------------------------------------------
struct Arps
{
    const double *A; // (Nx rows, Ny cols, Nt) normalised, median-filtered window
    int Nx, Ny, Nt, bs, timeIter, timeWindow, mw, nImages;
    int nxMbs, nyMbs;
    int64_t vecSize;
    double OoBlockSizeSq;
    std::vector<int64_t> patches, motions; // (2, vecSize, 2tw+1), (2, vecSize, 2tw)
    long long nCost = 0;

    Arps(const double *A_, int N, int Nt_, int bs_, int timeIter_, int tw, int mw_, int nImages_)
        : A(A_), Nx(N), Ny(N), Nt(Nt_), bs(bs_), timeIter(timeIter_), timeWindow(tw), mw(mw_), nImages(nImages_)
    {
        nxMbs = Nx - bs;
        nyMbs = Ny - bs;
        OoBlockSizeSq = 1.0 / (bs * bs);
        vecSize = (int64_t)(1 + nxMbs) * (1 + nyMbs);
        patches.assign((size_t)2 * vecSize * (2 * tw + 1), 0);
        motions.assign((size_t)2 * vecSize * (2 * tw), 0);
    }
    int64_t &P(int d, int64_t it, int s) { return patches[d + 2 * (it + vecSize * (size_t)s)]; }
    int64_t &M(int d, int64_t it, int s) { return motions[d + 2 * (it + vecSize * (size_t)s)]; }

    // arps.hpp:148-151  accu(square(A - B)) * OoBlockSizeSq over the bs x bs block, column-major
    double Cost(int ry, int rx, int f1, int py, int px, int f2)
    {
        nCost++;
        const double *a = A + (size_t)Nx * Ny * f1, *b = A + (size_t)Nx * Ny * f2;
        const int n = bs * bs;
        return accu2(n, [&](size_t e) {
                   const int r = (int)(e % bs), c = (int)(e / bs);
                   const double d = a[(ry + r) + (size_t)Nx * (rx + c)] - b[(py + r) + (size_t)Nx * (px + c)];
                   return d * d;
               }) *
               OoBlockSizeSq;
    }

    void Pair(int curFrame, int f1, int f2, int f3) // arps.hpp:153-375
    {
        (void)curFrame;
        const int W = 2 * mw + 1;
        std::vector<uint8_t> checkMat((size_t)W * W);
        const int SD[5][2] = {{0, -1}, {-1, 0}, {0, 0}, {1, 0}, {0, 1}};
        for (int64_t it = 0; it < vecSize; it++)
        {
            double costs[6];
            int LD[6][2];
            for (int k = 0; k < 6; k++)
                costs[k] = 1E9, LD[k][0] = LD[k][1] = 0;
            std::fill(checkMat.begin(), checkMat.end(), 0);
            const int i = (int)(it % (1 + nxMbs)), j = (int)(it / (1 + nyMbs));
            int x = j, y = i;
            costs[2] = Cost(i, j, f1, i, j, f2);
            checkMat[mw + (size_t)W * mw] = 1;
            int maxIdx, stepSize;
            if (j == 0)
            {
                stepSize = 2;
                maxIdx = 5;
            }
            else
            {
                const int yTmp = (int)std::llabs(M(0, it, f3)), xTmp = (int)std::llabs(M(1, it, f3));
                stepSize = (xTmp <= yTmp) ? yTmp : xTmp;
                if (((yTmp == 0) && (xTmp == stepSize)) || ((xTmp == 0) && (yTmp == stepSize)))
                    maxIdx = 5;
                else
                {
                    maxIdx = 6;
                    LD[5][0] = (int)M(1, it, f3);
                    LD[5][1] = (int)M(0, it, f3);
                }
            }
            LD[0][0] = 0, LD[0][1] = -stepSize;
            LD[1][0] = -stepSize, LD[1][1] = 0;
            LD[2][0] = 0, LD[2][1] = 0;
            LD[3][0] = stepSize, LD[3][1] = 0;
            LD[4][0] = 0, LD[4][1] = stepSize;
            for (int k = 0; k < maxIdx; k++) // LDSP
            {
                const int ver = y + LD[k][1], hor = x + LD[k][0];
                const bool skip = (k == 2) || (stepSize == 0) || (hor < 0) || (ver < 0) || (hor + bs - 1) >= Ny ||
                                  (ver + bs - 1) >= Nx;
                if (!skip)
                {
                    costs[k] = Cost(i, j, f1, ver, hor, f2);
                    // arma bounds check would throw if |LD| > mw; callers keep stepSize <= mw
                    const int cy = LD[k][1] + mw, cx = LD[k][0] + mw;
                    if (cy >= 0 && cy < W && cx >= 0 && cx < W)
                        checkMat[cy + (size_t)W * cx] = 1;
                }
            }
            int point = 0; // find(costs == costs.min())(0): first index of the minimum
            for (int k = 1; k < 6; k++)
                if (costs[k] < costs[point])
                    point = k;
            x += LD[point][0];
            y += LD[point][1];
            double cost = costs[point];
            for (int k = 0; k < 6; k++)
                costs[k] = 1E9;
            costs[2] = cost;
            bool done = false;
            uint32_t nSDSP = 0;
            do // SDSP
            {
                for (int k = 0; k < 5; k++)
                {
                    const int ver = y + SD[k][1], hor = x + SD[k][0];
                    bool skip = (k == 2) || (hor < 0) || (ver < 0) || (hor + bs - 1) >= Ny || (ver + bs - 1) >= Nx ||
                                (hor < j - mw) || (hor > j + mw) || (ver < i - mw) || (ver > i + mw);
                    if (!skip)
                        skip = checkMat[(y - i + SD[k][1] + mw) + (size_t)W * (x - j + SD[k][0] + mw)] == 1;
                    if (!skip)
                    {
                        costs[k] = Cost(i, j, f1, ver, hor, f2);
                        checkMat[(y - i + SD[k][1] + mw) + (size_t)W * (x - j + SD[k][0] + mw)] = 1;
                    }
                }
                point = 0;
                for (int k = 1; k < 6; k++)
                    if (costs[k] < costs[point])
                        point = k;
                cost = costs[point];
                if (point == 2 || nSDSP >= 1000000u)
                    done = true;
                else
                {
                    x += SD[point][0];
                    y += SD[point][1];
                    for (int k = 0; k < 6; k++)
                        costs[k] = 1E9;
                    costs[2] = cost;
                }
                nSDSP++;
            } while (!done);
            M(0, it, f3) = y - i;
            M(1, it, f3) = x - j;
            P(0, it, f2) = y;
            P(1, it, f2) = x;
        }
    }

    void Estimate(bool estimateMotion) // arps.hpp:52-134
    {
        const int tw = timeWindow;
        auto seed = [&](int s) {
            for (int64_t i = 0; i < vecSize; i++)
            {
                P(0, i, s) = i % (1 + nyMbs);
                P(1, i, s) = i / (1 + nxMbs);
            }
        };
        if (timeIter < tw)
        {
            const int loopEnd = Nt - timeIter - 1;
            seed(timeIter);
            if (estimateMotion)
            {
                for (int i = 0; i < loopEnd; i++)
                    Pair(i, timeIter + i, timeIter + i + 1, timeIter + i);
                for (int i = 0; i < timeIter; i++)
                {
                    const int negInc = -1 * (i + 1);
                    Pair(negInc, timeIter + negInc + 1, timeIter + negInc, timeIter + negInc + 1);
                }
            }
        }
        else if (timeIter >= (nImages - tw))
        {
            const int endFrame = timeIter - (nImages - Nt);
            const int loopEnd = 2 * tw - endFrame;
            seed(endFrame);
            if (estimateMotion)
            {
                for (int i = 0; i < loopEnd; i++)
                    Pair(i, endFrame + i, endFrame + i + 1, endFrame + i);
                for (int i = 0; i < endFrame; i++)
                {
                    const int negInc = -1 * (i + 1);
                    if (2 * tw == endFrame)
                        Pair(negInc, endFrame + negInc + 1, endFrame + negInc, endFrame + negInc);
                    else
                        Pair(negInc, endFrame + negInc + 1, endFrame + negInc, endFrame + negInc + 1);
                }
            }
        }
        else
        {
            seed(tw);
            if (estimateMotion)
            {
                for (int i = 0; i < tw; i++)
                    Pair(i, tw + i, tw + i + 1, tw + i);
                for (int i = 0; i < tw; i++)
                {
                    const int negInc = -1 * (i + 1);
                    Pair(negInc, tw + negInc + 1, tw + negInc, tw + negInc + 1);
                }
            }
        }
    }
};

// w: (N,N,Nt) normalised window; patches out: int64 (2, vecSize, Nt); motions out (2, vecSize, Nt-1) or NULL
extern "C" long long orc_arps(const double *w, int N, int Nt, int bs, int timeIter, int timeWindow, int motionWindow,
                              int nImages, int estimateMotion, int64_t *patches, int64_t *motions)
{
    Arps a(w, N, Nt, bs, timeIter, timeWindow, motionWindow, nImages);
    a.Estimate(estimateMotion != 0);
    std::memcpy(patches, a.patches.data(), a.patches.size() * sizeof(int64_t));
    if (motions)
        std::memcpy(motions, a.motions.data(), a.motions.size() * sizeof(int64_t));
    return a.nCost;
}

// ---------------------------------------------------------------------------------------------
// SVT — svt.hpp:24-167
// ---------------------------------------------------------------------------------------------
struct Svt
{
    const int64_t *patches; // (2, vecSizeAll, Nt)
    int Nx, Ny, Nt, bs, bo;
    bool expW;
    int nxMbs, nyMbs;
    int64_t vecSizeAll;
    std::vector<int64_t> actual; // sorted unique patch ids (svt.hpp:61-92)
    std::vector<double> U, S, V; // per patch: m*n, n, n*n
    int m, n;

    Svt(const int64_t *p, int Nx_, int Ny_, int Nt_, int bs_, int bo_, bool e)
        : patches(p), Nx(Nx_), Ny(Ny_), Nt(Nt_), bs(bs_), bo(bo_), expW(e)
    {
        nxMbs = Nx - bs;
        nyMbs = Ny - bs;
        vecSizeAll = (int64_t)(1 + nxMbs) * (1 + nyMbs);
        m = bs * bs;
        n = Nt;
        std::vector<int64_t> ids;
        for (int64_t i = 0; i < 1 + nyMbs; i += bo)
            for (int64_t j = 0; j < 1 + nxMbs; j += bo)
                ids.push_back(i * nyMbs + j); // NB stride nyMbs, not nyMbs+1 (svt.hpp:68)
        for (int64_t i = 0; i < 1 + nyMbs; i += bo)
            ids.push_back((int64_t)(nyMbs + 1) * i + nxMbs); // "bottom edge" svt.hpp:78
        for (int64_t i = 0; i < 1 + nxMbs; i += bo)
            ids.push_back((int64_t)(nyMbs + 1) * nxMbs + i); // "right edge" svt.hpp:84
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        actual = ids;
    }
    inline int64_t PY(int64_t id, int k) const { return patches[0 + 2 * (id + vecSizeAll * (size_t)k)]; }
    inline int64_t PX(int64_t id, int k) const { return patches[1 + 2 * (id + vecSizeAll * (size_t)k)]; }

    void Decompose(const double *u) // svt.hpp:58-118
    {
        const size_t P = actual.size();
        U.resize(P * m * n);
        S.resize(P * n);
        V.resize(P * n * n);
        std::vector<double> block((size_t)m * n);
        for (size_t it = 0; it < P; it++)
        {
            for (int k = 0; k < n; k++)
            {
                const int64_t y = PY(actual[it], k), x = PX(actual[it], k);
                for (int c = 0; c < bs; c++)
                    for (int r = 0; r < bs; r++)
                        block[(r + bs * c) + (size_t)m * k] = u[(y + r) + (size_t)Nx * ((x + c) + (size_t)Ny * k)];
            }
            svd_econ(m, n, block.data(), &U[it * m * n], &S[it * n], &V[it * n * n]);
        }
    }

    // svt.hpp:121-167.  v: (Nx,Ny,Nt)
    void Reconstruct(double lambda, double *v) const
    {
        const size_t tot = (size_t)Nx * Ny * Nt;
        std::vector<double> weights(tot, 0.0), block((size_t)m * n), US((size_t)m * n), thr(n);
        std::fill(v, v + tot, 0.0);
        const size_t P = actual.size();
        for (size_t it = 0; it < P; it++)
        {
            const double *Ub = &U[it * m * n], *Sb = &S[it * n], *Vb = &V[it * n * n];
            if (expW)
            {
                double smax = Sb[0];
                for (int k = 1; k < n; k++)
                    smax = std::max(smax, Sb[k]);
                for (int k = 0; k < n; k++)
                {
                    const double w = std::fabs(smax * std::exp(-0.5 * lambda * (Sb[k] * Sb[k])));
                    const double sg = (Sb[k] > 0) - (Sb[k] < 0);
                    thr[k] = sg * std::max(std::fabs(Sb[k]) - w, 0.0); // utils.hpp:96-106
                }
            }
            else
                for (int k = 0; k < n; k++)
                {
                    const double sg = (Sb[k] > 0) - (Sb[k] < 0);
                    thr[k] = sg * std::max(std::fabs(Sb[k]) - lambda, 0.0);
                }
            // block = U * diagmat(thr) * V.t()   (svt.hpp:146)
            for (int k = 0; k < n; k++)
                for (int i = 0; i < m; i++)
                    US[i + (size_t)m * k] = Ub[i + (size_t)m * k] * thr[k];
            for (int j = 0; j < n; j++)
                for (int i = 0; i < m; i++)
                {
                    double acc = 0.0;
                    for (int k = 0; k < n; k++)
                        acc += US[i + (size_t)m * k] * Vb[j + (size_t)n * k];
                    block[i + (size_t)m * j] = acc;
                }
            for (int k = 0; k < n; k++)
            {
                const int64_t y = PY(actual[it], k), x = PX(actual[it], k);
                for (int c = 0; c < bs; c++)
                    for (int r = 0; r < bs; r++)
                    {
                        const size_t o = (y + r) + (size_t)Nx * ((x + c) + (size_t)Ny * k);
                        v[o] += block[(r + bs * c) + (size_t)m * k];
                        weights[o] += 1.0;
                    }
            }
        }
        for (size_t i = 0; i < tot; i++)
        {
            v[i] /= weights[i];
            if (!std::isfinite(v[i]))
                v[i] = 0.0;
        }
    }
};

extern "C" void *orc_svt_new(const int64_t *patches, int N, int Nt, int bs, int bo, int expW)
{
    return new Svt(patches, N, N, Nt, bs, bo, expW != 0);
}
extern "C" void orc_svt_free(void *h) { delete (Svt *)h; }
extern "C" int64_t orc_svt_npatches(void *h) { return (int64_t)((Svt *)h)->actual.size(); }
extern "C" void orc_svt_patch_ids(void *h, int64_t *out)
{
    Svt *s = (Svt *)h;
    std::copy(s->actual.begin(), s->actual.end(), out);
}
extern "C" void orc_svt_decompose(void *h, const double *u) { ((Svt *)h)->Decompose(u); }
extern "C" void orc_svt_reconstruct(void *h, double lambda, double *v) { ((Svt *)h)->Reconstruct(lambda, v); }
extern "C" void orc_svt_singular_values(void *h, double *S)
{
    Svt *s = (Svt *)h;
    std::copy(s->S.begin(), s->S.end(), S);
}

// ---------------------------------------------------------------------------------------------
// NLopt 2.6.2 LN_SBPLX restated for n = 1 (sbplx.c + nldrmd.c; PARITY UNPINNED — third-party,
// not vendored by the reference).  Call site: pgure.hpp:196-237.
// ---------------------------------------------------------------------------------------------
struct SbplxResult
{
    double x, minf;
    int status, nevals;
};

static inline bool nl_close(double a, double b) { return std::fabs(a - b) <= 1e-13 * (std::fabs(a) + std::fabs(b)); }
static inline bool nl_relstop(double vold, double vnew, double reltol, double abstol)
{
    if (std::isinf(vold))
        return false;
    return (std::fabs(vnew - vold) < abstol || std::fabs(vnew - vold) < reltol * (std::fabs(vnew) + std::fabs(vold)) * 0.5 ||
            (reltol > 0 && vnew == vold));
}
// reflectpt(): xnew = c + scale*(c - xold) pinned to [lb,ub]; returns false if coincident with c or xold
static inline bool nl_reflect(double &xnew, double c, double scale, double xold, double lb, double ub)
{
    double nx = c + scale * (c - xold);
    if (nx < lb)
        nx = lb;
    if (nx > ub)
        nx = ub;
    const bool equalc = nl_close(nx, c), equalold = nl_close(nx, xold);
    xnew = nx;
    return !(equalc || equalold);
}

enum
{
    NL_FAILURE = -1,
    NL_INVALID = -2,
    NL_SUCCESS = 1,
    NL_FTOL = 3,
    NL_XTOL = 4,
    NL_MAXEVAL = 5
};

struct NlStop
{
    double ftol_rel, xtol_abs;
    int maxeval, nevals;
};

// nldrmd_minimize_ with psi > 0, n = 1.  x/minf in-out (best point so far).
static int nldrmd1(const std::function<double(double)> &f, double lb, double ub, double &x, double &minf, double xstep,
                   NlStop &stop, double psi, double &fdiff)
{
    // pts[0] = start point, pts[1] = start + step (addresses order the tie-break: pts[0] < pts[1])
    double px[2], pf[2];
    fdiff = HUGE_VAL;
    px[0] = x;
    pf[0] = minf;
    px[1] = x + xstep;
    if (px[1] > ub)
    {
        if (ub - x > std::fabs(xstep) * 0.1)
            px[1] = ub;
        else
            px[1] = x - std::fabs(xstep);
    }
    if (px[1] < lb)
    {
        if (x - lb > std::fabs(xstep) * 0.1)
            px[1] = lb;
        else
        {
            px[1] = x + std::fabs(xstep);
            if (px[1] > ub)
                px[1] = 0.5 * ((ub - x > x - lb ? ub : lb) + x);
        }
    }
    if (nl_close(px[1], x))
        return NL_FAILURE;
#define CHECK_EVAL(xc, fc)      \
    stop.nevals++;              \
    if ((fc) <= minf)           \
    {                           \
        minf = (fc);            \
        x = (xc);               \
    }                           \
    if (stop.maxeval > 0 && stop.nevals >= stop.maxeval) \
        return NL_MAXEVAL;
    pf[1] = f(px[1]);
    CHECK_EVAL(px[1], pf[1]);
    double init_diam = 0;
    while (true)
    {
        // low/high by (f, address)
        int lo, hi;
        if (pf[0] < pf[1] || (pf[0] == pf[1]))
            lo = 0, hi = 1; // tie → lower address is "smaller"
        else
            lo = 1, hi = 0;
        const double fl = pf[lo], xl = px[lo];
        double fh = pf[hi], xh = px[hi];
        fdiff = fh - fl;
        if (init_diam == 0)
            init_diam += std::fabs(xl - xh);
        const double c = xl; // centroid of all points but xh, n = 1 (ninv = 1)
        {
            const double diam = std::fabs(xl - xh);
            if (diam < psi * init_diam)
                return NL_XTOL;
        }
        double xcur;
        if (!nl_reflect(xcur, c, 1.0, xh, lb, ub))
            return NL_XTOL;
        const double fr = f(xcur);
        CHECK_EVAL(xcur, fr);
        if (fr < fl)
        { // expand
            if (!nl_reflect(xh, c, 2.0, xh, lb, ub))
            {
                px[hi] = xh;
                return NL_XTOL;
            }
            fh = f(xh);
            CHECK_EVAL(xh, fh);
            if (fh >= fr)
            {
                fh = fr;
                xh = xcur;
            }
        }
        else if (fr < fl) // rb_tree_pred(high) == low for two points: never taken
        {
            xh = xcur;
            fh = fr;
        }
        else
        { // contract
            if (!nl_reflect(xcur, c, fh <= fr ? -0.5 : 0.5, xh, lb, ub))
                return NL_XTOL;
            const double fc = f(xcur);
            CHECK_EVAL(xcur, fc);
            if (fc < fr && fc < fh)
            {
                xh = xcur;
                fh = fc;
            }
            else
            { // shrink towards xl
                double np;
                if (!nl_reflect(np, xl, -0.5, xh, lb, ub))
                {
                    px[hi] = np;
                    return NL_XTOL;
                }
                xh = np;
                fh = f(xh);
                CHECK_EVAL(xh, fh);
            }
        }
        px[hi] = xh;
        pf[hi] = fh;
    }
#undef CHECK_EVAL
}

static SbplxResult sbplx1(const std::function<double(double)> &f, double x0, double lb, double ub, double xstep0,
                          double ftol_rel, double xtol_abs, int maxeval)
{
    SbplxResult R;
    const double psi = 0.25;
    NlStop stop{ftol_rel, xtol_abs, maxeval, 0};
    double x = x0, xstep = xstep0;
    if (xstep0 == 0.0 || !(x0 >= lb && x0 <= ub))
    { // nlopt_set_initial_step(0) → NLOPT_INVALID_ARGS → std::invalid_argument
        R.x = x0;
        R.minf = NAN;
        R.status = NL_INVALID;
        R.nevals = 0;
        return R;
    }
    double minf = f(x);
    stop.nevals++;
    int ret = NL_SUCCESS;
    if (stop.maxeval > 0 && stop.nevals >= stop.maxeval)
        ret = NL_MAXEVAL;
    else
        while (true)
        {
            const double xprev = x;
            double fdiff;
            ret = nldrmd1(f, lb, ub, x, minf, xstep, stop, psi, fdiff);
            const double fdiff_max = (fdiff > 0) ? fdiff : 0;
            if (ret == NL_FAILURE)
            {
                ret = NL_XTOL;
                break;
            }
            if (ret != NL_XTOL)
                break;
            if (nl_relstop(minf + fdiff_max, minf, stop.ftol_rel, 0.0))
            {
                ret = NL_FTOL;
                break;
            }
            if (nl_relstop(xprev, x, 0.0, stop.xtol_abs))
            {
                if (!(std::fabs(xstep) * psi > stop.xtol_abs && std::fabs(xstep) * psi > 0.0 * std::fabs(x)))
                {
                    ret = NL_XTOL;
                    break;
                }
            }
            const double dx = x - xprev;
            const double scale = psi; // nsubs == 1
            xstep = (dx == 0) ? -(xstep * scale) : std::copysign(xstep * scale, dx);
        }
    R.x = x;
    R.minf = minf;
    R.status = ret;
    R.nevals = stop.nevals;
    return R;
}

typedef double (*orc_obj_fn)(double, void *);
extern "C" int orc_sbplx_1d(orc_obj_fn f, void *data, double x0, double lb, double ub, double step, double ftol_rel,
                            double xtol_abs, int maxeval, double *xout, double *fout, int *nevals)
{
    SbplxResult r = sbplx1([&](double x) { return f(x, data); }, x0, lb, ub, step, ftol_rel, xtol_abs, maxeval);
    *xout = r.x;
    *fout = r.minf;
    *nevals = r.nevals;
    return r.status;
}

// ---------------------------------------------------------------------------------------------
// PGURE — pgure.hpp:26-237
// ---------------------------------------------------------------------------------------------
static int g_eps1_mode = 0; // 0 = as the reference computes it (integer-truncated eps1), 1 = intended
extern "C" void orc_set_eps1_mode(int v) { g_eps1_mode = v; }
struct Pgure
{
    std::vector<double> U;
    const int64_t *patches;
    double alpha, mu, sigma, sigmasq;
    int Nx, Ny, Nt, bs, bo;
    int64_t seed;
    bool expW, opt;
    double OoNxNyNt, eps1 = 0, eps2 = 0, lambda = 0;
    Svt *svt0 = nullptr, *svt1 = nullptr, *svt2p = nullptr, *svt2m = nullptr;
    std::vector<double> Uhat, U1, U2p, U2m, delta2;
    std::vector<int64_t> delta1;
    int nevals = 0, status = 0;

    Pgure(const double *U_, const int64_t *patches_, int N, int Nt_, double alpha_, double mu_, double sigma_, int bs_,
          int bo_, int64_t seed_, bool expW_, bool opt_)
        : patches(patches_), alpha(alpha_), mu(mu_), sigma(sigma_), Nx(N), Ny(N), Nt(Nt_), bs(bs_), bo(bo_), seed(seed_),
          expW(expW_), opt(opt_)
    {
        const size_t tot = (size_t)Nx * Ny * Nt;
        U.assign(U_, U_ + tot);
        OoNxNyNt = 1.0 / ((uint32_t)Nx * (uint32_t)Ny * (uint32_t)Nt);
        sigmasq = sigma * sigma;
        svt0 = new Svt(patches, Nx, Ny, Nt, bs, bo, expW);
        svt0->Decompose(U.data());
        if (opt)
        {
            double umax = U[0];
            for (size_t i = 1; i < tot; i++)
                umax = std::max(umax, U[i]);
            eps1 = umax * 0.0001;
            eps2 = 100 * eps1;
            delta1.resize(tot);
            delta2.resize(tot);
            int64_t s = seed;
            if (s < 0)
            {
                std::random_device rd;
                s = (int64_t)(((uint64_t)rd() << 32 | rd()) >> 1);
            }
            orc_perturbations(s, (int64_t)tot, delta1.data(), delta2.data());
            U1.resize(tot);
            U2p.resize(tot);
            U2m.resize(tot);
            Uhat.resize(tot);
            for (size_t i = 0; i < tot; i++)
            {
                // pgure.hpp:80  U1 = U + (delta1 * eps1) with delta1 an arma::icube: Armadillo's
                // `Cube<sword> * scalar` takes the scalar as sword, so eps1 (1e-4) is truncated to 0 and
                // U1 == U bit for bit (evidence: only this reading passes the
                // reference's own test_known_noise threshold).  g_eps1_mode = 1 gives the intended maths.
                U1[i] = g_eps1_mode ? U[i] + ((double)delta1[i] * eps1) : U[i] + (double)(delta1[i] * (int64_t)eps1);
                U2p[i] = U[i] + (delta2[i] * eps2);
                U2m[i] = U[i] - (delta2[i] * eps2);
            }
            svt1 = new Svt(patches, Nx, Ny, Nt, bs, bo, expW);
            svt2p = new Svt(patches, Nx, Ny, Nt, bs, bo, expW);
            svt2m = new Svt(patches, Nx, Ny, Nt, bs, bo, expW);
            svt1->Decompose(U1.data());
            svt2p->Decompose(U2p.data());
            svt2m->Decompose(U2m.data());
        }
    }
    ~Pgure()
    {
        delete svt0;
        delete svt1;
        delete svt2p;
        delete svt2m;
    }
    double Calculate(double x, double *terms = nullptr) // pgure.hpp:120-137
    {
        lambda = x;
        svt0->Reconstruct(lambda, Uhat.data());
        svt1->Reconstruct(lambda, U1.data());
        svt2p->Reconstruct(lambda, U2p.data());
        svt2m->Reconstruct(lambda, U2m.data());
        const size_t tot = (size_t)Nx * Ny * Nt;
        const double s1 = accu2(tot, [&](size_t i) { const double d = std::fabs(Uhat[i] - U[i]); return d * d; });
        const double s2 = accu2(tot, [&](size_t i) { return U[i]; });
        const double s3 = accu2(tot, [&](size_t i) {
            return ((double)delta1[i] * (alpha * U[i] - alpha * mu + sigmasq)) * (U1[i] - Uhat[i]);
        });
        const double s4 = accu2(tot, [&](size_t i) { return delta2[i] * (U2p[i] - 2 * Uhat[i] + U2m[i]); });
        const double s5 = accu2(tot, [&](size_t i) { return Uhat[i]; });
        if (terms)
        {
            terms[0] = s1, terms[1] = s2, terms[2] = s3, terms[3] = s4, terms[4] = s5;
        }
        return OoNxNyNt * (s1 - (alpha + mu) * s2 + (2 / eps1 * s3) - (2 * sigmasq * alpha / (eps2 * eps2) * s4) +
                           (2 * mu * s5) + mu) -
               sigmasq;
    }
    double Optimize(double tol, double start, double bound, int eval) // pgure.hpp:196-237
    {
        SbplxResult r = sbplx1([&](double x) { return Calculate(x); }, start, 0.0, bound, std::sqrt(start), tol, 1E-12, eval);
        nevals = r.nevals;
        status = r.status;
        if (r.status == NL_INVALID)
            lambda = start; // reference would throw; keep the start point
        return lambda; // LAST evaluated lambda
    }
};

extern "C" void *orc_pgure_new(const double *U, const int64_t *patches, int N, int Nt, double alpha, double mu,
                               double sigma, int bs, int bo, int64_t seed, int expW, int opt)
{
    return new Pgure(U, patches, N, Nt, alpha, mu, sigma, bs, bo, seed, expW != 0, opt != 0);
}
extern "C" void orc_pgure_free(void *h) { delete (Pgure *)h; }
extern "C" double orc_pgure_calc(void *h, double lambda, double *terms) { return ((Pgure *)h)->Calculate(lambda, terms); }
extern "C" double orc_pgure_optimize(void *h, double tol, double start, double bound, int maxeval, int *nevals, int *status)
{
    Pgure *p = (Pgure *)h;
    const double l = p->Optimize(tol, start, bound, maxeval);
    if (nevals)
        *nevals = p->nevals;
    if (status)
        *status = p->status;
    return l;
}
extern "C" void orc_pgure_reconstruct(void *h, double lambda, double *v) { ((Pgure *)h)->svt0->Reconstruct(lambda, v); }

// ---------------------------------------------------------------------------------------------
// Noise estimation — noise.hpp:35-458
// ---------------------------------------------------------------------------------------------
struct Noise
{
    uint32_t noiseMethod, size = 8, weightType = 0;
    long long nSplit = 0, nLeaves = 0, nIrls = 0;
    explicit Noise(uint32_t m) : noiseMethod(m) {}

    static double ftest0025(uint32_t N, bool &ok)
    {
        static const uint32_t dof[12] = {2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
        static const double f[12] = {15.4392, 2.86209, 1.64602, 1.27893, 1.13046, 1.06318,
                                     1.03110, 1.01543, 1.00769, 1.00384, 1.00192, 1.00096};
        for (int i = 0; i < 12; i++)
            if (dof[i] == N)
            {
                ok = true;
                return f[i];
            }
        ok = false;
        return 0;
    }

    // A: N x N column-major (ld = N).  noise.hpp:182-221
    bool SplitBlockQ(const std::vector<double> &A, uint32_t N)
    {
        nSplit++;
        if (N <= size)
            return false;
        const uint32_t l = 5;
        std::vector<double> resids((size_t)N * N);
        const double sc = std::sqrt((double)(l * l + l));
        for (uint32_t x = 0; x < N; x++)
            for (uint32_t y = 0; y < N; y++)
            {
                const int xp = ((x + 1) == N) ? 1 : (x + 1);
                const int yp = ((y + 1) == N) ? 1 : (y + 1);
                const int xm = (x == 0) ? (N - 2) : (x - 1);
                const int ym = (y == 0) ? (N - 2) : (y - 1);
                resids[y + (size_t)N * x] =
                    l * A[y + (size_t)N * x] -
                    (A[yp + (size_t)N * x] + A[ym + (size_t)N * x] + A[y + (size_t)N * xm] + A[y + (size_t)N * xp]);
            }
        for (auto &r : resids)
            r /= sc;
        const uint32_t R = N * N;
        const double OoR = 1.0 / R, OoRm1 = 1.0 / (R - 1);
        const double accuZ = accu2(R, [&](size_t i) { return A[i]; }) * OoR;
        const double Sz = accu2(R, [&](size_t i) { const double d = A[i] - accuZ; return d * d; }) * OoRm1;
        const double accuR = accu2(R, [&](size_t i) { return resids[i]; }) * OoR;
        const double Se = accu2(R, [&](size_t i) { const double d = resids[i] - accuR; return d * d; }) * OoRm1;
        const double stat = (Sz > Se) ? Sz / Se : Se / Sz;
        bool ok;
        const double value = ftest0025(N, ok);
        if (!ok)
            return false; // reference: out-of-bounds exception; Python restricts N to 2^k
        return stat > value;
    }

    static double Iqr(std::vector<double> v) // noise.hpp:223-230
    {
        std::sort(v.begin(), v.end());
        const uint32_t N = (uint32_t)v.size();
        const uint32_t m = (uint32_t)std::floor((std::floor((double)((N + 1) / 2)) + 1) / 2);
        return v[N - m - 1] - v[m - 1];
    }
    static double RobustVariance(const std::vector<double> &A) // noise.hpp:232-236
    {
        const double med = arma_median(A);
        std::vector<double> d(A.size());
        for (size_t i = 0; i < A.size(); i++)
            d[i] = std::fabs(A[i] - med);
        const double sig = 1.4826 * arma_median(d);
        return sig * sig;
    }
    void Weight(const std::vector<double> &x, std::vector<double> &w) const // noise.hpp:303-326
    {
        if (weightType == 0)
        {
            const double p = 0.75;
            for (size_t i = 0; i < x.size(); i++)
                w[i] = (std::fabs(x[i]) < p) ? 1. : p / std::fabs(x[i]);
        }
        else
        {
            const double p = 3.5, pp = 12.25;
            for (size_t i = 0; i < x.size(); i++)
                w[i] = (std::fabs(x[i]) > p) ? 0. : (pp - x[i] * x[i]) * (pp - x[i] * x[i]) / (pp * pp);
        }
    }
    double RobustMean(const std::vector<double> &A) // noise.hpp:238-271
    {
        const uint32_t I = 10000, N = (uint32_t)A.size();
        double e, tol = 1E-6, d, m = 0., m0 = 1E12, eps = 1E-12, aux;
        std::vector<double> w(N, 1.0), r(N);
        for (uint32_t i = 0; i < I; i++)
        {
            nIrls++;
            for (uint32_t k = 0; k < N; k++)
                r[k] = w[k] * A[k];
            m = accu2(N, [&](size_t k) { return r[k]; });
            aux = accu2(N, [&](size_t k) { return w[k]; });
            m = (std::fabs(aux) < eps) ? m0 : m / aux;
            for (uint32_t k = 0; k < N; k++)
                r[k] = A[k] - m;
            e = accu2(N, [&](size_t k) { return std::fabs(r[k]); }) / (double)N;
            if (std::fabs(m0 - m) < tol || e < tol)
                break;
            m0 = m;
            d = Iqr(r) + eps;
            const double inv = 1. / d;
            for (uint32_t k = 0; k < N; k++)
                r[k] *= inv;
            Weight(r, w);
        }
        return m;
    }
    // in: s x s column-major; out(y,x) per noise.hpp:395-417
    static std::vector<double> ConvolveFIR(const std::vector<double> &in, uint32_t N)
    {
        std::vector<double> out((size_t)N * N, 0.0);
        auto IN = [&](int a, int b) { return in[a + (size_t)N * b]; };
        for (uint32_t x = 0; x < N; x++)
            for (uint32_t y = 0; y < N; y++)
            {
                const int xp = ((x + 1) == N) ? 1 : (x + 1);
                const int yp = ((y + 1) == N) ? 1 : (y + 1);
                const int xm = (x == 0) ? (N - 2) : (x - 1);
                const int ym = (y == 0) ? (N - 2) : (y - 1);
                // neighbours (3x3, filled row-wise), accu over column-major with kernel -laplacian
                const double nb[3][3] = {{IN(xm, ym), IN(x, ym), IN(xp, ym)},
                                         {IN(xm, y), IN(x, y), IN(xp, y)},
                                         {IN(xm, yp), IN(x, yp), IN(xp, yp)}};
                double t[9];
                int e = 0;
                for (int c = 0; c < 3; c++)
                    for (int r = 0; r < 3; r++)
                    {
                        const double lap = (r == 1 && c == 1) ? -1.0 : 0.125;
                        t[e++] = nb[r][c] * (-1 * lap);
                    }
                out[y + (size_t)N * x] = accu2(9, [&](size_t i) { return t[i]; });
            }
        return out;
    }
    double ComputeMode(const std::vector<double> &A) const // noise.hpp:273-301
    {
        uint32_t maxCount = 0;
        double maxValue = 0.;
        const double M = *std::max_element(A.begin(), A.end());
        const uint32_t N = (uint32_t)A.size();
        const double dyn = 1. * N;
        std::vector<double> a(N);
        for (uint32_t i = 0; i < N; i++)
            a[i] = std::round(A[i] * dyn / M);
        for (uint32_t i = 0; i < N; i++)
        {
            uint32_t count = 0;
            for (uint32_t j = 0; j < N; j++)
                if (a[j] == a[i])
                    count++;
            if (count > maxCount)
            {
                maxCount = count;
                maxValue = a[i];
            }
        }
        maxValue *= M / dyn;
        return maxValue;
    }
    void WLSFit(const std::vector<double> &x, const std::vector<double> &y, double &p0, double &p1) // noise.hpp:328-383
    {
        const uint32_t I = 10000, N = (uint32_t)x.size();
        double e, tol = 1E-6, d, a0 = 1E12, b0 = 1E12, eps = 1E-12, aux, sw2, sw2x, sw2y;
        std::vector<double> w(N, 1.0), w2(N), r(N);
        p0 = p1 = 0;
        for (uint32_t i = 0; i < I; i++)
        {
            for (uint32_t k = 0; k < N; k++)
                w2[k] = w[k] * w[k];
            sw2 = accu2(N, [&](size_t k) { return w2[k]; });
            sw2x = accu2(N, [&](size_t k) { return w2[k] * x[k]; });
            sw2y = accu2(N, [&](size_t k) { return w2[k] * y[k]; });
            p0 = sw2 * accu2(N, [&](size_t k) { return w2[k] * (x[k] * y[k]); }) - sw2x * sw2y;
            aux = sw2 * accu2(N, [&](size_t k) { return w2[k] * (x[k] * x[k]); }) - sw2x * sw2x;
            p0 = (std::fabs(aux) < eps) ? a0 : p0 / aux;
            p1 = sw2y - p0 * sw2x;
            p1 = (std::fabs(aux) < eps) ? b0 : p1 / sw2;
            for (uint32_t k = 0; k < N; k++)
                r[k] = y[k] - (x[k] * p0 + p1);
            e = accu2(N, [&](size_t k) { return std::fabs(r[k]); }) / (double)N;
            if ((std::fabs(a0 - p0) < tol && std::fabs(b0 - p1) < tol) || e < tol)
                break;
            a0 = p0;
            b0 = p1;
            d = Iqr(r) + eps;
            for (uint32_t k = 0; k < N; k++)
                r[k] /= d;
            Weight(r, w);
        }
    }

    struct Node
    {
        uint32_t i, j, s;
    };
    std::vector<Node> tree;
    std::vector<uint32_t> dele;
    void QuadTree(const double *A, uint32_t N, uint32_t part) // noise.hpp:419-458
    {
        const Node nd = tree[part];
        std::vector<double> patch((size_t)nd.s * nd.s);
        for (uint32_t c = 0; c < nd.s; c++)
            for (uint32_t r = 0; r < nd.s; r++)
                patch[r + (size_t)nd.s * c] = A[(nd.i + r) + (size_t)N * (nd.j + c)];
        if (!SplitBlockQ(patch, nd.s))
            return;
        const uint32_t s = nd.s / 2;
        const uint32_t n = (uint32_t)tree.size() - 1;
        tree.push_back({nd.i, nd.j, s});
        tree.push_back({nd.i + s, nd.j, s});
        tree.push_back({nd.i, nd.j + s, s});
        tree.push_back({nd.i + s, nd.j + s, s});
        dele.push_back(part);
        uint32_t iter = n;
        do
        {
            QuadTree(A, N, iter);
            iter++;
        } while (iter < n + 4);
    }

    // kept node list for one slice (duplicates included), noise.hpp:55-72
    std::vector<Node> Leaves(const double *A, uint32_t N)
    {
        tree.clear();
        dele.clear();
        tree.push_back({0, 0, N});
        QuadTree(A, N, 0);
        std::vector<uint32_t> d = dele;
        std::sort(d.begin(), d.end());
        d.erase(std::unique(d.begin(), d.end()), d.end());
        std::vector<Node> kept = tree;
        if (!d.empty())
            for (size_t k = d.size() - 1; k > 0; k--)
                kept.erase(kept.begin() + d[k]);
        return kept;
    }

    void Estimate(const double *input, uint32_t N, uint32_t T, double &alpha_, double &mu_, double &sigma_)
    {
        double alpha = alpha_, mu = mu_, sigma = sigma_, dSi = 0.0;
        const uint32_t Nx = N, Ny = N;
        std::vector<double> means, vars;
        for (uint32_t i = 0; i < T; i++)
        {
            const double *A = input + (size_t)N * N * i;
            std::vector<Node> kept = Leaves(A, N);
            for (size_t n = 0; n < kept.size(); n++)
            {
                const uint32_t x = kept[n].i, y = kept[n].j, s = kept[n].s;
                std::vector<double> col((size_t)s * s);
                for (uint32_t c = 0; c < s; c++)
                    for (uint32_t r = 0; r < s; r++)
                        col[r + (size_t)s * c] = A[(x + r) + (size_t)N * (y + c)];
                nLeaves++;
                const double meanEst = RobustMean(col);
                std::vector<double> lap = ConvolveFIR(col, s);
                const double varEst = RobustVariance(lap);
                // means/vars are filtered independently with >= 0 (noise.hpp:103-104)
                if (meanEst >= 0.)
                    means.push_back(meanEst);
                if (varEst >= 0.)
                    vars.push_back(varEst);
            }
        }
        const size_t n = means.size();
        std::vector<size_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return means[a] < means[b]; });
        std::vector<double> rm(n), rv(n);
        for (size_t k = 0; k < n; k++)
        {
            rm[k] = means[idx[k]];
            rv[k] = vars[idx[k]];
        }
        double ab0, ab1;
        WLSFit(rm, rv, ab0, ab1);
        alpha = (alpha >= 0.) ? alpha : ab0;
        auto restrict_ = [&](const std::vector<double> &a, uint32_t Is, uint32_t Ie) {
            return std::vector<double>(a.begin() + Is, a.begin() + Ie + 1);
        };
        switch (noiseMethod)
        {
        case 1:
        {
            const int L = (int)std::floor(1. * (Nx * Ny / (uint32_t)n));
            mu = (mu >= 0.) ? mu : ComputeMode(restrict_(rm, 0, (uint32_t)std::round(0.05 * L)));
            dSi = ComputeMode(restrict_(rv, 0, (uint32_t)std::round(0.05 * L)));
            sigma = (sigma >= 0.) ? sigma : std::sqrt(dSi);
            break;
        }
        case 2:
            mu = (mu >= 0.) ? mu : ComputeMode(rm);
            dSi = ComputeMode(rv);
            sigma = (sigma >= 0.) ? sigma : std::sqrt(std::max(dSi, std::max(ab1 + ab0 * dSi, 0.)));
            break;
        case 3:
            mu = (mu >= 0.) ? mu : ComputeMode(rm);
            sigma = (sigma >= 0.) ? sigma : std::sqrt(std::fabs(ab1 + ab0 * mu));
            break;
        case 4:
        default:
            mu = (mu >= 0.) ? mu : rm[0];
            sigma = (sigma >= 0.) ? sigma : std::sqrt(std::fabs(ab1 + ab0 * mu));
            break;
        }
        alpha_ = alpha;
        mu_ = mu;
        sigma_ = sigma;
    }
};

extern "C" void orc_noise_estimate(const double *u, int N, int T, int method, double *alpha, double *mu, double *sigma,
                                   long long *stats)
{
    Noise ne((uint32_t)method);
    ne.Estimate(u, (uint32_t)N, (uint32_t)T, *alpha, *mu, *sigma);
    if (stats)
    {
        stats[0] = ne.nSplit;
        stats[1] = ne.nLeaves;
        stats[2] = ne.nIrls;
    }
}
// quadtree bookkeeping only, with an always-split / root-only-split predicate (golden counts)
extern "C" void orc_quadtree_counts(int N, int mode, int *created, int *kept)
{
    struct Q
    {
        std::vector<Noise::Node> tree;
        std::vector<uint32_t> dele;
        int mode;
        void go(uint32_t part)
        {
            Noise::Node nd = tree[part];
            const bool split = (nd.s > 8) && (mode == 0 || nd.s == tree[0].s);
            if (!split)
                return;
            const uint32_t s = nd.s / 2, n = (uint32_t)tree.size() - 1;
            tree.push_back({nd.i, nd.j, s});
            tree.push_back({nd.i + s, nd.j, s});
            tree.push_back({nd.i, nd.j + s, s});
            tree.push_back({nd.i + s, nd.j + s, s});
            dele.push_back(part);
            for (uint32_t it = n; it < n + 4; it++)
                go(it);
        }
    } q;
    q.mode = mode;
    q.tree.push_back({0, 0, (uint32_t)N});
    q.go(0);
    std::vector<uint32_t> d = q.dele;
    std::sort(d.begin(), d.end());
    d.erase(std::unique(d.begin(), d.end()), d.end());
    size_t k = q.tree.size();
    if (!d.empty())
        k -= d.size() - 1;
    *created = (int)q.tree.size();
    *kept = (int)k;
}

// ---------------------------------------------------------------------------------------------
// Hot-pixel filter — hotpixel.hpp:19-64 (uint16 instantiation used by the CLI; SURVEY Q22:
// interpretation chosen = median of the per-column medians, uint16 modular arithmetic, scalars
// truncated to uint16 before use).  PARITY UNPINNED (Armadillo-version dependent).
// ---------------------------------------------------------------------------------------------
static uint16_t median_u16(std::vector<uint16_t> v)
{
    const size_t n = v.size(), half = n / 2;
    std::nth_element(v.begin(), v.begin() + half, v.end());
    const uint16_t val1 = v[half];
    if (n % 2 == 0)
    {
        const uint16_t val2 = *std::max_element(v.begin(), v.begin() + half);
        return (uint16_t)(val1 + (val2 - val1) / 2);
    }
    return val1;
}
static uint16_t median_of_col_medians(const uint16_t *f, int Nx, int Ny)
{
    std::vector<uint16_t> cm(Ny);
    for (int c = 0; c < Ny; c++)
        cm[c] = median_u16(std::vector<uint16_t>(f + (size_t)Nx * c, f + (size_t)Nx * (c + 1)));
    return median_u16(cm);
}
extern "C" void orc_hotpixel_u16(uint16_t *seq, int Nx, int Ny, int Nt, double threshold)
{
    const double mad_scale = 1.0 / 0.6745;
    for (int i = 0; i < Nt; i++)
    {
        uint16_t *f = seq + (size_t)Nx * Ny * i;
        const double median = (double)median_of_col_medians(f, Nx, Ny);
        const uint16_t med16 = (uint16_t)median;
        std::vector<uint16_t> dev((size_t)Nx * Ny);
        for (size_t k = 0; k < dev.size(); k++)
            dev[k] = (uint16_t)(f[k] - med16);
        const double mad = (double)median_of_col_medians(dev.data(), Nx, Ny) * mad_scale;
        const double tv = threshold * mad;
        const uint16_t thr16 = (tv >= 65535.0) ? 65535 : (uint16_t)tv;
        std::vector<size_t> outliers;
        for (size_t k = 0; k < dev.size(); k++)
            if (dev[k] > thr16)
                outliers.push_back(k);
        for (size_t k : outliers)
        {
            const int r = (int)(k % Nx), c = (int)(k / Nx);
            if (r > 0 && r < Nx - 1 && c > 0 && c < Ny - 1)
            {
                double w[8] = {(double)f[(r - 1) + (size_t)Nx * (c - 1)], (double)f[(r - 1) + (size_t)Nx * c],
                               (double)f[(r - 1) + (size_t)Nx * (c + 1)], (double)f[r + (size_t)Nx * (c - 1)],
                               (double)f[r + (size_t)Nx * (c + 1)],       (double)f[(r + 1) + (size_t)Nx * (c - 1)],
                               (double)f[(r + 1) + (size_t)Nx * c],       (double)f[(r + 1) + (size_t)Nx * (c + 1)]};
                std::sort(w, w + 8);
                f[k] = (uint16_t)(0.5 * (w[3] + w[4]));
            }
            else
                f[k] = (uint16_t)median;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Driver — pguresvt.hpp:17-172 with utils.hpp:108-168 thread fan-out
// ---------------------------------------------------------------------------------------------
template <typename Func>
static void parallel(const Func &func, uint32_t first, uint32_t last, int nJobs) // utils.hpp:108-168
{
    const uint32_t totalCores = (nJobs > 0) ? (uint32_t)nJobs : std::thread::hardware_concurrency();
    if ((nJobs == 0) || (totalCores <= 1) || ((last - first) <= 1))
    {
        for (uint32_t a = first; a != last; ++a)
            func(a);
        return;
    }
    std::vector<std::thread> threads;
    if (last - first <= totalCores)
    {
        for (uint32_t index = first; index != last; ++index)
            threads.emplace_back([&func, index]() { func(index); });
        for (auto &th : threads)
            th.join();
        return;
    }
    auto jobSlice = [&func](uint32_t a, uint32_t b) {
        if (a >= b)
            return;
        while (a != b)
            func(a++);
    };
    const uint64_t tasksPerThread = (last - first + totalCores - 1) / totalCores;
    for (uint64_t index = 0; index != totalCores - 1; ++index)
    {
        uint32_t f = (uint32_t)std::min<uint64_t>(tasksPerThread * index + first, last);
        uint32_t l = (uint32_t)std::min<uint64_t>((uint64_t)f + tasksPerThread, last);
        threads.emplace_back(jobSlice, f, l);
    }
    jobSlice((uint32_t)std::min<uint64_t>(tasksPerThread * (totalCores - 1) + first, last), last); // reference: dimFirst == 0 always
    for (auto &th : threads)
        th.join();
}

struct OrcParams
{
    uint32_t trajLength, blockSize, blockOverlap, motionWindow;
    int64_t medianSize;
    uint32_t noiseMethod, maxIter;
    int64_t nJobs, randomSeed;
    int32_t optimizePGURE, expWeighting, motionEstimation;
    double lambdaEst, alphaEst, muEst, sigmaEst, tol;
};

// optional per-stage wall-clock accumulation (seconds), summed over threads: [median, noise, arps, svd, opt, recon]
static double g_stage_time[8];
static std::mutex g_stage_mu;
static inline void stage_add(int i, double dt)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    g_stage_time[i] += dt;
}
static inline double nowsec()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
extern "C" void orc_stage_times(double *out, int reset)
{
    for (int i = 0; i < 8; i++)
    {
        out[i] = g_stage_time[i];
        if (reset)
            g_stage_time[i] = 0;
    }
}

// frame_begin/frame_end: only frames in [frame_begin, frame_end) are processed (the full sequence X is
// still given, so windows/edge rules are those of the whole sequence).  The reference processes all.
template <typename T1>
static uint32_t PGURESVT_oracle(const T1 *X, uint32_t Ny, uint32_t Nx, uint32_t Nimgs, const OrcParams &p, double *Y,
                                double *estimates, uint32_t frame_begin, uint32_t frame_end)
{
    const size_t fsz = (size_t)Nx * Ny;
    std::fill(Y, Y + fsz * Nimgs, 0.0);
    std::fill(estimates, estimates + (size_t)Nimgs * 4, 0.0);
    const uint32_t bs = p.blockSize;
    const uint32_t Nt = (bs * bs < p.trajLength) ? (bs * bs) - 1 : p.trajLength; // pguresvt.hpp:57
    const uint32_t frameWindow = (uint32_t)std::floor(Nt / 2);
    const double OoNxNyNt = 1.0 / (Nx * Ny * Nt);
    const double lambda0 = (p.lambdaEst >= 0.0) ? p.lambdaEst : -1.0;
    const double alpha0 = (p.alphaEst >= 0.0) ? p.alphaEst : -1.0;
    const double mu0 = (p.muEst >= 0.0) ? p.muEst : -1.0;
    const double sigma0 = (p.sigmaEst >= 0.0) ? p.sigmaEst : -1.0;
    const uint32_t win = 2 * frameWindow + 1;
    // only the frames any requested window touches need Z
    uint32_t zlo = 0, zhi = Nimgs;
    if (frame_begin > 0 || frame_end < Nimgs)
    {
        zlo = (frame_begin > frameWindow) ? frame_begin - frameWindow : 0;
        zhi = std::min<uint32_t>(Nimgs, frame_end + frameWindow);
        if (frame_begin < frameWindow)
            zhi = std::max(zhi, std::min(Nimgs, win));
        if (frame_end > Nimgs - frameWindow)
            zlo = std::min(zlo, Nimgs - win);
    }
    std::vector<double> Z(fsz * Nimgs, 0.0);
    if (p.medianSize > 0)
    {
        auto medianFunc = [&](uint32_t i) {
            const double t0 = nowsec();
            std::vector<uint16_t> src(fsz), dst(fsz);
            for (size_t k = 0; k < fsz; k++)
                src[k] = (uint16_t)X[fsz * i + k]; // conv_to<Mat<uint16_t>>
            orc_median_u16(src.data(), dst.data(), (int)Ny, (int)Nx, (int)p.medianSize);
            for (size_t k = 0; k < fsz; k++)
                Z[fsz * i + k] = (double)dst[k];
            stage_add(0, nowsec() - t0);
        };
        parallel(medianFunc, zlo, zhi, (int)p.nJobs);
    }
    else
        for (size_t k = fsz * zlo; k < fsz * zhi; k++)
            Z[k] = (double)X[k];

    auto pgureFunc = [&](uint32_t timeIter) {
        double lambda = lambda0, alpha = alpha0, mu = mu0, sigma = sigma0;
        uint32_t a;
        if (timeIter < frameWindow)
            a = 0;
        else if (timeIter >= (Nimgs - frameWindow))
            a = Nimgs - 2 * frameWindow - 1;
        else
            a = timeIter - frameWindow;
        const size_t tot = fsz * win;
        std::vector<double> u(tot), w(tot), v(tot);
        for (size_t k = 0; k < tot; k++)
        {
            u[k] = (double)X[fsz * a + k];
            w[k] = Z[fsz * a + k];
        }
        double uMax = u[0], wMax = w[0];
        for (size_t k = 1; k < tot; k++)
        {
            uMax = std::max(uMax, u[k]);
            wMax = std::max(wMax, w[k]);
        }
        for (size_t k = 0; k < tot; k++)
        {
            u[k] /= uMax;
            w[k] /= wMax;
        }
        double t0 = nowsec();
        if (p.optimizePGURE)
        {
            Noise ne(p.noiseMethod);
            ne.Estimate(u.data(), Nx, win, alpha, mu, sigma);
        }
        double t1 = nowsec();
        stage_add(1, t1 - t0);
        Arps me(w.data(), (int)Nx, (int)win, (int)bs, (int)timeIter, (int)frameWindow, (int)p.motionWindow, (int)Nimgs);
        me.Estimate(p.motionEstimation != 0);
        t0 = nowsec();
        stage_add(2, t0 - t1);
        // NB sigma and mu swapped (pguresvt.hpp:133 vs pgure.hpp:26-30)
        Pgure opt(u.data(), me.patches.data(), (int)Nx, (int)win, alpha, sigma, mu, (int)bs, (int)p.blockOverlap,
                  p.randomSeed, p.expWeighting != 0, p.optimizePGURE != 0);
        t1 = nowsec();
        stage_add(3, t1 - t0);
        if (p.optimizePGURE)
        {
            double startPoint = (lambda >= 0.0) ? lambda : accu2(tot, [&](size_t k) { return u[k]; }) * OoNxNyNt;
            startPoint = std::max(0.0, startPoint);
            const double upperBound = std::max(100.0, startPoint);
            lambda = opt.Optimize(p.tol, startPoint, upperBound, (int)p.maxIter);
        }
        t0 = nowsec();
        stage_add(4, t0 - t1);
        opt.svt0->Reconstruct(lambda, v.data());
        for (size_t k = 0; k < tot; k++)
            v[k] *= uMax;
        stage_add(5, nowsec() - t0);
        estimates[timeIter + (size_t)Nimgs * 0] = lambda;
        estimates[timeIter + (size_t)Nimgs * 1] = alpha;
        estimates[timeIter + (size_t)Nimgs * 2] = mu;
        estimates[timeIter + (size_t)Nimgs * 3] = sigma;
        uint32_t sl;
        if (timeIter < frameWindow)
            sl = timeIter;
        else if (timeIter >= (Nimgs - frameWindow))
            sl = timeIter - (Nimgs - Nt);
        else
            sl = frameWindow;
        std::memcpy(Y + fsz * timeIter, v.data() + fsz * sl, fsz * sizeof(double));
    };
    parallel(pgureFunc, frame_begin, frame_end, (int)p.nJobs);
    return 0;
}

#define ORC_ENTRY(NAME, T)                                                                                             \
    extern "C" uint32_t NAME(const T *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, const OrcParams *p,      \
                             double *Y, double *estimates, uint32_t frame_begin, uint32_t frame_end)                   \
    {                                                                                                                  \
        return PGURESVT_oracle<T>(X, n_rows, n_cols, n_frames, *p, Y, estimates, frame_begin, frame_end);              \
    }
ORC_ENTRY(orc_pguresvt_u8, uint8_t)
ORC_ENTRY(orc_pguresvt_u16, uint16_t)
ORC_ENTRY(orc_pguresvt_f32, float)
ORC_ENTRY(orc_pguresvt_f64, double)
