// compact.cuh — truncated factor cache for Casorati shapes beyond 16x15 (BASELINE config 5: 64x31, 4096^2 frames).
//
// The full factor cache of SVT::Decompose (svt.hpp:111-116: U, S, V of every patch, 23.8 KB per 64x31 patch and SVT
// object) is 398 GB per object at config 5 — it cannot stay resident.  What the lambda search of
// PGURE::Optimize (pgure.hpp:189-237) actually consumes per patch is much smaller:
//   * the singular values S of the three SVT objects U, U + eps2*delta2, U - eps2*delta2 (thresholds for any lambda),
//   * the bilinear forms q_k = u_k^T (delta2/weights) v_k of every singular triplet of the three objects: the
//     second-difference term of the risk (pgure.hpp:136) is linear in the reconstructed blocks, so it collapses to
//     sum_patches sum_k f_k(lambda) q_k (same identity as k_qform3 / k_eval3 of the 16x15 path),
//   * the LEADING R singular triplets of object U: Uhat enters the risk non-linearly, so its block is rebuilt — the
//     singular values are sorted and the soft threshold is monotone, so the survivors are a prefix of the spectrum.
// 2 x 32 doubles per patch and object + R x (m + 32) doubles per patch: 1.5 KB + R x 768 B at 64x31.
// A probe at which MORE than R triplets of some patch survive is still answered exactly: those patches are collected
// (k_eval_c's overflow list), decomposed again in chunks into a scratch buffer of full records and reconstructed with
// the generic k_recon.
//
// k_svd_warp — one-sided Jacobi SVD of a (32*RPL) x n matrix (n <= 32) by ONE WARP with the matrix in registers:
// lane l keeps rows l, l+32, ... of all 32 column slots.  Ordering: odd-even transposition with the swap folded into the
// rotation (round E pairs slots (2p, 2p+1), round O pairs (2p+1, 2p+2); the rotated columns are written back exchanged),
// so every pair of columns meets exactly once in 32 rounds and the schedule needs NO register moves and only two round
// bodies — the loop stays inside the instruction cache.  Per round: 16 partial inner products per lane, one transposing
// butterfly (16 double shuffles) that leaves pair p's sum on lanes 2p, 2p+1, one fast (scaled) rotation per lane from TRACKED
// squared norms (jacobi_fast_givens, norms refreshed once per sweep), its two coefficients broadcast through shared memory.
// V is not accumulated: V = A0^T U / sigma is rebuilt from the re-gathered matrix at the end (as in k_svd16_l4), fused
// with the q-forms.  Replaces arma::svd_econ (svt.hpp:111) for bs = 8 (64 x 15 ... 64 x 31).
#pragma once

namespace pgs
{

__device__ __forceinline__ double trw_16(const double (&x)[16], int lane)
{ // transposing reduction of 16 values over 32 lanes: lanes 2p and 2p+1 return the warp-wide sum of x[p]
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0;
    double y8[8], y4[4], y2[2];
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const double send = b4 ? x[j] : x[j + 8], keep = b4 ? x[j + 8] : x[j];
        y8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const double send = b3 ? y8[j] : y8[j + 4], keep = b3 ? y8[j + 4] : y8[j];
        y4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
        const double send = b2 ? y4[j] : y4[j + 2], keep = b2 ? y4[j + 2] : y4[j];
        y2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const double send = b1 ? y2[0] : y2[1], keep = b1 ? y2[1] : y2[0];
    const double y1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    return y1 + __shfl_xor_sync(0xffffffffu, y1, 1);
}

__device__ __forceinline__ double trw_32(const double (&x)[32], int lane)
{ // transposing reduction of 32 values over 32 lanes: lane l returns the warp-wide sum of x[l]
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
    double y16[16], y8[8], y4[4], y2[2];
#pragma unroll
    for (int j = 0; j < 16; j++)
    {
        const double send = b4 ? x[j] : x[j + 16], keep = b4 ? x[j + 16] : x[j];
        y16[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const double send = b3 ? y16[j] : y16[j + 8], keep = b3 ? y16[j + 8] : y16[j];
        y8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const double send = b2 ? y8[j] : y8[j + 4], keep = b2 ? y8[j + 4] : y8[j];
        y4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
        const double send = b1 ? y4[j] : y4[j + 2], keep = b1 ? y4[j + 2] : y4[j];
        y2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const double send = b0 ? y2[0] : y2[1], keep = b0 ? y2[1] : y2[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// Outputs of the SVD kernels of this file.  MODE 0 writes the generic full record (U m x n | V ldv x n | S ldv) that
// k_recon reads; MODE 1 writes the compact cache described above (lead == nullptr or R == 0: S and q only).
struct SvdOut
{
    double *fac;
    size_t rec;
    int ldv;
    double *S[3];     // per object: 32 doubles per patch, descending (svt.hpp:111 order), zero padded
    double *Q[3];     // per object: 32 doubles per patch: q_k in the same order
    double *lead;     // R x (m + 32) doubles per patch: u_k (m) | v_k (32, zero padded)
    int R;
    const double *c4; // delta2 / weights per voxel (k_c4)
};

// one round of the odd-even ordering; ODD = 0: slot pairs (2p, 2p+1), p = 0..15; ODD = 1: (2p+1, 2p+2), p = 0..14.
// Lane l holds the state of the column in slot l: its tracked TRUE squared norm and — fast (scaled) rotations as in
// k_svd16_l4 — the scale of the stored column and its reciprocal (true column = sc * stored column), so that a rotation
// costs two FMAs per element pair:  x' = x - alpha y,  y' = y + beta x  (jacobi_fast_givens).
template <int RPL, int ODD>
__device__ __forceinline__ void svdw_round(double (&a)[RPL][32], double &nrm, double &sc, double &isc, double2 *csb, int lane,
                                           double tol2, double big2, bool &big)
{
    double pg[16];
#pragma unroll
    for (int p = 0; p < 16; p++)
    {
        double sg = 0.0;
        if (2 * p + ODD + 1 < 32)
        {
#pragma unroll
            for (int r = 0; r < RPL; r++)
                sg = fma(a[r][2 * p + ODD], a[r][2 * p + ODD + 1], sg);
        }
        pg[p] = sg;
    }
    double G = trw_16(pg, lane); // pair p on lanes 2p, 2p+1
    if (ODD)
        G = __shfl_up_sync(0xffffffffu, G, 1); // pair p on lanes 2p+1 (its low slot) and 2p+2 (its high slot)
    const bool is_lo = ((lane & 1) == ODD);
    const bool active = ODD ? (lane >= 1 && lane <= 30) : true;
    const int partner = (is_lo ? lane + 1 : lane - 1) & 31;
    const double o_nrm = __shfl_sync(0xffffffffu, nrm, partner);
    const double o_sc = __shfl_sync(0xffffffffu, sc, partner);
    const double o_isc = __shfl_sync(0xffffffffu, isc, partner);
    double A = is_lo ? nrm : o_nrm, B = is_lo ? o_nrm : nrm;
    double sx = is_lo ? sc : o_sc, sy = is_lo ? o_sc : sc;
    double isx = is_lo ? isc : o_isc, isy = is_lo ? o_isc : isc;
    if (!active)
    {
        G = 0.0;
        A = B = sx = sy = isx = isy = 1.0;
    }
    double alpha, beta;
    bool bg = false;
    jacobi_fast_givens(A, B, sx, sy, isx, isy, G, tol2, big2, alpha, beta, bg); // both lanes of a pair derive the identical rotation
    big = big || bg;
    // the rotated columns are written back exchanged: the low slot receives y' (norm B, scale sy), the high slot x'
    if (active)
    {
        nrm = is_lo ? B : A;
        sc = is_lo ? sy : sx;
        isc = is_lo ? isy : isx;
    }
    if (is_lo && active)
        csb[(lane - ODD) >> 1] = make_double2(alpha, beta);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 16 - ODD; p++)
    {
        const double2 ab = csb[p];
#pragma unroll
        for (int r = 0; r < RPL; r++)
        {
            const double x = a[r][2 * p + ODD], y = a[r][2 * p + ODD + 1];
            a[r][2 * p + ODD] = fma(ab.y, x, y);      // y' = y + beta x
            a[r][2 * p + ODD + 1] = fma(-ab.x, y, x); // x' = x - alpha y
        }
    }
}

#define SVDW_WARP_DOUBLES(RPL) (32 * (RPL)*32 + 32 * 32 + 64) /* W staging (m x 32) | V of object U (32 x 32) | two coefficient buffers of 16 double2 */

// MODE 0: one SVT object (pt.mode), full records.  MODE 1: one object, compact cache (o.S[0], o.Q[0], o.lead).
// MODE 2: the three PGURE objects U, U + eps2*delta2, U - eps2*delta2 of a patch back to back by the same warp
// (o.S[i], o.Q[i]; pt.eps = eps2): the perturbed matrices are WARM-STARTED — multiplied by the V of object U, which
// never leaves shared memory — so their Jacobi iteration starts from nearly orthogonal columns (the same idea as
// k_svd16_l4<WARM>, without the round trip of V through HBM and without the perturbed-window copies).  A rank-deficient
// object U (a zero patch, say) has no orthogonal V to offer: its perturbed objects start cold.
template <int RPL, int MODE>
__global__ void __launch_bounds__(128, 2)
    k_svd_warp(const double *__restrict__ u, Perturb pt, const short2 *__restrict__ pos, const int *__restrict__ ids, int P,
               int vecSize, int N, int bs, int n, SvdOut o, int max_sweeps, double tol2, double big2, int *__restrict__ sweeps_out)
{
    constexpr int M = 32 * RPL;
    extern __shared__ __align__(16) double smw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pidx = blockIdx.x * 4 + wib;
    if (pidx >= P) // warp-uniform; the kernel has no CTA-wide barrier
        return;
    double *Ws = smw + (size_t)wib * SVDW_WARP_DOUBLES(RPL);
    double *V0s = Ws + M * 32; // V of object U: V0s[s * 32 + k] = V(k, slot s)
    double2 *csb = reinterpret_cast<double2 *>(V0s + 32 * 32);
    const int id = ids[pidx];
    const size_t fsz = (size_t)N * N;
    int myoff = 0; // offset of slice `lane`'s block origin within its slice (svt.hpp:99-109)
    if (lane < n)
    {
        const short2 p = pos[(size_t)lane * vecSize + id];
        myoff = p.x + N * p.y;
    }
    int eoff[RPL];
#pragma unroll
    for (int r = 0; r < RPL; r++)
    {
        const int e = lane + 32 * r;
        eoff[r] = (e % bs) + N * (e / bs);
    }
    bool warm_ok = false;
#pragma unroll 1
    for (int oi = 0; oi < (MODE == 2 ? 3 : 1); oi++)
    {
        if (MODE == 2)
            pt.mode = (oi == 0) ? 0 : oi + 1;
        // Casorati gather (svt.hpp:99-109) through shared memory in a ROLLED loop: the perturbation code exists once instead
        // of 64 times (the straight-line version was 110 KB of instructions, fetched once per matrix and never reused —
        // ncu: 29 % of the stalls were "no instruction")
        double a[RPL][32];
        __syncwarp();
#pragma unroll 1 // (unrolling by 4 to keep more loads in flight made no difference: 405 vs 407 ms per 3.1 M SVDs)
        for (int k = 0; k < n; k++)
        {
            const int offk = __shfl_sync(0xffffffffu, myoff, k);
#pragma unroll
            for (int r = 0; r < RPL; r++)
                Ws[k * M + lane + 32 * r] = load_perturbed(u, (size_t)offk + eoff[r] + fsz * k, pt);
        }
#pragma unroll
        for (int k = 0; k < 32; k++)
#pragma unroll
            for (int r = 0; r < RPL; r++)
                a[r][k] = (k < n) ? Ws[k * M + lane + 32 * r] : 0.0; // own writes only: no barrier needed
        if (MODE == 2 && oi > 0 && warm_ok)
        { // rows of A times V0, one slot per iteration of a rolled loop, results parked in the (free) W buffer
            __syncwarp();
#pragma unroll 1
            for (int s = 0; s < 32; s++)
            {
                const double2 *vs = reinterpret_cast<const double2 *>(V0s + s * 32);
                double acc[RPL];
#pragma unroll
                for (int r = 0; r < RPL; r++)
                    acc[r] = 0.0;
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++)
                {
                    const double2 v = vs[k2];
#pragma unroll
                    for (int r = 0; r < RPL; r++)
                    {
                        acc[r] = fma(a[r][2 * k2], v.x, acc[r]);
                        acc[r] = fma(a[r][2 * k2 + 1], v.y, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < RPL; r++)
                    Ws[s * M + lane + 32 * r] = acc[r];
            }
#pragma unroll
            for (int s = 0; s < 32; s++)
#pragma unroll
                for (int r = 0; r < RPL; r++)
                    a[r][s] = Ws[s * M + lane + 32 * r];
        }

        int sweep = 0, quiet = 0;
        double nrm = 0.0, sc = 1.0, isc = 1.0;
#pragma unroll 1
        for (; sweep < max_sweeps;)
        {
            { // fresh squared norms at the start of every sweep (the tracked updates drift by rounding only)
                double n2[32];
#pragma unroll
                for (int s = 0; s < 32; s++)
                {
                    double sacc = 0.0;
#pragma unroll
                    for (int r = 0; r < RPL; r++)
                        sacc = fma(a[r][s], a[r][s], sacc);
                    n2[s] = sacc;
                }
                nrm = trw_32(n2, lane) * (sc * sc);
            }
#pragma unroll 1
            for (int rp = 0; rp < 16 && quiet < 32; rp++)
            {
                bool big = false;
                svdw_round<RPL, 0>(a, nrm, sc, isc, csb, lane, tol2, big2, big);
                quiet = __any_sync(0xffffffffu, big) ? 0 : quiet + 1;
                big = false;
                svdw_round<RPL, 1>(a, nrm, sc, isc, csb + 16, lane, tol2, big2, big);
                quiet = __any_sync(0xffffffffu, big) ? 0 : quiet + 1;
            }
            sweep++;
            if (quiet >= 32) // every pair met once in the last 32 rounds and none needed a rotation above `big`
                break;
        }

        // true columns = scale * stored columns; the scale of slot s sits in lane s
#pragma unroll
        for (int s = 0; s < 32; s++)
        {
            const double ss = __shfl_sync(0xffffffffu, sc, s);
#pragma unroll
            for (int r = 0; r < RPL; r++)
                a[r][s] *= ss;
        }
        // singular values: lane s owns slot s
        double my2;
        {
            double n2[32];
#pragma unroll
            for (int s = 0; s < 32; s++)
            {
                double sacc = 0.0;
#pragma unroll
                for (int r = 0; r < RPL; r++)
                    sacc = fma(a[r][s], a[r][s], sacc);
                n2[s] = sacc;
            }
            my2 = trw_32(n2, lane);
        }
        const double sig = sqrt(my2);
        int rk = 0; // descending order like LAPACK (svt.hpp:111); ties by slot
#pragma unroll
        for (int t = 0; t < 32; t++)
        {
            const double st = __shfl_sync(0xffffffffu, sig, t);
            rk += (st > sig || (st == sig && t < lane)) ? 1 : 0;
        }
        const double inv = (sig > 0.0) ? 1.0 / sig : 0.0;
        // W = U diag(sigma) staged in shared memory (column s at Ws + s*M); the registers of `a` are dead from here on
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 32; s++)
#pragma unroll
            for (int r = 0; r < RPL; r++)
                Ws[s * M + lane + 32 * r] = a[r][s];
        __syncwarp();
        // lane k re-gathers column k of the decomposed matrix A0 (and of C4) and contracts it with every W column:
        //   vw[s] = A0(:,k) . w_s = sigma_s^2 v_s[k],   cw[s] = C4(:,k) . w_s   (q_s = sum_k vw[s] cw[s] / sigma_s^3)
        double vw[32], cw[32];
#pragma unroll
        for (int s = 0; s < 32; s++)
            vw[s] = cw[s] = 0.0;
        const bool realcol = lane < n;
        const size_t colbase = (size_t)myoff + fsz * lane;
#pragma unroll 1
        for (int e0 = 0; e0 < M; e0 += 2)
        { // two rows per iteration (one 16-byte broadcast read per W column); kept rolled for the instruction cache
            double a0[2], c0[2];
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                const int e = e0 + i;
                const size_t vox = colbase + (e % bs) + (size_t)N * (e / bs);
                a0[i] = realcol ? load_perturbed(u, vox, pt) : 0.0;
                c0[i] = (MODE != 0 && realcol) ? o.c4[vox] : 0.0;
            }
#pragma unroll
            for (int s = 0; s < 32; s++)
            {
                const double2 w = *reinterpret_cast<const double2 *>(Ws + s * M + e0);
                vw[s] = fma(a0[0], w.x, vw[s]);
                vw[s] = fma(a0[1], w.y, vw[s]);
                if (MODE != 0)
                {
                    cw[s] = fma(c0[0], w.x, cw[s]);
                    cw[s] = fma(c0[1], w.y, cw[s]);
                }
            }
        }
        if (MODE != 0)
        {
            double pr[32];
#pragma unroll
            for (int s = 0; s < 32; s++)
                pr[s] = vw[s] * cw[s];
            const double qs = trw_32(pr, lane) * (inv * inv * inv);
            double *So = (oi == 0) ? o.S[0] : (oi == 1) ? o.S[1] : o.S[2];
            double *Qo = (oi == 0) ? o.Q[0] : (oi == 1) ? o.Q[1] : o.Q[2];
            So[(size_t)pidx * 32 + rk] = sig;
            Qo[(size_t)pidx * 32 + rk] = qs;
            const bool want_lead = oi == 0 && o.lead && o.R > 0;
            if (want_lead || (MODE == 2 && oi == 0))
            {
                double *L = want_lead ? o.lead + (size_t)pidx * o.R * (M + 32) : nullptr;
#pragma unroll
                for (int s = 0; s < 32; s++)
                {
                    const int rks = __shfl_sync(0xffffffffu, rk, s);
                    const double invs = __shfl_sync(0xffffffffu, inv, s);
                    const double vks = vw[s] * (invs * invs); // V(lane, slot s)
                    if (want_lead && rks < o.R)
                    {
                        double *Lk = L + (size_t)rks * (M + 32);
#pragma unroll
                        for (int r = 0; r < RPL; r++)
                            Lk[lane + 32 * r] = Ws[s * M + lane + 32 * r] * invs;
                        Lk[M + lane] = vks;
                    }
                    if (MODE == 2)
                        V0s[s * 32 + lane] = vks;
                }
                if (MODE == 2)
                { // V0 is orthogonal on the n real columns only if object U has full column rank
                    const double smax = warp_max(sig);
                    warm_ok = __popc(__ballot_sync(0xffffffffu, sig > smax * 1e-8)) >= n;
                    __syncwarp();
                }
            }
        }
        else
        {
            double *R = o.fac + o.rec * (size_t)pidx;
#pragma unroll
            for (int s = 0; s < 32; s++)
            {
                const int rks = __shfl_sync(0xffffffffu, rk, s);
                const double invs = __shfl_sync(0xffffffffu, inv, s);
                if (rks < n)
                {
#pragma unroll
                    for (int r = 0; r < RPL; r++)
                        R[(size_t)M * rks + lane + 32 * r] = Ws[s * M + lane + 32 * r] * invs;
                    if (lane < o.ldv)
                        R[(size_t)M * n + (size_t)o.ldv * rks + lane] = vw[s] * (invs * invs);
                }
            }
            if (rk < n)
                R[(size_t)M * n + (size_t)o.ldv * n + rk] = sig;
        }
        if (sweeps_out && lane == 0)
        {
            atomicMax(sweeps_out, sweep);
            atomicAdd(sweeps_out + 1 + (oi > 0 ? 1 : 0), sweep);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Generic shapes (any m, n <= 32): the shared-memory Jacobi of k_svd_smem with the compact epilogue.
// dynamic smem per warp: (m*n + n*n + n) doubles, as k_svd_smem.
// ------------------------------------------------------------------------------------------------------
__global__ void k_svd_smem_c(const double *__restrict__ u, Perturb pt, const short2 *__restrict__ pos, const int *__restrict__ ids,
                             int P, int vecSize, int N, int bs, int n, SvdOut o, int max_sweeps, double tol2,
                             int *__restrict__ sweeps_out)
{
    extern __shared__ double smd[];
    const int m = bs * bs;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pidx = blockIdx.x * (blockDim.x >> 5) + wib;
    if (pidx >= P)
        return;
    double *A = smd + (size_t)wib * (m * n + n * n + n);
    double *V = A + m * n;
    double *sig = V + n * n;
    const int id = ids[pidx];
    const size_t fsz = (size_t)N * N;
    int myoff = 0;
    if (lane < n)
    {
        const short2 p = pos[(size_t)lane * vecSize + id];
        myoff = p.x + N * p.y;
    }
    for (int k = 0; k < n; k++)
    {
        const int offk = __shfl_sync(0xffffffffu, myoff, k);
        for (int e = lane; e < m; e += 32)
            A[e + m * k] = load_perturbed(u, (size_t)offk + (e % bs) + (size_t)N * (e / bs) + fsz * k, pt);
    }
    for (int e = lane; e < n * n; e += 32)
        V[e] = ((e % n) == (e / n)) ? 1.0 : 0.0;
    __syncwarp();
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++)
    {
        bool rotated = false;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++)
            {
                double al = 0, be = 0, g = 0;
                for (int e = lane; e < m; e += 32)
                {
                    const double x = A[e + m * p], y = A[e + m * q];
                    al = fma(x, x, al);
                    be = fma(y, y, be);
                    g = fma(x, y, g);
                }
                al = warp_sum(al);
                be = warp_sum(be);
                g = warp_sum(g);
                double c, s;
                bool rot = false;
                jacobi_cs(al, be, g, tol2, c, s, rot);
                if (rot)
                {
                    rotated = true;
                    for (int e = lane; e < m; e += 32)
                    {
                        const double x = A[e + m * p], y = A[e + m * q];
                        A[e + m * p] = c * x - s * y;
                        A[e + m * q] = s * x + c * y;
                    }
                    for (int e = lane; e < n; e += 32)
                    {
                        const double x = V[e + n * p], y = V[e + n * q];
                        V[e + n * p] = c * x - s * y;
                        V[e + n * q] = s * x + c * y;
                    }
                }
                __syncwarp();
            }
        if (!rotated)
            break;
    }
    for (int j = 0; j < n; j++)
    {
        double al = 0;
        for (int e = lane; e < m; e += 32)
            al = fma(A[e + m * j], A[e + m * j], al);
        al = warp_sum(al);
        if (lane == 0)
            sig[j] = sqrt(al);
    }
    __syncwarp();
    // lane j: sigma, rank and q of column j
    const double sj = (lane < n) ? sig[lane] : 0.0;
    int rk = 0;
    for (int t = 0; t < n; t++)
        rk += (sig[t] > sj || (sig[t] == sj && t < lane)) ? 1 : 0;
    const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
    double qmine = 0.0; // sigma_j * q_j = w_j^T C4 v_j
    for (int e0 = 0; e0 < m; e0 += 32)
    {
        const int e = e0 + lane;
        const bool re = e < m;
        const int eo = re ? (e % bs) + N * (e / bs) : 0;
        double cwv[32];
#pragma unroll
        for (int k = 0; k < 32; k++)
        {
            const int offk = __shfl_sync(0xffffffffu, myoff, k);
            cwv[k] = (re && k < n) ? o.c4[(size_t)offk + eo + fsz * k] : 0.0;
        }
        for (int j = 0; j < n; j++)
        {
            const double *vj = V + n * j;
            double z = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k++)
                if (k < n)
                    z = fma(cwv[k], vj[k], z);
            double val = re ? A[e + m * j] * z : 0.0;
            val = warp_sum(val);
            if (lane == j)
                qmine += val;
        }
    }
    if (lane < n)
    {
        o.S[0][(size_t)pidx * 32 + rk] = sj;
        o.Q[0][(size_t)pidx * 32 + rk] = qmine * inv;
    }
    if (o.lead && o.R > 0)
    {
        double *L = o.lead + (size_t)pidx * o.R * (m + 32);
        for (int j = 0; j < n; j++)
        {
            const int rkj = __shfl_sync(0xffffffffu, rk, j);
            const double invj = __shfl_sync(0xffffffffu, inv, j);
            if (rkj < o.R)
            {
                double *Lk = L + (size_t)rkj * (m + 32);
                for (int e = lane; e < m; e += 32)
                    Lk[e] = A[e + m * j] * invj;
                Lk[m + lane] = (lane < n) ? V[lane + n * j] : 0.0;
            }
        }
    }
    if (lane == 0 && sweeps_out)
        atomicMax(sweeps_out, sweep + 1);
}

// ------------------------------------------------------------------------------------------------------
// K_eval_c — one PGURE evaluation (pgure.hpp:120-137 = SVT::Reconstruct svt.hpp:121-167 + risk sums) from the compact
// cache.  One warp per patch (grid-stride): thresholds of the three objects from S (lane k = slot k), the
// second-difference partial sum from the q-forms, and — if at most R triplets of object U survive — the block
// sum_k f_k u_k v_k^T rebuilt from the leading triplets and overlap-added along the trajectory with FP64 REDs
// (svt.hpp:146-155).  Patches with more than R survivors are appended to the overflow list (macroblock ids) for the
// exact chunked fallback on the host side.  only_k >= 0 restricts the overlap-add to one slice (final reconstruction).
// partial / kpart: one entry per warp of the grid (s4 part, triplets used).
// ------------------------------------------------------------------------------------------------------
#define EVC_RMAX 32
__global__ void __launch_bounds__(128)
    k_eval_c(const double *__restrict__ S0, const double *__restrict__ S2, const double *__restrict__ S3, const double *__restrict__ Q0,
             const double *__restrict__ Q2, const double *__restrict__ Q3, const double *__restrict__ lead, int R, int m, int n, int bs,
             const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, double lambda, int expw, int only_k,
             int want_s4, double *__restrict__ acc, const double *__restrict__ accs, double *__restrict__ partial, int *__restrict__ kpart,
             int *__restrict__ ovf)
{
    const double ascale = __ldg(accs);
    __shared__ __align__(16) double sv[4][EVC_RMAX * 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nw = gridDim.x * 4, w = blockIdx.x * 4 + wib;
    const size_t fsz = (size_t)N * N;
    double s4tot = 0.0;
    int ktot = 0;
    for (int pidx = w; pidx < P; pidx += nw)
    {
        const size_t b = (size_t)pidx * 32 + lane;
        const double s0 = S0[b];
        const double f0 = soft_f(s0, __shfl_sync(0xffffffffu, s0, 0), lambda, expw);
        if (want_s4)
        {
            const double s2 = S2[b], s3 = S3[b];
            const double f2 = soft_f(s2, __shfl_sync(0xffffffffu, s2, 0), lambda, expw);
            const double f3 = soft_f(s3, __shfl_sync(0xffffffffu, s3, 0), lambda, expw);
            s4tot += fma(f2, Q2[b], fma(f3, Q3[b], -2.0 * f0 * Q0[b]));
        }
        const int r0 = __popc(__ballot_sync(0xffffffffu, f0 != 0.0)); // survivors are a prefix of the sorted spectrum
        if (r0 == 0)
            continue;
        const int id = ids[pidx];
        if (r0 > R)
        {
            if (lane == 0)
                ovf[1 + atomicAdd(ovf, 1)] = id;
            continue;
        }
        ktot += r0;
        int myoff = 0;
        if (lane < n)
        {
            const short2 p = pos[(size_t)lane * vecSize + id];
            myoff = p.x + N * p.y;
        }
        const double *L = lead + (size_t)pidx * R * (m + 32);
        __syncwarp();
        for (int j = 0; j < r0; j++)
            sv[wib][j * 32 + lane] = L[(size_t)j * (m + 32) + m + lane] * __shfl_sync(0xffffffffu, f0, j); // f_j v_j[k]
        __syncwarp();
        for (int e0 = 0; e0 < m; e0 += 32)
        {
            const int e = e0 + lane;
            const bool re = e < m;
            double a[32];
#pragma unroll
            for (int k = 0; k < 32; k++)
                a[k] = 0.0;
            for (int j = 0; j < r0; j++)
            {
                const double uj = re ? L[(size_t)j * (m + 32) + e] : 0.0;
                const double2 *vj = reinterpret_cast<const double2 *>(&sv[wib][j * 32]);
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++)
                {
                    const double2 v = vj[k2];
                    a[2 * k2] = fma(uj, v.x, a[2 * k2]);
                    a[2 * k2 + 1] = fma(uj, v.y, a[2 * k2 + 1]);
                }
            }
            const int eo = re ? (e % bs) + N * (e / bs) : 0;
#pragma unroll
            for (int k = 0; k < 32; k++)
            {
                const int offk = __shfl_sync(0xffffffffu, myoff, k);
                if (re && k < n && (only_k < 0 || k == only_k))
                    acc_add(acc, (size_t)offk + eo + fsz * k, a[k], ascale);
            }
        }
    }
    s4tot = warp_sum(s4tot);
    if (lane == 0)
    {
        partial[w] = s4tot;
        kpart[w] = ktot;
    }
}

} // namespace pgs
