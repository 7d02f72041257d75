// Noise estimation stage — NoiseEstimator::Estimate (src/noise.hpp:35-153) on the GPU.
//
//   1. k_noise_split   F-test split decision of SplitBlockQ (noise.hpp:182-221) for EVERY dyadic block of every
//                      level (s = N … 16) of every window slice in one launch per level — the decision depends only
//                      on the block, so the quirky recursion (noise.hpp:419-458) is replayed afterwards
//                      on the host against this table, which reproduces the reference's node multiset exactly
//                      (root kept, tree built twice, duplicates counted twice).
//   2. k_noise_leaf    one CTA per DISTINCT kept region: robust mean (IRLS with Huber weights and the interquartile
//                      scale, noise.hpp:238-271 — the two order statistics of A are found once by radix select, since
//                      sorting r = A - m is sorting A), Laplacian pseudo-residual (noise.hpp:395-417) and the MAD
//                      variance (noise.hpp:232-236, medians by radix select).
//   3. k_noise_wls     iteratively re-weighted least-squares line fit var = a*mean + b (noise.hpp:328-383) in one
//                      CTA; the per-iteration IQR of the residuals is again a radix select.
// The host only walks the quadtree (scalar, pointer-chasing) and applies the method-4 formulas (noise.hpp:139-146).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace pgs
{

struct NoiseRegion
{
    int i, j, s, slice; // rows i.., cols j.., side s (noise.hpp:75-77)
    long long off;      // offset of this region's Laplacian scratch
};

// ---- block-wide helpers (blockDim.x threads, deterministic order) --------------------------------------
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++)
        r += sm[w];
    return r;
}

__device__ __forceinline__ unsigned long long key_of(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double val_of(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) of n values produced by get(e), e in [0,n): MSB-first radix select, 8 bits per pass.
// All threads of the CTA must call it; hist = 256 ints + 2 words of shared memory.
template <int NT, typename Get>
__device__ double block_select(Get get, int n, int k, unsigned *hist, unsigned long long *spre)
{
    unsigned long long prefix = 0;
    int kk = k;
    for (int pass = 0; pass < 8; pass++)
    {
        const int shift = 56 - 8 * pass;
        for (int b = threadIdx.x; b < 256; b += NT)
            hist[b] = 0u;
        __syncthreads();
        // warp-aggregated histogram: values of one block share their top bytes, so un-aggregated shared-memory
        // atomics would serialise on one or two bins
        for (int e0 = 0; e0 < n; e0 += NT)
        {
            const int e = e0 + threadIdx.x;
            unsigned bin = 256u; // "does not take part"
            if (e < n)
            {
                const unsigned long long key = key_of(get(e));
                const bool match = (pass == 0) || (((key ^ prefix) >> (shift + 8)) == 0ull);
                if (match)
                    bin = (unsigned)(key >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (bin < 256u && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
                atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 32)
        { // warp 0: each lane owns 8 consecutive bins
            unsigned loc[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                loc[q] = hist[threadIdx.x * 8 + q];
                tot += loc[q];
            }
            unsigned inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o)
                    inc += t;
            }
            unsigned before = inc - tot; // elements in bins of lower lanes
            if ((unsigned)kk >= before && (unsigned)kk < inc)
            {
#pragma unroll
                for (int q = 0; q < 8; q++)
                {
                    if ((unsigned)kk >= before && (unsigned)kk < before + loc[q])
                    {
                        spre[0] = prefix | ((unsigned long long)(threadIdx.x * 8 + q) << shift);
                        spre[1] = (unsigned long long)((unsigned)kk - before);
                    }
                    before += loc[q];
                }
            }
        }
        __syncthreads();
        prefix = spre[0];
        kk = (int)spre[1];
        __syncthreads();
    }
    return val_of(prefix);
}

// window slices analysed by one batch of launches (blockIdx.y / .z selects the entry)
struct NoiseSliceList
{
    int n;
    int idx[64]; // index of the slice inside the window cube
};

// ---- 1. split decisions ------------------------------------------------------------------------------
// SplitBlockQ statistics: Sz = variance of the block, Se = variance of its 5-point residual (wrapping at the block's
// own border, so the residual of a border pixel depends on the level).
__device__ __forceinline__ double split_residual(const double *__restrict__ A, int N, int s, int y, int x)
{
    const int xp = (x + 1 == s) ? 1 : x + 1, yp = (y + 1 == s) ? 1 : y + 1;
    const int xm = (x == 0) ? s - 2 : x - 1, ym = (y == 0) ? s - 2 : y - 1;
    const double a = A[y + (size_t)N * x];
    return (5.0 * a - (((A[yp + (size_t)N * x] + A[ym + (size_t)N * x]) + A[y + (size_t)N * xm]) + A[y + (size_t)N * xp])) / sqrt(30.0);
}

// small levels (s <= 64): grid = (blocks per slice at this level, batch entries); one CTA per block, two passes.
template <int NT>
__global__ void __launch_bounds__(NT) k_noise_split(const double *__restrict__ u, int N, int s, double ftest, unsigned char *__restrict__ flags,
                                                    int flags_per_slice, int level_off, NoiseSliceList sl)
{
    __shared__ double sm[NT / 32];
    const int nb = N / s;
    const int bi = blockIdx.x % nb, bj = blockIdx.x / nb;
    const double *A = u + (size_t)N * N * sl.idx[blockIdx.y] + (size_t)(bi * s) + (size_t)N * (bj * s);
    const int R = s * s;
    double sa = 0.0, sr = 0.0;
    for (int e = threadIdx.x; e < R; e += NT)
    {
        const int y = e % s, x = e / s;
        sa += A[y + (size_t)N * x];
        sr += split_residual(A, N, s, y, x);
    }
    const double accuZ = block_sum<NT>(sa, sm) * (1.0 / R);
    const double accuR = block_sum<NT>(sr, sm) * (1.0 / R);
    double vz = 0.0, ve = 0.0;
    for (int e = threadIdx.x; e < R; e += NT)
    {
        const int y = e % s, x = e / s;
        const double dz = A[y + (size_t)N * x] - accuZ, de = split_residual(A, N, s, y, x) - accuR;
        vz = fma(dz, dz, vz);
        ve = fma(de, de, ve);
    }
    const double Sz = block_sum<NT>(vz, sm) * (1.0 / (R - 1));
    const double Se = block_sum<NT>(ve, sm) * (1.0 / (R - 1));
    if (threadIdx.x == 0)
    {
        const double stat = (Sz > Se) ? Sz / Se : Se / Sz;
        flags[(size_t)flags_per_slice * blockIdx.y + level_off + blockIdx.x] = (stat > ftest) ? 1 : 0;
    }
}

// large levels (s >= 128): a block is far too big for one CTA, so 64x64 tiles produce shifted one-pass partial sums
// (sum (a-c), sum (a-c)^2, sum res, sum res^2; c = first pixel of the block) for every large level at once —
// grid = (tiles per slice, large levels, batch entries) — and k_noise_split_fin folds the tiles of each block in a
// fixed order and takes the decision.
#define NOISE_TILE 64
__global__ void __launch_bounds__(256) k_noise_split_tiles(const double *__restrict__ u, int N, NoiseSliceList sl, double *__restrict__ part)
{
    __shared__ double sm[8];
    const int nt = N / NOISE_TILE, ti = blockIdx.x % nt, tj = blockIdx.x / nt;
    const int s = 128 << blockIdx.y;
    const int bi = (ti * NOISE_TILE) / s * s, bj = (tj * NOISE_TILE) / s * s;
    const double *A = u + (size_t)N * N * sl.idx[blockIdx.z] + (size_t)bi + (size_t)N * bj;
    const int oy = ti * NOISE_TILE - bi, ox = tj * NOISE_TILE - bj;
    const double c = A[0];
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
    for (int e = threadIdx.x; e < NOISE_TILE * NOISE_TILE; e += 256)
    {
        const int y = oy + (e % NOISE_TILE), x = ox + (e / NOISE_TILE);
        const double d = A[y + (size_t)N * x] - c, r = split_residual(A, N, s, y, x);
        q0 += d;
        q1 = fma(d, d, q1);
        q2 += r;
        q3 = fma(r, r, q3);
    }
    const double Q0 = block_sum<256>(q0, sm), Q1 = block_sum<256>(q1, sm), Q2 = block_sum<256>(q2, sm), Q3 = block_sum<256>(q3, sm);
    if (threadIdx.x == 0)
    {
        double *o = part + 4 * ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        o[0] = Q0, o[1] = Q1, o[2] = Q2, o[3] = Q3;
    }
}

// grid = (blocks per slice at level s, batch entries), one warp per block
__global__ void __launch_bounds__(32) k_noise_split_fin(const double *__restrict__ part, int N, int lvl_big, int nlev_big, double ftest,
                                                        unsigned char *__restrict__ flags, int flags_per_slice, int level_off)
{
    const int s = 128 << lvl_big, nb = N / s, nt = N / NOISE_TILE, tpb = s / NOISE_TILE;
    const int bi = blockIdx.x % nb, bj = blockIdx.x / nb;
    const double *P = part + 4 * ((size_t)(blockIdx.y * nlev_big + lvl_big) * nt * nt);
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < tpb * tpb; t += 32)
    {
        const int ti = bi * tpb + (t % tpb), tj = bj * tpb + (t / tpb);
        const double *o = P + 4 * ((size_t)ti + (size_t)nt * tj);
#pragma unroll
        for (int k = 0; k < 4; k++)
            q[k] += o[k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    if (threadIdx.x == 0)
    {
        const double R = (double)s * (double)s;
        const double Sz = (q[1] - q[0] * q[0] / R) / (R - 1.0), Se = (q[3] - q[2] * q[2] / R) / (R - 1.0);
        const double stat = (Sz > Se) ? Sz / Se : Se / Sz;
        flags[(size_t)flags_per_slice * blockIdx.y + level_off + blockIdx.x] = (stat > ftest) ? 1 : 0;
    }
}

// ---- 2. leaf statistics ------------------------------------------------------------------------------
// Laplacian pseudo-residual of ConvolveFIR (noise.hpp:395-417) at (x, y) of an s x s region: neighbours (rows filled
// first) times -laplacian, accumulated in column-major two-accumulator order
__device__ __forceinline__ double laplace_residual(const double *__restrict__ A, int ld, int s, int x, int y)
{
    const int xp = (x + 1 == s) ? 1 : x + 1, yp = (y + 1 == s) ? 1 : y + 1;
    const int xm = (x == 0) ? s - 2 : x - 1, ym = (y == 0) ? s - 2 : y - 1;
#define IN_(a, b) A[(a) + (size_t)ld * (b)]
    const double t0 = IN_(xm, ym) * -0.125, t1 = IN_(xm, y) * -0.125, t2 = IN_(xm, yp) * -0.125;
    const double t3 = IN_(x, ym) * -0.125, t4 = IN_(x, y) * 1.0, t5 = IN_(x, yp) * -0.125;
    const double t6 = IN_(xp, ym) * -0.125, t7 = IN_(xp, y) * -0.125, t8 = IN_(xp, yp) * -0.125;
#undef IN_
    const double v1 = (((t0 + t2) + t4) + t6) + t8;
    const double v2 = ((t1 + t3) + t5) + t7;
    return v1 + v2;
}

// ascending bitonic sort of 64 doubles held two per lane (element i = lane + 32 r in v_r)
__device__ __forceinline__ void warp_sort64(double &v0, double &v1, int lane)
{
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            if (j == 32)
            { // k == 64: partner is the other register of the same lane, ascending
                const double lo = fmin(v0, v1), hi = fmax(v0, v1);
                v0 = lo, v1 = hi;
            }
            else
            {
                const double p0 = __shfl_xor_sync(0xffffffffu, v0, j), p1 = __shfl_xor_sync(0xffffffffu, v1, j);
                const bool lower = (lane & j) == 0;
                const bool up0 = (lane & k) == 0, up1 = ((lane + 32) & k) == 0;
                v0 = (lower == up0) ? fmin(v0, p0) : fmax(v0, p0);
                v1 = (lower == up1) ? fmin(v1, p1) : fmax(v1, p1);
            }
        }
}
__device__ __forceinline__ double warp_pick64(double v0, double v1, int q)
{
    const double a = __shfl_sync(0xffffffffu, v0, q & 31), b = __shfl_sync(0xffffffffu, v1, q & 31);
    return (q >> 5) ? b : a;
}

// 8x8 regions (nearly all of them on noisy data): one WARP per region, two pixels per lane, everything in registers
// — order statistics by a 64-element bitonic sort, sums by shuffles — except the 3x3 stencil, which reads a
// 64-double shared tile.
__global__ void __launch_bounds__(256) k_noise_leaf8(const double *__restrict__ u, int N, const NoiseRegion *__restrict__ regions,
                                                     const int *__restrict__ list, int nlist, double *__restrict__ out)
{
    __shared__ double tile[8][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * 8 + wib;
    if (w >= nlist)
        return;
    const int ridx = list[w];
    const NoiseRegion rg = regions[ridx];
    const double *A = u + (size_t)N * N * rg.slice + (size_t)rg.i + (size_t)N * rg.j;
    // element e = y + 8 x (column-major inside the region); lane holds e = lane and e = lane + 32
    const double a0 = A[(lane & 7) + (size_t)N * (lane >> 3)], a1 = A[(lane & 7) + (size_t)N * ((lane >> 3) + 4)];
    tile[wib][lane] = a0;
    tile[wib][lane + 32] = a1;
    double s0 = a0, s1 = a1;
    warp_sort64(s0, s1, lane);
    const int n = 64, mq = 16; // floor((floor((n+1)/2)+1)/2)
    const double Alo = warp_pick64(s0, s1, mq - 1), Ahi = warp_pick64(s0, s1, n - mq - 1);
    double m = 0.0, m0 = 1E12, mprev = 0.0, inv = 0.0;
    const double tol = 1E-6, eps = 1E-12;
    bool first = true;
    for (int it = 0; it < 10000; it++)
    {
        double w0 = 1.0, w1 = 1.0;
        if (!first)
        {
            const double r0 = fabs((a0 - mprev) * inv), r1 = fabs((a1 - mprev) * inv);
            w0 = (r0 < 0.75) ? 1.0 : 0.75 / r0;
            w1 = (r1 < 0.75) ? 1.0 : 0.75 / r1;
        }
        const double S1 = warp_sum(w0 * a0 + w1 * a1), S0 = warp_sum(w0 + w1);
        m = (fabs(S0) < eps) ? m0 : S1 / S0;
        const double ee = warp_sum(fabs(a0 - m) + fabs(a1 - m)) / (double)n;
        if (fabs(m0 - m) < tol || ee < tol)
            break;
        m0 = m;
        const double d = ((Ahi - m) - (Alo - m)) + eps;
        inv = 1. / d;
        mprev = m;
        first = false;
    }
    __syncwarp();
    const double *T = tile[wib];
    const double L0 = laplace_residual(T, 8, 8, lane >> 3, lane & 7), L1 = laplace_residual(T, 8, 8, (lane >> 3) + 4, lane & 7);
    s0 = L0, s1 = L1;
    warp_sort64(s0, s1, lane);
    const double l1 = warp_pick64(s0, s1, 32), l2 = warp_pick64(s0, s1, 31);
    const double med = l1 + (l2 - l1) / 2.0;
    s0 = fabs(L0 - med), s1 = fabs(L1 - med);
    warp_sort64(s0, s1, lane);
    const double d1 = warp_pick64(s0, s1, 32), d2 = warp_pick64(s0, s1, 31);
    const double mad = d1 + (d2 - d1) / 2.0;
    if (lane == 0)
    {
        const double sig = 1.4826 * mad;
        out[2 * (size_t)ridx] = m;
        out[2 * (size_t)ridx + 1] = sig * sig;
    }
}

// regions of side 16..64: one CTA each
template <int NT>
__global__ void __launch_bounds__(NT) k_noise_leaf(const double *__restrict__ u, int N, const NoiseRegion *__restrict__ regions,
                                                   const int *__restrict__ list, double *__restrict__ scratch,
                                                   double *__restrict__ out /* 2 per region: mean, var */)
{
    __shared__ double sm[NT / 32];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long spre[2];
    const int ridx = list[blockIdx.x];
    const NoiseRegion rg = regions[ridx];
    const int s = rg.s, n = s * s;
    const double *A = u + (size_t)N * N * rg.slice + (size_t)rg.i + (size_t)N * rg.j;
    auto getA = [&](int e) { return A[(e % s) + (size_t)N * (e / s)]; };

    // interquartile order statistics of A (InterquartileDistance, noise.hpp:223-230)
    const int mq = (int)floor((floor((double)((n + 1) / 2)) + 1) / 2);
    const double Alo = block_select<NT>(getA, n, mq - 1, hist, spre);
    const double Ahi = block_select<NT>(getA, n, n - mq - 1, hist, spre);

    // robust mean (RobustMeanEstimate, noise.hpp:238-271)
    double m = 0.0, m0 = 1E12, mprev = 0.0, inv = 0.0;
    const double tol = 1E-6, eps = 1E-12;
    bool first = true;
    for (int it = 0; it < 10000; it++)
    {
        double s1 = 0.0, s0 = 0.0;
        for (int e = threadIdx.x; e < n; e += NT)
        {
            const double a = getA(e);
            double w = 1.0;
            if (!first)
            {
                const double r = fabs((a - mprev) * inv);
                w = (r < 0.75) ? 1.0 : 0.75 / r;
            }
            s1 += w * a;
            s0 += w;
        }
        const double S1 = block_sum<NT>(s1, sm), S0 = block_sum<NT>(s0, sm);
        m = (fabs(S0) < eps) ? m0 : S1 / S0;
        double se = 0.0;
        for (int e = threadIdx.x; e < n; e += NT)
            se += fabs(getA(e) - m);
        const double ee = block_sum<NT>(se, sm) / (double)n;
        if (fabs(m0 - m) < tol || ee < tol)
            break;
        m0 = m;
        const double d = ((Ahi - m) - (Alo - m)) + eps;
        inv = 1. / d;
        mprev = m;
        first = false;
    }

    double *L = scratch + rg.off;
    for (int e = threadIdx.x; e < n; e += NT)
        L[e] = laplace_residual(A, N, s, e / s, e % s); // (x, y) as in the reference's loops
    __syncthreads();
    auto getL = [&](int e) { return L[e]; };
    // arma::median of an even-length vector: nth = n/2, plus the largest of the lower half
    const double l1 = block_select<NT>(getL, n, n / 2, hist, spre);
    const double l2 = block_select<NT>(getL, n, n / 2 - 1, hist, spre);
    const double med = l1 + (l2 - l1) / 2.0;
    auto getD = [&](int e) { return fabs(L[e] - med); };
    const double d1 = block_select<NT>(getD, n, n / 2, hist, spre);
    const double d2 = block_select<NT>(getD, n, n / 2 - 1, hist, spre);
    const double mad = d1 + (d2 - d1) / 2.0;
    if (threadIdx.x == 0)
    {
        const double sig = 1.4826 * mad;
        out[2 * (size_t)ridx] = m;
        out[2 * (size_t)ridx + 1] = sig * sig;
    }
}

// ---- grid-wide (cooperative launch) versions for the large regions and the line fit -------------------
// Large regions (the whole-frame root node of every slice is always kept) and the ~5e5-sample line fit
// are latency-bound inside one CTA; here the same algorithms run on one CTA per SM with grid-wide reductions:
// per-CTA partials in a rotating slot + grid.sync, and radix selects that find TWO order statistics per sweep
// (both quartiles, or the two middle elements of a median) with histograms merged through a small ring of global
// 2x256-bin histograms (a slot is re-zeroed one pass after it was read).
namespace cg = cooperative_groups;

struct GridScratch
{
    double *partials; // [4][8][grid]
    unsigned *ghist;  // [4][512], zero at launch
};

template <int NT, int NQ>
__device__ __forceinline__ void grid_sum(const double (&v)[NQ], double (&out)[NQ], cg::grid_group &grid, const GridScratch &gs,
                                         unsigned &slot, double *sm)
{
    double *P = gs.partials + (size_t)(slot & 3u) * 8 * gridDim.x;
#pragma unroll
    for (int q = 0; q < NQ; q++)
    {
        const double b = block_sum<NT>(v[q], sm);
        if (threadIdx.x == 0)
            P[(size_t)q * gridDim.x + blockIdx.x] = b;
    }
    grid.sync();
#pragma unroll
    for (int q = 0; q < NQ; q++)
    {
        double r = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++)
            r += P[(size_t)q * gridDim.x + b];
        out[q] = r;
    }
    slot++;
}

// kA-th and kB-th smallest (0-based) of n values in one MSB-first radix sweep (8 bits per pass).  hist: 512 words,
// spre: 4 words of shared memory.
template <int NT, typename Get>
__device__ void grid_select2(Get get, int n, int kA, int kB, double &outA, double &outB, cg::grid_group &grid, const GridScratch &gs,
                             unsigned &hslot, unsigned *hist, unsigned long long *spre)
{
    unsigned long long pA = 0, pB = 0;
    int ka = kA, kb = kB;
    const int stride = gridDim.x * NT;
    for (int pass = 0; pass < 8; pass++)
    {
        const int shift = 56 - 8 * pass;
        const bool same = (pA == pB); // both targets still share their prefix: one histogram serves both
        unsigned *G = gs.ghist + (size_t)(hslot & 3u) * 512;
        for (int b = threadIdx.x; b < 512; b += NT)
            hist[b] = 0u;
        __syncthreads();
        for (int e0 = blockIdx.x * NT; e0 < n; e0 += stride)
        {
            const int e = e0 + threadIdx.x;
            unsigned bin = 512u; // "does not take part"
            if (e < n)
            {
                const unsigned long long key = key_of(get(e));
                const unsigned digit = (unsigned)(key >> shift) & 255u;
                if (pass == 0)
                    bin = digit;
                else
                {
                    const unsigned long long hi = key >> (shift + 8);
                    if (hi == (pA >> (shift + 8)))
                        bin = digit;
                    else if (!same && hi == (pB >> (shift + 8)))
                        bin = 256u + digit;
                }
            }
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (bin < 512u && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
                atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        for (int b = threadIdx.x; b < 512; b += NT)
            if (hist[b])
                atomicAdd(&G[b], hist[b]);
        grid.sync();
        // the slot read one pass ago is no longer in use by anybody: clear it for its next turn
        if (blockIdx.x == 0)
            for (int b = threadIdx.x; b < 512; b += NT)
                gs.ghist[(size_t)((hslot + 3u) & 3u) * 512 + b] = 0u;
        if (threadIdx.x < 64)
        { // warp 0 resolves target A, warp 1 target B; each lane owns 8 consecutive bins
            const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const unsigned *H = G + ((which && !same) ? 256 : 0);
            const unsigned long long prefix = which ? pB : pA;
            const unsigned kk = (unsigned)(which ? kb : ka);
            unsigned loc[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                loc[q] = H[lane * 8 + q];
                tot += loc[q];
            }
            unsigned inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += t;
            }
            unsigned before = inc - tot; // elements in bins of lower lanes
            if (kk >= before && kk < inc)
            {
#pragma unroll
                for (int q = 0; q < 8; q++)
                {
                    if (kk >= before && kk < before + loc[q])
                    {
                        spre[2 * which] = prefix | ((unsigned long long)(lane * 8 + q) << shift);
                        spre[2 * which + 1] = (unsigned long long)(kk - before);
                        spre[4 + which] = (unsigned long long)loc[q]; // population of the chosen bin
                    }
                    before += loc[q];
                }
            }
        }
        __syncthreads();
        pA = spre[0];
        ka = (int)spre[1];
        pB = spre[2];
        kb = (int)spre[3];
        const bool single = (spre[4] == 1ull && spre[5] == 1ull);
        __syncthreads();
        hslot++;
        if (single && pass < 7)
        { // both targets are alone in their bins (typical after 4-5 passes on continuous data): one fetch pass reads their
          // remaining bits instead of up to four more histogram passes.  The decision is uniform over the grid.
            unsigned long long *K = reinterpret_cast<unsigned long long *>(gs.ghist + 4 * 512) + 2 * (hslot & 3u);
            for (int e0 = blockIdx.x * NT; e0 < n; e0 += stride)
            {
                const int e = e0 + threadIdx.x;
                if (e < n)
                {
                    const unsigned long long key = key_of(get(e));
                    if ((key >> shift) == (pA >> shift))
                        K[0] = key;
                    if ((key >> shift) == (pB >> shift))
                        K[1] = key;
                }
            }
            grid.sync();
            pA = __ldcg(K);
            pB = __ldcg(K + 1);
            break;
        }
    }
    outA = val_of(pA);
    outB = val_of(pB);
}

// all large regions, one after the other, each spread over the whole grid.  Same maths as k_noise_leaf.
template <int NT>
__global__ void __launch_bounds__(NT) k_noise_big(const double *__restrict__ u, int N, const NoiseRegion *__restrict__ regions,
                                                  const int *__restrict__ big_idx, int nbig, double *__restrict__ scratch,
                                                  double *__restrict__ out, GridScratch gs)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[NT / 32];
    __shared__ unsigned hist[512];
    __shared__ unsigned long long spre[6];
    unsigned slot = 0, hslot = 0;
    const int gstride = gridDim.x * NT;
    for (int bq = 0; bq < nbig; bq++)
    {
        const int ridx = big_idx[bq];
        const NoiseRegion rg = regions[ridx];
        const int s = rg.s, n = s * s, sh = 31 - __clz(s); // s is a power of two
        const double *A = u + (size_t)N * N * rg.slice + (size_t)rg.i + (size_t)N * rg.j;
        auto getA = [&](int e) { return A[(e & (s - 1)) + (size_t)N * (e >> sh)]; };
        const int mq = (int)floor((floor((double)((n + 1) / 2)) + 1) / 2);
        double Alo, Ahi;
        grid_select2<NT>(getA, n, mq - 1, n - mq - 1, Alo, Ahi, grid, gs, hslot, hist, spre);
        double m = 0.0, m0 = 1E12, mprev = 0.0, inv = 0.0;
        const double tol = 1E-6, eps = 1E-12;
        bool first = true;
        for (int it = 0; it < 10000; it++)
        {
            double v2[2] = {0.0, 0.0}, o2[2];
            for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
            {
                const double a = getA(e);
                double w = 1.0;
                if (!first)
                {
                    const double r = fabs((a - mprev) * inv);
                    w = (r < 0.75) ? 1.0 : 0.75 / r;
                }
                v2[0] += w * a;
                v2[1] += w;
            }
            grid_sum<NT, 2>(v2, o2, grid, gs, slot, sm);
            m = (fabs(o2[1]) < eps) ? m0 : o2[0] / o2[1];
            double v1[1] = {0.0}, o1[1];
            for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
                v1[0] += fabs(getA(e) - m);
            grid_sum<NT, 1>(v1, o1, grid, gs, slot, sm);
            const double ee = o1[0] / (double)n;
            if (fabs(m0 - m) < tol || ee < tol)
                break;
            m0 = m;
            const double d = ((Ahi - m) - (Alo - m)) + eps;
            inv = 1. / d;
            mprev = m;
            first = false;
        }
        double *L = scratch + rg.off;
        for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
            L[e] = laplace_residual(A, N, s, e >> sh, e & (s - 1));
        grid.sync();
        auto getL = [&](int e) { return L[e]; };
        double l1, l2, d1, d2;
        grid_select2<NT>(getL, n, n / 2, n / 2 - 1, l1, l2, grid, gs, hslot, hist, spre);
        const double med = l1 + (l2 - l1) / 2.0;
        auto getD = [&](int e) { return fabs(L[e] - med); };
        grid_select2<NT>(getD, n, n / 2, n / 2 - 1, d1, d2, grid, gs, hslot, hist, spre);
        const double mad = d1 + (d2 - d1) / 2.0;
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            const double sig = 1.4826 * mad;
            out[2 * (size_t)ridx] = m;
            out[2 * (size_t)ridx + 1] = sig * sig;
        }
    }
}

// grid-wide WLSFit (noise.hpp:328-383) over the (mean, variance) samples of all window slices
template <int NT>
__global__ void __launch_bounds__(NT) k_noise_wls_grid(const double *__restrict__ xs, const double *__restrict__ ys, int n,
                                                       double *__restrict__ out, GridScratch gs)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[NT / 32];
    __shared__ unsigned hist[512];
    __shared__ unsigned long long spre[6];
    unsigned slot = 0, hslot = 0;
    const int gstride = gridDim.x * NT;
    const double tol = 1E-6, eps = 1E-12;
    double a0 = 1E12, b0 = 1E12, p0 = 0.0, p1 = 0.0, pa = 0.0, pb = 0.0, pd = 1.0;
    bool first = true;
    int it = 0;
    auto X = [&](int e) { return xs[e]; };
    auto Y = [&](int e) { return ys[e]; };
    double mn = INFINITY;
    for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
        mn = fmin(mn, X(e));
    for (; it < 10000; it++)
    {
        double q[5] = {0, 0, 0, 0, 0}, Q[5];
        for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
        {
            const double xe = X(e), ye = Y(e);
            double w = 1.0;
            if (!first)
            {
                const double r = fabs((ye - (xe * pa + pb)) / pd);
                w = (r < 0.75) ? 1.0 : 0.75 / r;
            }
            const double w2 = w * w;
            q[0] += w2;
            q[1] += w2 * xe;
            q[2] += w2 * ye;
            q[3] += w2 * (xe * ye);
            q[4] += w2 * (xe * xe);
        }
        grid_sum<NT, 5>(q, Q, grid, gs, slot, sm);
        p0 = Q[0] * Q[3] - Q[1] * Q[2];
        const double aux = Q[0] * Q[4] - Q[1] * Q[1];
        p0 = (fabs(aux) < eps) ? a0 : p0 / aux;
        p1 = Q[2] - p0 * Q[1];
        p1 = (fabs(aux) < eps) ? b0 : p1 / Q[0];
        double v1[1] = {0.0}, o1[1];
        for (int e = blockIdx.x * NT + threadIdx.x; e < n; e += gstride)
            v1[0] += fabs(Y(e) - (X(e) * p0 + p1));
        grid_sum<NT, 1>(v1, o1, grid, gs, slot, sm);
        const double ee = o1[0] / (double)n;
        if ((fabs(a0 - p0) < tol && fabs(b0 - p1) < tol) || ee < tol)
            break;
        a0 = p0;
        b0 = p1;
        auto getR = [&](int e) { return Y(e) - (X(e) * p0 + p1); };
        const int mq = (int)floor((floor((double)((n + 1) / 2)) + 1) / 2);
        double rhi, rlo;
        grid_select2<NT>(getR, n, n - mq - 1, mq - 1, rhi, rlo, grid, gs, hslot, hist, spre);
        pd = (rhi - rlo) + eps;
        pa = p0;
        pb = p1;
        first = false;
    }
    // smallest robust mean (robustMeans(0) after the sort, noise.hpp:141)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double r = sm[0];
        for (int w = 1; w < NT / 32; w++)
            r = fmin(r, sm[w]);
        gs.partials[(size_t)(slot & 3u) * 8 * gridDim.x + blockIdx.x] = r;
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        double r = INFINITY;
        for (unsigned b = 0; b < gridDim.x; b++)
            r = fmin(r, gs.partials[(size_t)(slot & 3u) * 8 * gridDim.x + b]);
        out[0] = p0;
        out[1] = p1;
        out[2] = (double)it;
        out[3] = r;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// Per-slice results are cached: a window slice is the frame divided by the window maximum, so its quadtree and its
// (mean, variance) samples depend only on (global frame, uMax).  Consecutive windows share 2*fw of their slices and
// very often the maximum, so in steady state one new slice is analysed per frame; the line fit always runs over
// the samples of all slices of the window.  All slices that do need analysing go through ONE batch of launches.
struct NoiseSliceSamples
{
    double umax = 0;
    std::vector<double> xs, ys; // robust means / variances in the reference's node order (duplicates included)
};

struct NoiseWorkspace
{
    double *dFit = nullptr, *dPartials = nullptr, *dX = nullptr, *dY = nullptr, *dSplitPart = nullptr, *dScratch = nullptr, *dLeaf = nullptr;
    unsigned *dHist = nullptr;
    unsigned char *dFlags = nullptr;
    NoiseRegion *dRegions = nullptr;
    int *dLists = nullptr;
    size_t capSamples = 0, capFlags = 0, capRegions = 0, capScratch = 0, capSplitPart = 0;
    int grid = 0, device = 0;
    long long slices_analysed = 0, slices_reused = 0, n_regions = 0, n_big = 0, n_mid = 0;
    double fit_iters = 0, t_fit = 0, t_slices = 0, t_split = 0, t_replay = 0, t_leaf = 0;
    std::unordered_map<long long, NoiseSliceSamples> cache;
    void release()
    {
        auto F = [](void *p) {
            if (p)
                cudaFree(p);
        };
        F(dFit), F(dPartials), F(dHist), F(dX), F(dY), F(dSplitPart), F(dScratch), F(dLeaf), F(dFlags), F(dRegions), F(dLists);
        dFit = dPartials = dX = dY = dSplitPart = dScratch = dLeaf = nullptr, dHist = nullptr, dFlags = nullptr, dRegions = nullptr,
        dLists = nullptr;
        capSamples = capFlags = capRegions = capScratch = capSplitPart = 0;
        cache.clear();
    }
};

static double noise_ftest0025(int s)
{ // fTest0025 / degOfFreePlus1, noise.hpp:163-178
    static const int dof[12] = {2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
    static const double f[12] = {15.4392, 2.86209, 1.64602, 1.27893, 1.13046, 1.06318, 1.03110, 1.01543, 1.00769, 1.00384, 1.00192, 1.00096};
    for (int i = 0; i < 12; i++)
        if (dof[i] == s)
            return f[i];
    return -1.0;
}

#define NCU(call)                                                                                         \
    do                                                                                                    \
    {                                                                                                     \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
        {                                                                                                 \
            char b_[256];                                                                                 \
            snprintf(b_, sizeof b_, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            err = b_;                                                                                     \
            return 2;                                                                                     \
        }                                                                                                 \
    } while (0)

#define NOISE_BIG_SIDE 128 /* regions with side >= this run grid-wide */

template <typename T>
static int noise_reserve(T *&p, size_t &cap, size_t need, std::string &err)
{
    if (cap >= need)
        return 0;
    if (p)
        cudaFree(p);
    p = nullptr;
    cap = 0;
    NCU(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return 0;
}

static int noise_init(NoiseWorkspace &ws, int sm_count, std::string &err)
{
    if (ws.dFit)
        return 0;
    ws.grid = sm_count > 0 ? sm_count : 148;
    NCU(cudaGetDevice(&ws.device));
    NCU(cudaMalloc(&ws.dFit, 8 * sizeof(double)));
    NCU(cudaMalloc(&ws.dPartials, (size_t)4 * 8 * ws.grid * sizeof(double)));
    NCU(cudaMalloc(&ws.dHist, 4 * 512 * sizeof(unsigned) + 8 * sizeof(unsigned long long))); // + [4][2] fetched keys
    return 0;
}

// QuadTree() (noise.hpp:419-458) of one slice replayed against its split-decision table `fl`.  Distinct kept regions
// are appended to `regions` (slice field = `slice`); sample_region lists, in the reference's node order, the region
// of every kept node (duplicates included).
struct NoiseReplay
{
    std::vector<NoiseRegion> regions;
    std::vector<int> sample_region;
};
static void noise_replay(const unsigned char *fl, int N, const std::vector<int> &offs, int per_slice, int slice, NoiseReplay &out)
{
    struct Node
    {
        int i, j, s, lvl;
    };
    struct Frame
    {
        int n, k;
    };
    std::vector<int> region_of((size_t)per_slice, -1), dele;
    std::vector<Node> tree;
    std::vector<Frame> fr;
    tree.reserve((size_t)per_slice * 2 + 8);
    dele.reserve((size_t)per_slice);
    out.regions.clear();
    out.sample_region.clear();
    out.regions.reserve((size_t)(N / 8) * (N / 8) + 8);
    out.sample_region.reserve((size_t)(N / 8) * (N / 8) * 2 + 8);
    tree.push_back({0, 0, N, 0});
    auto enter = [&](int part) {
        const Node nd = tree[part];
        if (nd.s <= 8)
            return;
        if (!fl[offs[nd.lvl] + (nd.i / nd.s) + (N / nd.s) * (nd.j / nd.s)])
            return;
        const int s = nd.s / 2;
        const int n = (int)tree.size() - 1; // index of the LAST EXISTING node
        tree.push_back({nd.i, nd.j, s, nd.lvl + 1});
        tree.push_back({nd.i + s, nd.j, s, nd.lvl + 1});
        tree.push_back({nd.i, nd.j + s, s, nd.lvl + 1});
        tree.push_back({nd.i + s, nd.j + s, s, nd.lvl + 1});
        dele.push_back(part);
        fr.push_back({n, 0});
    };
    enter(0);
    while (!fr.empty())
    {
        Frame &f = fr.back();
        if (f.k == 4)
        {
            fr.pop_back();
            continue;
        }
        const int part = f.n + f.k;
        f.k++;
        enter(part); // may push a new frame (f is not used afterwards)
    }
    std::vector<char> removed(tree.size(), 0);
    if (!dele.empty())
    { // every split node is shed except the smallest index (the root): the reference's loop stops at k > 0
        const int keep = *std::min_element(dele.begin(), dele.end());
        for (int d : dele)
            removed[d] = 1;
        removed[keep] = 0;
    }
    long long scratch = 0;
    for (size_t n = 0; n < tree.size(); n++)
    {
        if (removed[n])
            continue;
        const Node &nd = tree[n];
        int &ridx = region_of[offs[nd.lvl] + (nd.i / nd.s) + (N / nd.s) * (nd.j / nd.s)];
        if (ridx < 0)
        {
            ridx = (int)out.regions.size();
            out.regions.push_back({nd.i, nd.j, nd.s, slice, nd.s > 8 ? scratch : -1});
            if (nd.s > 8)
                scratch += (long long)nd.s * nd.s; // Laplacian scratch (8x8 regions keep theirs on chip); rebased by the caller
        }
        out.sample_region.push_back(ridx);
    }
}

// quadtree + leaf statistics of a batch of window slices (indices `todo` into dU) -> samples in the reference's order
static int noise_analyse_batch(NoiseWorkspace &ws, const double *dU, int N, const std::vector<int> &todo,
                               const std::vector<NoiseSliceSamples *> &parts, cudaStream_t st, long long *launches, std::string &err)
{
    const int S = (int)todo.size();
    NoiseSliceList sl;
    sl.n = S;
    for (int q = 0; q < S; q++)
        sl.idx[q] = todo[q];
    // level table: sides N, N/2, ..., 8 (side 8 never splits but is needed to address regions)
    std::vector<int> sides, offs;
    int per_slice = 0, per_slice_flags = 0;
    for (int s = N; s >= 8; s >>= 1)
    {
        sides.push_back(s);
        offs.push_back(per_slice);
        per_slice += (N / s) * (N / s);
        if (s >= 16)
            per_slice_flags = per_slice;
    }
    const size_t nflags = (size_t)per_slice_flags * S;
    int rc;
    if ((rc = noise_reserve(ws.dFlags, ws.capFlags, nflags, err)))
        return rc;
    const auto tp0 = std::chrono::steady_clock::now();
    int nlev_big = 0;
    for (int s = 128; s <= N; s <<= 1)
        nlev_big++;
    if (nlev_big)
    {
        const int nt = N / NOISE_TILE;
        if ((rc = noise_reserve(ws.dSplitPart, ws.capSplitPart, (size_t)4 * S * nlev_big * nt * nt, err)))
            return rc;
        k_noise_split_tiles<<<dim3(nt * nt, nlev_big, S), 256, 0, st>>>(dU, N, sl, ws.dSplitPart);
        if (launches)
            (*launches)++;
    }
    for (size_t l = 0; l < sides.size(); l++)
    {
        const int s = sides[l], nb = (N / s) * (N / s);
        if (s < 16)
            break;
        const double ft = noise_ftest0025(s);
        if (s >= 128)
        {
            int lb = 0;
            while ((128 << lb) < s)
                lb++;
            k_noise_split_fin<<<dim3(nb, S), 32, 0, st>>>(ws.dSplitPart, N, lb, nlev_big, ft, ws.dFlags, per_slice_flags, offs[l]);
        }
        else
            k_noise_split<128><<<dim3(nb, S), 128, 0, st>>>(dU, N, s, ft, ws.dFlags, per_slice_flags, offs[l], sl);
        if (launches)
            (*launches)++;
    }
    std::vector<unsigned char> fl(nflags);
    NCU(cudaMemcpyAsync(fl.data(), ws.dFlags, nflags, cudaMemcpyDeviceToHost, st));
    NCU(cudaStreamSynchronize(st));
    NCU(cudaGetLastError());
    const auto tp1 = std::chrono::steady_clock::now();
    ws.t_split += std::chrono::duration<double>(tp1 - tp0).count();

    // replay QuadTree() against the decision tables (host, one thread per slice when there are several)
    std::vector<NoiseReplay> rep(S);
    {
        auto work = [&](int w, int nw) {
            for (int q = w; q < S; q += nw)
                noise_replay(fl.data() + (size_t)per_slice_flags * q, N, offs, per_slice, todo[q], rep[q]);
        };
        const int nw = std::min(S, 16);
        if (nw <= 1)
            work(0, 1);
        else
        {
            std::vector<std::thread> th;
            for (int w = 0; w < nw; w++)
                th.emplace_back(work, w, nw);
            for (auto &t : th)
                t.join();
        }
    }
    // concatenate: region index base and scratch base per slice; lists of 8x8 / mid / big regions
    std::vector<size_t> rbase(S + 1, 0);
    for (int q = 0; q < S; q++)
        rbase[q + 1] = rbase[q] + rep[q].regions.size();
    const size_t nreg = rbase[S];
    std::vector<NoiseRegion> regions(nreg);
    std::vector<int> l8, lmid, lbig;
    l8.reserve(nreg);
    long long scratch_need = 0;
    for (int q = 0; q < S; q++)
        for (size_t r = 0; r < rep[q].regions.size(); r++)
        {
            NoiseRegion rg = rep[q].regions[r];
            const int gi = (int)(rbase[q] + r);
            if (rg.s == 8)
                l8.push_back(gi);
            else
            {
                rg.off = scratch_need;
                scratch_need += (long long)rg.s * rg.s;
                (rg.s >= NOISE_BIG_SIDE ? lbig : lmid).push_back(gi);
            }
            regions[gi] = rg;
        }
    const auto tp2 = std::chrono::steady_clock::now();
    ws.t_replay += std::chrono::duration<double>(tp2 - tp1).count();
    ws.n_regions += (long long)nreg;
    ws.n_big += (long long)lbig.size();
    ws.n_mid += (long long)lmid.size();
    {
        const size_t old = ws.capRegions;
        if ((rc = noise_reserve(ws.dRegions, ws.capRegions, nreg, err)))
            return rc;
        if (ws.capRegions != old)
        { // out (2 doubles) and list entries (1 int) per region travel with the region table
            if (ws.dLeaf)
                cudaFree(ws.dLeaf);
            if (ws.dLists)
                cudaFree(ws.dLists);
            ws.dLeaf = nullptr, ws.dLists = nullptr;
            NCU(cudaMalloc(&ws.dLeaf, ws.capRegions * 2 * sizeof(double)));
            NCU(cudaMalloc(&ws.dLists, ws.capRegions * sizeof(int)));
        }
    }
    if ((rc = noise_reserve(ws.dScratch, ws.capScratch, (size_t)std::max<long long>(scratch_need, 1), err)))
        return rc;
    std::vector<int> lists;
    lists.reserve(nreg);
    lists.insert(lists.end(), l8.begin(), l8.end());
    lists.insert(lists.end(), lmid.begin(), lmid.end());
    lists.insert(lists.end(), lbig.begin(), lbig.end());
    NCU(cudaMemcpyAsync(ws.dRegions, regions.data(), nreg * sizeof(NoiseRegion), cudaMemcpyHostToDevice, st));
    NCU(cudaMemcpyAsync(ws.dLists, lists.data(), nreg * sizeof(int), cudaMemcpyHostToDevice, st));
    const int *d8 = ws.dLists, *dMid = ws.dLists + l8.size(), *dBig = dMid + lmid.size();
    if (!l8.empty())
    {
        k_noise_leaf8<<<(unsigned)((l8.size() + 7) / 8), 256, 0, st>>>(dU, N, ws.dRegions, d8, (int)l8.size(), ws.dLeaf);
        if (launches)
            (*launches)++;
    }
    if (!lmid.empty())
    {
        k_noise_leaf<128><<<(unsigned)lmid.size(), 128, 0, st>>>(dU, N, ws.dRegions, dMid, ws.dScratch, ws.dLeaf);
        if (launches)
            (*launches)++;
    }
    if (!lbig.empty())
    {
        GridScratch gs;
        gs.partials = ws.dPartials;
        gs.ghist = ws.dHist;
        NCU(cudaMemsetAsync(ws.dHist, 0, 4 * 512 * sizeof(unsigned), st));
        const double *a0 = dU;
        int a1 = N, a4 = (int)lbig.size();
        const NoiseRegion *a2 = ws.dRegions;
        const int *a3 = dBig;
        double *a5 = ws.dScratch, *a6 = ws.dLeaf;
        void *args[] = {(void *)&a0, (void *)&a1, (void *)&a2, (void *)&a3, (void *)&a4, (void *)&a5, (void *)&a6, (void *)&gs};
        // the whole grid (round 1 gave the steady-state case a quarter of the SMs so that the long Jacobi launches kept their
        // CTAs; beside the 2.3-ms dominant-triplet launches of round 2 the estimate itself became the exposed part: measured
        // 32.2 -> 33.9 frames/s, profiles/r02/noise_grid_sweep.txt)
        static const int big_div_env = getenv("PGURESVT_BIG_GRID_DIV") ? atoi(getenv("PGURESVT_BIG_GRID_DIV")) : 0;
        const int big_div = big_div_env > 0 ? big_div_env : 1;
        NCU(cudaLaunchCooperativeKernel((void *)k_noise_big<512>, dim3(std::max(1, ws.grid / big_div)), dim3(512), args, 0, st));
        if (launches)
            (*launches)++;
    }
    std::vector<double> leaf(nreg * 2);
    NCU(cudaMemcpyAsync(leaf.data(), ws.dLeaf, nreg * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NCU(cudaStreamSynchronize(st));
    NCU(cudaGetLastError());
    ws.t_leaf += std::chrono::duration<double>(std::chrono::steady_clock::now() - tp2).count();
    for (int q = 0; q < S; q++)
    {
        NoiseSliceSamples &o = *parts[todo[q]];
        o.xs.clear();
        o.ys.clear();
        o.xs.reserve(rep[q].sample_region.size());
        o.ys.reserve(rep[q].sample_region.size());
        for (int r : rep[q].sample_region)
        { // means and variances are filtered independently with >= 0 (noise.hpp:103-104)
            const double mm = leaf[2 * (rbase[q] + (size_t)r)], vv = leaf[2 * (rbase[q] + (size_t)r) + 1];
            if (mm >= 0.)
                o.xs.push_back(mm);
            if (vv >= 0.)
                o.ys.push_back(vv);
        }
    }
    return 0;
}

// Estimate (alpha, mu, sigma) of one window dU (N,N,T) whose slice k is global frame frame0+k divided by umax.
// In/out values < 0 are estimated (noise.hpp:113-147).  Everything runs on stream `st`.
static int noise_estimate_window(NoiseWorkspace &ws, const double *dU, int N, int T, int method, int sm_count, cudaStream_t st,
                                 double &alpha, double &mu, double &sigma, long long *launches, std::string &err,
                                 long long frame0 = -1, double umax = 0.0)
{
    if (N < 16 || (N & (N - 1)) != 0 || N > 4096)
    {
        err = "quadtree noise estimation requires square frames with a power-of-two side between 16 and 4096";
        return 1;
    }
    int rc = noise_init(ws, sm_count, err);
    if (rc)
        return rc;
    std::vector<NoiseSliceSamples *> parts(T);
    std::vector<NoiseSliceSamples> local(frame0 >= 0 ? 0 : T);
    if (frame0 >= 0)
    { // drop cache entries that can no longer be part of a window
        for (auto it = ws.cache.begin(); it != ws.cache.end();)
            it = (it->first < frame0 - T || it->first > frame0 + 2 * T) ? ws.cache.erase(it) : std::next(it);
    }
    std::vector<int> todo;
    for (int k = 0; k < T; k++)
    {
        if (frame0 >= 0)
        {
            NoiseSliceSamples &e = ws.cache[frame0 + k]; // references into unordered_map stay valid across inserts
            parts[k] = &e;
            if (e.umax != umax || e.xs.empty())
            {
                e.umax = umax;
                todo.push_back(k);
            }
            else
                ws.slices_reused++;
        }
        else
        {
            parts[k] = &local[k];
            todo.push_back(k);
        }
    }
    if (!todo.empty())
    {
        const auto ts0 = std::chrono::steady_clock::now();
        for (size_t b = 0; b < todo.size(); b += 64)
        {
            const std::vector<int> chunk(todo.begin() + b, todo.begin() + std::min(todo.size(), b + 64));
            if ((rc = noise_analyse_batch(ws, dU, N, chunk, parts, st, launches, err)))
                return rc;
        }
        ws.slices_analysed += (long long)todo.size();
        ws.t_slices += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count();
    }
    size_t n = 0, ny = 0;
    for (int k = 0; k < T; k++)
    {
        n += parts[k]->xs.size();
        ny += parts[k]->ys.size();
    }
    if (n != ny || n == 0)
    {
        err = "noise estimation produced a negative robust mean or variance (the reference's sample filtering would misalign)";
        return 1;
    }
    std::vector<double> xs, ys;
    xs.reserve(n);
    ys.reserve(n);
    for (int k = 0; k < T; k++)
    {
        xs.insert(xs.end(), parts[k]->xs.begin(), parts[k]->xs.end());
        ys.insert(ys.end(), parts[k]->ys.begin(), parts[k]->ys.end());
    }
    if (ws.capSamples < n)
    {
        if (ws.dX)
            cudaFree(ws.dX);
        if (ws.dY)
            cudaFree(ws.dY);
        ws.dX = ws.dY = nullptr;
        ws.capSamples = 0;
        NCU(cudaMalloc(&ws.dX, n * sizeof(double)));
        NCU(cudaMalloc(&ws.dY, n * sizeof(double)));
        ws.capSamples = n;
    }
    const auto tf0 = std::chrono::steady_clock::now();
    NCU(cudaMemcpyAsync(ws.dX, xs.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
    NCU(cudaMemcpyAsync(ws.dY, ys.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        GridScratch gs;
        gs.partials = ws.dPartials;
        gs.ghist = ws.dHist;
        NCU(cudaMemsetAsync(ws.dHist, 0, 4 * 512 * sizeof(unsigned), st));
        const double *a0 = ws.dX, *a1 = ws.dY;
        int a2 = (int)n;
        double *a3 = ws.dFit;
        void *args[] = {(void *)&a0, (void *)&a1, (void *)&a2, (void *)&a3, (void *)&gs};
        // a third of the SMs: the fit is bound by its ~290 grid-wide barriers, which get cheaper with fewer CTAs, and the
        // other three quarters keep both of their SVD CTAs while it runs (measured +2.3 % frames/s against one CTA per SM)
        static const int wls_div = getenv("PGURESVT_WLS_GRID_DIV") ? atoi(getenv("PGURESVT_WLS_GRID_DIV")) : 3;
        NCU(cudaLaunchCooperativeKernel((void *)k_noise_wls_grid<512>, dim3(std::max(1, ws.grid / wls_div)), dim3(512), args, 0, st));
        if (launches)
            (*launches)++;
    }
    double fit[4];
    NCU(cudaMemcpyAsync(fit, ws.dFit, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NCU(cudaStreamSynchronize(st));
    NCU(cudaGetLastError());
    ws.t_fit += std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count();
    ws.fit_iters = fit[2];
    if (getenv("PGURESVT_NOISE_TIMING"))
        fprintf(stderr,
                "[noise] slices %.2f ms (analysed %lld, reused %lld; split %.2f, replay %.2f, leaves %.2f ms; regions %lld, mid %lld, big %lld), "
                "fit %.2f ms (%g iters), %zu samples\n",
                ws.t_slices * 1e3, ws.slices_analysed, ws.slices_reused, ws.t_split * 1e3, ws.t_replay * 1e3, ws.t_leaf * 1e3, ws.n_regions,
                ws.n_mid, ws.n_big, ws.t_fit * 1e3, ws.fit_iters, n);
    alpha = (alpha >= 0.) ? alpha : fit[0]; // noise.hpp:113
    if (method >= 1 && method <= 3)
    {
        // noise.hpp:115-137: mode-based variants.  They work on the samples sorted by mean (stable) — scalar
        // post-processing of the GPU-computed samples, done on the host.
        std::vector<size_t> idx(n);
        for (size_t k = 0; k < n; k++)
            idx[k] = k;
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return xs[a] < xs[b]; });
        std::vector<double> rm(n), rv(n);
        for (size_t k = 0; k < n; k++)
        {
            rm[k] = xs[idx[k]];
            rv[k] = ys[idx[k]];
        }
        auto compute_mode = [](const std::vector<double> &A) { // ComputeMode, noise.hpp:273-301 (first maximal count wins)
            const size_t nn = A.size();
            const double M = *std::max_element(A.begin(), A.end()), dyn = 1. * nn;
            std::vector<double> a(nn);
            for (size_t i = 0; i < nn; i++)
                a[i] = std::round(A[i] * dyn / M);
            std::unordered_map<double, std::pair<unsigned, size_t>> cnt; // value -> (count, first index)
            for (size_t i = 0; i < nn; i++)
            {
                auto it = cnt.find(a[i]);
                if (it == cnt.end())
                    cnt.emplace(a[i], std::make_pair(1u, i));
                else
                    it->second.first++;
            }
            unsigned best = 0;
            size_t first = 0;
            double val = 0.;
            for (auto &kv : cnt)
                if (kv.second.first > best || (kv.second.first == best && kv.second.second < first))
                {
                    best = kv.second.first;
                    first = kv.second.second;
                    val = kv.first;
                }
            return val * (M / dyn);
        };
        auto head = [](const std::vector<double> &a, size_t last) { return std::vector<double>(a.begin(), a.begin() + last + 1); };
        if (method == 1)
        {
            const int L = (int)std::floor(1. * ((unsigned)(N * N) / (unsigned)n));
            const size_t last = (size_t)std::round(0.05 * L);
            mu = (mu >= 0.) ? mu : compute_mode(head(rm, last));
            const double dSi = compute_mode(head(rv, last));
            sigma = (sigma >= 0.) ? sigma : std::sqrt(dSi);
        }
        else if (method == 2)
        {
            mu = (mu >= 0.) ? mu : compute_mode(rm);
            const double dSi = compute_mode(rv);
            sigma = (sigma >= 0.) ? sigma : std::sqrt(std::max(dSi, std::max(fit[1] + fit[0] * dSi, 0.)));
        }
        else
        {
            mu = (mu >= 0.) ? mu : compute_mode(rm);
            sigma = (sigma >= 0.) ? sigma : std::sqrt(std::fabs(fit[1] + fit[0] * mu));
        }
        return 0;
    }
    // noise.hpp:139-146 (method 4; the default branch is the same)
    mu = (mu >= 0.) ? mu : fit[3];
    sigma = (sigma >= 0.) ? sigma : std::sqrt(std::fabs(fit[1] + fit[0] * mu));
    return 0;
}
#undef NCU
} // namespace pgs
