// Hand-written sm_100a kernels of the PGURE-SVT hot path.  Each kernel cites the reference code it replaces
// (file:line relative to tjof2/pgure-svt v0.6.4).  All arithmetic the reference does in double stays FP64;
// tensor cores are not used (the per-patch matrices are 16x15 … 256x15, not a dense contraction).
//
// Device data layout (column-major like Armadillo):
//   frames      X[r + N*(c + N*f)]                      native input type (u8/u16/f32/f64)
//   window      u, w  double (N, N, win)                normalised by the window maximum (pguresvt.hpp:116-120)
//   pos         short2[k*vecSize + id]  (.x=row,.y=col) trajectories = arma::icube patches(2, vecSize, win)
//   mot         short2[s*vecSize + id]                  motions(2, vecSize, win-1)
//   factors     per patch one record of REC doubles: U (m x n) | V (ldv x n, rows >= n zero) | S (ldv, descending)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgs
{

// ------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// v / weights with non-finite -> 0 (svt.hpp:163-164)
__device__ __forceinline__ double norm_or_zero(double a, unsigned c)
{
    const double v = a / (double)c;
    return isfinite(v) ? v : 0.0;
}

template <typename T>
__device__ __forceinline__ double to_double(T v)
{
    return (double)v;
}

// Overlap-add accumulators (svt.hpp:148-155) are 64-bit FIXED POINT: every contribution is rounded once to a multiple of
// 1/scale and added as an integer, so the sum does not depend on the order the REDs land in — the reference's sequential
// overlap-add is deterministic, and so is this one (bit-identical objective, lambda and pixels from run to run).
// scale = 2^e is chosen per frame on the device (k_acc_scale) from the largest weight and the bound
// |block entry| <= ||A||_F <= sqrt(m n) max|a|, so that the sum cannot overflow; for 16 x 15 patches without motion pile-up
// e = 53, i.e. a resolution of 1.1e-16 on window-normalised values.  accs[0] = scale, accs[1] = 1/scale (exact).
__device__ __forceinline__ void acc_add(double *acc, size_t i, double v, double scale)
{
    atomicAdd(reinterpret_cast<unsigned long long *>(acc) + i, (unsigned long long)__double2ll_rn(v * scale));
}
__device__ __forceinline__ double acc_val(const double *acc, size_t i, double inv_scale)
{
    return (double)reinterpret_cast<const long long *>(acc)[i] * inv_scale;
}

// largest weight of cnt[off .. off+n) -> *out (atomicMax; *out cleared by the caller)
__global__ void k_cnt_max(const unsigned *__restrict__ cnt, size_t off, size_t n, unsigned *__restrict__ out)
{
    unsigned m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, cnt[off + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m)
        atomicMax(out, m);
}
// accs = {2^e, 2^-e} with maxcnt * entry_bound * 2^e < 2^62
__global__ void k_acc_scale(const unsigned *__restrict__ maxcnt, double entry_bound, double *__restrict__ accs)
{
    const double tot = fmax((double)*maxcnt, 1.0) * entry_bound;
    int e = 61 - ilogb(tot);
    e = min(max(e, -900), 60);
    accs[0] = ldexp(1.0, e);
    accs[1] = ldexp(1.0, -e);
}

// ------------------------------------------------------------------------------------------------------
// frame maxima  (u.max(), w.max(): pguresvt.hpp:116-117) — per-frame partial maxima, combined per window
// on the host.  grid = (bpf, nframes)
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_frame_max(const T *__restrict__ X, size_t fsz, double *__restrict__ partial, double *__restrict__ partial_min = nullptr)
{
    const T *f = X + fsz * blockIdx.y;
    double m = -INFINITY, lo = INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < fsz; i += (size_t)gridDim.x * blockDim.x)
    {
        const double v = to_double(f[i]);
        m = fmax(m, v);
        lo = fmin(lo, v);
    }
    m = warp_max(m);
    lo = -warp_max(-lo);
    __shared__ double sm[2][32];
    if ((threadIdx.x & 31) == 0)
    {
        sm[0][threadIdx.x >> 5] = m;
        sm[1][threadIdx.x >> 5] = lo;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sm[0][threadIdx.x] : -INFINITY;
        lo = (threadIdx.x < (blockDim.x >> 5)) ? sm[1][threadIdx.x] : INFINITY;
        m = warp_max(m);
        lo = -warp_max(-lo);
        if (threadIdx.x == 0)
        {
            partial[blockIdx.y * gridDim.x + blockIdx.x] = m;
            if (partial_min) // (only the overflow bound of the fixed-point accumulators needs it: signed floating-point input)
                partial_min[blockIdx.y * gridDim.x + blockIdx.x] = lo;
        }
    }
}

// conv_to<Mat<uint16_t>>::from(X.slice(i))  (pguresvt.hpp:75)
template <typename T>
__global__ void k_to_u16(const T *__restrict__ X, uint16_t *__restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (uint16_t)X[i];
}

// ------------------------------------------------------------------------------------------------------
// K_median — (2r+1)^2 clamp-to-edge median of uint16, the result of ConstantTimeMedianFilter
// (medfilter.hpp:478-539 as called at pguresvt.hpp:75-76).  One CTA = 32x32 output tile staged
// (with halo) in shared memory; each pixel selects its median by a 16-step bitwise search over the value
// (count of window entries below the candidate), so no per-thread sort or histogram is needed.
// grid = (ceil(N/32), ceil(N/32), nframes), block = (32, 8)
// ------------------------------------------------------------------------------------------------------
__global__ void k_median_u16(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int nr, int nc, int r)
{
    extern __shared__ uint16_t tile[];
    const int tw = 32 + 2 * r;
    const size_t fsz = (size_t)nr * nc;
    const uint16_t *s = src + fsz * blockIdx.z;
    uint16_t *d = dst + fsz * blockIdx.z;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int idx = tid; idx < tw * tw; idx += 256)
    {
        const int tr = idx % tw, tc = idx / tw;
        const int gr = min(max(r0 + tr - r, 0), nr - 1);
        const int gc = min(max(c0 + tc - r, 0), nc - 1);
        tile[idx] = s[gr + (size_t)nr * gc];
    }
    __syncthreads();
    const int K = (2 * r + 1) * (2 * r + 1), kth = K / 2, wdt = 2 * r + 1;
    const int lr = threadIdx.x;
#pragma unroll 1
    for (int cc = 0; cc < 4; cc++)
    {
        const int lc = threadIdx.y + 8 * cc;
        if (r0 + lr >= nr || c0 + lc >= nc)
            continue;
        unsigned prefix = 0;
#pragma unroll 1
        for (int bit = 15; bit >= 0; bit--)
        {
            const unsigned cand = prefix | (1u << bit);
            int cnt = 0;
            for (int dc = 0; dc < wdt; dc++)
            {
                const uint16_t *col = tile + lr + tw * (lc + dc);
                for (int dr = 0; dr < wdt; dr++)
                    cnt += (col[dr] < cand) ? 1 : 0;
            }
            if (cnt <= kth)
                prefix = cand;
        }
        d[(r0 + lr) + (size_t)nr * (c0 + lc)] = (uint16_t)prefix;
    }
}

// ------------------------------------------------------------------------------------------------------
// K_window — u = conv_to<cube>(X.slices(a,b)) / uMax  (pguresvt.hpp:100-120).  True division, like arma's
// `cube /= scalar`.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_window(const T *__restrict__ X, double *__restrict__ u, size_t n, double vmax)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        u[i] = __ddiv_rn(to_double(X[i]), vmax);
}

// ------------------------------------------------------------------------------------------------------
// K_perturb — Bernoulli perturbations of PGURE::GenerateRandomPerturbations (pgure.hpp:167-186) drawn from
// pcg64 = setseq_xsl_rr_128_64 (pcg_random.hpp:166-169,427-451,1085-1113) through libstdc++'s
// bernoulli_distribution, bit-identically: draw g (all of delta1 first, then delta2, column-major) uses the
// generator state after g+1 steps, reached by the O(log g) LCG jump-ahead (pcg_random.hpp:471-474), so the
// stream is generated in parallel.
// ------------------------------------------------------------------------------------------------------
struct U128
{
    unsigned long long lo, hi;
};
__device__ __forceinline__ U128 mul128(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo * b.lo;
    r.hi = __umul64hi(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
    return r;
}
__device__ __forceinline__ U128 add128(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
#define PCG_MULT_HI 2549297995355413924ull
#define PCG_MULT_LO 4865540595714422341ull
#define PCG_INC_HI 6364136223846793005ull
#define PCG_INC_LO 1442695040888963407ull

__global__ void k_perturb(int8_t *__restrict__ d1, int8_t *__restrict__ d2neg, long long n, unsigned long long seed,
                          double vP, int chunk)
{
    const long long start = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (long long)chunk;
    if (start >= 2 * n)
        return;
    const U128 MULT = {PCG_MULT_LO, PCG_MULT_HI}, INC = {PCG_INC_LO, PCG_INC_HI};
    U128 st = {seed, 0ull};
    st = add128(mul128(add128(st, INC), MULT), INC); // engine(seed): state = (seed + inc)*mult + inc
    { // advance by `start` steps
        U128 acc_mult = {1ull, 0ull}, acc_plus = {0ull, 0ull}, cur_mult = MULT, cur_plus = INC;
        unsigned long long delta = (unsigned long long)start;
        while (delta > 0)
        {
            if (delta & 1)
            {
                acc_mult = mul128(acc_mult, cur_mult);
                acc_plus = add128(mul128(acc_plus, cur_mult), cur_plus);
            }
            U128 one = {1ull, 0ull};
            cur_plus = mul128(add128(cur_mult, one), cur_plus);
            cur_mult = mul128(cur_mult, cur_mult);
            delta >>= 1;
        }
        st = add128(mul128(acc_mult, st), acc_plus);
    }
    for (int j = 0; j < chunk; j++)
    {
        const long long g = start + j;
        if (g >= 2 * n)
            break;
        st = add128(mul128(st, MULT), INC);
        const unsigned rot = (unsigned)(st.hi >> 58);
        const unsigned long long x = st.hi ^ st.lo;
        const unsigned long long out = (x >> rot) | (x << ((64 - rot) & 63));
        double uu = __ull2double_rn(out) * 5.42101086242752217e-20; // * 2^-64
        if (uu >= 1.0)
            uu = 0.99999999999999989;
        if (g < n)
            d1[g] = (uu < 0.5) ? (int8_t)-1 : (int8_t)1;
        else
            d2neg[g - n] = (uu < vP) ? (int8_t)1 : (int8_t)0;
    }
}

// ------------------------------------------------------------------------------------------------------
// K_arps — adaptive rood pattern search, one thread per macroblock, one launch per frame pair in the
// reference's order (arps.hpp:52-134 schedules, :153-375 search).  The block cost is evaluated exactly like
// arma::accu(square(A-B)) * (1/bs^2) (arps.hpp:148-151): two running sums over even/odd column-major linear
// indices, added at the end, no FMA contraction — motion vectors must be bit-exact.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double arps_cost(const double *__restrict__ A1, const double *__restrict__ A2, int N, int bs,
                                            int ry, int rx, int py, int px, double oobs2)
{
    double v1 = 0.0, v2 = 0.0;
    int e = 0;
    for (int c = 0; c < bs; c++)
    {
        const double *a = A1 + ry + (size_t)N * (rx + c);
        const double *b = A2 + py + (size_t)N * (px + c);
        for (int r = 0; r < bs; r++, e++)
        {
            const double d = __dsub_rn(a[r], b[r]);
            const double sq = __dmul_rn(d, d);
            if (e & 1)
                v2 = __dadd_rn(v2, sq);
            else
                v1 = __dadd_rn(v1, sq);
        }
    }
    return __dmul_rn(__dadd_rn(v1, v2), oobs2);
}

#define ARPS_MAX_MW 15
// `pred` (optional) is the trajectory slice whose displacement from the grid position is this pair's predictor,
// i.e. motions(:, it, f3) of the reference when an earlier pair of the same frame wrote it (only the first backward
// pair, which reads what the first forward pair wrote: arps.hpp:68-77); NULL = zero predictor.
// `out` is the trajectory slice of the target frame f2.
__global__ void k_arps_pair(const double *__restrict__ w, int N, int bs, int mw, int f1, int f2,
                            const short2 *__restrict__ pred, short2 *__restrict__ out, int vecSize, double oobs2,
                            unsigned long long *__restrict__ ncost)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= vecSize)
        return;
    const int M1 = N - bs + 1;
    const int i = it % M1, j = it / M1;
    const size_t fsz = (size_t)N * N;
    const double *A1 = w + fsz * f1, *A2 = w + fsz * f2;
    const int W = 2 * mw + 1;
    unsigned chk[(2 * ARPS_MAX_MW + 1) * (2 * ARPS_MAX_MW + 1) / 32 + 1];
#pragma unroll
    for (int q = 0; q < (2 * ARPS_MAX_MW + 1) * (2 * ARPS_MAX_MW + 1) / 32 + 1; q++)
        chk[q] = 0u;
#define CHK_SET(cy, cx)                                \
    {                                                  \
        const int b_ = (cy) + W * (cx);                \
        chk[b_ >> 5] |= 1u << (b_ & 31);               \
    }
#define CHK_GET(cy, cx) ((chk[((cy) + W * (cx)) >> 5] >> (((cy) + W * (cx)) & 31)) & 1u)
    const int SDx[5] = {0, -1, 0, 1, 0}, SDy[5] = {-1, 0, 0, 0, 1}; // (hor, ver) offsets, arps.hpp:174-184
    double costs[6];
    int LDx[6], LDy[6];
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        costs[k] = 1E9;
        LDx[k] = 0;
        LDy[k] = 0;
    }
    int x = j, y = i;
    unsigned nc = 1;
    costs[2] = arps_cost(A1, A2, N, bs, i, j, i, j, oobs2);
    CHK_SET(mw, mw);
    int maxIdx, stepSize;
    if (j == 0)
    {
        stepSize = 2;
        maxIdx = 5;
    }
    else
    {
        short2 pm = make_short2(0, 0);
        if (pred)
        {
            const short2 pp = pred[it];
            pm = make_short2((short)(pp.x - i), (short)(pp.y - j));
        }
        const int yTmp = abs((int)pm.x), xTmp = abs((int)pm.y);
        stepSize = (xTmp <= yTmp) ? yTmp : xTmp;
        if (((yTmp == 0) && (xTmp == stepSize)) || ((xTmp == 0) && (yTmp == stepSize)))
            maxIdx = 5;
        else
        {
            maxIdx = 6;
            LDx[5] = pm.y;
            LDy[5] = pm.x;
        }
    }
    LDx[0] = 0, LDy[0] = -stepSize;
    LDx[1] = -stepSize, LDy[1] = 0;
    LDx[2] = 0, LDy[2] = 0;
    LDx[3] = stepSize, LDy[3] = 0;
    LDx[4] = 0, LDy[4] = stepSize;
    for (int k = 0; k < maxIdx; k++) // LDSP, arps.hpp:247-287
    {
        const int ver = y + LDy[k], hor = x + LDx[k];
        const bool skip = (k == 2) || (stepSize == 0) || (hor < 0) || (ver < 0) || (hor + bs - 1) >= N || (ver + bs - 1) >= N;
        if (!skip)
        {
            costs[k] = arps_cost(A1, A2, N, bs, i, j, ver, hor, oobs2);
            nc++;
            const int cy = LDy[k] + mw, cx = LDx[k] + mw;
            if (cy >= 0 && cy < W && cx >= 0 && cx < W)
                CHK_SET(cy, cx);
        }
    }
    int point = 0;
#pragma unroll
    for (int k = 1; k < 6; k++)
        if (costs[k] < costs[point])
            point = k;
    x += LDx[point];
    y += LDy[point];
    double cost = costs[point];
#pragma unroll
    for (int k = 0; k < 6; k++)
        costs[k] = 1E9;
    costs[2] = cost;
    bool done = false;
    unsigned nSDSP = 0;
    do // SDSP, arps.hpp:299-368
    {
        for (int k = 0; k < 5; k++)
        {
            const int ver = y + SDy[k], hor = x + SDx[k];
            bool skip = (k == 2) || (hor < 0) || (ver < 0) || (hor + bs - 1) >= N || (ver + bs - 1) >= N || (hor < j - mw) ||
                        (hor > j + mw) || (ver < i - mw) || (ver > i + mw);
            if (!skip)
                skip = CHK_GET(y - i + SDy[k] + mw, x - j + SDx[k] + mw) != 0u;
            if (!skip)
            {
                costs[k] = arps_cost(A1, A2, N, bs, i, j, ver, hor, oobs2);
                nc++;
                CHK_SET(y - i + SDy[k] + mw, x - j + SDx[k] + mw);
            }
        }
        point = 0;
#pragma unroll
        for (int k = 1; k < 6; k++)
            if (costs[k] < costs[point])
                point = k;
        cost = costs[point];
        if (point == 2 || nSDSP >= 1000000u)
            done = true;
        else
        {
            x += SDx[point];
            y += SDy[point];
#pragma unroll
            for (int k = 0; k < 6; k++)
                costs[k] = 1E9;
            costs[2] = cost;
        }
        nSDSP++;
    } while (!done);
    out[it] = make_short2((short)y, (short)x);
    if (ncost)
        atomicAdd(ncost, (unsigned long long)nc);
#undef CHK_SET
#undef CHK_GET
}

// K_arps, patch size 4 — same search, specialised: the 4x4 reference block lives in registers (it is compared against
// ~16 candidate blocks), the visited-positions matrix is a 4 x 64-bit register bitmap, and the cost loop is fully
// unrolled in the reference's accumulation order (even/odd column-major linear index -> two accumulators).
__device__ __forceinline__ double arps_cost4(const double (&ref)[16], const double *__restrict__ A2, int N, int py, int px,
                                             double oobs2)
{
    double v1 = 0.0, v2 = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        const double *b = A2 + py + (size_t)N * (px + c);
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const double d = __dsub_rn(ref[4 * c + r], b[r]);
            const double sq = __dmul_rn(d, d);
            if ((4 * c + r) & 1)
                v2 = __dadd_rn(v2, sq);
            else
                v1 = __dadd_rn(v1, sq);
        }
    }
    return __dmul_rn(__dadd_rn(v1, v2), oobs2);
}

__device__ __forceinline__ void bm_set(unsigned long long (&bm)[4], int idx)
{
    const unsigned long long bit = 1ull << (idx & 63);
    const int w = idx >> 6;
    bm[0] |= (w == 0) ? bit : 0ull;
    bm[1] |= (w == 1) ? bit : 0ull;
    bm[2] |= (w == 2) ? bit : 0ull;
    bm[3] |= (w == 3) ? bit : 0ull;
}
__device__ __forceinline__ bool bm_get(const unsigned long long (&bm)[4], int idx)
{
    const int w = idx >> 6;
    const unsigned long long word = (w == 0) ? bm[0] : (w == 1) ? bm[1] : (w == 2) ? bm[2] : bm[3];
    return (word >> (idx & 63)) & 1ull;
}

// requires mw <= 7 (W*W <= 225 bits)
__global__ void __launch_bounds__(128)
    k_arps_pair4(const double *__restrict__ w, int N, int mw, int f1, int f2, const short2 *__restrict__ pred,
                 short2 *__restrict__ out, int vecSize, unsigned long long *__restrict__ ncost)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= vecSize)
        return;
    const int M1 = N - 3;
    const int i = it % M1, j = it / M1;
    const size_t fsz = (size_t)N * N;
    const double *A1 = w + fsz * f1, *A2 = w + fsz * f2;
    const int W = 2 * mw + 1;
    const double oobs2 = 1.0 / 16.0;
    double ref[16];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++)
            ref[4 * c + r] = A1[(i + r) + (size_t)N * (j + c)];
    unsigned long long bm[4] = {0ull, 0ull, 0ull, 0ull};
    double costs[6];
    int LDx[6], LDy[6];
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        costs[k] = 1E9;
        LDx[k] = 0;
        LDy[k] = 0;
    }
    int x = j, y = i;
    unsigned nc = 1;
    costs[2] = arps_cost4(ref, A2, N, i, j, oobs2);
    bm_set(bm, mw + W * mw);
    int maxIdx, stepSize;
    if (j == 0)
    {
        stepSize = 2;
        maxIdx = 5;
    }
    else
    {
        short2 pm = make_short2(0, 0);
        if (pred)
        {
            const short2 pp = pred[it];
            pm = make_short2((short)(pp.x - i), (short)(pp.y - j));
        }
        const int yTmp = abs((int)pm.x), xTmp = abs((int)pm.y);
        stepSize = (xTmp <= yTmp) ? yTmp : xTmp;
        if (((yTmp == 0) && (xTmp == stepSize)) || ((xTmp == 0) && (yTmp == stepSize)))
            maxIdx = 5;
        else
        {
            maxIdx = 6;
            LDx[5] = pm.y;
            LDy[5] = pm.x;
        }
    }
    LDx[0] = 0, LDy[0] = -stepSize;
    LDx[1] = -stepSize, LDy[1] = 0;
    LDx[3] = stepSize, LDy[3] = 0;
    LDx[4] = 0, LDy[4] = stepSize;
    if (stepSize != 0)
    {
#pragma unroll
        for (int k = 0; k < 6; k++) // LDSP, arps.hpp:247-287
        {
            if (k == 2 || k >= maxIdx)
                continue;
            const int ver = y + LDy[k], hor = x + LDx[k];
            const bool skip = (hor < 0) || (ver < 0) || (hor + 3) >= N || (ver + 3) >= N;
            if (!skip)
            {
                costs[k] = arps_cost4(ref, A2, N, ver, hor, oobs2);
                nc++;
                const int cy = LDy[k] + mw, cx = LDx[k] + mw;
                if (cy >= 0 && cy < W && cx >= 0 && cx < W)
                    bm_set(bm, cy + W * cx);
            }
        }
    }
    int point = 0;
#pragma unroll
    for (int k = 1; k < 6; k++)
        if (costs[k] < costs[point])
            point = k;
    {
        int dx = 0, dy = 0;
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (k == point)
            {
                dx = LDx[k];
                dy = LDy[k];
            }
        x += dx;
        y += dy;
    }
    double cost = costs[0];
#pragma unroll
    for (int k = 1; k < 6; k++)
        cost = (k == point) ? costs[k] : cost;
    // SDSP, arps.hpp:299-368: offsets (hor, ver) = (0,-1), (-1,0), centre, (1,0), (0,1); ties -> lowest index
    bool done = false;
    unsigned nSDSP = 0;
    do
    {
        double cs[5] = {1E9, 1E9, cost, 1E9, 1E9};
        const int sdx[5] = {0, -1, 0, 1, 0}, sdy[5] = {-1, 0, 0, 0, 1};
#pragma unroll
        for (int k = 0; k < 5; k++)
        {
            if (k == 2)
                continue;
            const int ver = y + sdy[k], hor = x + sdx[k];
            bool skip = (hor < 0) || (ver < 0) || (hor + 3) >= N || (ver + 3) >= N || (hor < j - mw) || (hor > j + mw) ||
                        (ver < i - mw) || (ver > i + mw);
            const int bidx = (ver - i + mw) + W * (hor - j + mw);
            if (!skip)
                skip = bm_get(bm, bidx);
            if (!skip)
            {
                cs[k] = arps_cost4(ref, A2, N, ver, hor, oobs2);
                nc++;
                bm_set(bm, bidx);
            }
        }
        int pt = 0;
#pragma unroll
        for (int k = 1; k < 5; k++)
            if (cs[k] < cs[pt])
                pt = k;
        // the sixth cost slot of the reference is 1e9 here and can never win against the carried centre cost
        if (pt == 2 || nSDSP >= 1000000u)
            done = true;
        else
        {
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (k == pt)
                {
                    x += sdx[k];
                    y += sdy[k];
                    cost = cs[k];
                }
        }
        nSDSP++;
    } while (!done);
    out[it] = make_short2((short)y, (short)x);
    if (ncost)
        atomicAdd(ncost, (unsigned long long)nc);
}

// reference coordinates of the window's reference slice (arps.hpp:60-64)
// up to 32 trajectory slices copied in ONE launch (cached ARPS pairs into the window's trajectory array and back): a frame
// moves ~14 slices of 4 MB, and 14 cudaMemcpyAsync calls cost more in launch gaps than the 56 MB cost in bandwidth
struct SliceCopies
{
    const int *src[32];
    int *dst[32];
    int n;
};
__global__ void k_copy_slices(SliceCopies sc, int words)
{
    const int *s = sc.src[blockIdx.y];
    int *d = sc.dst[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x)
        d[i] = s[i];
}

__global__ void k_seed_pos(short2 *__restrict__ pos, int vecSize, int M1, int ref)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it < vecSize)
        pos[(size_t)ref * vecSize + it] = make_short2((short)(it % M1), (short)(it / M1));
}

// ------------------------------------------------------------------------------------------------------
// K_count — `weights` of SVT::Reconstruct (svt.hpp:156-159): number of patches covering each voxel.  It is
// independent of lambda and of the SVT object, so it is computed once per frame.  One thread per (patch, k).
// ------------------------------------------------------------------------------------------------------
__global__ void k_count(const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, int bs,
                        int win, int only_k, unsigned *__restrict__ cnt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nk = (only_k >= 0) ? 1 : win;
    if (t >= (long long)P * nk)
        return;
    const int pi = (int)(t % P);
    const int k = (only_k >= 0) ? only_k : (int)(t / P);
    const short2 p = pos[(size_t)k * vecSize + ids[pi]];
    unsigned *c0 = cnt + (size_t)N * N * k;
    for (int c = 0; c < bs; c++)
        for (int r = 0; r < bs; r++)
            atomicAdd(c0 + (p.x + r) + (size_t)N * (p.y + c), 1u);
}

// ------------------------------------------------------------------------------------------------------
// Casorati gather shared by the SVD kernels (svt.hpp:99-109) with the PGURE perturbations applied on the fly
// (pgure.hpp:80-82): mode 0: U; 1: U + delta1*eps1; 2: U + delta2*eps2; 3: U - delta2*eps2.
// ------------------------------------------------------------------------------------------------------
struct Perturb
{
    const int8_t *d1;
    const int8_t *d2neg;
    int mode;
    double eps;  // eps1 or eps2
    double dNeg; // -sqrt(vQ/vP)
    double dPos; // +sqrt(vP/vQ)
};
__device__ __forceinline__ double load_perturbed(const double *__restrict__ u, size_t vox, const Perturb &pt)
{
    double v = u[vox];
    if (pt.mode == 1)
        v = __dadd_rn(v, __dmul_rn((double)pt.d1[vox], pt.eps));
    else if (pt.mode >= 2)
    {
        const double t = __dmul_rn(pt.d2neg[vox] ? pt.dNeg : pt.dPos, pt.eps);
        v = (pt.mode == 2) ? __dadd_rn(v, t) : __dsub_rn(v, t);
    }
    return v;
}

// window of a perturbed SVT object, U + eps*delta (pgure.hpp:80-82), written out once per object so that the register
// SVD kernel gathers plain doubles (same expression as load_perturbed, hence bit-identical to perturbing on the fly)
__global__ void k_perturb_window(const double *__restrict__ u, Perturb pt, size_t n, double *__restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = load_perturbed(u, i, pt);
}

__device__ __forceinline__ void jacobi_cs(double A, double B, double G, double tol2, double &c, double &s, bool &rot)
{
    c = 1.0;
    s = 0.0;
    if (G * G > tol2 * A * B)
    {
        const double d = B - A;
        const double h = sqrt(d * d + 4.0 * G * G);
        const double t = ((d >= 0.0) ? 2.0 * G : -2.0 * G) / (fabs(d) + h);
        c = rsqrt(1.0 + t * t);
        s = c * t;
        rot = true;
    }
}

// ------------------------------------------------------------------------------------------------------
// K_svd (generic) — thin SVD of every patch's Casorati matrix (arma::svd_econ, svt.hpp:111) by one-sided
// Jacobi, one warp per matrix, matrix and V in shared memory.  Any m = bs^2, n = win.  Fallback for the
// shapes the register kernel below does not cover (e.g. 256x15, 64x31).
// dynamic smem per warp: (m*n + n*n + n) doubles.
// ------------------------------------------------------------------------------------------------------
__global__ void k_svd_smem(const double *__restrict__ u, Perturb pt, const short2 *__restrict__ pos,
                           const int *__restrict__ ids, int P, int vecSize, int N, int bs, int n, int ldv,
                           double *__restrict__ fac, size_t rec, int max_sweeps, double tol2, int *__restrict__ sweeps_out)
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
    for (int k = 0; k < n; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        for (int e = lane; e < m; e += 32)
        {
            const int r = e % bs, c = e / bs;
            A[e + m * k] = load_perturbed(u, (size_t)(p.x + r) + (size_t)N * (p.y + c) + fsz * k, pt);
        }
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
                double a = 0, b = 0, g = 0;
                for (int e = lane; e < m; e += 32)
                {
                    const double x = A[e + m * p], y = A[e + m * q];
                    a = fma(x, x, a);
                    b = fma(y, y, b);
                    g = fma(x, y, g);
                }
                a = warp_sum(a);
                b = warp_sum(b);
                g = warp_sum(g);
                double c, s;
                bool rot = false;
                jacobi_cs(a, b, g, tol2, c, s, rot);
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
        double a = 0;
        for (int e = lane; e < m; e += 32)
            a = fma(A[e + m * j], A[e + m * j], a);
        a = warp_sum(a);
        if (lane == 0)
            sig[j] = sqrt(a);
    }
    __syncwarp();
    double *R = fac + rec * (size_t)pidx;
    for (int j = 0; j < n; j++)
    {
        const double sj = sig[j];
        int rk = 0;
        for (int t = 0; t < n; t++)
            rk += (sig[t] > sj || (sig[t] == sj && t < j)) ? 1 : 0;
        const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
        for (int e = lane; e < m; e += 32)
            R[e + (size_t)m * rk] = A[e + m * j] * inv;
        for (int e = lane; e < n; e += 32)
            R[(size_t)m * n + e + (size_t)ldv * rk] = V[e + n * j];
        if (lane == 0)
            R[(size_t)m * n + (size_t)ldv * n + rk] = sj;
    }
    if (lane == 0 && sweeps_out)
        atomicMax(sweeps_out, sweep + 1);
}

// ------------------------------------------------------------------------------------------------------
// K_svd (register) — 16x15 Casorati matrices (bs = 4, win = 15: the headline configuration).
// Eight lanes per matrix, four matrices per warp.  Lane j of a group keeps rows 2j, 2j+1 of A (16 x 16 with a
// zero 16th column) and of V in registers.  A sweep is 15 rounds of the round-robin (Brent–Luk) ordering: in
// every round the 8 slot pairs (2i, 2i+1) are orthogonalised at once —
//   * 24 partial dot products per lane (alpha, beta, gamma of the 8 pairs over the lane's two rows),
//   * ONE transposing butterfly over the 8 lanes (7 shuffles per quantity) that leaves lane i with the three
//     sums of pair i, so each lane derives a single rotation (one sqrt, one divide, one rsqrt),
//   * (c, s) of pair i broadcast from lane i, rotations applied to the lane's rows of A and V,
//   * columns move one slot along the round-robin cycle (register renaming; slot 0 is fixed).
// After 15 rounds every column is back in its home slot, so convergence is tested per sweep with one vote.
// Singular values are sorted descending and U = A·V/sigma is written with V and S (LAPACK's order, svt.hpp:111).
// ------------------------------------------------------------------------------------------------------
#define SVD16_M 16
#define SVD16_N 15
#define SVD16_LDV 16
#define SVD16_REC (SVD16_M * SVD16_N + SVD16_LDV * SVD16_N + SVD16_LDV) /* 496 doubles */

__device__ __forceinline__ double tr8(const double (&x)[8], int sub)
{
    // transposing reduction of 8 values over the 8 lanes of a group: lane `sub` returns sum over lanes of x[sub]
    double y[4], z[2];
    const bool h4 = (sub & 4) != 0, h2 = (sub & 2) != 0, h1 = (sub & 1) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const double send = h4 ? x[j] : x[j + 4];
        const double keep = h4 ? x[j + 4] : x[j];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
        const double send = h2 ? y[j] : y[j + 2];
        const double keep = h2 ? y[j + 2] : y[j];
        z[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const double send = h1 ? z[0] : z[1];
    const double keep = h1 ? z[1] : z[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__global__ void __launch_bounds__(128)
    k_svd_16x15(const double *__restrict__ u, Perturb pt, const short2 *__restrict__ pos, const int *__restrict__ ids,
                int P, int vecSize, int N, double *__restrict__ fac, int max_sweeps, double tol2,
                int *__restrict__ sweeps_out)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = threadIdx.x & 7;
    int pidx = gtid >> 3;
    const bool valid = pidx < P;
    if (!valid)
        pidx = P - 1; // keep the lanes alive for the shuffles; nothing is written
    const int id = ids[pidx];
    const size_t fsz = (size_t)N * N;
    const int e0 = 2 * sub;
    const int pr = e0 & 3, pc = e0 >> 2;

    double a0[16], a1[16], v0[16], v1[16];
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        const size_t vox = (size_t)(p.x + pr) + (size_t)N * (p.y + pc) + fsz * k;
        a0[k] = load_perturbed(u, vox, pt);
        a1[k] = load_perturbed(u, vox + 1, pt);
    }
    a0[15] = 0.0;
    a1[15] = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j++)
    {
        v0[j] = (j == e0) ? 1.0 : 0.0;
        v1[j] = (j == e0 + 1) ? 1.0 : 0.0;
    }

    int sweep = 0;
#pragma unroll 1
    for (; sweep < max_sweeps; sweep++)
    {
        bool rot = false;
#pragma unroll 1
        for (int round = 0; round < 15; round++)
        {
            double pa[8], pb[8], pg[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const double x0 = a0[2 * i], y0 = a0[2 * i + 1], x1 = a1[2 * i], y1 = a1[2 * i + 1];
                pa[i] = fma(x1, x1, x0 * x0);
                pb[i] = fma(y1, y1, y0 * y0);
                pg[i] = fma(x1, y1, x0 * y0);
            }
            const double A = tr8(pa, sub), B = tr8(pb, sub), G = tr8(pg, sub);
            double c, s;
            jacobi_cs(A, B, G, tol2, c, s, rot);
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const double ci = __shfl_sync(0xffffffffu, c, i, 8);
                const double si = __shfl_sync(0xffffffffu, s, i, 8);
                double x, y;
                x = a0[2 * i], y = a0[2 * i + 1];
                a0[2 * i] = fma(ci, x, -si * y);
                a0[2 * i + 1] = fma(si, x, ci * y);
                x = a1[2 * i], y = a1[2 * i + 1];
                a1[2 * i] = fma(ci, x, -si * y);
                a1[2 * i + 1] = fma(si, x, ci * y);
                x = v0[2 * i], y = v0[2 * i + 1];
                v0[2 * i] = fma(ci, x, -si * y);
                v0[2 * i + 1] = fma(si, x, ci * y);
                x = v1[2 * i], y = v1[2 * i + 1];
                v1[2 * i] = fma(ci, x, -si * y);
                v1[2 * i + 1] = fma(si, x, ci * y);
            }
            // round-robin move: top slots t_i = 2i, bottom b_i = 2i+1; t0 fixed;
            // t1 <- b0, t_i <- t_{i-1} (i>=2), b_i <- b_{i+1} (i<=6), b7 <- t7
#define RR_MOVE(X)                 \
    {                              \
        const double b0_ = X[1];   \
        const double t7_ = X[14];  \
        X[14] = X[12];             \
        X[12] = X[10];             \
        X[10] = X[8];              \
        X[8] = X[6];               \
        X[6] = X[4];               \
        X[4] = X[2];               \
        X[2] = b0_;                \
        X[1] = X[3];               \
        X[3] = X[5];               \
        X[5] = X[7];               \
        X[7] = X[9];               \
        X[9] = X[11];              \
        X[11] = X[13];             \
        X[13] = X[15];             \
        X[15] = t7_;               \
    }
            RR_MOVE(a0)
            RR_MOVE(a1)
            RR_MOVE(v0)
            RR_MOVE(v1)
#undef RR_MOVE
        }
        if (!__any_sync(0xffffffffu, rot))
        {
            sweep++;
            break;
        }
    }

    // singular values: lane `sub` ends up with the squared norms of columns 2*sub and 2*sub+1
    double n2[16];
#pragma unroll
    for (int j = 0; j < 16; j++)
        n2[j] = fma(a1[j], a1[j], a0[j] * a0[j]);
    double q8[8], q4[4], q2[2];
    {
        const bool h4 = (sub & 4) != 0, h2 = (sub & 2) != 0, h1 = (sub & 1) != 0;
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const double send = h4 ? n2[j] : n2[j + 8];
            const double keep = h4 ? n2[j + 8] : n2[j];
            q8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const double send = h2 ? q8[j] : q8[j + 4];
            const double keep = h2 ? q8[j + 4] : q8[j];
            q4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
        for (int j = 0; j < 2; j++)
        {
            const double send = h1 ? q4[j] : q4[j + 2];
            const double keep = h1 ? q4[j + 2] : q4[j];
            q2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
    const double s_lo = sqrt(q2[0]), s_hi = sqrt(q2[1]); // columns 2*sub, 2*sub+1
    double sig[16];
#pragma unroll
    for (int j = 0; j < 16; j++)
        sig[j] = __shfl_sync(0xffffffffu, (j & 1) ? s_hi : s_lo, j >> 1, 8);

    double *R = fac + (size_t)SVD16_REC * pidx;
#pragma unroll
    for (int j = 0; j < SVD16_N; j++)
    {
        int rk = 0;
#pragma unroll
        for (int t = 0; t < SVD16_N; t++)
            rk += (sig[t] > sig[j] || (sig[t] == sig[j] && t < j)) ? 1 : 0;
        const double inv = (sig[j] > 0.0) ? 1.0 / sig[j] : 0.0;
        if (valid)
        {
            *reinterpret_cast<double2 *>(R + e0 + SVD16_M * rk) = make_double2(a0[j] * inv, a1[j] * inv);
            *reinterpret_cast<double2 *>(R + SVD16_M * SVD16_N + e0 + SVD16_LDV * rk) = make_double2(v0[j], v1[j]);
            if ((j >> 1) == sub)
                R[SVD16_M * SVD16_N + SVD16_LDV * SVD16_N + rk] = sig[j];
        }
    }
    if (sweeps_out && (threadIdx.x & 31) == 0)
        atomicMax(sweeps_out, sweep);
}


// ------------------------------------------------------------------------------------------------------
// K_svd (register, v2) — 16x15 Casorati matrices, FOUR lanes per matrix (eight matrices per warp).
// Lane j of a group owns patch column j, i.e. rows 4j..4j+3 of A (16 x 16 with a zero 16th column), in
// registers; V is NOT accumulated during the sweeps.  Per round of the round-robin ordering:
//   * 24 partial dot products over the lane's four rows,
//   * one transposing butterfly over the 4 lanes (6 shuffles per quantity) leaving lane j with the sums of slot
//     pairs 2j and 2j+1, for which it derives the two rotations with approximate-reciprocal/rsqrt + Newton
//     (no IEEE divide/sqrt slow paths),
//   * (c, s) broadcast, rotations applied to the lane's rows, columns moved along the round-robin cycle.
// A sweep in which no pair had |cos| > 1e-6 is the last one (quadratic convergence puts every pair below
// ~1e-12 afterwards), so no extra "check" sweep is spent.
// WARM = 1 pre-multiplies the gathered matrix by the V of SVT object 0 (the unperturbed data): the perturbed
// matrices U +- eps2*delta2 then start with nearly orthogonal columns and converge in about half the sweeps.
// After convergence  U = W / sigma  (W = A V = rotated columns) is written, and  V = A^T U / sigma  is rebuilt
// from the re-gathered A — accurate in the product f(sigma_k) u_k v_k^T that SVT consumes, because the soft
// threshold gives f(sigma_k) <= sigma_k.  Singular values are stored in slot (column) order with
// S[15] = sigma_max; U, V column-major with leading dimension 16.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r;
}
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = r * fma(-hx * r, r, 1.5);
    r = r * fma(-hx * r, r, 1.5);
    return r;
}
// rotation for a slot pair with squared norms A, B and inner product G; flags: rotate if cos^2 > tol2,
// `big` if cos^2 > big2
__device__ __forceinline__ void jacobi_cs_fast(double A, double B, double G, double tol2, double big2, double &c, double &s,
                                               bool &big)
{
    const double g2 = G * G, ab = A * B;
    const bool rot = g2 > tol2 * ab;
    big = big || (g2 > big2 * ab);
    const double d = B - A;
    const double q = fma(d, d, 4.0 * g2);
    const double h = q * rsqrt_fast(q); // sqrt(d^2 + 4 G^2)
    const double t = ((d >= 0.0) ? 2.0 * G : -2.0 * G) * rcp_fast(fabs(d) + h);
    const double cc = rsqrt_fast(fma(t, t, 1.0));
    c = rot ? cc : 1.0;
    s = rot ? cc * t : 0.0;
}

// Rotation for a slot pair whose squared norms A, B are TRACKED (updated after every rotation instead of being
// recomputed from the columns).  The tangent only steers convergence, so it is taken from the unrefined MUFU seeds
// (~2^-20 relative): a rotation then shrinks the pair's cosine by ~1e-6 instead of annihilating it, which the
// quadratically convergent end game does not notice.  The rotation itself stays orthonormal to rounding because
// c = 1/sqrt(1+t^2) is refined to full precision and s = c t.  w is the exchange of squared norm for ANY orthonormal
// (c, s):  A' = c^2 A - 2 c s G + s^2 B = A + w,  B' = B - w  with  w = s (s (B - A) - 2 c G).
__device__ __forceinline__ void jacobi_cs_track(double &A, double &B, double G, double tol2, double big2, double &c, double &s,
                                                bool &big)
{
    const double g2 = G * G, ab = A * B;
    const bool rot = g2 > tol2 * ab;
    big = big || (g2 > big2 * ab);
    const double d = B - A;
    const double q = fma(d, d, 4.0 * g2);
    double rq, rd;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rq) : "d"(q));
    const double den = fma(q, rq, fabs(d)); // |d| + sqrt(d^2 + 4 G^2)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(den));
    const double G2 = G + G;
    const double t = ((d >= 0.0) ? G2 : -G2) * rd;
    const double cc = rsqrt_fast(fma(t, t, 1.0));
    c = rot ? cc : 1.0;
    s = rot ? cc * t : 0.0;
    const double w = s * fma(s, d, -c * G2);
    A += w;
    B -= w;
}

// Fast (scaled) rotation for a slot pair: the columns are kept UNNORMALISED, true column = scale * stored column, so that
// applying  x' = c (x - t y),  y' = c (y + t x)  costs two FMAs per element pair,
//     x_st' = x_st - alpha y_st,   y_st' = y_st + beta x_st,   alpha = t sy / sx,  beta = t sx / sy,
// with the factor c absorbed into the two scales (and 1/c into their tracked reciprocals).  A, B: tracked TRUE squared norms;
// G: inner product of the STORED columns.  c >= 1/sqrt(2), so over the at most 30 x 16 rotations of a column the scales stay
// within 2^-240 .. 1 — no renormalisation is needed.
__device__ __forceinline__ void jacobi_fast_givens(double &A, double &B, double &sx, double &sy, double &isx, double &isy, double G,
                                                   double tol2, double big2, double &alpha, double &beta, bool &big)
{
    const double Gt = G * (sx * sy);
    const double g2 = Gt * Gt, ab = A * B;
    const bool rot = g2 > tol2 * ab;
    big = big || (g2 > big2 * ab);
    const double d = B - A;
    const double q = fma(d, d, 4.0 * g2);
    double rq, rd;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rq) : "d"(q));
    const double den = fma(q, rq, fabs(d)); // |d| + sqrt(d^2 + 4 G^2)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(den));
    const double G2 = Gt + Gt;
    const double t = rot ? ((d >= 0.0) ? G2 : -G2) * rd : 0.0;
    const double x1 = fma(t, t, 1.0);
    const double cc = rsqrt_fast(x1);
    const double c = rot ? cc : 1.0;
    const double h = rot ? x1 * cc : 1.0; // 1 / c
    const double s = c * t;
    const double w = s * fma(s, d, -c * G2);
    A += w;
    B -= w;
    alpha = t * (sy * isx);
    beta = t * (sx * isy);
    sx *= c;
    sy *= c;
    isx *= h;
    isy *= h;
}

// transposing reduction of 8 values over the 4 lanes of a group: lane `sub` gets the sums of x[2*sub], x[2*sub+1]
__device__ __forceinline__ void tr4_8(const double (&x)[8], int sub, double &o0, double &o1)
{
    double y[4];
    const bool h2 = (sub & 2) != 0, h1 = (sub & 1) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const double send = h2 ? x[j] : x[j + 4];
        const double keep = h2 ? x[j + 4] : x[j];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const double send0 = h1 ? y[0] : y[2], send1 = h1 ? y[1] : y[3];
        const double keep0 = h1 ? y[2] : y[0], keep1 = h1 ? y[3] : y[1];
        o0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
        o1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
    }
}
// same for 16 values: lane `sub` gets the sums of x[4*sub .. 4*sub+3]
__device__ __forceinline__ void tr4_16(const double (&x)[16], int sub, double (&o)[4])
{
    double y[8];
    const bool h2 = (sub & 2) != 0, h1 = (sub & 1) != 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const double send = h2 ? x[j] : x[j + 8];
        const double keep = h2 ? x[j + 8] : x[j];
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const double send = h1 ? y[j] : y[j + 4];
        const double keep = h1 ? y[j + 4] : y[j];
        o[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}

// cp.async helpers (16-byte global -> shared copies that bypass L1; completion tracked per thread in groups)
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NN>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(NN) : "memory");
}

#define SVD16_V0_STRIDE 242 /* doubles per matrix in shared memory: 15 x 16 + 2 padding (bank spread) */

// EPI selects what the epilogue leaves behind:
//   0  the full record U | V | S (svt.hpp:111-116) — generic consumers, exact fallback of the lean mode;
//   1  ONLY the head entries of this object: its three largest singular values and the bilinear forms q_k = u_k^T C4 v_k of the
//      two leading triplets (see k_qform3) — what the lambda search consumes of the perturbed objects U +- eps2*delta2.  No U/V
//      stores (3,968 B per patch), V rebuilt for two columns instead of fifteen, no separate q-form pass over the records;
//   2  full record AND head entries (object U: the search rebuilds Uhat from its leading triplets).
// head: 16 doubles per patch = S0[0..2] | S2[0..2] | S3[0..2] | q0[0..1] | q2[0..1] | q3[0..1] | pad; part = 0, 1, 2 for objects U, U2p, U2m.
template <int WARM, int TRACK, int EPI>
__global__ void __launch_bounds__(128, 2)
    k_svd16_l4(const double *__restrict__ u, const short2 *__restrict__ pos, const int *__restrict__ ids, int P,
               int vecSize, int N, double *__restrict__ fac, const double *__restrict__ fac0, int max_sweeps, double tol2,
               double big2, int *__restrict__ sweeps_out, const double *__restrict__ c4, double *__restrict__ head, int part,
               const int *__restrict__ plist = nullptr)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = threadIdx.x & 3; // patch column owned by this lane
    int pidx = gtid >> 2;
    const bool valid = pidx < P; // P = number of matrices of this launch
    if (!valid)
        pidx = P - 1;
    if (plist) // the launch walks a list of patch indices (exact re-decomposition of a few patches)
        pidx = plist[pidx];
    const int id = ids[pidx];
    const size_t fsz = (size_t)N * N;
    extern __shared__ __align__(16) double sv0[]; // WARM: V of object 0 for the block's 32 matrices
    if (WARM)
    { // request V0 (15 columns x 16 doubles) of this lane's matrix now; it is consumed after the gather below
        const double *V0g = fac0 + (size_t)SVD16_REC * pidx + SVD16_M * SVD16_N;
        double *dst = sv0 + (size_t)(threadIdx.x >> 2) * SVD16_V0_STRIDE;
#pragma unroll
        for (int q = 0; q < 30; q++) // 120 16-byte pieces per matrix, 30 per lane
        {
            const int piece = q * 4 + sub;
            cp_async16(dst + 2 * piece, V0g + 2 * piece);
        }
        cp_async_commit();
    }

    // slot 0 is the fixed seat of the round-robin schedule: it holds the zero padding column, so slot pair 0 is a
    // no-op in every round and is skipped statically; real column k lives in slot k+1
    double a[4][16];
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        const size_t vox = (size_t)p.x + (size_t)N * (p.y + sub) + fsz * k;
#pragma unroll
        for (int r = 0; r < 4; r++)
            a[r][k + 1] = __ldg(u + vox + r);
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
        a[r][0] = 0.0;

    if (WARM)
    { // rows of A times V0 (column-major, ld 16): each row independently; the 15 results of a row are parked in a
      // lane-private shared-memory column so that the register file never holds two copies of the matrix
        __shared__ double stage[SVD16_N][128];
        cp_async_wait<0>();
        __syncwarp();
        const double *V0 = sv0 + (size_t)(threadIdx.x >> 2) * SVD16_V0_STRIDE;
        // V0 is orthogonal only if object 0 has full column rank (its V is rebuilt as A^T U / sigma, so a vanishing sigma
        // leaves a zero column — a zero patch, say): such a matrix starts cold instead
        bool v0_ok;
        {
            const double *S0 = fac0 + (size_t)SVD16_REC * pidx + SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
            double lo = fmin(fmin(S0[4 * sub], S0[4 * sub + 1]), fmin(S0[4 * sub + 2], (sub == 3) ? S0[4 * sub + 2] : S0[4 * sub + 3]));
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, 2));
            v0_ok = lo > S0[15] * 1e-8;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            if (!v0_ok)
                break;
#pragma unroll 1
            for (int j = 0; j < SVD16_N; j++)
            {
                const double2 *col = reinterpret_cast<const double2 *>(V0 + SVD16_LDV * j);
                double acc = 0.0;
#pragma unroll
                for (int i2 = 0; i2 < 8; i2++)
                {
                    const double2 v = col[i2];
                    acc = fma(a[r][2 * i2 + 1], v.x, acc);
                    if (2 * i2 + 1 < SVD16_N)
                        acc = fma(a[r][2 * i2 + 2], v.y, acc);
                }
                stage[j][threadIdx.x] = acc;
            }
#pragma unroll
            for (int j = 0; j < SVD16_N; j++)
                a[r][j + 1] = stage[j][threadIdx.x];
        }
    }

    int sweep = 0;
    double nr[4] = {0.0, 0.0, 0.0, 0.0}; // TRACK: squared norms of slots 4*sub .. 4*sub+3
    double sc[4] = {1.0, 1.0, 1.0, 1.0}, isc[4] = {1.0, 1.0, 1.0, 1.0}; // TRACK == 2: column scales of those slots and reciprocals
    int quiet = 0; // consecutive rounds (warp-wide) without a rotation above the `big` threshold
#pragma unroll 1
    for (; sweep < max_sweeps;)
    {
        bool big = false;
        if (TRACK)
        { // fresh squared norms at the start of every sweep (the updates below drift by rounding only)
            double n2s[16];
            n2s[0] = 0.0;
#pragma unroll
            for (int j = 1; j < 16; j++)
            {
                double sacc = 0.0;
#pragma unroll
                for (int r = 0; r < 4; r++)
                    sacc = fma(a[r][j], a[r][j], sacc);
                n2s[j] = sacc;
            }
            tr4_16(n2s, sub, nr);
            if (TRACK == 2)
            {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    nr[j] *= sc[j] * sc[j];
            }
        }
        // 16 rounds = the 15-round cycle plus its first pair set again: an even count lets the body be unrolled by 2, which
        // removes the register moves of RR_MOVE; the convergence exit is taken between such double rounds only (an exit
        // inside the unrolled body would bring the moves back)
#pragma unroll 1
        for (int rp = 0; rp < 8 && quiet < 15; rp++)
#pragma unroll
        for (int half = 0; half < 2; half++)
        {
            double c0, s0, c1, s1;
            if (TRACK)
            {
                double pg[8];
                pg[0] = 0.0;
#pragma unroll
                for (int i = 1; i < 8; i++)
                {
                    double sg = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        sg = fma(a[r][2 * i], a[r][2 * i + 1], sg);
                    pg[i] = sg;
                }
                double G0, G1;
                tr4_8(pg, sub, G0, G1);
                if (TRACK == 2)
                { // c0/s0, c1/s1 carry (alpha, beta) of the two pairs
                    jacobi_fast_givens(nr[0], nr[1], sc[0], sc[1], isc[0], isc[1], G0, tol2, big2, c0, s0, big);
                    jacobi_fast_givens(nr[2], nr[3], sc[2], sc[3], isc[2], isc[3], G1, tol2, big2, c1, s1, big);
                }
                else
                {
                    jacobi_cs_track(nr[0], nr[1], G0, tol2, big2, c0, s0, big);
                    jacobi_cs_track(nr[2], nr[3], G1, tol2, big2, c1, s1, big);
                }
                // the norms travel with their columns (RR_MOVE below): slot 4s <- 4s-2, 4s+2 <- 4s, 4s+1 <- 4s+3, 4s+3 <- 4s+5
#define SLOT_STATE_MOVE(X, PAD)                                      \
    {                                                                \
        const double up = __shfl_up_sync(0xffffffffu, X[2], 1, 4);   \
        const double dn = __shfl_down_sync(0xffffffffu, X[1], 1, 4); \
        const double o0 = X[0], o1 = X[1], o2 = X[2], o3 = X[3];     \
        X[0] = (sub == 0) ? PAD : up;                                \
        X[1] = o3;                                                   \
        X[2] = (sub == 0) ? o1 : o0;                                 \
        X[3] = (sub == 3) ? o2 : dn;                                 \
    }
                SLOT_STATE_MOVE(nr, 0.0)
                if (TRACK == 2)
                {
                    SLOT_STATE_MOVE(sc, 1.0)
                    SLOT_STATE_MOVE(isc, 1.0)
                }
#undef SLOT_STATE_MOVE
            }
            else
            {
                double pa[8], pb[8], pg[8];
                pa[0] = pb[0] = pg[0] = 0.0;
#pragma unroll
                for (int i = 1; i < 8; i++)
                {
                    double sa = 0.0, sb = 0.0, sg = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                    {
                        const double x = a[r][2 * i], y = a[r][2 * i + 1];
                        sa = fma(x, x, sa);
                        sb = fma(y, y, sb);
                        sg = fma(x, y, sg);
                    }
                    pa[i] = sa;
                    pb[i] = sb;
                    pg[i] = sg;
                }
                double A0, A1, B0, B1, G0, G1;
                tr4_8(pa, sub, A0, A1);
                tr4_8(pb, sub, B0, B1);
                tr4_8(pg, sub, G0, G1);
                jacobi_cs_fast(A0, B0, G0, tol2, big2, c0, s0, big);
                jacobi_cs_fast(A1, B1, G1, tol2, big2, c1, s1, big);
            }
#pragma unroll
            for (int i = 1; i < 8; i++)
            {
                const double ci = __shfl_sync(0xffffffffu, (i & 1) ? c1 : c0, i >> 1, 4);
                const double si = __shfl_sync(0xffffffffu, (i & 1) ? s1 : s0, i >> 1, 4);
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    const double x = a[r][2 * i], y = a[r][2 * i + 1];
                    if (TRACK == 2)
                    { // (ci, si) = (alpha, beta) of the fast rotation
                        a[r][2 * i] = fma(-ci, y, x);
                        a[r][2 * i + 1] = fma(si, x, y);
                    }
                    else
                    {
                        a[r][2 * i] = fma(ci, x, -si * y);
                        a[r][2 * i + 1] = fma(si, x, ci * y);
                    }
                }
            }
#define RR_MOVE(X)                 \
    {                              \
        const double b0_ = X[1];   \
        const double t7_ = X[14];  \
        X[14] = X[12];             \
        X[12] = X[10];             \
        X[10] = X[8];              \
        X[8] = X[6];               \
        X[6] = X[4];               \
        X[4] = X[2];               \
        X[2] = b0_;                \
        X[1] = X[3];               \
        X[3] = X[5];               \
        X[5] = X[7];               \
        X[7] = X[9];               \
        X[9] = X[11];              \
        X[11] = X[13];             \
        X[13] = X[15];             \
        X[15] = t7_;               \
    }
            RR_MOVE(a[0])
            RR_MOVE(a[1])
            RR_MOVE(a[2])
            RR_MOVE(a[3])
#undef RR_MOVE
            // every pair is visited once in any 15 consecutive rounds: 15 quiet rounds in a row mean that all pairs were below
            // the threshold when last rotated, i.e. the state "a whole sweep without a big rotation" — reached here at round
            // granularity instead of at the next sweep boundary
            {
                const bool any_big = __any_sync(0xffffffffu, big);
                quiet = any_big ? 0 : quiet + 1;
                big = false;
            }
        }
        sweep++;
        if (quiet >= 15)
            break;
    }

    if (TRACK == 2)
    { // true columns = scale * stored columns; the scale of slot j sits in lane j >> 2
#pragma unroll
        for (int j = 1; j < 16; j++)
        {
            const double sj = __shfl_sync(0xffffffffu, sc[j & 3], j >> 2, 4);
#pragma unroll
            for (int r = 0; r < 4; r++)
                a[r][j] *= sj;
        }
    }
    // squared column norms: lane `sub` gets columns 4*sub .. 4*sub+3, then everybody gets all 16 sigmas
    double n2[16], q[4];
#pragma unroll
    for (int j = 0; j < 16; j++)
    { // n2[j] = squared norm of column j (slot j+1); n2[15] = 0
        double sacc = 0.0;
        if (j < SVD16_N)
        {
#pragma unroll
            for (int r = 0; r < 4; r++)
                sacc = fma(a[r][j + 1], a[r][j + 1], sacc);
        }
        n2[j] = sacc;
    }
    tr4_16(n2, sub, q);
    double sig[16], so[4];
#pragma unroll
    for (int jl = 0; jl < 4; jl++)
        so[jl] = sqrt(q[jl]); // sigma of the lane's own columns 4*sub .. 4*sub+3
#pragma unroll
    for (int j = 0; j < 16; j++)
        sig[j] = __shfl_sync(0xffffffffu, so[j & 3], j >> 2, 4);
    double smax = 0.0;
#pragma unroll
    for (int j = 0; j < SVD16_N; j++)
        smax = fmax(smax, sig[j]);

    // descending order like LAPACK (svt.hpp:111): every lane ranks its own four columns, then the ranks are exchanged
    int rk[SVD16_N], rko[4];
#pragma unroll
    for (int jl = 0; jl < 4; jl++)
    {
        const int j = 4 * sub + jl;
        int rr = 0;
#pragma unroll
        for (int t2 = 0; t2 < SVD16_N; t2++)
            rr += (sig[t2] > so[jl] || (sig[t2] == so[jl] && t2 < j)) ? 1 : 0;
        rko[jl] = rr;
    }
#pragma unroll
    for (int j = 0; j < SVD16_N; j++)
        rk[j] = __shfl_sync(0xffffffffu, rko[j & 3], j >> 2, 4);
    double *R = fac + (size_t)SVD16_REC * pidx;
    // U = W / sigma (columns with sigma below 1e-20 sigma_max carry nothing after thresholding: set to zero);
    // afterwards a[r][j+1] holds z = w / sigma^2 for the V rebuild
#pragma unroll
    for (int j = 0; j < SVD16_N; j++)
    {
        const double inv = (sig[j] > smax * 1e-20 && sig[j] > 0.0) ? 1.0 / sig[j] : 0.0;
        double uu[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            uu[r] = a[r][j + 1] * inv;
            a[r][j + 1] = uu[r] * inv;
        }
        if (valid)
        {
            if (EPI != 1)
            {
                double2 *dst = reinterpret_cast<double2 *>(R + SVD16_M * rk[j] + 4 * sub);
                dst[0] = make_double2(uu[0], uu[1]);
                dst[1] = make_double2(uu[2], uu[3]);
                if ((j >> 2) == sub)
                    R[SVD16_M * SVD16_N + SVD16_LDV * SVD16_N + rk[j]] = sig[j];
            }
            if (EPI != 0 && (j >> 2) == sub && rk[j] < 3)
                head[(size_t)16 * pidx + 3 * part + rk[j]] = sig[j];
        }
    }
    if (EPI != 1 && valid && sub == 3)
        R[SVD16_M * SVD16_N + SVD16_LDV * SVD16_N + 15] = smax; // slot 15 carries sigma_max
    // slots and singular values of the two leading triplets (head entries)
    int lead0 = 0, lead1 = 0;
    double sl0 = 0.0, sl1 = 0.0;
    if (EPI != 0)
    {
#pragma unroll
        for (int j = 0; j < SVD16_N; j++)
        {
            if (rk[j] == 0)
                lead0 = j, sl0 = sig[j];
            if (rk[j] == 1)
                lead1 = j, sl1 = sig[j];
        }
    }
    // V(i, k) = sum_rows A(row, i) * z(row, k).  z is handed over through shared memory (the V0 buffer of the warm start is
    // dead by now; the cold kernel is launched with the same allocation): lane `sub` then owns rows i = 4*sub .. 4*sub+3 of V,
    // re-gathers those four complete columns of A (16 values each) and forms all 15 inner products per column locally —
    // no cross-lane reduction (the shuffle/select tree this replaces cost as much as seven Jacobi rounds).
    double *zs = sv0 + (size_t)(threadIdx.x >> 2) * SVD16_V0_STRIDE; // z(row, k) at zs[16 * k + row]
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        double2 *dst = reinterpret_cast<double2 *>(zs + 16 * k + 4 * sub);
        dst[0] = make_double2(a[0][k + 1], a[1][k + 1]);
        dst[1] = make_double2(a[2][k + 1], a[3][k + 1]);
    }
    double ac[4][16]; // [column 4*sub + ii of A][row]
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
    {
        const int i = min(4 * sub + ii, SVD16_N - 1); // (lane 3's fourth column is the zero padding row of V: value unused)
        const short2 p = pos[(size_t)i * vecSize + id];
        const size_t vox = (size_t)p.x + (size_t)N * p.y + fsz * i;
#pragma unroll
        for (int c2 = 0; c2 < 4; c2++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                ac[ii][4 * c2 + r] = __ldg(u + vox + (size_t)N * c2 + r);
    }
    __syncwarp();
    // head entries: t_k(i) = sigma_k sum_rows C4(row, i) z_k(row) for the two leading triplets and this lane's four slices;
    // q_k = sum_i v_k(i) t_k(i) is finished below once v_k(i) is known
    double t0[4] = {0.0, 0.0, 0.0, 0.0}, t1[4] = {0.0, 0.0, 0.0, 0.0};
    if (EPI != 0)
    {
        double z0[16], z1[16];
#pragma unroll
        for (int h2 = 0; h2 < 8; h2++)
        {
            const double2 v0 = reinterpret_cast<const double2 *>(zs + 16 * lead0)[h2];
            const double2 v1 = reinterpret_cast<const double2 *>(zs + 16 * lead1)[h2];
            z0[2 * h2] = v0.x, z0[2 * h2 + 1] = v0.y;
            z1[2 * h2] = v1.x, z1[2 * h2 + 1] = v1.y;
        }
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
        {
            const int i = min(4 * sub + ii, SVD16_N - 1);
            const short2 p = pos[(size_t)i * vecSize + id];
            const size_t vox = (size_t)p.x + (size_t)N * p.y + fsz * i;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c2 = 0; c2 < 4; c2++)
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    const double cw = __ldg(c4 + vox + (size_t)N * c2 + r);
                    s0 = fma(cw, z0[4 * c2 + r], s0);
                    s1 = fma(cw, z1[4 * c2 + r], s1);
                }
            const bool real = 4 * sub + ii < SVD16_N;
            t0[ii] = real ? s0 * sl0 : 0.0;
            t1[ii] = real ? s1 * sl1 : 0.0;
        }
    }
    double *Vg = R + SVD16_M * SVD16_N;
    double q0acc = 0.0, q1acc = 0.0;
    if (EPI == 1)
    { // only the two leading columns of V are formed (never stored)
#pragma unroll 1
        for (int w2 = 0; w2 < 2; w2++)
        {
            const int k = w2 ? lead1 : lead0;
            double zk[16];
#pragma unroll
            for (int h2 = 0; h2 < 8; h2++)
            {
                const double2 v = reinterpret_cast<const double2 *>(zs + 16 * k)[h2];
                zk[2 * h2] = v.x;
                zk[2 * h2 + 1] = v.y;
            }
            double vo[4];
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
            {
                double sacc = 0.0;
#pragma unroll
                for (int row = 0; row < 16; row++)
                    sacc = fma(ac[ii][row], zk[row], sacc);
                vo[ii] = sacc;
            }
            if (sub == 3)
                vo[3] = 0.0;
            if (w2 == 0)
                q0acc = fma(vo[0], t0[0], fma(vo[1], t0[1], fma(vo[2], t0[2], vo[3] * t0[3])));
            else
                q1acc = fma(vo[0], t1[0], fma(vo[1], t1[1], fma(vo[2], t1[2], vo[3] * t1[3])));
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < SVD16_N; k++)
        {
            double zk[16];
#pragma unroll
            for (int h2 = 0; h2 < 8; h2++)
            {
                const double2 v = reinterpret_cast<const double2 *>(zs + 16 * k)[h2];
                zk[2 * h2] = v.x;
                zk[2 * h2 + 1] = v.y;
            }
            double vo[4];
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
            {
                double sacc = 0.0;
#pragma unroll
                for (int row = 0; row < 16; row++)
                    sacc = fma(ac[ii][row], zk[row], sacc);
                vo[ii] = sacc;
            }
            if (sub == 3)
                vo[3] = 0.0; // row 15 of V is padding
            if (EPI != 0)
            {
                if (k == lead0)
                    q0acc = fma(vo[0], t0[0], fma(vo[1], t0[1], fma(vo[2], t0[2], vo[3] * t0[3])));
                if (k == lead1)
                    q1acc = fma(vo[0], t1[0], fma(vo[1], t1[1], fma(vo[2], t1[2], vo[3] * t1[3])));
            }
            if (valid)
            {
                double2 *dst = reinterpret_cast<double2 *>(Vg + SVD16_LDV * rk[k] + 4 * sub);
                dst[0] = make_double2(vo[0], vo[1]);
                dst[1] = make_double2(vo[2], vo[3]);
            }
        }
    }
    if (EPI != 0)
    {
        q0acc += __shfl_xor_sync(0xffffffffu, q0acc, 1);
        q0acc += __shfl_xor_sync(0xffffffffu, q0acc, 2);
        q1acc += __shfl_xor_sync(0xffffffffu, q1acc, 1);
        q1acc += __shfl_xor_sync(0xffffffffu, q1acc, 2);
        if (valid && sub == 0)
        {
            head[(size_t)16 * pidx + 9 + 2 * part] = q0acc;
            head[(size_t)16 * pidx + 9 + 2 * part + 1] = q1acc;
        }
    }
    if (sweeps_out && (threadIdx.x & 31) == 0)
    {
        atomicMax(sweeps_out, sweep);
        atomicAdd(sweeps_out + 1, sweep);
    }
}

// ------------------------------------------------------------------------------------------------------
// K_top1 — the perturbed objects U +- eps2*delta2 of the lean PGURE path (pgure.hpp:81-82,136).  The lambda search consumes of
// them: the thresholded spectrum (through the second-difference sum) and the q-forms of the surviving triplets.  With the
// reference's exponential weighting (and on Poisson-like data in general) ONE triplet survives at the probed lambdas, so the
// full 16x15 SVD (~100 Jacobi rounds) is replaced by
//   * the dominant triplet by power iteration on A^T A, started from object U's v_1 (the perturbation is ~1 % of the data):
//     sigma_1, u_1, v_1 to full FP64 accuracy, and its q-form u_1^T C4 v_1;
//   * a RIGOROUS upper bound on every other singular value (Weyl):  sigma_k(A + E) <= sigma_2(A) + ||E||_F for k >= 2, with
//     sigma_2(A) from object U's exact spectrum and ||E||_F = eps2 * sqrt(sum delta2^2) over the patch's 240 entries.
// The head record gets S = (sigma_1, bound, -1): the -1 marks "S[1] is a bound".  At a probe where the bound survives the
// threshold (k_lean_check) — or if the iteration did not converge — the patch is decomposed exactly by k_svd16_l4 before the
// evaluation, so every probe is answered exactly (soft_f is monotone in s: a bound that does not survive proves that no
// singular value below it does).  4 lanes per matrix like k_svd16_l4.
// ------------------------------------------------------------------------------------------------------
// MODE 0: perturbed object, object U decomposed exactly (bound = min(Weyl, Gram), start and shift from U's record);
// MODE 1: perturbed object, object U itself in top-1 form (bound from the Gram matrix; start from U's v_1);
// MODE 2: object U: cold start from the constant vector; leaves a minimal record (u_1, v_1, S = (sigma_1, 0, ..), slot 15 =
//         sigma_1) for the consumers of the leading triplet, plus the head entries.
// Gram bound: with G = A^T A - sigma_1^2 v_1 v_1^T (eigenvalues sigma_2^2 .. sigma_15^2),  sigma_2^4 <= sum_k>=2 sigma_k^4 =
// ||G||_F^2, i.e. sigma_2 <= ||G||_F^(1/2) — 1.2 x sigma_2 on noise-dominated patches (the Frobenius norm of the residual
// itself is 1.9 x: too loose, the bound would "survive" on a sixth of the patches at the lambdas the search settles on).
template <int MODE>
__global__ void __launch_bounds__(128, 2)
    k_top1_l4(const double *__restrict__ u, const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N,
              double *__restrict__ fac0, const int8_t *__restrict__ d2neg, double eps2, double dNeg, double dPos,
              const double *__restrict__ c4, double *__restrict__ head, int part, int max_iters, int *__restrict__ iters_out)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = threadIdx.x & 3;
    int pidx = gtid >> 2;
    const bool valid = pidx < P;
    if (!valid)
        pidx = P - 1;
    const int id = ids[pidx];
    const size_t fsz = (size_t)N * N;
    double a[4][SVD16_N];
    int nneg = 0;
    double fro2 = 0.0;
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        const size_t vox = (size_t)p.x + (size_t)N * (p.y + sub) + fsz * k;
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            a[r][k] = __ldg(u + vox + r);
            fro2 = fma(a[r][k], a[r][k], fro2);
            if (MODE == 0)
                nneg += d2neg[vox + r] ? 1 : 0;
        }
    }
    fro2 += __shfl_xor_sync(0xffffffffu, fro2, 1);
    fro2 += __shfl_xor_sync(0xffffffffu, fro2, 2);
    double *R0 = fac0 + (size_t)SVD16_REC * pidx;
    double x[SVD16_N];
    double xn2 = 0.0;
#pragma unroll
    for (int j = 0; j < SVD16_N; j++)
    {
        x[j] = (MODE == 2) ? 0.0 : R0[SVD16_M * SVD16_N + j]; // v_1 of object U
        xn2 = fma(x[j], x[j], xn2);
    }
    if (!(xn2 > 0.5))
    { // object U itself, or a rank-deficient object U (zero patch): start from the constant vector
#pragma unroll
        for (int j = 0; j < SVD16_N; j++)
            x[j] = 0.2581988897471611; // 1 / sqrt(15)
    }
    // shifted iteration x <- (A^T A - mu) x with mu near the centre of the unwanted spectrum [sigma_15^2, sigma_2^2]: contraction
    // (sigma_2^2 - sigma_15^2) / (2 sigma_1^2 - sigma_2^2 - sigma_15^2) per step instead of (sigma_2 / sigma_1)^2.  The shift only
    // steers convergence; sigma_1 and u_1 come from the final A v_1.
    double mu_shift = 0.0;
    if (MODE == 0)
    {
        const double *S0 = R0 + SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
        mu_shift = 0.5 * (S0[1] * S0[1] + S0[14] * S0[14]);
        if (!(mu_shift < 0.25 * S0[0] * S0[0]))
            mu_shift = 0.0;
    }
    int it = 0;
    bool conv = false;
    double y[4];
#pragma unroll 1
    for (; it < max_iters; it++)
    {
        double y2 = 0.0;
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j = 0; j < SVD16_N; j += 3)
            {
                s0 = fma(a[r][j], x[j], s0);
                s1 = fma(a[r][j + 1], x[j + 1], s1);
                s2 = fma(a[r][j + 2], x[j + 2], s2);
            }
            y[r] = (s0 + s1) + s2;
            y2 = fma(y[r], y[r], y2);
        }
        if (MODE != 0 && it == 0)
        { // no spectrum of object U to take the shift from: mean of the unwanted eigenvalues from ||A||_F^2 - ||A x0||^2
            y2 += __shfl_xor_sync(0xffffffffu, y2, 1);
            y2 += __shfl_xor_sync(0xffffffffu, y2, 2);
            mu_shift = fmax(fro2 - y2, 0.0) * (1.0 / 14.0);
            if (!(mu_shift < 0.25 * y2))
                mu_shift = 0.0;
        }
        double z[SVD16_N], n2 = 0.0;
#pragma unroll
        for (int j = 0; j < SVD16_N; j++)
        {
            double t = fma(a[0][j], y[0], fma(a[1][j], y[1], fma(a[2][j], y[2], a[3][j] * y[3])));
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t = fma(-mu_shift, x[j], t);
            z[j] = t;
            n2 = fma(t, t, n2);
        }
        if (!(n2 > 0.0))
        { // zero matrix: nothing to iterate on
            conv = true;
        }
        else if (!conv)
        {
            const double inv = rsqrt(n2);
            double diff = 0.0;
#pragma unroll
            for (int j = 0; j < SVD16_N; j++)
            {
                const double xn = z[j] * inv;
                diff = fmax(diff, fabs(xn - x[j]));
                x[j] = xn;
            }
            conv = diff <= 1.0e-15;
        }
        if (__all_sync(0xffffffffu, conv))
            break;
    }
    // final pass with the converged v_1: sigma_1 = ||A v_1||, u_1 = A v_1 / sigma_1
    double s2sum = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < SVD16_N; j += 3)
        {
            s0 = fma(a[r][j], x[j], s0);
            s1 = fma(a[r][j + 1], x[j + 1], s1);
            s2 = fma(a[r][j + 2], x[j + 2], s2);
        }
        y[r] = (s0 + s1) + s2;
        s2sum = fma(y[r], y[r], s2sum);
    }
    s2sum += __shfl_xor_sync(0xffffffffu, s2sum, 1);
    s2sum += __shfl_xor_sync(0xffffffffu, s2sum, 2);
    const double sig1 = sqrt(s2sum);
    const double isig = sig1 > 0.0 ? 1.0 / sig1 : 0.0;
    // ||G||_F^2 = ||A^T A||_F^2 - sigma_1^4 (A^T A v_1 = sigma_1^2 v_1): the 120 entries of the upper triangle of A^T A in groups
    // of 16 partial inner products, each group summed over the four lanes by one transposing reduction (tr4_16); off-diagonal
    // entries carry a factor sqrt(2) so that their squares count twice.  The subtraction cancels (sigma_2 / sigma_1)^4 of the
    // leading digits: 1e-13 sigma_1^4 is added back as the rounding allowance that keeps the bound rigorous.
    double g2 = 0.0;
    {
        double part16[16];
        int cnt16 = 0;
#pragma unroll
        for (int i = 0; i < SVD16_N; i++)
#pragma unroll
            for (int j = i; j < SVD16_N; j++)
            {
                const double pij = fma(a[0][i], a[0][j], fma(a[1][i], a[1][j], fma(a[2][i], a[2][j], a[3][i] * a[3][j])));
                part16[cnt16] = (i == j) ? pij : 1.4142135623730951 * pij;
                cnt16++;
                if (cnt16 == 16 || (i == SVD16_N - 1 && j == SVD16_N - 1))
                {
#pragma unroll
                    for (int q = cnt16; q < 16; q++)
                        part16[q] = 0.0;
                    double tot[4];
                    tr4_16(part16, sub, tot);
                    g2 = fma(tot[0], tot[0], fma(tot[1], tot[1], fma(tot[2], tot[2], fma(tot[3], tot[3], g2))));
                    cnt16 = 0;
                }
            }
    }
    g2 += __shfl_xor_sync(0xffffffffu, g2, 1);
    g2 += __shfl_xor_sync(0xffffffffu, g2, 2);
    g2 = fmax(g2 * (1.0 + 1e-13) - s2sum * s2sum, 0.0) + 1e-13 * s2sum * s2sum;
    // q-form u_1^T C4 v_1 (C4 gathered along the trajectory, the same voxels as the matrix)
    double qf = 0.0;
    {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < SVD16_N; k++)
        {
            const short2 p = pos[(size_t)k * vecSize + id];
            const size_t vox = (size_t)p.x + (size_t)N * (p.y + sub) + fsz * k;
#pragma unroll
            for (int r = 0; r < 4; r++)
                t[r] = fma(__ldg(c4 + vox + r), x[k], t[r]);
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
            qf = fma(t[r], y[r] * isig, qf);
    }
    qf += __shfl_xor_sync(0xffffffffu, qf, 1);
    qf += __shfl_xor_sync(0xffffffffu, qf, 2);
    if (MODE == 0)
    {
        nneg += __shfl_xor_sync(0xffffffffu, nneg, 1);
        nneg += __shfl_xor_sync(0xffffffffu, nneg, 2);
    }
    double bound = sqrt(sqrt(g2)) * (1.0 + 1e-9);
    if (MODE == 0)
    { // Weyl: sigma_k(A + E) <= sigma_2(A) + ||E||_F for k >= 2, with object U's exact sigma_2
        const double e2 = (double)nneg * dNeg * dNeg + (double)(SVD16_M * SVD16_N - nneg) * dPos * dPos;
        const double s2u = R0[SVD16_M * SVD16_N + SVD16_LDV * SVD16_N + 1];
        bound = fmin(bound, (s2u + eps2 * sqrt(e2)) * (1.0 + 1e-12));
    }
    if (!conv)
        bound = INFINITY; // not converged within max_iters (dominant pair not separated): exact decomposition on first use
    if (MODE == 2 && valid)
    { // minimal record of object U: u_1, v_1, S = (sigma_1, 0, ...), slot 15 = sigma_max
        double2 *ud = reinterpret_cast<double2 *>(R0 + 4 * sub);
        ud[0] = make_double2(y[0] * isig, y[1] * isig);
        ud[1] = make_double2(y[2] * isig, y[3] * isig);
        double *Vd = R0 + SVD16_M * SVD16_N, *Sd = R0 + SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
#pragma unroll
        for (int e = 0; e < 4; e++)
        {
            const int j = 4 * sub + e;
            double xv = 0.0;
#pragma unroll
            for (int q = 0; q < SVD16_N; q++)
                xv = (q == j) ? x[q] : xv;
            Vd[j] = (j < SVD16_N) ? xv : 0.0;
            Sd[j] = (j == 0 || j == 15) ? sig1 : 0.0;
        }
    }
    if (valid && sub == 0)
    {
        double *hd = head + (size_t)16 * pidx;
        hd[3 * part] = sig1;
        hd[3 * part + 1] = bound;
        hd[3 * part + 2] = -1.0;
        hd[9 + 2 * part] = qf;
        hd[9 + 2 * part + 1] = 0.0;
    }
    if (iters_out && (threadIdx.x & 31) == 0)
    {
        atomicMax(iters_out, it + 1);
        atomicAdd(iters_out + 1, it + 1);
    }
}

// Lean path, per frame: the lambda beyond which (exponential weighting) / below which (plain thresholding) the first bound of
// a marked head record survives the threshold — probes on the safe side of it need no check at all.
//   exponential weighting: f(B) > 0  <=>  B > s1 exp(-lambda B^2 / 2)  <=>  lambda > 2 ln(s1 / B) / B^2;  out[0] = min over patches
//   plain:                 f(B) > 0  <=>  lambda < B;                                                    out[1] = max over patches
// out must be preset to {+inf, 0}; doubles >= 0 order like their bit patterns, so atomicMin/Max on the bits is exact.
__global__ void k_lean_crit(const double *__restrict__ head, int P, double *__restrict__ out)
{
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    double lmin = INFINITY, bmax = 0.0;
    if (pidx < P)
    {
        const double *hd = head + (size_t)16 * pidx;
#pragma unroll
        for (int part = 0; part <= 2; part++)
            if (hd[3 * part + 2] < 0.0)
            {
                const double s1 = hd[3 * part], B = hd[3 * part + 1];
                bmax = fmax(bmax, B);
                double lc = 0.0; // B >= s1 or B not finite: survives at any lambda
                if (B < s1 && B > 0.0)
                    lc = 2.0 * log(s1 / B) / (B * B) * (1.0 - 1e-9);
                else if (B <= 0.0)
                    lc = INFINITY;
                lmin = fmin(lmin, lc);
            }
    }
    lmin = -warp_max(-lmin);
    bmax = warp_max(bmax);
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(fmax(lmin, 0.0)));
        atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)__double_as_longlong(fmin(bmax, 1.7e308)));
    }
}

// Lean path, per probe beyond the critical lambda: patches whose bound survives at this lambda -> list (list[0] = count)
__global__ void k_lean_check(const double *__restrict__ head, int P, double lambda, int expw, int *__restrict__ list)
{
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= P)
        return;
    const double *hd = head + (size_t)16 * pidx;
    bool off = false;
#pragma unroll
    for (int part = 0; part <= 2; part++)
        if (hd[3 * part + 2] < 0.0)
        {
            const double s1 = hd[3 * part], B = hd[3 * part + 1];
            const double w = expw ? fabs(s1 * exp(-0.5 * lambda * (B * B))) : lambda;
            off = off || !(B - w <= 0.0);
        }
    if (off)
        list[1 + atomicAdd(list, 1)] = pidx;
}

// ------------------------------------------------------------------------------------------------------
// K_eval3 — one PGURE objective evaluation for the 16x15 / reference-eps1 configuration, fused over the
// three SVT objects (U, U+eps2*delta2, U-eps2*delta2) of every patch (pgure.hpp:130-136, svt.hpp:121-164).
// Only Uhat enters the risk non-linearly (sum (Uhat-U)^2); U2p and U2m enter through
//   s4 = sum_voxels delta2 * (U2p - 2 Uhat + U2m) = sum_patches sum_entries (delta2/weights)(voxel) * (b2p - 2 b0 + b2m)
// so their blocks never need the overlap-add: one group of 16 lanes rebuilds the three 16x15 blocks of a patch
// (rank-adaptively), adds b0 into the Uhat accumulator (FP64 RED) and folds the second difference straight
// into a per-block partial sum.  partial: gridDim.x doubles (s4 part).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double soft_f(double s, double smax, double lambda, int expw)
{
    const double w = expw ? fabs(smax * exp(-0.5 * lambda * (s * s))) : lambda;
    return fmax(s - w, 0.0); // s >= 0 from the Jacobi kernels
}

#define EV_C 3                  /* singular triplets of object 0 staged per chunk */
#define EV_CH (EV_C * 32)       /* doubles per chunk: C columns x (16 of U | 16 of V) */
#define EV_GRP (96 + 2 * EV_CH) /* doubles of shared memory per patch: S and q of 3 objects | two chunk buffers */

// c4 = delta2 / weights (0 where no patch covers the voxel: svt.hpp:163-164 sets those voxels to 0), once per frame
__global__ void k_c4(const unsigned *__restrict__ cnt, const int8_t *__restrict__ d2neg, double dNeg, double dPos, size_t n,
                     double *__restrict__ c4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        c4[i] = cnt[i] ? (d2neg[i] ? dNeg : dPos) / (double)cnt[i] : 0.0;
}

// K_qform — once per frame: q[obj][patch][k] = u_k^T C4_patch v_k for the three SVT objects, where C4_patch is the
// per-voxel multiplier delta2 / weights gathered along the patch trajectory (a 16 x 15 matrix).  The second-difference
// term of the risk,  s4 = sum_voxels delta2 (U2p - 2 Uhat + U2m)  (pgure.hpp:136), is LINEAR in the reconstructed
// blocks b = sum_k f_k(lambda) u_k v_k^T, so for every lambda it collapses to  sum_patches sum_k f_k q_k  — the
// perturbed objects never have to be rebuilt or overlap-added during the lambda search, only thresholded.
// 16 lanes per patch (lane g = block row g), 8 patches per CTA; the U and V of the three objects (11.5 KB per patch)
// are streamed into shared memory with cp.async while the C4 gather is in flight.
// q: 16 doubles per patch and object (slot 15 = 0).  dynamic smem: 8 * 3 * 480 doubles.
__global__ void __launch_bounds__(128, 2)
    k_qform3(const double *__restrict__ fac0, const double *__restrict__ fac2, const double *__restrict__ fac3,
             const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, const double *__restrict__ c4,
             double *__restrict__ q0, double *__restrict__ q2, double *__restrict__ q3, int kmax)
{
    extern __shared__ __align__(16) double sq[];
    const int g = threadIdx.x & 15;
    double *sg = sq + (size_t)(threadIdx.x >> 4) * (3 * 480);
    int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = pidx < P;
    if (!valid)
        pidx = P - 1;
    const size_t roff = (size_t)SVD16_REC * pidx;
    const double *R[3] = {fac0 + roff, fac2 + roff, fac3 + roff};
    // only the first kmax singular triplets (descending order) are needed: kmax columns of U (pieces 0..kmax/2*...) and of V
#pragma unroll
    for (int o = 0; o < 3; o++)
#pragma unroll
        for (int piece = 0; piece < 15; piece++) // 480 doubles = 240 16-byte pieces per object, 15 per lane
        {
            const int d0 = 2 * (piece * 16 + g);          // first double of this 16-byte piece within the record
            const int col = (d0 < 240) ? (d0 >> 4) : ((d0 - 240) >> 4); // column of U (first 240 doubles) or of V
            if (col < kmax)
                cp_async16(sg + o * 480 + d0, R[o] + d0);
        }
    cp_async_commit();
    const int id = ids[pidx];
    const int r = g & 3, c = g >> 2;
    const int fsz = N * N;
    double cw[SVD16_N];
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        cw[k] = c4[(p.x + r) + N * (p.y + c) + fsz * k];
    }
    cp_async_wait<0>();
    __syncwarp();
    double *qd[3] = {q0, q2, q3};
#pragma unroll
    for (int o = 0; o < 3; o++)
    {
        const double *Us = sg + o * 480, *Vs = Us + SVD16_M * SVD16_N;
        double mine = 0.0;
#pragma unroll 1
        for (int kk = 0; kk < kmax; kk++)
        {
            const double2 *v = reinterpret_cast<const double2 *>(Vs + SVD16_LDV * kk);
            double z = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++)
            {
                const double2 x = v[k2];
                z = fma(cw[2 * k2], x.x, z);
                if (2 * k2 + 1 < SVD16_N)
                    z = fma(cw[2 * k2 + 1], x.y, z);
            }
            double val = Us[SVD16_M * kk + g] * z;
#pragma unroll
            for (int sh = 8; sh > 0; sh >>= 1)
                val += __shfl_xor_sync(0xffffffffu, val, sh);
            if (g == kk)
                mine = val;
        }
        if (valid)
            qd[o][(size_t)16 * pidx + g] = mine; // slots >= kmax are written as 0 and flagged by k_eval3 if ever needed
    }
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}

#define EV_STAGE 144 /* doubles per prefetch stage: S,q of 3 objects (96) | leading triplet U,V (32) | 15 positions (16 ints) | pad */
#define EV_GRP2 (2 * EV_STAGE + 2 * EV_CH)

// K_qform for the leading KQ triplets only (the lazy default): same arithmetic as k_qform3 with a compact staging buffer
// (3 x KQ x 32 doubles per patch instead of the whole U and V of three objects), which lifts the residency from 2 to 8 CTAs
// per SM — the kernel is bound by the latency of the C4 gather.
template <int KQ>
__global__ void __launch_bounds__(128, 8)
    k_qform3_lead(const double *__restrict__ fac0, const double *__restrict__ fac2, const double *__restrict__ fac3,
                  const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, const double *__restrict__ c4,
                  double *__restrict__ q0, double *__restrict__ q2, double *__restrict__ q3)
{
    __shared__ __align__(16) double sq[8 * 3 * KQ * 32];
    const int g = threadIdx.x & 15;
    double *sg = sq + (size_t)(threadIdx.x >> 4) * (3 * KQ * 32);
    int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = pidx < P;
    if (!valid)
        pidx = P - 1;
    const size_t roff = (size_t)SVD16_REC * pidx;
    const double *R[3] = {fac0 + roff, fac2 + roff, fac3 + roff};
#pragma unroll
    for (int o = 0; o < 3; o++)
#pragma unroll
        for (int kk = 0; kk < KQ; kk++)
        { // lanes 0..7: U column kk, lanes 8..15: V column kk
            const double *src = (g < 8) ? R[o] + SVD16_M * kk + 2 * g : R[o] + SVD16_M * SVD16_N + SVD16_LDV * kk + 2 * (g - 8);
            cp_async16(sg + (o * KQ + kk) * 32 + 2 * g, src);
        }
    cp_async_commit();
    const int id = ids ? ids[pidx] : pidx;
    const int r = g & 3, c = g >> 2;
    const int fsz = N * N;
    double cw[SVD16_N];
#pragma unroll
    for (int k = 0; k < SVD16_N; k++)
    {
        const short2 p = pos[(size_t)k * vecSize + id];
        cw[k] = c4[(p.x + r) + N * (p.y + c) + fsz * k];
    }
    cp_async_wait<0>();
    __syncwarp();
    double *qd[3] = {q0, q2, q3};
#pragma unroll
    for (int o = 0; o < 3; o++)
    {
        double mine = 0.0;
#pragma unroll
        for (int kk = 0; kk < KQ; kk++)
        {
            const double *b = sg + (o * KQ + kk) * 32;
            const double2 *v = reinterpret_cast<const double2 *>(b + 16);
            double z = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++)
            {
                const double2 x = v[k2];
                z = fma(cw[2 * k2], x.x, z);
                if (2 * k2 + 1 < SVD16_N)
                    z = fma(cw[2 * k2 + 1], x.y, z);
            }
            double val = b[g] * z;
#pragma unroll
            for (int sh = 8; sh > 0; sh >>= 1)
                val += __shfl_xor_sync(0xffffffffu, val, sh);
            if (g == kk)
                mine = val;
        }
        if (valid)
            qd[o][(size_t)16 * pidx + g] = mine; // slots >= KQ are written as 0 and flagged by k_eval3 if ever needed
    }
}

// LEAN: S and q of the three objects come from the 128-byte head record the SVD kernels leave behind (three singular values
// and two q-forms per object); a surviving THIRD singular value of any object flags need_more_q and the caller repeats
// the evaluation through the general path after decomposing the perturbed objects in full.
template <int MINB, int PPG, int LEAN>
__global__ void __launch_bounds__(128, MINB)
    k_eval3(const double *__restrict__ fac0, const double *__restrict__ fac2, const double *__restrict__ fac3,
            const double *__restrict__ q0, const double *__restrict__ q2, const double *__restrict__ q3,
            const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, double lambda, int expw,
            double *__restrict__ acc0, const double *__restrict__ accs, double *__restrict__ partial, int *__restrict__ kpart,
            int qmax, int *__restrict__ need_more_q, int tiled)
{
    // One PGURE evaluation for 16 x 15 patches and the three SVT objects U, U +- eps2*delta2.  16 lanes per patch; lane g
    // owns block row g (pixel (g&3, g>>2) of the patch) for all 15 slices and thresholds slot g of every object.
    //  * s4 (second difference) needs no block at all: sum_k (f2p_k q2p_k + f2m_k q2m_k - 2 f0_k q0_k), see k_qform.
    //  * Uhat enters the risk non-linearly, so object 0's block is rebuilt (rank-adaptively: singular values are sorted
    //    and the soft threshold is monotone, so the survivors are a prefix) and overlap-added with fire-and-forget
    //    FP64 REDs — per slice the 16 lanes cover the patch's 4 x 4 footprint.
    // A CTA covers 8 * PPG consecutive patches; each 16-lane group walks PPG of them and prefetches the next patch's
    // S, q, leading triplet and trajectory with cp.async (two stages) while the REDs of the current one are issued.
    // Triplets beyond the leading one (rare) are fetched on demand in chunks of EV_C, double-buffered.
    __shared__ __align__(16) double smem[8 * EV_GRP2];
    const int lane = threadIdx.x & 31;
    const int g = threadIdx.x & 15;
    const int grp = threadIdx.x >> 4;
    double *sg = smem + grp * EV_GRP2;
    const int soff = SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
    const int r = g & 3, c = g >> 2;
    const int fsz = N * N;
    const int base = blockIdx.x * (8 * PPG) + grp;
    const double ascale = __ldg(accs);

    auto issue = [&](int j, int st) {
        int pidx = base + 8 * j;
        if (pidx >= P)
            pidx = P - 1;
        const size_t roff = (size_t)SVD16_REC * pidx;
        double *dst = sg + st * EV_STAGE;
        if (LEAN == 2)
        { // accumulate-only pass over ONE object (eps1_mode 1: object U1): its S, U and V column 0
            if (g < 8)
            {
                cp_async16(dst + 0 + 2 * g, fac0 + roff + soff + 2 * g);
                cp_async16(dst + 96 + 2 * g, fac0 + roff + 2 * g);
            }
            else
                cp_async16(dst + 112 + 2 * (g - 8), fac0 + roff + SVD16_M * SVD16_N + 2 * (g - 8));
        }
        else if (LEAN)
        { // head record (q0 carries it in this mode), U and V column 0 of object U
            if (g < 8)
            {
                cp_async16(dst + 0 + 2 * g, q0 + (size_t)16 * pidx + 2 * g);
                cp_async16(dst + 96 + 2 * g, fac0 + roff + 2 * g);
            }
            else
                cp_async16(dst + 112 + 2 * (g - 8), fac0 + roff + SVD16_M * SVD16_N + 2 * (g - 8));
        }
        else if (g < 8)
        { // S of the three objects, U column 0
            cp_async16(dst + 0 + 2 * g, fac0 + roff + soff + 2 * g);
            cp_async16(dst + 16 + 2 * g, fac2 + roff + soff + 2 * g);
            cp_async16(dst + 32 + 2 * g, fac3 + roff + soff + 2 * g);
            cp_async16(dst + 96 + 2 * g, fac0 + roff + 2 * g);
        }
        else
        { // q of the three objects, V column 0
            const int h2 = 2 * (g - 8);
            cp_async16(dst + 48 + h2, q0 + (size_t)16 * pidx + h2);
            cp_async16(dst + 64 + h2, q2 + (size_t)16 * pidx + h2);
            cp_async16(dst + 80 + h2, q3 + (size_t)16 * pidx + h2);
            cp_async16(dst + 112 + h2, fac0 + roff + SVD16_M * SVD16_N + h2);
        }
        if (g < SVD16_N)
        { // trajectory position of slice g
            const int id = ids ? ids[pidx] : pidx; // ids == nullptr: the patch set is the full macroblock grid (patch_overlap 1)
            cp_async4(reinterpret_cast<int *>(dst + 128) + g, pos + (size_t)g * vecSize + id);
        }
        cp_async_commit();
    };

    double s4tot = 0.0;
    int ktot = 0;
    issue(0, 0);
#pragma unroll 1
    for (int j = 0; j < PPG; j++)
    {
        const int st = j & 1;
        if (j + 1 < PPG)
            issue(j + 1, st ^ 1);
        else
            cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        int pidx = base + 8 * j;
        const bool valid = pidx < P;
        if (!valid)
            pidx = P - 1;
        const double *R0 = fac0 + (size_t)SVD16_REC * pidx;
        const double *ss = sg + st * EV_STAGE;
        const short2 *sp = reinterpret_cast<const short2 *>(ss + 128);
        double f0 = 0.0, s4 = 0.0;
        if (LEAN == 2)
        {
            if (g < SVD16_N)
                f0 = soft_f(ss[g], ss[15], lambda, expw);
        }
        else if (LEAN)
        {
            if (g < 3)
            {
                f0 = soft_f(ss[g], ss[0], lambda, expw);
                const double f2 = soft_f(ss[3 + g], ss[3], lambda, expw);
                const double f3 = soft_f(ss[6 + g], ss[6], lambda, expw);
                if (g < 2)
                    s4 = fma(f2, ss[11 + g], fma(f3, ss[13 + g], -2.0 * f0 * ss[9 + g]));
                else if ((f0 != 0.0 || f2 != 0.0 || f3 != 0.0) && valid)
                    *need_more_q = 1;
            }
        }
        else if (g < SVD16_N)
        {
            f0 = soft_f(ss[g], ss[15], lambda, expw);
            const double f2 = soft_f(ss[16 + g], ss[16 + 15], lambda, expw);
            const double f3 = soft_f(ss[32 + g], ss[32 + 15], lambda, expw);
            s4 = fma(f2, ss[64 + g], fma(f3, ss[80 + g], -2.0 * f0 * ss[48 + g]));
            // q-forms are prepared lazily for the leading qmax triplets only; a survivor beyond them invalidates this pass
            if (g >= qmax && (f0 != 0.0 || f2 != 0.0 || f3 != 0.0) && valid)
                *need_more_q = 1;
        }
        unsigned m = __ballot_sync(0xffffffffu, f0 != 0.0);
        m = (m | (m >> 16)) & 0xffffu;
        const int Kw = 32 - __clz(m); // one past the largest surviving index over both patches of the warp (0 if none)
        double a0[SVD16_N];
#pragma unroll
        for (int k = 0; k < SVD16_N; k++)
            a0[k] = 0.0;
        if (Kw > 0)
        { // leading triplet from the prefetch stage
            const double fk0 = __shfl_sync(0xffffffffu, f0, lane & 16);
            const double u0 = ss[96 + g] * fk0;
            const double2 *v0 = reinterpret_cast<const double2 *>(ss + 112);
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++)
            {
                const double2 x0 = v0[k2];
                a0[2 * k2] = fma(u0, x0.x, a0[2 * k2]);
                if (2 * k2 + 1 < SVD16_N)
                    a0[2 * k2 + 1] = fma(u0, x0.y, a0[2 * k2 + 1]);
            }
        }
        if (Kw > 1)
        { // further survivors: chunks of EV_C triplets, double-buffered (waits also drain the prefetch of the next patch)
            double *cbuf = sg + 2 * EV_STAGE;
            auto request_chunk = [&](int c0, int buf) {
                double *dst = cbuf + buf * EV_CH;
#pragma unroll
                for (int cc = 0; cc < EV_C; cc++)
                {
                    const int kk = c0 + cc;
                    if (kk < SVD16_N)
                    {
                        // lanes 0..7: U column kk (16 doubles), lanes 8..15: V column kk
                        const double *src = (g < 8) ? R0 + SVD16_M * kk + 2 * g : R0 + SVD16_M * SVD16_N + SVD16_LDV * kk + 2 * (g - 8);
                        cp_async16(dst + cc * 32 + 2 * g, src);
                    }
                }
            };
            int buf = 0;
            request_chunk(1, 0);
            cp_async_commit();
            for (int c0 = 1; c0 < Kw; c0 += EV_C)
            {
                if (c0 + EV_C < Kw)
                    request_chunk(c0 + EV_C, buf ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const double *cb = cbuf + buf * EV_CH;
#pragma unroll
                for (int cc = 0; cc < EV_C; cc++)
                {
                    const int kk = c0 + cc;
                    if (kk < Kw)
                    {
                        const double fk0 = __shfl_sync(0xffffffffu, f0, (lane & 16) | kk);
                        const double *b0 = cb + cc * 32;
                        const double u0 = b0[g] * fk0;
                        const double2 *v0 = reinterpret_cast<const double2 *>(b0 + 16);
#pragma unroll
                        for (int k2 = 0; k2 < 8; k2++)
                        {
                            const double2 x0 = v0[k2];
                            a0[2 * k2] = fma(u0, x0.x, a0[2 * k2]);
                            if (2 * k2 + 1 < SVD16_N)
                                a0[2 * k2 + 1] = fma(u0, x0.y, a0[2 * k2 + 1]);
                        }
                    }
                }
                __syncwarp();
                buf ^= 1;
            }
            cp_async_wait<0>();
        }
        if (valid)
        {
#pragma unroll
            for (int k = 0; k < SVD16_N; k++)
            {
                const short2 p = sp[k];
                // tiled accumulator (N even): every 2 x 2 pixel block is one 32-byte sector, so a 4 x 4 footprint at a random
                // offset touches 6.25 sectors on average instead of 7 (four columns x 1.75) — the kernel is bound by RED sectors
                const int row = p.x + r, col = p.y + c;
                const int idx = tiled ? 4 * ((row >> 1) + (N >> 1) * (col >> 1)) + (row & 1) + 2 * (col & 1) : row + N * col;
                acc_add(acc0, (size_t)(idx + fsz * k), a0[k], ascale);
            }
            s4tot += s4;
        }
        ktot += 2 * Kw; // triplets of object 0 fetched for the warp's two patches (upper bound per patch)
        __syncwarp(); // the stage is overwritten by the prefetch issued in the next iteration
    }
    cp_async_wait<0>();
    // per-warp partials (no CTA barrier): partial[4 * blockIdx.x + warp] = s4 part, kpart[...] = triplets fetched
    s4tot = warp_sum(s4tot);
    if (lane == 0 && LEAN != 2)
    {
        const int w = 4 * blockIdx.x + (threadIdx.x >> 5);
        partial[w] = s4tot;
        kpart[w] = ktot;
    }
}

// Final reconstruction of ONE slice (the only one the driver consumes, pguresvt.hpp:155-166) for 16 x 15 patches in the
// l4 record format: 16 lanes per patch, lane g owns block row g; the thresholded spectrum is sorted, so only the
// surviving leading triplets are read (S: 128 B, then 136 B per survivor instead of the whole 3,968-byte record).
__global__ void __launch_bounds__(128)
    k_final16(const double *__restrict__ fac0, const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N,
              double lambda, int expw, int kref, double *__restrict__ acc, const double *__restrict__ accs)
{
    const int lane = threadIdx.x & 31, g = threadIdx.x & 15;
    int pidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = pidx < P;
    if (!valid)
        pidx = P - 1;
    const double *R = fac0 + (size_t)SVD16_REC * pidx;
    const double *S = R + SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
    const double f = (g < SVD16_N) ? soft_f(S[g], S[15], lambda, expw) : 0.0;
    unsigned m = __ballot_sync(0xffffffffu, f != 0.0);
    m = (m | (m >> 16)) & 0xffffu;
    const int Kw = 32 - __clz(m); // one past the largest surviving index over both patches of the warp
    double a = 0.0;
    for (int kk = 0; kk < Kw; kk++)
    {
        const double fk = __shfl_sync(0xffffffffu, f, (lane & 16) | kk);
        if (fk != 0.0)
            a = fma(R[SVD16_M * kk + g] * fk, R[SVD16_M * SVD16_N + SVD16_LDV * kk + kref], a);
    }
    if (valid)
    {
        const int id = ids ? ids[pidx] : pidx;
        const short2 p = pos[(size_t)kref * vecSize + id];
        acc_add(acc, (size_t)(p.x + (g & 3)) + (size_t)N * (p.y + (g >> 2)) + (size_t)N * N * kref, a, __ldg(accs));
    }
}

// voxel pass of the fused evaluation: Uhat = acc0 / weights (non-finite -> 0, svt.hpp:163-164);
// s1 = sum (Uhat - U)^2, s5 = sum Uhat; the accumulator is cleared on the way for the next evaluation, and the
// per-CTA s4 partials of k_eval3 are folded in (third sum) so that one fixed-order reduction finishes all three.
// partial: gridDim.x * 4 doubles (s1, s5, s4, triplets streamed)
template <int EPS1>
__global__ void __launch_bounds__(256, 8) k_risk_uhat(const double *__restrict__ u, const unsigned *__restrict__ cnt, double *__restrict__ acc0, size_t tot,
                            const double *__restrict__ accs, const double *__restrict__ s4part, const int *__restrict__ kpart, int ns4,
                            double *__restrict__ partial, int tiledN = 0, double *__restrict__ acc1 = nullptr,
                            const int8_t *__restrict__ d1 = nullptr, double e_alpha = 0.0, double e_const = 0.0,
                            double *__restrict__ partial3 = nullptr)
{
    // acc1 != nullptr (eps1_mode 1): third sum of pgure.hpp:136, s3 = sum delta1 (alpha U - alpha mu + sigma^2)(U1 - Uhat), with
    // U1's accumulator filled by the accumulate-only pass of k_eval3; e_alpha = alpha, e_const = sigma^2 - alpha mu
    double s1 = 0, s5 = 0, s4 = 0, sk = 0, s3 = 0;
    const double ainv = __ldg(accs + 1);
    const unsigned N_ = (unsigned)tiledN, fsz_ = N_ * N_, hN = N_ >> 1;
    // (an unrolled variant with more loads in flight per thread needs 58 registers, halves the resident CTAs and is slower)
    auto voxel = [&](size_t ia, size_t i) { // ia: index into the accumulator(s), i: voxel in image order
        const double v0 = norm_or_zero(acc_val(acc0, ia, ainv), cnt[i]);
        acc0[ia] = 0.0;
        const double d = v0 - u[i];
        s1 = fma(d, d, s1);
        s5 += v0;
        if (EPS1)
        {
            const double v1 = norm_or_zero(acc_val(acc1, ia, ainv), cnt[i]);
            acc1[ia] = 0.0;
            s3 += ((double)d1[i] * (e_alpha * u[i] + e_const)) * (v1 - v0);
        }
    };
    if (!tiledN)
    { // image-order accumulator
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
            voxel(i, i);
    }
    else
    { // 2 x 2-tiled accumulator (k_eval3): the tiles of one column pair of one slice are 2N contiguous doubles, so the loop walks
      // such strips — one division per strip instead of two per voxel (the per-voxel index arithmetic cost 0.02 ms of a 0.09-ms
      // kernel); weights and u stay in image order
        const unsigned nstrips = (unsigned)(tot / fsz_) * hN, per = 2u * N_;
        for (unsigned strip = blockIdx.x; strip < nstrips; strip += gridDim.x)
        {
            const unsigned k = strip / hN, tc = strip - k * hN;
            const size_t abase = (size_t)k * fsz_ + 4u * (size_t)hN * tc, ibase = (size_t)k * fsz_ + (size_t)N_ * (2u * tc);
            for (unsigned e = threadIdx.x; e < per; e += blockDim.x)
                voxel(abase + e, ibase + (2u * (e >> 2) + (e & 1u)) + (size_t)N_ * ((e >> 1) & 1u));
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns4; i += gridDim.x * blockDim.x)
    {
        s4 += s4part[i];
        sk += (double)kpart[i];
    }
    __shared__ double sm[5][32];
    s1 = warp_sum(s1);
    s5 = warp_sum(s5);
    s4 = warp_sum(s4);
    sk = warp_sum(sk);
    s3 = warp_sum(s3);
    if ((threadIdx.x & 31) == 0)
    {
        sm[0][threadIdx.x >> 5] = s1;
        sm[1][threadIdx.x >> 5] = s5;
        sm[2][threadIdx.x >> 5] = s4;
        sm[3][threadIdx.x >> 5] = sk;
        sm[4][threadIdx.x >> 5] = s3;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
#pragma unroll
        for (int q = 0; q < 5; q++)
        {
            double r = (threadIdx.x < (blockDim.x >> 5)) ? sm[q][threadIdx.x] : 0.0;
            r = warp_sum(r);
            if (threadIdx.x == 0)
            {
                if (q < 4)
                    partial[(size_t)blockIdx.x * 4 + q] = r;
                else if (partial3)
                    partial3[blockIdx.x] = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K_recon — SVT::Reconstruct (svt.hpp:121-160) for one SVT object: threshold the cached singular values
// (plain lambda, or the exponential weighting of svt.hpp:135-143 with SoftThreshold of utils.hpp:96-106),
// rebuild block = U diag(Sthr) V^T (svt.hpp:146) and overlap-add it along the patch trajectory into `acc`
// (svt.hpp:148-155).  A group of G lanes (G = 16 for bs = 4, else 32) owns one patch: V and the thresholded
// spectrum are staged in shared memory, each lane rebuilds whole rows of the block in registers.  Columns
// whose thresholded singular value is zero are skipped (rank-adaptive: they contribute nothing).
// only_k >= 0 restricts the overlap-add to one slice (the only slice the driver consumes for the final
// reconstruction, pguresvt.hpp:155-166).
// ------------------------------------------------------------------------------------------------------
template <int NMAX>
__global__ void k_recon(const double *__restrict__ fac, size_t rec, int m, int n, int ldv, int bs,
                        const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N,
                        double lambda, int expw, int only_k, double *__restrict__ acc, const double *__restrict__ accs, int G)
{
    extern __shared__ double smr[];
    const double ascale = __ldg(accs);
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G, g = threadIdx.x % G;
    const int pidx = blockIdx.x * gpb + gl;
    const bool valid = pidx < P;
    const size_t per = (size_t)ldv * n + NMAX + NMAX; // V | f | pos (short2 packed in doubles' space)
    double *sV = smr + per * gl;
    double *sf = sV + (size_t)ldv * n;
    short2 *sp = reinterpret_cast<short2 *>(sf + NMAX);
    const double *R = fac + rec * (size_t)(valid ? pidx : 0);
    if (valid)
    {
        const double *S = R + (size_t)m * n + (size_t)ldv * n;
        const int id = ids[pidx];
        double smax = S[0];
        for (int k = 1; k < n; k++)
            smax = fmax(smax, S[k]);
        for (int k = g; k < n; k += G)
        {
            const double s = S[k];
            double f;
            if (expw)
            {
                const double w = fabs(smax * exp(-0.5 * lambda * (s * s)));
                f = fmax(fabs(s) - w, 0.0);
            }
            else
                f = fmax(fabs(s) - lambda, 0.0);
            sf[k] = (s < 0.0) ? -f : f;
            sp[k] = pos[(size_t)k * vecSize + id];
        }
        const double *Vg = R + (size_t)m * n;
        for (int e = g; e < ldv * n; e += G)
            sV[e] = Vg[e];
    }
    __syncthreads();
    if (!valid)
        return;
    const size_t fsz = (size_t)N * N;
    for (int e = g; e < m; e += G)
    {
        double a[NMAX];
#pragma unroll
        for (int k = 0; k < NMAX; k++)
            a[k] = 0.0;
        for (int kk = 0; kk < n; kk++)
        {
            const double fk = sf[kk];
            if (fk != 0.0)
            {
                const double uf = R[e + (size_t)m * kk] * fk;
                const double *vc = sV + (size_t)ldv * kk;
#pragma unroll
                for (int k = 0; k < NMAX; k++)
                    if (k < n)
                        a[k] = fma(uf, vc[k], a[k]);
            }
        }
        const int r = e % bs, c = e / bs;
#pragma unroll
        for (int k = 0; k < NMAX; k++)
            if (k < n && (only_k < 0 || k == only_k))
            {
                const short2 p = sp[k];
                acc_add(acc, (size_t)(p.x + r) + (size_t)N * (p.y + c) + fsz * k, a[k], ascale);
            }
    }
}

// ------------------------------------------------------------------------------------------------------
// K_risk — `v /= weights; non-finite -> 0` (svt.hpp:163-164) for the four reconstructions fused with the five
// global sums of PGURE::CalculatePGURE (pgure.hpp:136).  Fixed grid, fixed-order block reduction → the sums
// are deterministic given the accumulators.  partial: gridDim.x * 5 doubles.
// ------------------------------------------------------------------------------------------------------

__global__ void k_risk(const double *__restrict__ u, const int8_t *__restrict__ d1, const int8_t *__restrict__ d2neg,
                       const unsigned *__restrict__ cnt, const double *__restrict__ acc0, const double *__restrict__ acc1,
                       const double *__restrict__ acc2p, const double *__restrict__ acc2m, const double *__restrict__ accs, size_t tot,
                       double alpha, double mu, double sigmasq, double dNeg, double dPos, double *__restrict__ partial)
{
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    const double ainv = __ldg(accs + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
    {
        const unsigned c = cnt[i];
        const double U = u[i];
        const double v0 = norm_or_zero(acc_val(acc0, i, ainv), c);
        const double v2p = norm_or_zero(acc_val(acc2p, i, ainv), c);
        const double v2m = norm_or_zero(acc_val(acc2m, i, ainv), c);
        const double d = fabs(v0 - U);
        s1 += d * d;
        s2 += U;
        if (acc1)
        {
            const double v1 = norm_or_zero(acc_val(acc1, i, ainv), c);
            s3 += ((double)d1[i] * (alpha * U - alpha * mu + sigmasq)) * (v1 - v0);
        }
        s4 += (d2neg[i] ? dNeg : dPos) * (v2p - 2 * v0 + v2m);
        s5 += v0;
    }
    __shared__ double sm[5][32];
    double vals[5] = {s1, s2, s3, s4, s5};
#pragma unroll
    for (int q = 0; q < 5; q++)
    {
        const double r = warp_sum(vals[q]);
        if ((threadIdx.x & 31) == 0)
            sm[q][threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
#pragma unroll
        for (int q = 0; q < 5; q++)
        {
            double r = (threadIdx.x < (blockDim.x >> 5)) ? sm[q][threadIdx.x] : 0.0;
            r = warp_sum(r);
            if (threadIdx.x == 0)
                partial[(size_t)blockIdx.x * 5 + q] = r;
        }
    }
}

// final fixed-order reduction of the per-block partial sums: one block, nq quantities
__global__ void k_reduce_partials(const double *__restrict__ partial, int nblocks, int nq, double *__restrict__ out)
{
    __shared__ double sm[32];
    for (int q = 0; q < nq; q++)
    {
        double r = 0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
            r += partial[(size_t)b * nq + q];
        r = warp_sum(r);
        if ((threadIdx.x & 31) == 0)
            sm[threadIdx.x >> 5] = r;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
            r = warp_sum(r);
            if (threadIdx.x == 0)
                out[q] = r;
        }
        __syncthreads();
    }
}

// sum of a double array (accu(u), pguresvt.hpp:139) → partial[gridDim.x]
__global__ void k_sum(const double *__restrict__ x, size_t n, double *__restrict__ partial)
{
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        s += x[i];
    s = warp_sum(s);
    __shared__ double sm[32];
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0)
            partial[blockIdx.x] = s;
    }
}

// v = acc / weights, non-finite -> 0, times scale (svt.hpp:163-164, pguresvt.hpp:147); n voxels starting at
// offset `off` of acc/cnt, written to out[0..n)
__global__ void k_finalize(const double *__restrict__ acc, const double *__restrict__ accs, const unsigned *__restrict__ cnt, size_t off,
                           size_t n, double scale, double *__restrict__ out)
{
    const double ainv = __ldg(accs + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = norm_or_zero(acc_val(acc, off + i, ainv), cnt[off + i]) * scale;
}

} // namespace pgs
