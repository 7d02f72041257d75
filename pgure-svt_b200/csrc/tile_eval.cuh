// Atomics-free, order-deterministic evaluation of the PGURE objective for 16 x 15 patches (the production path of
// BASELINE configs 3 and 4): SVT::Reconstruct's overlap-add (svt.hpp:148-160) turned from a scatter into a GATHER BY
// DESTINATION, fused with `v /= weights` (svt.hpp:163-164) and the risk sums of PGURE::CalculatePGURE (pgure.hpp:130-136).
//
// Trajectories are fixed while a frame's ~24 evaluations run, so the scatter is inverted once per frame: for every slice k
// a CSR "patches whose block lands with its origin on pixel b" (k_bin_count / scan / k_bin_fill / k_bin_sort; entries of
// a bin sorted by patch index, so every voxel's sum has ONE order).  An evaluation is then
//   k_thresh     per patch: thresholded leading singular values f0, f1 of object U, the patch's term of the
//                second-difference sum from the q-forms (see k_qform3), "a third triplet survives" flag;
//   k_tile_eval  one CTA per 61 x 29 output tile of one slice: the tile accumulates in shared memory — 16 colour steps
//                (bin origin mod 4 in both directions: blocks of one colour never overlap, so plain read-modify-write
//                needs no atomics), 8 lanes per bin, each lane two entries of the 4 x 4 block; then Uhat = tile / weights,
//                sum (Uhat - U)^2 and sum Uhat straight from shared memory.  No accumulator cube in HBM, no clearing pass,
//                no REDs; the per-CTA partial sums are reduced in a fixed order.
// The same kernel in MODE 1 writes the denoised output slice (pguresvt.hpp:147,155-166).
#pragma once
#include "kernels.cuh"

namespace pgs
{

#define TG_RR 64            /* bins (block origins) per CTA region: rows */
#define TG_RC 32            /*                                     cols */
#define TG_VR (TG_RR - 3)   /* output rows a region completes: every covering origin lies inside it */
#define TG_VC (TG_RC - 3)
#define TG_TR (TG_RR + 3)   /* tile rows touched */
#define TG_TC (TG_RC + 3)
#define TG_RG 17            /* row groups of 4: ceil(TG_TR / 4) */
#define TG_LDC (4 * TG_RG)  /* doubles per tile column */

struct __align__(16) TgEntry
{
    int pidx;  // patch index (record number)
    int pad;
    double v;  // leading right singular vector of object U at this slice: v_0[k]
};

// rows de-interleaved mod 4: the 16 lanes of a half-warp (2 bins x 8 lanes) hit 16 different bank pairs
__device__ __forceinline__ int tg_idx(int row, int col) { return (row >> 2) + TG_RG * (row & 3) + TG_LDC * col; }

// ---- per-frame CSR build -------------------------------------------------------------------------------------------
// binc[k * N*N + row + N*col] = number of patches whose slice-k block origin is (row, col)
__global__ void k_bin_count(const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, int win,
                            int *__restrict__ binc)
{
    const long long tot = (long long)P * win;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x)
    {
        const int k = (int)(i / P), pidx = (int)(i - (long long)k * P);
        const short2 p = pos[(size_t)k * vecSize + (ids ? ids[pidx] : pidx)];
        atomicAdd(binc + (size_t)k * N * N + p.x + (size_t)N * p.y, 1);
    }
}

// exclusive scan of n ints in three passes (block sums, scan of the sums, apply): SCAN_ITEMS per block
#define SCAN_THREADS 256
#define SCAN_PER_THREAD 16
#define SCAN_ITEMS (SCAN_THREADS * SCAN_PER_THREAD)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const int *__restrict__ in, size_t n, int *__restrict__ sums)
{
    const size_t base = (size_t)blockIdx.x * SCAN_ITEMS;
    int s = 0;
    for (int q = 0; q < SCAN_PER_THREAD; q++)
    {
        const size_t i = base + (size_t)q * SCAN_THREADS + threadIdx.x;
        s += (i < n) ? in[i] : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ int sm[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 0)
        sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++)
            t += sm[w];
        sums[blockIdx.x] = t;
    }
}
// one block: exclusive scan of the block sums in place (nb up to a few thousand), total -> sums[nb]
__global__ void __launch_bounds__(1024) k_scan_block_sums(int *__restrict__ sums, int nb)
{
    __shared__ int sm[1024];
    __shared__ int carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024)
    {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? sums[i] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1)
        {
            const int t = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb)
            sums[i] = carry + sm[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        sums[nb] = carry;
}
// start[i] = exclusive prefix of in[i]; start[n] = total.  Each thread owns SCAN_PER_THREAD CONSECUTIVE items.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int *__restrict__ in, size_t n, const int *__restrict__ sums, int nb,
                                                             int *__restrict__ start)
{
    const size_t base = (size_t)blockIdx.x * SCAN_ITEMS + (size_t)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_PER_THREAD; q++)
    {
        v[q] = (base + q < n) ? in[base + q] : 0;
        s += v[q];
    }
    // exclusive scan of the per-thread totals across the block
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    __shared__ int wsum[SCAN_THREADS / 32];
    if (lane == 31)
        wsum[w] = inc;
    __syncthreads();
    int off = sums[blockIdx.x];
    for (int q = 0; q < w; q++)
        off += wsum[q];
    off += inc - s;
#pragma unroll
    for (int q = 0; q < SCAN_PER_THREAD; q++)
    {
        if (base + q < n)
            start[base + q] = off;
        off += v[q];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        start[n] = sums[nb];
}

// entries: slot = start[bin] + (position within the bin, handed out by atomics — put in order by k_bin_sort)
__global__ void k_bin_fill(const short2 *__restrict__ pos, const int *__restrict__ ids, int P, int vecSize, int N, int win,
                           const int *__restrict__ start, int *__restrict__ cursor, const double *__restrict__ fac0,
                           TgEntry *__restrict__ ent)
{
    const long long tot = (long long)P * win;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x)
    {
        const int k = (int)(i / P), pidx = (int)(i - (long long)k * P);
        const short2 p = pos[(size_t)k * vecSize + (ids ? ids[pidx] : pidx)];
        const size_t b = (size_t)k * N * N + p.x + (size_t)N * p.y;
        const int slot = start[b] + atomicAdd(cursor + b, 1);
        TgEntry e;
        e.pidx = pidx;
        e.pad = 0;
        e.v = fac0[(size_t)SVD16_REC * pidx + SVD16_M * SVD16_N + k]; // V column 0 (leading triplet), row k
        ent[slot] = e;
    }
}
// bins with more than one entry: insertion sort by patch index (fixes the order of every voxel's sum)
__global__ void k_bin_sort(const int *__restrict__ start, size_t nbins, TgEntry *__restrict__ ent)
{
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbins; b += (size_t)gridDim.x * blockDim.x)
    {
        const int e0 = start[b], e1 = start[b + 1];
        for (int i = e0 + 1; i < e1; i++)
        {
            const TgEntry x = ent[i];
            int j = i - 1;
            while (j >= e0 && ent[j].pidx > x.pidx)
            {
                ent[j + 1] = ent[j];
                j--;
            }
            ent[j + 1] = x;
        }
    }
}

// ---- per-frame head records ----------------------------------------------------------------------------------------
// head[p][16] = S0[0..2] | S2[0..2] | S3[0..2] | q0[0..1] | q2[0..1] | q3[0..1] | 0   (one 128-byte line per patch)
#define TG_HEAD 16
__global__ void k_head_pack(const double *__restrict__ fac0, const double *__restrict__ fac2, const double *__restrict__ fac3,
                            const double *__restrict__ q0, const double *__restrict__ q2, const double *__restrict__ q3, int P,
                            double *__restrict__ head, double *__restrict__ u0c)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int pidx = t >> 2, part = t & 3; // 4 lanes per patch: objects 0, 2, 3 and the tail
    if (pidx >= P)
        return;
    const size_t soff = (size_t)SVD16_REC * pidx + SVD16_M * SVD16_N + SVD16_LDV * SVD16_N;
    double *hd = head + (size_t)TG_HEAD * pidx;
    if (!head)
    { // lean mode: the SVD kernels wrote the head entries already
    }
    else if (part < 3)
    {
        const double *S = (part == 0 ? fac0 : part == 1 ? fac2 : fac3) + soff;
        const double *q = (part == 0 ? q0 : part == 1 ? q2 : q3) + (size_t)16 * pidx;
        hd[3 * part] = S[0], hd[3 * part + 1] = S[1], hd[3 * part + 2] = S[2];
        hd[9 + 2 * part] = q[0], hd[9 + 2 * part + 1] = q[1];
    }
    else
        hd[15] = 0.0;
    // leading left singular vector of object U, packed: the gather kernel reads 128 contiguous bytes per patch from a 133 MB
    // array instead of one line out of every 3,968-byte record (TLB reach, DRAM page locality, L2 footprint)
    const double2 *src = reinterpret_cast<const double2 *>(fac0 + (size_t)SVD16_REC * pidx) + 2 * part;
    double2 *dst = reinterpret_cast<double2 *>(u0c + (size_t)16 * pidx) + 2 * part;
    dst[0] = src[0];
    dst[1] = src[1];
}

// ---- per evaluation ------------------------------------------------------------------------------------------------
// fth[p] = (f0, f1) of object U; s4 partial per warp; *need_more = 1 if the third singular value of any object survives
// (the q-forms and this path cover two triplets: the caller then answers the probe through the general path)
__global__ void __launch_bounds__(256) k_thresh(const double *__restrict__ head, int P, double lambda, int expw, double2 *__restrict__ fth,
                                                double *__restrict__ s4part, int *__restrict__ kpart, int *__restrict__ need_more)
{
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    double s4 = 0.0;
    int nk = 0;
    if (pidx < P)
    {
        const double2 *hd = reinterpret_cast<const double2 *>(head + (size_t)TG_HEAD * pidx);
        const double2 a = hd[0], b = hd[1], c = hd[2], d = hd[3], e = hd[4], f = hd[5], g = hd[6], hq = hd[7];
        // a = S0_0 S0_1 | b = S0_2 S2_0 | c = S2_1 S2_2 | d = S3_0 S3_1 | e = S3_2 q0_0 | f = q0_1 q2_0 | g = q2_1 q3_0 | hq = q3_1 0
        const double f00 = soft_f(a.x, a.x, lambda, expw), f01 = soft_f(a.y, a.x, lambda, expw), f02 = soft_f(b.x, a.x, lambda, expw);
        const double f20 = soft_f(b.y, b.y, lambda, expw), f21 = soft_f(c.x, b.y, lambda, expw), f22 = soft_f(c.y, b.y, lambda, expw);
        const double f30 = soft_f(d.x, d.x, lambda, expw), f31 = soft_f(d.y, d.x, lambda, expw), f32 = soft_f(e.x, d.x, lambda, expw);
        if (f02 != 0.0 || f22 != 0.0 || f32 != 0.0)
            *need_more = 1;
        // slot order of k_eval3: fma(f2, q2, fma(f3, q3, -2 f0 q0)) per triplet
        s4 = fma(f20, f.y, fma(f30, g.y, -2.0 * f00 * e.y)) + fma(f21, g.x, fma(f31, hq.x, -2.0 * f01 * f.x));
        fth[pidx] = make_double2(f00, f01);
        nk = (f00 != 0.0) + (f01 != 0.0);
    }
    s4 = warp_sum(s4);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        nk += __shfl_xor_sync(0xffffffffu, nk, o);
    if ((threadIdx.x & 31) == 0)
    {
        const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        s4part[w] = s4;
        kpart[w] = nk;
    }
}

// MODE 0: partial[2 * cta + {0,1}] = sum (Uhat - U)^2, sum Uhat over the CTA's output tile.
// MODE 1: outY = Uhat * scale for slice kfix (grid.x = 1).
// Per colour step an octet (8 lanes) owns eight bins of one bin column:
//   M  lane q fetches the metadata of "its" bin — offsets, up to TG_K entries, their thresholds: three dependent loads, but one
//      instruction each for the octet's eight bins;
//   F  the entries are flattened into a per-octet work list in shared memory (prefix sum over the octet);
//   P  the eight lanes walk the list four entries at a time: one 128-byte u line per entry (16 bytes per lane), then a plain
//      read-modify-write of the lane's two tile elements.  Entries of one bin are adjacent in the list, so the adds to a
//      voxel happen in a fixed order; bins of one colour never overlap, so no atomics.
// (First version: bin after bin through four dependent loads with one line in flight per octet — 2.5 ms per evaluation;
//  second: lane-parallel metadata + register accumulation over bins x rounds — 1.26 ms, 490 M warp instructions.)
#define TG_K 3 /* entries per bin fetched lane-parallel; longer bins and second triplets finish in a serial tail */
struct __align__(16) TgWork
{
    int pidx, row; // patch, bin row index (0..15) of the region
    double g;      // f0 * v0[k]
};
template <int MODE>
__global__ void __launch_bounds__(128, 8)
    k_tile_eval(const int *__restrict__ start, const TgEntry *__restrict__ ent, const double *__restrict__ fac0,
                const double *__restrict__ u0c, const double2 *__restrict__ fth, const double *__restrict__ u,
                const unsigned *__restrict__ cnt, int N, int kfix, double scale, double *__restrict__ outY, double *__restrict__ partial)
{
    __shared__ __align__(16) double tile[TG_LDC * TG_TC];
    __shared__ TgWork wl[16][8 * TG_K];
    // grid = (slices, tile rows, tile columns): the CTAs of one spatial tile run together and share its u vectors through L2
    const int k = MODE ? kfix : blockIdx.x;
    const int r_org = blockIdx.y * TG_VR - 3, c_org = blockIdx.z * TG_VC - 3;
    for (int i = threadIdx.x; i < TG_LDC * TG_TC; i += 128)
        tile[i] = 0.0;
    __syncthreads();
    const int q = threadIdx.x & 7;    // lane within the octet: entries (2q, 2q+1) of the 4 x 4 block
    const int oct = threadIdx.x >> 3; // 16 octets: bin column j = oct >> 1, bin rows i = 2 b + (oct & 1), b = 0..7
    const int dr = 2 * (q & 1), dc = q >> 1;
    const int j = oct >> 1, ipar = oct & 1;
    const size_t kbase = (size_t)k * N * N;
    const int M = N - 4; // largest block origin
    TgWork *mywl = wl[oct];
#pragma unroll 1
    for (int color = 0; color < 16; color++)
    {
        const int cr = color & 3, cc = color >> 2;
        // tile index of this lane's two elements for a bin in row i of the region: i + L0, i + L1
        const int L0 = ((cr + dr) >> 2) + TG_RG * ((cr + dr) & 3) + TG_LDC * (4 * j + cc + dc);
        const int L1 = ((cr + dr + 1) >> 2) + TG_RG * ((cr + dr + 1) & 3) + TG_LDC * (4 * j + cc + dc);
        // ---- M: metadata of this lane's bin (row 2q + ipar of the octet's column) ----
        int e0 = 0, ne = 0;
        {
            const int br = r_org + 4 * (2 * q + ipar) + cr, bc = c_org + 4 * j + cc;
            if (br >= 0 && bc >= 0 && br <= M && bc <= M)
            {
                const size_t b = kbase + br + (size_t)N * bc;
                e0 = __ldg(start + b);
                ne = __ldg(start + b + 1) - e0;
            }
        }
        int pidx[TG_K];
        double g0[TG_K];
        bool second = false; // some entry of the bin keeps a second triplet (rare): the whole bin goes through the serial tail
#pragma unroll
        for (int kk = 0; kk < TG_K; kk++)
        {
            pidx[kk] = 0;
            g0[kk] = 0.0;
            if (kk < ne)
            {
                const int4 raw = __ldg(reinterpret_cast<const int4 *>(ent + e0 + kk));
                pidx[kk] = raw.x;
                g0[kk] = __hiloint2double(raw.w, raw.z); // v for now
            }
        }
#pragma unroll
        for (int kk = 0; kk < TG_K; kk++)
            if (kk < ne)
            {
                const double2 f = __ldg(fth + pidx[kk]);
                g0[kk] *= f.x;
                second = second || f.y != 0.0;
            }
        const int nfast = second ? 0 : min(ne, TG_K); // entries that go through the work list
        // ---- F: flatten into the octet's work list ----
        int off = nfast;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, off, o, 8);
            if (q >= o)
                off += t;
        }
        const int T = __shfl_sync(0xffffffffu, off, 7, 8);
        off -= nfast;
#pragma unroll
        for (int kk = 0; kk < TG_K; kk++)
            if (kk < nfast)
            {
                TgWork w;
                w.pidx = pidx[kk];
                w.row = 2 * q + ipar;
                w.g = g0[kk];
                mywl[off + kk] = w;
            }
        __syncwarp();
        // ---- P: walk the list, four u lines in flight per octet ----
        for (int t0 = 0; t0 < T; t0 += 4)
        {
            TgWork w[4];
            double2 uu[4];
#pragma unroll
            for (int x = 0; x < 4; x++)
            {
                w[x].g = 0.0;
                w[x].row = 0;
                w[x].pidx = 0;
                if (t0 + x < T)
                {
                    const int4 raw = *reinterpret_cast<const int4 *>(&mywl[t0 + x]);
                    w[x].pidx = raw.x;
                    w[x].row = raw.y;
                    w[x].g = __hiloint2double(raw.w, raw.z);
                }
                uu[x] = make_double2(0.0, 0.0);
                if (w[x].g != 0.0)
                    uu[x] = __ldg(reinterpret_cast<const double2 *>(u0c + (size_t)16 * w[x].pidx) + q);
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
                if (t0 + x < T)
                {
                    tile[w[x].row + L0] = fma(w[x].g, uu[x].x, tile[w[x].row + L0]);
                    tile[w[x].row + L1] = fma(w[x].g, uu[x].y, tile[w[x].row + L1]);
                }
        }
        // ---- tail: bins with more than TG_K entries or a second surviving triplet, entry by entry (rare) ----
        const bool slow = ne > nfast;
        if (__any_sync(0xffffffffu, slow))
        {
#pragma unroll 1
            for (int b = 0; b < 8; b++)
            {
                const int eb0 = __shfl_sync(0xffffffffu, e0, b, 8), neb = __shfl_sync(0xffffffffu, ne, b, 8);
                const int nfb = __shfl_sync(0xffffffffu, nfast, b, 8);
                double a0 = 0.0, a1 = 0.0;
                for (int e = eb0 + nfb; e < eb0 + neb; e++)
                {
                    const int4 raw = __ldg(reinterpret_cast<const int4 *>(ent + e));
                    const double2 f = __ldg(fth + raw.x);
                    const double *R = fac0 + (size_t)SVD16_REC * raw.x;
                    if (f.x != 0.0)
                    {
                        const double2 u0 = __ldg(reinterpret_cast<const double2 *>(R) + q);
                        const double g = f.x * __hiloint2double(raw.w, raw.z);
                        a0 = fma(g, u0.x, a0);
                        a1 = fma(g, u0.y, a1);
                    }
                    if (f.y != 0.0)
                    { // second surviving triplet: u_1 and v_1[k] straight from the record
                        const double2 u1 = __ldg(reinterpret_cast<const double2 *>(R + SVD16_M) + q);
                        const double g1 = f.y * __ldg(R + SVD16_M * SVD16_N + SVD16_LDV + k);
                        a0 = fma(g1, u1.x, a0);
                        a1 = fma(g1, u1.y, a1);
                    }
                }
                if (neb > nfb)
                {
                    tile[2 * b + ipar + L0] += a0;
                    tile[2 * b + ipar + L1] += a1;
                }
            }
        }
        __syncthreads();
    }
    double s1 = 0.0, s5 = 0.0;
    for (int idx = threadIdx.x; idx < TG_VR * TG_VC; idx += 128)
    {
        const int lr = idx % TG_VR + 3, lc = idx / TG_VR + 3;
        const int gr = r_org + lr, gc = c_org + lc;
        if (gr < N && gc < N)
        {
            const size_t vox = kbase + gr + (size_t)N * gc;
            const double v0 = norm_or_zero(tile[tg_idx(lr, lc)], cnt[vox]);
            if (MODE)
                outY[gr + (size_t)N * gc] = v0 * scale;
            else
            {
                const double d = v0 - u[vox];
                s1 = fma(d, d, s1);
                s5 += v0;
            }
        }
    }
    if (!MODE)
    {
        __shared__ double sm[2][4];
        s1 = warp_sum(s1);
        s5 = warp_sum(s5);
        if ((threadIdx.x & 31) == 0)
        {
            sm[0][threadIdx.x >> 5] = s1;
            sm[1][threadIdx.x >> 5] = s5;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const size_t cta = blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
            partial[2 * cta] = (sm[0][0] + sm[0][1]) + (sm[0][2] + sm[0][3]);
            partial[2 * cta + 1] = (sm[1][0] + sm[1][1]) + (sm[1][2] + sm[1][3]);
        }
    }
}

// fixed-order reduction of an evaluation's partial sums: out = {s1, s5, s4, triplets}
__global__ void __launch_bounds__(1024) k_reduce_eval(const double *__restrict__ tpart, int ntile, const double *__restrict__ s4part,
                                                      const int *__restrict__ kpart, int nw, double *__restrict__ out)
{
    double s1 = 0, s5 = 0, s4 = 0, sk = 0;
    for (int i = threadIdx.x; i < ntile; i += 1024)
    {
        s1 += tpart[2 * i];
        s5 += tpart[2 * i + 1];
    }
    for (int i = threadIdx.x; i < nw; i += 1024)
    {
        s4 += s4part[i];
        sk += (double)kpart[i];
    }
    __shared__ double sm[4][32];
    double v[4] = {s1, s5, s4, sk};
#pragma unroll
    for (int qn = 0; qn < 4; qn++)
    {
        const double r = warp_sum(v[qn]);
        if ((threadIdx.x & 31) == 0)
            sm[qn][threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
#pragma unroll
        for (int qn = 0; qn < 4; qn++)
        {
            const double r = warp_sum(sm[qn][threadIdx.x]);
            if (threadIdx.x == 0)
                out[qn] = r;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// K_accu_seq — arma::accu(u) BIT FOR BIT.  The lambda search starts at accu(u) / (Nx Ny Nt) (pguresvt.hpp:139) and its
// whole probe sequence — hence the lambda it ends on inside the flat basin of the objective — follows from the last bits of
// that number (tools/diag_lambda.py: with a tree-ordered sum 5 of 16 frames of the bench sample ended 1e-4 .. 3e-3 away from
// the oracle, with the exact sum all of them coincide).  Armadillo sums a cube with TWO sequential accumulators (even / odd
// elements, acc1 + acc2 at the end); a floating-point running sum cannot be re-associated, but it can be EMULATED in integer
// arithmetic: while the accumulator stays in one binade (acc = A 2^(e-52), 2^52 <= A < 2^53) and the addend has a smaller
// exponent, fl(acc + a) = (A + q + round) 2^(e-52) with q = m >> sh, and round-to-nearest-even is "+1 if the remainder
// exceeds half an ulp; on an exact tie round A + q up to even".  A tie makes the result even whatever came before, so a run
// of elements is summarised by (c1, tie?, c2): A -> A + c1 without a tie, A -> roundup_even(A + c1) + c2 with one — an
// associative composition, i.e. a parallel scan.  One CTA per accumulator walks its elements in chunks; where the emulation's
// preconditions fail (binade crossing, addend not smaller than the accumulator, zero accumulator) the offending thread's
// few elements are added with real FP64 adds and the scan resumes behind it.
// ------------------------------------------------------------------------------------------------------------------------
#define AS_THREADS 256
#define AS_E 32 /* elements per thread and chunk (the block scan is amortised over them: 8 per thread took 15 ms per window) */
struct AsFn
{
    unsigned long long c1, c2;
    int tie;
};
#define AS_SAT (1ull << 62)
__device__ __forceinline__ unsigned long long as_sadd(unsigned long long a, unsigned long long b)
{
    const unsigned long long s = a + b;
    return (s < a || s > AS_SAT) ? AS_SAT : s;
}
// f first, then g
__device__ __forceinline__ AsFn as_compose(const AsFn &f, const AsFn &g)
{
    AsFn r;
    if (!g.tie)
    {
        r.tie = f.tie;
        r.c1 = f.tie ? f.c1 : as_sadd(f.c1, g.c1);
        r.c2 = f.tie ? as_sadd(f.c2, g.c1) : 0ull;
    }
    else if (f.tie)
    { // value after f is even + f.c2
        r.tie = 1;
        r.c1 = f.c1;
        r.c2 = as_sadd((as_sadd(f.c2, g.c1) + 1ull) & ~1ull, g.c2);
    }
    else
    {
        r.tie = 1;
        r.c1 = as_sadd(f.c1, g.c1);
        r.c2 = g.c2;
    }
    return r;
}
__device__ __forceinline__ unsigned long long as_apply(const AsFn &f, unsigned long long A)
{
    return f.tie ? as_sadd((as_sadd(A, f.c1) + 1ull) & ~1ull, f.c2) : as_sadd(A, f.c1);
}

// out[blockIdx.x] = sequential sum of u[blockIdx.x], u[blockIdx.x + 2], ... (n elements in total, both parities)
__global__ void __launch_bounds__(AS_THREADS) k_accu_seq(const double *__restrict__ u, size_t n, double *__restrict__ out)
{
    const int par = blockIdx.x;
    const size_t nel = (n + 1 - par) / 2; // elements of this accumulator
    __shared__ AsFn wfn[AS_THREADS / 32];
    __shared__ double s_acc;
    __shared__ int s_stop;
    extern __shared__ double stage[]; // the chunk: AS_THREADS * (AS_E + 1) doubles, element j at j + j / AS_E (conflict-free runs)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double acc = 0.0;
    for (size_t base = 0; base < nel && acc == acc; base += (size_t)AS_THREADS * AS_E)
    {
        // coalesced 16-byte loads of (even, odd) pairs, this accumulator's half staged in shared memory.  All loads of a full
        // chunk are issued before the first one is consumed (one DRAM latency per chunk, not one per element: the first version
        // selected and stored right behind every load and took 21 us per chunk)
        if (2 * (base + (size_t)AS_THREADS * AS_E) <= n)
        {
            double2 pr[AS_E];
#pragma unroll
            for (int r = 0; r < AS_E; r++)
                pr[r] = __ldg(reinterpret_cast<const double2 *>(u) + base + r * AS_THREADS + threadIdx.x);
#pragma unroll
            for (int r = 0; r < AS_E; r++)
            {
                const int jl = r * AS_THREADS + threadIdx.x;
                stage[jl + jl / AS_E] = par ? pr[r].y : pr[r].x;
            }
        }
        else
        {
#pragma unroll 1
            for (int r = 0; r < AS_E; r++)
            {
                const int jl = r * AS_THREADS + threadIdx.x;
                const size_t j = base + jl;
                double v = 0.0;
                if (j < nel)
                {
                    if (2 * j + 1 < n)
                    {
                        const double2 pr = reinterpret_cast<const double2 *>(u)[j];
                        v = par ? pr.y : pr.x;
                    }
                    else
                        v = u[2 * j + par];
                }
                stage[jl + jl / AS_E] = v;
            }
        }
        __syncthreads();
        // this thread's AS_E consecutive elements of the chunk
        double a[AS_E];
#pragma unroll
        for (int k = 0; k < AS_E; k++)
            a[k] = stage[threadIdx.x * (AS_E + 1) + k];
        int seg = 0; // threads below seg are done
        for (;;)
        {
            const long long abits = __double_as_longlong(acc);
            const int ebias = (int)((abits >> 52) & 0x7ff);
            const bool acc_ok = abits > 0 && ebias > 0 && ebias < 0x7ff; // positive normal accumulator
            const unsigned long long A0 = ((unsigned long long)abits & ((1ull << 52) - 1)) | (1ull << 52);
            AsFn f;
            f.c1 = f.c2 = 0ull;
            f.tie = 0;
            bool bad = false;
            if ((int)threadIdx.x >= seg)
            {
                // the elements' increments are independent of each other; only an exact tie (rare: the dropped bits equal
                // half an ulp exactly) makes the composition order-dependent, so the common case is a plain integer sum
                unsigned long long incs[AS_E];
                unsigned tiemask = 0u;
#pragma unroll
                for (int k = 0; k < AS_E; k++)
                {
                    const long long b = __double_as_longlong(a[k]);
                    const int eb = (int)((b >> 52) & 0x7ff);
                    const int sh = ebias - eb;
                    incs[k] = 0ull;
                    if (b == 0)
                        continue; // +0.0: identity
                    if (!acc_ok || b < 0 || eb == 0 || sh < 1)
                    { // zero / non-finite accumulator, negative, subnormal or not-smaller addend: real adds for this thread
                        bad = true;
                        continue;
                    }
                    if (sh < 64)
                    {
                        const unsigned long long m = ((unsigned long long)b & ((1ull << 52) - 1)) | (1ull << 52);
                        const unsigned long long rem = m & ((1ull << sh) - 1), half = 1ull << (sh - 1);
                        incs[k] = (m >> sh) + (rem > half ? 1ull : 0ull);
                        if (rem == half)
                            tiemask |= 1u << k;
                    }
                }
                if (tiemask == 0u)
                {
                    unsigned long long t = 0ull;
#pragma unroll
                    for (int k = 0; k < AS_E; k++)
                        t += incs[k]; // (each below 2^52: no overflow)
                    f.c1 = t > AS_SAT ? AS_SAT : t;
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < AS_E; k++)
                    {
                        AsFn g;
                        g.tie = (tiemask >> k) & 1u;
                        g.c1 = incs[k];
                        g.c2 = 0ull;
                        f = as_compose(f, g);
                    }
                }
            }
            // inclusive scan of the functions over the block (identity for finished threads)
            AsFn inc = f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                AsFn p;
                p.c1 = __shfl_up_sync(0xffffffffu, inc.c1, o);
                p.c2 = __shfl_up_sync(0xffffffffu, inc.c2, o);
                p.tie = __shfl_up_sync(0xffffffffu, inc.tie, o);
                if (lane >= o)
                    inc = as_compose(p, inc);
            }
            if (lane == 31)
                wfn[w] = inc;
            if (threadIdx.x == 0)
                s_stop = AS_THREADS;
            __syncthreads();
            AsFn pre;
            pre.c1 = pre.c2 = 0ull;
            pre.tie = 0;
            for (int q2 = 0; q2 < w; q2++)
                pre = as_compose(pre, wfn[q2]);
            const AsFn incl = as_compose(pre, inc);
            // exclusive prefix of this thread = inclusive of the previous one
            AsFn excl;
            excl.c1 = __shfl_up_sync(0xffffffffu, inc.c1, 1);
            excl.c2 = __shfl_up_sync(0xffffffffu, inc.c2, 1);
            excl.tie = __shfl_up_sync(0xffffffffu, inc.tie, 1);
            if (lane == 0)
                excl = pre;
            else
                excl = as_compose(pre, excl);
            const unsigned long long Aex = as_apply(excl, A0), Ain = as_apply(incl, A0);
            const bool flagged = (int)threadIdx.x >= seg && (bad || Ain >= (1ull << 53));
            if (flagged)
                atomicMin(&s_stop, (int)threadIdx.x);
            __syncthreads();
            const int stop = s_stop;
            if (stop == AS_THREADS)
            { // whole remainder of the chunk emulated: the last thread holds the new accumulator
                if (threadIdx.x == AS_THREADS - 1)
                    s_acc = acc_ok ? __longlong_as_double((long long)(((unsigned long long)ebias << 52) | (Ain & ((1ull << 52) - 1)))) : acc;
                __syncthreads();
                acc = s_acc;
                __syncthreads();
                break;
            }
            if ((int)threadIdx.x == stop)
            { // everything before this thread is valid: continue from its exclusive value with real adds
                double x = acc_ok ? __longlong_as_double((long long)(((unsigned long long)ebias << 52) | (Aex & ((1ull << 52) - 1)))) : acc;
#pragma unroll
                for (int k = 0; k < AS_E; k++)
                    x = __dadd_rn(x, a[k]);
                s_acc = x;
            }
            __syncthreads();
            acc = s_acc;
            seg = stop + 1;
            __syncthreads();
            if (seg >= AS_THREADS)
                break;
        }
    }
    if (threadIdx.x == 0)
        out[par] = acc;
}

} // namespace pgs
