// Hot-pixel prefilter — HotPixelFilter<uint16_t> (src/hotpixel.hpp:19-64, called by the CLI at
// src/PGURE-SVT.cpp:171-179) on the GPU.  Per frame:
//   median = median(median(frame))                      → per-column medians, then their median
//   mad    = median(median(|frame - median|)) / 0.6745  → in uint16 modular arithmetic
//   outliers = |frame - median| > threshold * mad       → found on the ORIGINAL values, column-major order
//   interior outlier ← mean of the 4th and 5th of its sorted 8 neighbours, edge outlier ← frame median,
//   applied sequentially in place, so later outliers see earlier replacements (hotpixel.hpp:36-58).
// Medians are radix selects (block_select of noise.cuh); the sparse, order-dependent replacement runs as one
// thread per frame over a compacted, ordered outlier list.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "noise.cuh"

namespace pgs
{

__device__ __forceinline__ int median_from_two(int v1, int v2, int n)
{ // arma::median of integers: nth element, averaged with the largest of the lower half when n is even (robust_mean)
    return (n % 2 == 0) ? v1 + (v2 - v1) / 2 : v1;
}

// grid = (columns, frames); column median of the frame (mode 0) or of (uint16)(frame - med[frame]) (mode 1)
template <int NT>
__global__ void __launch_bounds__(NT) k_hp_colmed(const uint16_t *__restrict__ seq, int nr, int nc, int mode,
                                                  const int *__restrict__ med, int *__restrict__ colmed)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long spre[2];
    const uint16_t *col = seq + (size_t)nr * nc * blockIdx.y + (size_t)nr * blockIdx.x;
    const uint16_t m16 = mode ? (uint16_t)med[blockIdx.y] : (uint16_t)0;
    auto get = [&](int e) { return (double)(uint16_t)(col[e] - m16); };
    const int v1 = (int)block_select<NT>(get, nr, nr / 2, hist, spre);
    const int v2 = (nr % 2 == 0) ? (int)block_select<NT>(get, nr, nr / 2 - 1, hist, spre) : v1;
    if (threadIdx.x == 0)
        colmed[(size_t)nc * blockIdx.y + blockIdx.x] = (int)(uint16_t)median_from_two(v1, v2, nr);
}

// grid = frames; median of the column medians
template <int NT>
__global__ void __launch_bounds__(NT) k_hp_medofmed(const int *__restrict__ colmed, int nc, int *__restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long spre[2];
    const int *c = colmed + (size_t)nc * blockIdx.x;
    auto get = [&](int e) { return (double)c[e]; };
    const int v1 = (int)block_select<NT>(get, nc, nc / 2, hist, spre);
    const int v2 = (nc % 2 == 0) ? (int)block_select<NT>(get, nc, nc / 2 - 1, hist, spre) : v1;
    if (threadIdx.x == 0)
        out[blockIdx.x] = (int)(uint16_t)median_from_two(v1, v2, nc);
}

// flag outliers of the ORIGINAL frame and count them per column.  grid-stride over pixels, grid.y = frame
__global__ void k_hp_flag(const uint16_t *__restrict__ seq, int nr, int nc, const int *__restrict__ med, const int *__restrict__ madraw,
                          double threshold, uint8_t *__restrict__ flags, int *__restrict__ colcnt)
{
    const size_t fsz = (size_t)nr * nc;
    const uint16_t m16 = (uint16_t)med[blockIdx.y];
    const double mad = (double)madraw[blockIdx.y] * (1.0 / 0.6745);
    const double tv = threshold * mad;
    const uint16_t thr16 = (tv >= 65535.0) ? (uint16_t)65535 : (uint16_t)tv;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < fsz; i += (size_t)gridDim.x * blockDim.x)
    {
        const uint16_t dev = (uint16_t)(seq[fsz * blockIdx.y + i] - m16);
        const uint8_t o = dev > thr16;
        flags[fsz * blockIdx.y + i] = o;
        if (o)
            atomicAdd(&colcnt[(size_t)nc * blockIdx.y + (i / nr)], 1);
    }
}

// exclusive scan of the per-column counts (one CTA per frame, serial over chunks of NT columns)
template <int NT>
__global__ void __launch_bounds__(NT) k_hp_scan(const int *__restrict__ colcnt, int nc, int *__restrict__ coloff, int *__restrict__ total)
{
    __shared__ int sm[NT];
    __shared__ int carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    const int *c = colcnt + (size_t)nc * blockIdx.x;
    int *o = coloff + (size_t)nc * blockIdx.x;
    for (int base = 0; base < nc; base += NT)
    {
        const int idx = base + threadIdx.x;
        const int v = (idx < nc) ? c[idx] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < NT; d <<= 1)
        {
            const int t = (threadIdx.x >= d) ? sm[threadIdx.x - d] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        if (idx < nc)
            o[idx] = carry + sm[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += sm[NT - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        total[blockIdx.x] = carry;
}

// ordered outlier list: thread per (column, frame) walks its column top to bottom
__global__ void k_hp_list(const uint8_t *__restrict__ flags, int nr, int nc, const int *__restrict__ coloff, int *__restrict__ list)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc)
        return;
    const size_t fsz = (size_t)nr * nc;
    const uint8_t *f = flags + fsz * blockIdx.y + (size_t)nr * c;
    int *l = list + fsz * blockIdx.y;
    int o = coloff[(size_t)nc * blockIdx.y + c];
    for (int r = 0; r < nr; r++)
        if (f[r])
            l[o++] = r + nr * c;
}

// sequential in-place replacement (one thread per frame)
__global__ void k_hp_fix(uint16_t *__restrict__ seq, int nr, int nc, const int *__restrict__ list, const int *__restrict__ total,
                         const int *__restrict__ med, int nframes)
{
    const int fr = blockIdx.x * blockDim.x + threadIdx.x;
    if (fr >= nframes)
        return;
    const size_t fsz = (size_t)nr * nc;
    uint16_t *f = seq + fsz * fr;
    const int *l = list + fsz * fr;
    const int n = total[fr];
    const double median = (double)med[fr];
    for (int q = 0; q < n; q++)
    {
        const int k = l[q], r = k % nr, c = k / nr;
        if (r > 0 && r < nr - 1 && c > 0 && c < nc - 1)
        {
            double w[8] = {(double)f[(r - 1) + (size_t)nr * (c - 1)], (double)f[(r - 1) + (size_t)nr * c],
                           (double)f[(r - 1) + (size_t)nr * (c + 1)], (double)f[r + (size_t)nr * (c - 1)],
                           (double)f[r + (size_t)nr * (c + 1)],       (double)f[(r + 1) + (size_t)nr * (c - 1)],
                           (double)f[(r + 1) + (size_t)nr * c],       (double)f[(r + 1) + (size_t)nr * (c + 1)]};
            for (int a = 1; a < 8; a++)
            { // insertion sort of 8
                const double v = w[a];
                int b = a - 1;
                while (b >= 0 && w[b] > v)
                {
                    w[b + 1] = w[b];
                    b--;
                }
                w[b + 1] = v;
            }
            f[k] = (uint16_t)(0.5 * (w[3] + w[4]));
        }
        else
            f[k] = (uint16_t)median;
    }
}

#define HCU(call)                                                                                         \
    do                                                                                                    \
    {                                                                                                     \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
        {                                                                                                 \
            char b_[256];                                                                                 \
            snprintf(b_, sizeof b_, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            err = b_;                                                                                     \
            rc = 2;                                                                                       \
            goto done;                                                                                    \
        }                                                                                                 \
    } while (0)

static int hotpixel_filter_u16(uint16_t *seq, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, double threshold, int device,
                               std::string &err)
{
    int rc = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        err = "no CUDA device available (the PGURE-SVT hot path has no CPU fallback)";
        return 2;
    }
    if (n_rows < 3 || n_cols < 3 || n_rows > 32767 || n_cols > 32767)
    {
        err = "hot-pixel filter: unsupported frame size";
        return 1;
    }
    const size_t fsz = (size_t)n_rows * n_cols;
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_frames, ((size_t)512 << 20) / (fsz * 7)));
    uint16_t *dSeq = nullptr;
    uint8_t *dFlags = nullptr;
    int *dList = nullptr, *dColmed = nullptr, *dColcnt = nullptr, *dColoff = nullptr, *dMed = nullptr, *dMad = nullptr, *dTotal = nullptr;
    HCU(cudaSetDevice(device));
    HCU(cudaMalloc(&dSeq, fsz * chunk * sizeof(uint16_t)));
    HCU(cudaMalloc(&dFlags, fsz * chunk));
    HCU(cudaMalloc(&dList, fsz * chunk * sizeof(int)));
    HCU(cudaMalloc(&dColmed, (size_t)n_cols * chunk * sizeof(int)));
    HCU(cudaMalloc(&dColcnt, (size_t)n_cols * chunk * sizeof(int)));
    HCU(cudaMalloc(&dColoff, (size_t)n_cols * chunk * sizeof(int)));
    HCU(cudaMalloc(&dMed, chunk * sizeof(int)));
    HCU(cudaMalloc(&dMad, chunk * sizeof(int)));
    HCU(cudaMalloc(&dTotal, chunk * sizeof(int)));
    for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk)
    {
        const uint32_t nf = std::min(chunk, n_frames - f0);
        HCU(cudaMemcpy(dSeq, seq + fsz * f0, fsz * nf * sizeof(uint16_t), cudaMemcpyHostToDevice));
        k_hp_colmed<128><<<dim3(n_cols, nf), 128>>>(dSeq, (int)n_rows, (int)n_cols, 0, nullptr, dColmed);
        k_hp_medofmed<256><<<nf, 256>>>(dColmed, (int)n_cols, dMed);
        k_hp_colmed<128><<<dim3(n_cols, nf), 128>>>(dSeq, (int)n_rows, (int)n_cols, 1, dMed, dColmed);
        k_hp_medofmed<256><<<nf, 256>>>(dColmed, (int)n_cols, dMad);
        HCU(cudaMemset(dColcnt, 0, (size_t)n_cols * nf * sizeof(int)));
        k_hp_flag<<<dim3(std::min<unsigned>((unsigned)((fsz + 255) / 256), 1184u), nf), 256>>>(dSeq, (int)n_rows, (int)n_cols, dMed, dMad,
                                                                                              threshold, dFlags, dColcnt);
        k_hp_scan<256><<<nf, 256>>>(dColcnt, (int)n_cols, dColoff, dTotal);
        k_hp_list<<<dim3((n_cols + 127) / 128, nf), 128>>>(dFlags, (int)n_rows, (int)n_cols, dColoff, dList);
        k_hp_fix<<<(nf + 31) / 32, 32>>>(dSeq, (int)n_rows, (int)n_cols, dList, dTotal, dMed, (int)nf);
        HCU(cudaGetLastError());
        HCU(cudaMemcpy(seq + fsz * f0, dSeq, fsz * nf * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    }
done:
    cudaFree(dSeq), cudaFree(dFlags), cudaFree(dList), cudaFree(dColmed), cudaFree(dColcnt), cudaFree(dColoff), cudaFree(dMed),
        cudaFree(dMad), cudaFree(dTotal);
    return rc;
}
#undef HCU
} // namespace pgs
