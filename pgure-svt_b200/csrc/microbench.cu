// Micro-benchmarks that supply the roofline denominators MEASURED_PEAKS.json lacks:
//   * FP64 DFMA peak of the vector pipe (the SVD kernel's bound),
//   * FP64 atomicAdd (RED.ADD.F64) throughput to an L2/HBM-resident cube (the overlap-add's bound),
//   * plain HBM copy bandwidth for cross-checking MEASURED_PEAKS.json.
// Prints one JSON object.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a microbench.cu -o ../microbench
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++)
    {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_atomic(double *acc, size_t n, int reps, int stride)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; r++)
    {
        size_t j = (i * stride + (size_t)r * 4099) % n;
        atomicAdd(acc + j, 1.0);
    }
}
// integer (fixed-point) and FP32 variants of the RED pattern the overlap-add uses: 16 lanes cover 4 columns x 4 rows
template <typename T>
__global__ void k_atomic_patch(T *acc, int N, int reps)
{
    const int g = threadIdx.x & 15;
    const size_t patch = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int M1 = N - 3;
    const int y = (int)(patch % M1), x = (int)((patch / M1) % M1);
    for (int k = 0; k < reps; k++)
    {
        const int yy = min(max(y + ((int)(patch * 7 + k) % 5) - 2, 0), M1 - 1);
        const size_t vox = (size_t)(yy + (g & 3)) + (size_t)N * (x + (g >> 2)) + (size_t)N * N * k;
        atomicAdd(acc + vox, (T)1);
    }
}
// The same overlap-add pattern through the TMA: every patch slice is one 4 x 4 x 1 box of a 3-D FP64 tensor map,
// added to global memory with cp.reduce.async.bulk.tensor (arbitrary element coordinates, no 16-byte alignment
// requirement on the destination).  16 lanes per patch stage the 15 boxes in shared memory, lanes 0..14 issue one box each.
template <int MODE>
__global__ void __launch_bounds__(128) k_tma_patch(const __grid_constant__ CUtensorMap tmap, int N, int reps)
{
    __shared__ __align__(128) double sbox[8][15][16];
    const int g = threadIdx.x & 15, grp = threadIdx.x >> 4;
    const size_t patch = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int M1 = N - 3;
    const int y = (int)(patch % M1), x = (int)((patch / M1) % M1);
    for (int k = 0; k < reps; k++)
        sbox[grp][k][g] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (g < reps)
    {
        const int k = g;
        // NB the box start must be 16-byte aligned in the innermost dimension: odd FP64 rows raise "illegal instruction"
        // (measured), so this variant only visits even rows; arbitrary rows would need 6-row boxes padded with zeros.
        const int yy = min(max(y + ((int)(patch * 7 + k) % 5) - 2, 0), M1 - 1) & ~1;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(&sbox[grp][k][0]);
        if (MODE == 2)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap), "r"(yy),
                         "r"(x), "r"(k), "r"(sa)
                         : "memory");
        else
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap),
                         "r"(yy), "r"(x), "r"(k), "r"(sa)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
__global__ void k_copy(const double4 *a, double4 *b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b[i] = a[i];
}
static float timeit(cudaEvent_t e0, cudaEvent_t e1)
{
    float ms;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main()
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double *out;
    cudaMalloc(&out, 148 * 16 * 256 * sizeof(double));
    const int iters = 20000;
    double best = 0;
    for (int rep = 0; rep < 5; rep++)
    {
        cudaEventRecord(e0);
        k_dfma<<<148 * 16, 256>>>(out, iters);
        cudaEventRecord(e1);
        float ms = timeit(e0, e1);
        double tf = 2.0 * 8 * iters * 148.0 * 16 * 256 / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    // sustained: ~2 s back to back
    double sustained = 0;
    {
        cudaEventRecord(e0);
        int n = 0;
        for (; n < 60; n++)
            k_dfma<<<148 * 16, 256>>>(out, iters);
        cudaEventRecord(e1);
        float ms = timeit(e0, e1);
        sustained = 2.0 * 8 * iters * 148.0 * 16 * 256 * n / (ms * 1e-3) / 1e12;
    }
    // atomics: 126 MB cube (1024^2 x 15 doubles), 16 adds per thread, consecutive lanes -> consecutive doubles
    size_t n = (size_t)1024 * 1024 * 15;
    double *acc;
    cudaMalloc(&acc, n * sizeof(double));
    cudaMemset(acc, 0, n * sizeof(double));
    double atom_g[2];
    for (int mode = 0; mode < 2; mode++)
    {
        const int stride = mode == 0 ? 1 : 17;
        const size_t threads = (size_t)1 << 24;
        k_atomic<<<(unsigned)(threads / 256), 256>>>(acc, n, 2, stride);
        cudaEventRecord(e0);
        k_atomic<<<(unsigned)(threads / 256), 256>>>(acc, n, 16, stride);
        cudaEventRecord(e1);
        float ms = timeit(e0, e1);
        atom_g[mode] = threads * 16.0 / (ms * 1e-3) / 1e9;
    }
    // patch-pattern REDs: double vs unsigned long long vs float (1024^2 x 15 cube, 1.04M patches x 15 slices x 16 lanes)
    double patch_g[3];
    {
        const int N = 1024;
        const size_t np = (size_t)(N - 3) * (N - 3);
        const unsigned blocks = (unsigned)((np * 16 + 127) / 128);
        for (int v = 0; v < 3; v++)
        {
            cudaMemset(acc, 0, n * sizeof(double));
            for (int rep = 0; rep < 2; rep++)
            {
                cudaEventRecord(e0);
                if (v == 0) k_atomic_patch<double><<<blocks, 128>>>(acc, N, 15);
                if (v == 1) k_atomic_patch<unsigned long long><<<blocks, 128>>>((unsigned long long *)acc, N, 15);
                if (v == 2) k_atomic_patch<float><<<blocks, 128>>>((float *)acc, N, 15);
                cudaEventRecord(e1);
            }
            float ms = timeit(e0, e1);
            patch_g[v] = np * 16.0 * 15 / (ms * 1e-3) / 1e9;
        }
    }
    // TMA reduce-add variant of the patch pattern
    double tma_g = 0, tma_mode_g[3] = {0, 0, 0};
    const char *tma_err = "ok";
    {
        const int N = 1024;
        typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            tma_err = "no cuTensorMapEncodeTiled";
        else
        {
            for (int mode = 2; mode >= 0; mode--)
            {
                CUtensorMap tm;
                cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)N, 15};
                cuuint64_t strides[2] = {(cuuint64_t)N * 8, (cuuint64_t)N * N * 8};
                cuuint32_t box[3] = {4, 4, 1};
                cuuint32_t estr[3] = {1, 1, 1};
                CUresult r = ((encode_t)fn)(&tm, mode == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, acc, dims,
                                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS)
                {
                    tma_err = "encode failed";
                    continue;
                }
                const size_t np = (size_t)(N - 3) * (N - 3);
                const unsigned blocks = (unsigned)((np * 16 + 127) / 128);
                cudaMemset(acc, 0, n * sizeof(double));
                for (int rep = 0; rep < 2; rep++)
                {
                    cudaEventRecord(e0);
                    if (mode == 0) k_tma_patch<0><<<blocks, 128>>>(tm, N, 15);
                    if (mode == 1) k_tma_patch<1><<<blocks, 128>>>(tm, N, 15);
                    if (mode == 2) k_tma_patch<2><<<blocks, 128>>>(tm, N, 15);
                    cudaEventRecord(e1);
                }
                float ms = timeit(e0, e1);
                cudaError_t ce = cudaDeviceSynchronize();
                fprintf(stderr, "tma mode %d: %.3f ms, %s\n", mode, ms, cudaGetErrorString(ce));
                if (ce != cudaSuccess)
                {
                    tma_err = "kernel failed";
                    break;
                }
                tma_mode_g[mode] = np * 16.0 * 15 / (ms * 1e-3) / 1e9;
                if (mode == 0)
                {
                    tma_g = tma_mode_g[0];
                    double *hacc = (double *)malloc(n * sizeof(double));
                    cudaMemcpy(hacc, acc, n * sizeof(double), cudaMemcpyDeviceToHost);
                    double tot = 0;
                    for (size_t i = 0; i < n; i++)
                        tot += hacc[i];
                    free(hacc);
                    if (tot != 2.0 * np * 240.0)
                        tma_err = "sum mismatch";
                }
            }
        }
    }
    // copy
    size_t cn = (size_t)1 << 27; // 128M double4 = 4 GB
    double4 *a, *b;
    cudaMalloc(&a, cn * sizeof(double4) / 4);
    cudaMalloc(&b, cn * sizeof(double4) / 4);
    cn /= 4;
    double copy = 0;
    for (int rep = 0; rep < 5; rep++)
    {
        cudaEventRecord(e0);
        k_copy<<<148 * 8, 512>>>(a, b, cn);
        cudaEventRecord(e1);
        float ms = timeit(e0, e1);
        double gbs = 2.0 * cn * 32 / (ms * 1e-3) / 1e9;
        if (gbs > copy) copy = gbs;
    }
    printf("{\"dfma_tflops\": %.2f, \"dfma_tflops_sustained\": %.2f, \"atomic_f64_gops_coalesced\": %.1f, "
           "\"atomic_f64_gops_strided\": %.1f, \"patch_red_f64_gops\": %.1f, \"patch_red_u64_gops\": %.1f, \"patch_red_f32_gops\": %.1f, "
           "\"patch_tma_reduce_f64_gops\": %.1f, \"tma\": \"%s\", \"patch_tma_u64_gops\": %.1f, \"patch_tma_store_gops\": %.1f, \"copy_gbs\": %.1f, \"cuda_error\": \"%s\"}\n",
           best, sustained, atom_g[0], atom_g[1], patch_g[0], patch_g[1], patch_g[2], tma_g, tma_err, tma_mode_g[1], tma_mode_g[2], copy, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
