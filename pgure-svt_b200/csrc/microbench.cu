// Micro-benchmarks that supply the roofline denominators MEASURED_PEAKS.json lacks:
//   * FP64 DFMA peak of the vector pipe (the SVD kernel's bound),
//   * FP64 atomicAdd (RED.ADD.F64) throughput to an L2/HBM-resident cube (the overlap-add's bound),
//   * plain HBM copy bandwidth for cross-checking MEASURED_PEAKS.json.
// Prints one JSON object.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a microbench.cu -o ../microbench
#include <cstdio>
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
           "\"atomic_f64_gops_strided\": %.1f, \"copy_gbs\": %.1f, \"cuda_error\": \"%s\"}\n",
           best, sustained, atom_g[0], atom_g[1], copy, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
