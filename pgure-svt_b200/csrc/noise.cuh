// Noise estimation stage (src/noise.hpp:35-458) — GPU implementation.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace pgs
{
struct NoiseWorkspace
{
    void release() {}
};

static int noise_estimate_window(NoiseWorkspace &, const double *, int, int, int, int, cudaStream_t, double &, double &,
                                 double &, long long *, std::string &err)
{
    err = "noise estimation (noise_alpha/mu/sigma < 0) is not available on the GPU path yet";
    return 3; // PGS_ERR_UNSUPPORTED
}
} // namespace pgs
