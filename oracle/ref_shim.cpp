// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C entry points over the two pieces of the reference that compile standalone in this image
// (SURVEY §8c): the vendored PCG generator + libstdc++ Bernoulli draw exactly as pgure.hpp:167-186
// uses them, and the constant-time median filter of medfilter.hpp:478-539 exactly as
// pguresvt.hpp:75-76 calls it.  The reference sources are compiled where they lie
// (-I/root/reference/src); nothing is copied.  Output: oracle/_ref/libpguresvt_ref.so (git-ignored).
#include <cmath>
#include <cstdint>
#include <random>

#include "pcg/pcg_random.hpp"
#include "medfilter.hpp"

extern "C" void ref_pcg64_raw(int64_t seed, uint64_t *out, int n)
{
    pcg64 RNG;
    RNG.seed(seed); // pgure.hpp:58
    for (int i = 0; i < n; i++)
        out[i] = RNG();
}

extern "C" void ref_perturbations(int64_t seed, int64_t n, int64_t *delta1, double *delta2)
{
    pcg64 RNG;
    RNG.seed(seed);
    auto bernoulliFunc = [&](std::bernoulli_distribution &dist, auto value1, auto value2) {
        return (dist(RNG)) ? value1 : value2;
    };
    double kappa = 1.;
    double vP = 0.5 + 0.5 * kappa / std::sqrt(kappa * kappa + 4);
    double vQ = 1 - vP;
    double vQvP = std::sqrt(vQ / vP);
    double vPvQ = std::sqrt(vP / vQ);
    std::bernoulli_distribution binary_dist1(0.5);
    std::bernoulli_distribution binary_dist2(vP);
    for (int64_t i = 0; i < n; i++)
        delta1[i] = bernoulliFunc(binary_dist1, -1, 1);
    for (int64_t i = 0; i < n; i++)
        delta2[i] = static_cast<double>(bernoulliFunc(binary_dist2, -1 * vQvP, vPvQ));
}

// pguresvt.hpp:75-76: ConstantTimeMedianFilter(src, dst, Nx, Ny, Nx, Nx, medianSize, 1, 1024*1024)
extern "C" void ref_ctmf(const uint16_t *src, uint16_t *dst, int nx, int ny, int r)
{
    ConstantTimeMedianFilter(src, dst, nx, ny, nx, nx, r, 1, 1024 * 1024);
}
