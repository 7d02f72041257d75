// PGURESVT<T1,T2>() — drop-in for the reference's C++ entry point (src/pguresvt.hpp:17-172): same template
// name, same argument order and meaning, same outputs (Y sized like X and zero-initialised, estimates
// (n_frames, 4) = [lambda, alpha, mu, sigma]).  The body only marshals into the C ABI
// (include/pguresvt_b200.h); all computation happens in libpguresvt_b200.so on the GPU.
// Return value: 0 on success like the reference (pguresvt.hpp:171); a PGS_ERR_* code otherwise (the reference
// would have thrown / terminated) — pguresvt_last_error() has the message.
#ifndef PGURESVT_B200_PGURESVT_HPP
#define PGURESVT_B200_PGURESVT_HPP

#include <cstdint>
#include <cstdlib>
#include <type_traits>

#ifdef PGURESVT_USE_ARMADILLO
#include <armadillo>
#else
#include "arma_shim.hpp"
#endif
#include "../../include/pguresvt_b200.h"

template <typename T1, typename T2>
uint32_t PGURESVT(arma::Cube<T2> &Y,
                  arma::Mat<T2> &estimates,
                  const arma::Cube<T1> &X,
                  const uint32_t trajLength,
                  const uint32_t blockSize,
                  const uint32_t blockOverlap,
                  const uint32_t motionWindow,
                  const int64_t medianSize,
                  const uint32_t noiseMethod,
                  const uint32_t maxIter,
                  const int64_t nJobs,
                  const int64_t randomSeed,
                  const bool optimizePGURE,
                  const bool expWeighting,
                  const bool motionEstimation,
                  const double lambdaEst,
                  const double alphaEst,
                  const double muEst,
                  const double sigmaEst,
                  const double tol,
                  const bool verbose = false)
{
    static_assert(std::is_same<T2, double>::value, "the reference instantiates T2 = double only (_pguresvt.pyx:181-358)");
    static_assert(std::is_same<T1, uint8_t>::value || std::is_same<T1, uint16_t>::value || std::is_same<T1, float>::value ||
                      std::is_same<T1, double>::value,
                  "T1 must be uint8_t, uint16_t, float or double");
    (void)verbose;
    Y.set_size(X.n_rows, X.n_cols, X.n_slices);
    Y.zeros();
    estimates.set_size(X.n_slices, 4);
    estimates.zeros();

    pguresvt_params p = {};
    p.traj_length = trajLength;
    p.block_size = blockSize;
    p.block_overlap = blockOverlap;
    p.motion_window = motionWindow;
    p.median_size = medianSize;
    p.noise_method = noiseMethod;
    p.max_iter = maxIter;
    p.n_jobs = nJobs;
    p.random_seed = randomSeed;
    p.optimize_pgure = optimizePGURE;
    p.exp_weighting = expWeighting;
    p.motion_estimation = motionEstimation;
    p.lambda_est = lambdaEst;
    p.alpha_est = alphaEst;
    p.mu_est = muEst;
    p.sigma_est = sigmaEst;
    p.tol = tol;
    // extensions: the call fans out over every visible GPU by default (n_gpus = 0), like nJobs = -1 over host threads
    if (const char *e = std::getenv("PGURESVT_DEVICE"))
        p.device = std::atoi(e);
    if (const char *e = std::getenv("PGURESVT_NGPUS"))
        p.n_gpus = std::atoi(e);

    const uint32_t nr = (uint32_t)X.n_rows, nc = (uint32_t)X.n_cols, nf = (uint32_t)X.n_slices;
    int rc;
    if (std::is_same<T1, uint8_t>::value)
        rc = pguresvt_run_u8((const uint8_t *)X.memptr(), nr, nc, nf, &p, Y.memptr(), estimates.memptr());
    else if (std::is_same<T1, uint16_t>::value)
        rc = pguresvt_run_u16((const uint16_t *)X.memptr(), nr, nc, nf, &p, Y.memptr(), estimates.memptr());
    else if (std::is_same<T1, float>::value)
        rc = pguresvt_run_f32((const float *)X.memptr(), nr, nc, nf, &p, Y.memptr(), estimates.memptr());
    else
        rc = pguresvt_run_f64((const double *)X.memptr(), nr, nc, nf, &p, Y.memptr(), estimates.memptr());
    return (uint32_t)rc;
}
#endif
