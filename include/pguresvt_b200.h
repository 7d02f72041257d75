/* pguresvt_b200.h — C ABI of the B200-native PGURE-SVT denoising hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point names
 * the reference interface it replaces (file:line relative to tjof2/pgure-svt v0.6.4).  The host-side
 * mirrors of the reference API (C++ template PGURESVT<T1,T2>, the _pguresvt bridge, the SVT class, the
 * PGURE-SVT CLI) all sit ABOVE this header and only marshal arguments; see INTEGRATION.md.
 *
 * Memory layout everywhere is Armadillo's: column-major, element (r, c, t) of an (n_rows, n_cols,
 * n_frames) sequence at r + n_rows * (c + n_cols * t).
 *
 * There is no CPU fallback: every call fails with PGS_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef PGURESVT_B200_H
#define PGURESVT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scalar arguments of PGURESVT<T1,T2>() in the reference's order (src/pguresvt.hpp:17-38). */
typedef struct pguresvt_params
{
    uint32_t traj_length;      /* trajLength        (default 15) */
    uint32_t block_size;       /* blockSize         (default 4)  */
    uint32_t block_overlap;    /* blockOverlap      (default 1)  */
    uint32_t motion_window;    /* motionWindow      (default 7)  */
    int64_t median_size;       /* medianSize: CTMF *radius*, <= 0 disables (pguresvt.hpp:69-88) */
    uint32_t noise_method;     /* noiseMethod 1..4  (noise.hpp:115-147) */
    uint32_t max_iter;         /* maxIter → nlopt maxeval (pgure.hpp:209) */
    int64_t n_jobs;            /* nJobs: accepted, results never depend on it (utils.hpp:108-168) */
    int64_t random_seed;       /* randomSeed; < 0 → std::random_device (pgure.hpp:52-59) */
    int32_t optimize_pgure;    /* optimizePGURE */
    int32_t exp_weighting;     /* expWeighting */
    int32_t motion_estimation; /* motionEstimation */
    double lambda_est;         /* lambdaEst: fixed λ, or start point if optimising (< 0 → mean(u)) */
    double alpha_est;          /* alphaEst  (< 0 → estimate) */
    double mu_est;             /* muEst     (< 0 → estimate) */
    double sigma_est;          /* sigmaEst  (< 0 → estimate) */
    double tol;                /* tol → nlopt ftol_rel */
    /* ---- extensions (zero = reference behaviour) ---- */
    int32_t device;            /* CUDA device ordinal */
    int32_t eps1_mode;         /* 0: as the reference computes it (eps1*delta1 integer-truncated to 0,
                                  pgure.hpp:80); 1: intended first-order perturbation */
    int32_t svd_kernel;        /* 0: auto (4-lane register Jacobi, tracked norms, fast scaled rotations); 1: generic shared-memory Jacobi; 2: 8-lane register Jacobi; 3: 4-lane without norm tracking; 4: 4-lane tracked norms, full rotations */
    int32_t rank_cache;        /* shapes other than 16x15 with optimize_pgure (eps1_mode 0): the factor cache keeps S and the
                                  q-forms of every singular triplet plus the leading rank_cache triplets of object U
                                  (DESIGN.md "truncated factor cache"); probes at which more triplets survive are answered
                                  exactly by re-decomposing those patches.  0: automatic (as many as half of the free HBM holds, up to all);
                                  < 0: keep the full U, S, V of every patch (svt.hpp:111-116) */
    int32_t n_gpus;            /* one-shot entry points only: how many CUDA devices the call fans out over — the role of nJobs /
                                  pguresvt::parallel inside PGURESVT() (pguresvt.hpp:169, utils.hpp:108-168): device, device+1, …
                                  each take one contiguous block of frames (plus halo) on their own host thread.
                                  0: automatic (every visible device from `device` on, at least 8 frames each); 1: `device` only */
} pguresvt_params;

enum
{
    PGS_OK = 0,
    PGS_ERR_ARG = 1,         /* invalid argument (non-square frames, too few frames, bad sizes …) */
    PGS_ERR_CUDA = 2,        /* CUDA runtime error / no usable device */
    PGS_ERR_UNSUPPORTED = 3, /* parameter combination not implemented on the GPU path */
    PGS_ERR_OPT = 4          /* optimiser could not start (e.g. zero initial step) */
};

enum
{
    PGS_U8 = 0,
    PGS_U16 = 1,
    PGS_F32 = 2,
    PGS_F64 = 3
};

/* ---------------------------------------------------------------------------------------------
 * One-shot entry points.  Replace the four instantiations PGURESVT<T1,double> bound by
 * pguresvt/_pguresvt.pyx:181,240,299,358 and PGURESVT<uint16_t,double> at src/PGURE-SVT.cpp:187.
 * X: host, (n_rows, n_cols, n_frames) column-major of the named type.
 * Y: host, same shape, double (pguresvt.hpp:44-45).  estimates: host, (n_frames, 4) column-major:
 * columns lambda, alpha, mu, sigma (pguresvt.hpp:47,150-153).  Returns PGS_OK (the reference always
 * returns 0, pguresvt.hpp:171) or an error code; never throws.
 * ------------------------------------------------------------------------------------------- */
int pguresvt_run_u8(const uint8_t *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                    const pguresvt_params *p, double *Y, double *estimates);
int pguresvt_run_u16(const uint16_t *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                     const pguresvt_params *p, double *Y, double *estimates);
int pguresvt_run_f32(const float *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                     const pguresvt_params *p, double *Y, double *estimates);
int pguresvt_run_f64(const double *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                     const pguresvt_params *p, double *Y, double *estimates);

/* The one-shot entry points keep one handle per device between calls (device buffers of the last frame size / parameter
 * set, page-locked staging): a later call with the same configuration re-targets it instead of re-allocating.  This frees
 * them.  PGURESVT_NO_CACHE=1 in the environment disables the cache. */
void pguresvt_release_cached(void);

/* Message of the last error on the calling thread ("" if none).  Replaces the C++ exceptions that escape
 * the reference's worker threads (SURVEY §5 "Failure detection"). */
const char *pguresvt_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Handle API: the same path split into upload / process / download so that (a) a contiguous block of
 * frames [frame_begin, frame_end) of a longer sequence can be processed per GPU exactly like one slice of
 * pguresvt::parallel (src/utils.hpp:150-166) and (b) throughput can be timed with inputs resident in HBM.
 * The window / edge rules always use the GLOBAL n_frames (pguresvt.hpp:100-114,155-166).
 * ------------------------------------------------------------------------------------------- */
typedef struct pguresvt_handle pguresvt_handle;

/* dtype: PGS_U8..PGS_F64.  The handle owns device storage for frames [frame_begin - fw, frame_end + fw)
 * clamped per the reference's first/last-window rule. */
pguresvt_handle *pguresvt_create(int dtype, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                                 const pguresvt_params *p, uint32_t frame_begin, uint32_t frame_end);
void pguresvt_destroy(pguresvt_handle *h);

/* First/last global frame index (half-open) the handle keeps on the device (block + halo frames). */
int pguresvt_resident_range(const pguresvt_handle *h, uint32_t *first, uint32_t *last);

/* Host→device copy of the resident frames.  X_full points at frame 0 of the WHOLE sequence. */
int pguresvt_upload(pguresvt_handle *h, const void *X_full);
/* Same, but the source is already a device pointer to frame `first` of the resident range
 * (n_rows*n_cols*(last-first) elements of the handle's dtype). */
int pguresvt_upload_device(pguresvt_handle *h, const void *dX_resident);

/* Re-target an existing handle at another block [frame_begin, frame_end) of the same sequence without re-allocating
 * (the block and its halo must not be longer than the ones the handle was created for).  The streaming loop of the
 * one-shot entry points uses this; upload must follow. */
int pguresvt_retarget(pguresvt_handle *h, uint32_t frame_begin, uint32_t frame_end);

/* Stream the denoised frames to the host while the block is processed: with Y_full set (whole-sequence array as in
 * pguresvt_download, pinned or pageable), pguresvt_process copies every frame out as soon as it is final — device→host on
 * a copy stream into a pinned ring, then (pageable targets) host→host on a helper thread — overlapped with the SVDs of the
 * following frames; all copies have landed when pguresvt_process returns.  NULL switches streaming off. */
int pguresvt_stream_output(pguresvt_handle *h, double *Y_full);

/* Run median prefilter + per-frame pipeline for frames [frame_begin, frame_end); results stay on device. */
int pguresvt_process(pguresvt_handle *h);

/* Device pointers to the block's results: Y_block (n_rows, n_cols, frame_end-frame_begin) doubles and
 * estimates_block (4, frame_end-frame_begin) doubles [row-per-quantity, i.e. the transposed layout that
 * one ncclAllGather over frames concatenates correctly]. */
double *pguresvt_device_output(pguresvt_handle *h);
double *pguresvt_device_estimates(pguresvt_handle *h);

/* Device→host copy.  Y_full / estimates_full are the WHOLE-sequence arrays (same shapes as the one-shot
 * call); only the handle's frames are written. */
int pguresvt_download(pguresvt_handle *h, double *Y_full, double *estimates_full);

/* Counters of the last pguresvt_process(): all doubles.
 *  [0] kernel launches   [1] patch SVDs computed   [2] PGURE objective evaluations
 *  [3] ms median   [4] ms ARPS   [5] ms SVD   [6] ms lambda search (reconstruct+risk)   [7] ms final reconstruct
 *  [8] ms noise estimation   [9] ms total (device timeline)   [10] SVD sweeps (sum over launches, max per launch)
 *  [11] bytes of SVD factors resident per frame   [12]/[13] mean Jacobi sweeps (object 0 / warm-started objects)
 *  [14]/[15] ARPS frame pairs computed / reused from the cross-window cache
 *  [16] singular triplets streamed by all evaluations   [17] ms search preparation (weights, multipliers, q-forms)
 *  [18] evaluations redone because a triplet beyond the lazily prepared q-forms survived
 *  [19] optimiser probes answered from the per-frame memo (a lambda already evaluated bit for bit)
 *  [20] patches re-decomposed because more than rank_cache triplets survived a probe (compact cache; [18] counts those probes)
 *  [21] leading triplets of object U kept per patch by the compact cache (0: full factor cache)
 *  [22] probes beyond the frame's critical lambda that ran the bound check of the lean (dominant-triplet) mode;
 *  [23] patch SVDs redone exactly because a bound survived the threshold there ([20] counts the patches) */
#define PGS_NSTATS 24
int pguresvt_get_stats(const pguresvt_handle *h, double *stats);

/* ---------------------------------------------------------------------------------------------
 * Stage probes (diagnostics for the parity tests; each runs the production kernels of one stage for
 * ONE global frame index t of an uploaded handle and copies the intermediate to the host).
 * ------------------------------------------------------------------------------------------- */
/* Median-filtered frame t as uint16 (pguresvt.hpp:73-81 / medfilter.hpp:478-539). */
int pguresvt_probe_median(pguresvt_handle *h, uint32_t t, uint16_t *Z);
/* ARPS trajectories of frame t's window (arps.hpp:52-134): patches int32 (2, (N-bs+1)^2, 2*fw+1)
 * column-major, [0]=row, [1]=col; optional n_cost (number of block-cost evaluations). */
int pguresvt_probe_arps(pguresvt_handle *h, uint32_t t, int32_t *patches);
/* Singular values (descending) of every patch of SVT object `obj` (0:U, 1:U1, 2:U2p, 3:U2m) for frame t:
 * S is (n_t, n_patches) column-major, n_t = 2*fw+1 (svt.hpp:58-118). */
int pguresvt_probe_singular_values(pguresvt_handle *h, uint32_t t, int obj, double *S, int64_t *n_patches);
/* PGURE objective (pgure.hpp:120-137) of frame t at each of n lambdas, with the noise parameters given
 * in estimate order (alpha, mu, sigma) — the sigma/mu swap of pguresvt.hpp:133 is applied inside.
 * values[n]; terms[5*n] = the five global sums (may be NULL). */
int pguresvt_probe_pgure(pguresvt_handle *h, uint32_t t, double alpha, double mu, double sigma, int n,
                         const double *lambdas, double *values, double *terms);
/* Full reconstructed window v (n_rows, n_cols, 2*fw+1) of frame t at lambda, BEFORE the *uMax rescale
 * (svt.hpp:121-167). */
int pguresvt_probe_reconstruct(pguresvt_handle *h, uint32_t t, double lambda, double *v);
/* Bernoulli perturbation signs (pgure.hpp:167-186): delta1 int8 in {-1,+1}; delta2neg int8 1 where the
 * negative branch -sqrt(vQ/vP) was drawn; n = n_rows*n_cols*(2*fw+1) each. */
int pguresvt_probe_perturbations(pguresvt_handle *h, int8_t *delta1, int8_t *delta2neg);
/* Noise estimate of frame t's window (noise.hpp:35-153): in/out alpha, mu, sigma (< 0 = estimate). */
int pguresvt_probe_noise(pguresvt_handle *h, uint32_t t, double *alpha, double *mu, double *sigma);

/* arma::accu(u) of frame t's max-normalised window — the start point of the lambda search times Nx*Ny*Nt
 * (pguresvt.hpp:139) and the second sum of the risk (pgure.hpp:136) — in Armadillo's order (two sequential
 * accumulators), bit for bit. */
int pguresvt_probe_window_sum(pguresvt_handle *h, uint32_t t, double *sum);

/* Hot-pixel prefilter (src/hotpixel.hpp:19-64, called from src/PGURE-SVT.cpp:171-179): in place on a host
 * uint16 sequence. */
int pguresvt_hotpixel_u16(uint16_t *seq, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, double threshold,
                          int device);

/* Host-side logic exposed for CPU-only tests (no device needed).
 * pguresvt_host_sbplx: the 1-D subplex driver used for the lambda search, i.e. the stand-in for the
 * reference's nlopt::opt(LN_SBPLX, 1).optimize() call (pgure.hpp:206-216); returns the NLopt-style status
 * (3 = ftol, 4 = xtol, 5 = maxeval, -2 = invalid arguments).
 * pguresvt_host_patch_ids: sorted unique patch-id set of SVT::Decompose (svt.hpp:61-97); returns its size. */
int pguresvt_host_sbplx(double (*f)(double, void *), void *data, double x0, double lb, double ub, double step,
                        double ftol_rel, double xtol_abs, int maxeval, double *xbest, double *fbest, int *nevals);
int64_t pguresvt_host_patch_ids(uint32_t N, uint32_t bs, uint32_t bo, int32_t *out, int64_t cap);

/* Number of devices a one-shot call with these parameters would use for an n_frames sequence (host logic of the fan-out;
 * n_visible < 0: ask the CUDA runtime).  pguresvt_host_frame_block: the contiguous block [*begin, *end) of rank `part` of
 * `parts` — the partition of pguresvt::parallel (utils.hpp:150-166). */
int pguresvt_host_plan_gpus(const pguresvt_params *p, uint32_t n_frames, int n_visible);
int pguresvt_host_frame_block(uint32_t n_frames, int parts, int part, uint32_t *begin, uint32_t *end);

/* out (n_rows, n_cols, n_frames) C-order = in (n_frames, n_cols, n_rows) C-order with the axes reversed: the copy
 * pguresvt/svt.py:329 makes of the bridge's result, cache-blocked over host threads (n_threads <= 0: automatic). */
int pguresvt_host_transpose_f64(const double *in, uint32_t n_frames, uint32_t n_cols, uint32_t n_rows, double *out, int n_threads);

/* Measured FP64 DFMA throughput of the device's vector pipe in TFLOP/s (burst = best single launch, sustained = ~1 s back
 * to back): the roofline denominator of the SVD kernels, taken in the same job as the bench (bench.py). */
int pguresvt_bench_dfma(int device, double *tflops_burst, double *tflops_sustained);

/* Library / device info: fills name (up to len bytes), returns SM count or -1. */
int pguresvt_device_info(int device, char *name, int len);

#ifdef __cplusplus
}
#endif
#endif /* PGURESVT_B200_H */
