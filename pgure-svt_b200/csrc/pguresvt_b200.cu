// C-ABI + host driver of the B200-native PGURE-SVT hot path (include/pguresvt_b200.h).
//
// The driver restates the control flow of PGURESVT<T1,T2>() (src/pguresvt.hpp:17-172): median prefilter of
// every frame, then per output frame: window normalisation, (noise estimation), ARPS trajectories, the patch
// SVDs of up to four SVT objects cached in HBM, the lambda search where every objective evaluation is
// threshold + reconstruct + aggregate + risk reduction on the device, the final reconstruction and the copy
// of the reference slice.  All array work is CUDA; the host only sequences kernels and runs the scalar 1-D
// optimiser.  There is no CPU fallback: without a CUDA device every entry point fails.
#include "../../include/pguresvt_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "compact.cuh"
#include "tile_eval.cuh"
#include "noise.cuh"
#include "sbplx1d.hpp"

using namespace pgs;

// ------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
extern "C" const char *pguresvt_last_error(void) { return g_err.c_str(); }

#define CU(call)                                                                                          \
    do                                                                                                    \
    {                                                                                                     \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(PGS_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------------
struct pguresvt_handle
{
    pguresvt_params p{};
    int dtype = 0;
    uint32_t N = 0, nframes = 0, fb = 0, fe = 0; // square frames N x N; block [fb, fe)
    uint32_t Nt = 0, fw = 0, win = 0;            // pguresvt.hpp:57-58; window has 2*fw+1 slices
    uint32_t r0 = 0, r1 = 0;                     // resident frames [r0, r1)
    uint32_t cap_res = 0, cap_blk = 0;           // capacity the device buffers were allocated for (pguresvt_retarget)
    size_t fsz = 0, esz = 0;
    int m = 0, n = 0, ldv = 0, M1 = 0, vecSize = 0, P = 0;
    size_t rec = 0;
    int nobj = 1;
    int objs[4] = {0, 0, 0, 0}; // SVT objects present: 0:U 1:U1 2:U2p 3:U2m
    bool use_reg_svd = false; // any register-resident 16x15 kernel
    bool use_l4 = false;      // 4-lanes-per-matrix kernel (S in slot order, S[15] = sigma_max)
    bool use_fused_eval = false;
    bool lean = false;           // fused path: perturbed objects leave only head entries behind (k_svd16_l4 EPI 1), no q-form pass
    bool frame_full = false;     // lean mode: the current frame has been decomposed in full after all (a third triplet survived)
    bool eps1f = false;          // eps1_mode 1 through the fused evaluation (object U and U1 exact, U2p / U2m lean)
    double *dPartial3 = nullptr; // eps1f: per-block partials of s3
    bool acc1_clean = false;
    bool top1_all = false;       // top1 for object U as well (bound from the Gram matrix): no Jacobi launch in the common case
    bool top1 = false;           // lean mode: perturbed objects by k_top1_l4 (dominant triplet + rigorous bound), exact on demand
    double *dUp3 = nullptr;      // top1: perturbed window of object U2m (dUp keeps U2p's)
    int *dLeanList = nullptr;    // top1: [0] count, [1..] patches whose bound survives at the current probe
    double *dCrit = nullptr;     // top1: {critical lambda (exp weighting), largest bound (plain)} of the current frame
    double crit_lambda = 0.0, crit_bound = 0.0;
    double *hLean = nullptr;     // pinned: critical values / list count
    bool full_mode = false;      // stage_svd: write full records for the perturbed objects (fallback of the lean mode, probes)
    bool use_tile = false;       // atomics-free gather evaluation (tile_eval.cuh) on top of the fused path
    bool frame_fallback = false; // a third triplet survived at some probe of this frame: general k_eval3 path from there on
    int *dBinCnt = nullptr, *dBinStart = nullptr, *dScanSums = nullptr;
    TgEntry *dEnt = nullptr;
    double *dHead = nullptr, *dTilePart = nullptr, *dU0c = nullptr;
    double2 *dFth = nullptr;
    int tile_r = 0, tile_c = 0;
    // arma::accu(u) bit for bit (k_accu_seq) on its own stream beside the ARPS / SVD stages
    cudaStream_t sum_st = nullptr;
    cudaEvent_t evU = nullptr, evSum = nullptr;
    double *dSum2 = nullptr, *hSum2 = nullptr;
    bool sum_pending = false;
    bool use_warp_svd = false; // 64 x n Casorati matrices: warp-per-matrix register Jacobi (compact.cuh)
    bool use_warp3 = false;    // compact cache + warp kernel: the three PGURE objects of a patch in one launch, warm-started
    bool use_compact = false;  // truncated factor cache (S, q-forms, leading Rc triplets of object 0), compact.cuh
    int Rc = 0;                // leading singular triplets of object 0 kept per patch
    int evc_warps = 0;         // warps of the k_eval_c grid (one s4 partial each)
    int eval_blocks = 0;
    int sm_count = 148;
    double vP = 0, d2Neg = 0, d2Pos = 0;
    int64_t seed_used = 0;

    // device
    void *dX = nullptr;
    uint16_t *dZ = nullptr, *dTmp16 = nullptr;
    double *dU = nullptr, *dW = nullptr, *dUp = nullptr; // dUp: window of the perturbed object being decomposed
    short2 *dPos = nullptr, *dArpsF = nullptr, *dArpsB = nullptr;
    struct ArpsTag
    {
        long long frame = -1;
        double wmax = 0;
    };
    std::vector<ArpsTag> tagF, tagB;
    SliceCopies pend{}; // trajectory slice copies waiting for one k_copy_slices launch
    int arps_ring = 0;
    int *dIds = nullptr;
    unsigned *dCnt = nullptr;
    double *dAccScale = nullptr; // {2^e, 2^-e} of the fixed-point overlap-add accumulators (k_acc_scale), + the largest weight
    double *dAcc[4] = {nullptr, nullptr, nullptr, nullptr};
    double *dFac[4] = {nullptr, nullptr, nullptr, nullptr};
    int8_t *dD1 = nullptr, *dD2 = nullptr;
    double *dPartial = nullptr, *dOut = nullptr, *dMaxPartial = nullptr, *dC4 = nullptr, *dPartialE = nullptr, *dQ[3] = {nullptr, nullptr, nullptr};
    double *dY = nullptr, *dEst = nullptr, *dV = nullptr;
    double *dSc[4] = {nullptr, nullptr, nullptr, nullptr}, *dQc[4] = {nullptr, nullptr, nullptr, nullptr}; // compact cache: S, q per object
    double *dLead = nullptr;       // compact cache: leading triplets of object 0
    int *dOvf = nullptr;           // [0] number of patches whose survivors exceed Rc at the current probe, [1..] their macroblock ids
    double *dFacScratch = nullptr; // full records of one chunk of overflow patches (allocated on first use)
    int scratch_patches = 0;
    int *hOvf = nullptr;           // pinned
    bool attr_svdw = false;
    int *dSweeps = nullptr;
    unsigned long long *dNcost = nullptr;
    NoiseWorkspace noise_ws;
    double *hOut = nullptr; // pinned
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // noise estimation of the current window runs beside the ARPS / SVD stages (it only feeds the lambda search)
    cudaStream_t noise_st = nullptr;
    cudaEvent_t evWin = nullptr;
    std::thread noise_thread;
    bool noise_req = false, noise_started = false;
    // the estimate of frame t+1's window is started while frame t's lambda search runs (L2-bound, leaves the SMs room):
    // dUn holds that window, noise_pre_t the frame it belongs to (-1: none outstanding)
    double *dUn = nullptr;
    long long noise_pre_t = -1;
    int noise_rc = 0;
    double noise_val[3] = {-1., -1., -1.};
    long long noise_launches = 0;
    std::string noise_err;

    std::vector<double> xmax, zmax, xmin; // per resident frame
    std::vector<double> est;        // (fe-fb) x 4 row-per-quantity
    bool uploaded = false, prefiltered = false, perturbed = false, attr_warm = false, attr_qform = false, acc0_clean = false;
    int q_k = 0;           // number of leading triplets whose q-forms exist for the current frame
    int *dNeedQ = nullptr; // set by k_eval3 when a triplet beyond q_k survives
    int *dKpart = nullptr; // per-warp count of singular triplets streamed by k_eval3
    long long cur_t = -1;
    double cur_uMax = 0, cur_wMax = 0, cur_sumU = 0, cur_absmax = 1.0; // cur_absmax: max |u| of the normalised window
    int cur_ref = 0, cur_sl = 0, cur_a = 0;
    double stats[PGS_NSTATS] = {0};
    long long launches = 0;

    // stage timers: event pairs recorded on the stream and resolved once at the end of pguresvt_process (no mid-stream syncs)
    struct TimerRec
    {
        int slot;
        cudaEvent_t a, b;
    };
    std::vector<cudaEvent_t> evpool;
    size_t evused = 0;
    std::vector<TimerRec> trecs;

    // output streaming (pguresvt_stream_output): frames leave on copy_st as soon as they are final
    double *sinkY = nullptr;
    bool sink_pinned = false;
    cudaStream_t copy_st = nullptr;
    cudaEvent_t evFrame = nullptr;
    static constexpr int RING = 4;
    double *ring[RING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ring_done[RING] = {nullptr, nullptr, nullptr, nullptr};
    bool ring_busy[RING] = {false, false, false, false};
    struct CopyJob
    {
        int slot;
        uint32_t t;
    };
    std::deque<CopyJob> jobs;
    int jobs_inflight = 0;
    bool copier_stop = false;
    std::mutex mx;
    std::condition_variable cv;
    std::thread copier;
    uint64_t frames_streamed = 0;
    // page-locked staging of the one-shot entry's input sub-blocks (kept with the handle so that a cached handle keeps them)
    void *stg[2] = {nullptr, nullptr};
    size_t stg_bytes = 0;
};

#define RISK_BLOCKS 1184
#define EVAL_PPG 8 /* patches each 16-lane group of k_eval3 walks (software-pipelined) */
#define LAUNCHED(h) ((h)->launches++)

static size_t dtype_size(int dt) { return dt == PGS_U8 ? 1 : dt == PGS_U16 ? 2 : dt == PGS_F32 ? 4 : 8; }

// sorted unique patch ids of SVT::Decompose (svt.hpp:61-97)
static std::vector<int> patch_ids(int N, int bs, int bo)
{
    const long long M = N - bs;
    std::vector<long long> ids;
    for (long long i = 0; i < 1 + M; i += bo)
        for (long long j = 0; j < 1 + M; j += bo)
            ids.push_back(i * M + j);
    for (long long i = 0; i < 1 + M; i += bo)
        ids.push_back((M + 1) * i + M);
    for (long long i = 0; i < 1 + M; i += bo)
        ids.push_back((M + 1) * M + i);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    return std::vector<int>(ids.begin(), ids.end());
}

static uint32_t window_start(const pguresvt_handle *h, uint32_t t) // pguresvt.hpp:100-114
{
    if (t < h->fw)
        return 0;
    if (t >= h->nframes - h->fw)
        return h->nframes - 2 * h->fw - 1;
    return t - h->fw;
}

static void free_all(pguresvt_handle *h)
{
    auto F = [](void *p) {
        if (p)
            cudaFree(p);
    };
    // helper threads first: they use the streams and buffers released below
    if (h->noise_thread.joinable())
        h->noise_thread.join();
    if (h->copier.joinable())
    {
        {
            std::lock_guard<std::mutex> lk(h->mx);
            h->copier_stop = true;
        }
        h->cv.notify_all();
        h->copier.join();
    }
    if (h->st)
        cudaStreamSynchronize(h->st);
    if (h->noise_st)
        cudaStreamSynchronize(h->noise_st);
    if (h->copy_st)
    {
        cudaStreamSynchronize(h->copy_st);
        cudaStreamDestroy(h->copy_st);
    }
    for (int i = 0; i < pguresvt_handle::RING; i++)
    {
        if (h->ring[i])
            cudaFreeHost(h->ring[i]);
        if (h->ring_done[i])
            cudaEventDestroy(h->ring_done[i]);
    }
    if (h->evFrame)
        cudaEventDestroy(h->evFrame);
    for (int i = 0; i < 2; i++)
        if (h->stg[i])
            cudaFreeHost(h->stg[i]);
    for (cudaEvent_t e : h->evpool)
        cudaEventDestroy(e);
    F(h->dX), F(h->dZ), F(h->dTmp16), F(h->dU), F(h->dUp), F(h->dW), F(h->dPos), F(h->dArpsF), F(h->dArpsB), F(h->dIds), F(h->dCnt);
    for (int i = 0; i < 4; i++)
        F(h->dAcc[i]), F(h->dFac[i]);
    for (int i = 0; i < 4; i++)
        F(h->dSc[i]), F(h->dQc[i]);
    F(h->dLead), F(h->dOvf), F(h->dFacScratch), F(h->dUn), F(h->dAccScale);
    if (h->sum_st)
    {
        cudaStreamSynchronize(h->sum_st);
        cudaStreamDestroy(h->sum_st);
    }
    if (h->evU)
        cudaEventDestroy(h->evU);
    if (h->evSum)
        cudaEventDestroy(h->evSum);
    if (h->hSum2)
        cudaFreeHost(h->hSum2);
    F(h->dSum2);
    F(h->dUp3), F(h->dLeanList), F(h->dCrit), F(h->dPartial3);
    if (h->hLean)
        cudaFreeHost(h->hLean);
    F(h->dBinCnt), F(h->dBinStart), F(h->dScanSums), F(h->dEnt), F(h->dHead), F(h->dTilePart), F(h->dFth), F(h->dU0c);
    if (h->hOvf)
        cudaFreeHost(h->hOvf);
    F(h->dD1), F(h->dD2), F(h->dC4), F(h->dPartialE), F(h->dKpart), F(h->dQ[0]), F(h->dQ[1]), F(h->dQ[2]), F(h->dPartial), F(h->dOut), F(h->dMaxPartial), F(h->dY), F(h->dEst), F(h->dV), F(h->dSweeps),
        F(h->dNcost);
    h->noise_ws.release();
    if (h->noise_st)
        cudaStreamDestroy(h->noise_st);
    if (h->evWin)
        cudaEventDestroy(h->evWin);
    if (h->hOut)
        cudaFreeHost(h->hOut);
    if (h->st)
        cudaStreamDestroy(h->st);
    for (int i = 0; i < 2; i++)
        if (h->ev[i])
            cudaEventDestroy(h->ev[i]);
}

extern "C" void pguresvt_destroy(pguresvt_handle *h)
{
    if (!h)
        return;
    cudaSetDevice(h->p.device);
    free_all(h);
    delete h;
}

static int create_impl(pguresvt_handle *h)
{
    const pguresvt_params &p = h->p;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PGS_ERR_CUDA, "no CUDA device available (the PGURE-SVT hot path has no CPU fallback)");
    if (p.device < 0 || p.device >= ndev)
        return fail(PGS_ERR_ARG, "device %d out of range (%d devices)", p.device, ndev);
    CU(cudaSetDevice(p.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, p.device));
    h->sm_count = prop.multiProcessorCount;

    const uint32_t bs = p.block_size;
    if (bs < 1 || bs > h->N)
        return fail(PGS_ERR_ARG, "patch_size %u invalid for %ux%u frames", bs, h->N, h->N);
    if (p.block_overlap < 1)
        return fail(PGS_ERR_ARG, "patch_overlap must be >= 1");
    h->Nt = (bs * bs < p.traj_length) ? bs * bs - 1 : p.traj_length;
    h->fw = h->Nt / 2;
    h->win = 2 * h->fw + 1;
    if (h->Nt < 1)
        return fail(PGS_ERR_ARG, "trajectory length / patch size give an empty window");
    if (h->nframes < h->win)
        return fail(PGS_ERR_ARG, "sequence has %u frames, fewer than the %u-frame window", h->nframes, h->win);
    if (h->fb >= h->fe || h->fe > h->nframes)
        return fail(PGS_ERR_ARG, "invalid frame block [%u, %u) of %u", h->fb, h->fe, h->nframes);
    if (p.motion_estimation && (p.motion_window > ARPS_MAX_MW))
        return fail(PGS_ERR_UNSUPPORTED, "motion_window %u > %d not supported by the ARPS kernel", p.motion_window,
                    ARPS_MAX_MW);
    if (p.median_size > 60)
        return fail(PGS_ERR_UNSUPPORTED, "median radius %lld > 60 not supported", (long long)p.median_size);
    if (h->win > 32)
        return fail(PGS_ERR_UNSUPPORTED, "window of %u slices > 32 not supported", h->win);
    if (h->N > 32767)
        return fail(PGS_ERR_UNSUPPORTED, "frames larger than 32767 px not supported");
    h->m = bs * bs;
    h->n = h->win;
    h->ldv = (h->n + 1) & ~1;
    h->rec = (size_t)h->m * h->n + (size_t)h->ldv * h->n + h->ldv;
    h->M1 = h->N - bs + 1;
    h->vecSize = h->M1 * h->M1;
    std::vector<int> ids = patch_ids(h->N, bs, p.block_overlap);
    h->P = (int)ids.size();
    h->fsz = (size_t)h->N * h->N;
    h->esz = dtype_size(h->dtype);
    h->r0 = window_start(h, h->fb);
    h->r1 = window_start(h, h->fe - 1) + h->win;
    h->cap_blk = h->fe - h->fb;
    // any block of cap_blk frames fits: its halo is at most win - 1 frames
    h->cap_res = std::min<uint32_t>(h->nframes, h->cap_blk + h->win - 1);
    h->nobj = 0;
    h->objs[h->nobj++] = 0;
    if (p.optimize_pgure)
    {
        if (p.eps1_mode == 1)
            h->objs[h->nobj++] = 1;
        h->objs[h->nobj++] = 2;
        h->objs[h->nobj++] = 3;
    }
    h->use_reg_svd = (h->m == 16 && h->n == 15 && p.svd_kernel != 1);
    if (p.svd_kernel >= 2 && !h->use_reg_svd)
        return fail(PGS_ERR_UNSUPPORTED, "register SVD kernels only cover 16x15 Casorati matrices");
    h->use_l4 = h->use_reg_svd && p.svd_kernel != 2; // 0 / 3: 4-lane kernel with tracked / recomputed pair norms
    // eps1_mode 1 (4 SVT objects): fused evaluation too, with object U1 decomposed exactly and overlap-added by an accumulate-only
    // pass (k_eval3<..., 2>) and the first-order sum s3 taken in the voxel pass; PGURESVT_EPS1_FUSED=0 keeps the generic path
    h->eps1f = h->use_l4 && p.optimize_pgure && p.eps1_mode == 1 && !(getenv("PGURESVT_EPS1_FUSED") && atoi(getenv("PGURESVT_EPS1_FUSED")) == 0);
    h->use_fused_eval = h->use_l4 && p.optimize_pgure && (p.eps1_mode == 0 || h->eps1f);
    h->lean = h->use_fused_eval && !(getenv("PGURESVT_LEAN") && atoi(getenv("PGURESVT_LEAN")) == 0);
    h->top1 = h->lean && !(getenv("PGURESVT_TOP1") && atoi(getenv("PGURESVT_TOP1")) == 0);
    h->top1_all = h->top1 && !h->eps1f && !(getenv("PGURESVT_TOP1_ALL") && atoi(getenv("PGURESVT_TOP1_ALL")) == 0);
    h->use_warp_svd = (h->m == 64 && h->n <= 32 && p.svd_kernel != 1);
    // rank_cache: 0 = automatic, > 0 = that many leading triplets, < 0 = keep the full factor cache (generic path)
    h->use_compact = !h->use_l4 && p.optimize_pgure && p.eps1_mode == 0 && p.rank_cache >= 0;
    h->use_warp3 = h->use_compact && h->use_warp_svd && !getenv("PGURESVT_NO_WARP3");
    {
        const double kappa = 1.;
        h->vP = 0.5 + 0.5 * kappa / std::sqrt(kappa * kappa + 4);
        const double vQ = 1 - h->vP;
        h->d2Neg = -1 * std::sqrt(vQ / h->vP);
        h->d2Pos = std::sqrt(h->vP / vQ);
    }

    const size_t wtot = h->fsz * h->win;
    const uint32_t nres = h->cap_res, nblk = h->cap_blk;
    if (h->use_l4 && p.optimize_pgure && p.rank_cache >= 0)
    { // the full factor cache of the register-SVD path is 3,968 B per patch and object: 12 GB at 1024^2, 200 GB at 4096^2.
      // Where it cannot fit, 16x15 goes through the truncated cache (S, q-forms, leading triplets; exact overflow path) with
      // the generic shared-memory SVD — slower, but it runs (default-parameter sequences at 4096^2).
        size_t freeb = 0, totb = 0;
        CU(cudaMemGetInfo(&freeb, &totb));
        const size_t need = h->rec * (size_t)h->P * sizeof(double) * (h->lean ? (h->eps1f ? 2 : 1) : h->nobj) + wtot * 60 + h->fsz * ((size_t)nres * (h->esz + 2) + (size_t)nblk * 8);
        if (need > freeb - freeb / 10)
        {
            h->use_l4 = h->use_reg_svd = h->use_fused_eval = h->lean = h->eps1f = false;
            h->use_compact = p.eps1_mode == 0;
            if (!h->use_compact)
                return fail(PGS_ERR_UNSUPPORTED, "the SVD factors of %d patches x %d objects (%.1f GB) do not fit on device %d (%.1f GB free)", h->P,
                            h->nobj, need / 1e9, p.device, freeb / 1e9);
        }
    }
    CU(cudaStreamCreate(&h->st));
    CU(cudaEventCreate(&h->ev[0]));
    CU(cudaEventCreate(&h->ev[1]));
    CU(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->evFrame, cudaEventDisableTiming));
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&h->noise_st, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&h->evWin, cudaEventDisableTiming));
        CU(cudaStreamCreateWithPriority(&h->sum_st, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&h->evU, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->evSum, cudaEventDisableTiming));
        CU(cudaMalloc(&h->dSum2, 2 * sizeof(double)));
        CU(cudaMallocHost(&h->hSum2, 2 * sizeof(double)));
    }
    CU(cudaMalloc(&h->dX, h->fsz * nres * h->esz));
    if (p.median_size > 0)
    {
        CU(cudaMalloc(&h->dZ, h->fsz * nres * sizeof(uint16_t)));
        if (h->dtype != PGS_U16)
            CU(cudaMalloc(&h->dTmp16, h->fsz * nres * sizeof(uint16_t)));
    }
    CU(cudaMalloc(&h->dU, wtot * sizeof(double)));
    if (p.motion_estimation)
        CU(cudaMalloc(&h->dW, wtot * sizeof(double)));
    CU(cudaMalloc(&h->dPos, (size_t)h->win * h->vecSize * sizeof(short2)));
    if (p.motion_estimation)
    {
        h->arps_ring = (int)h->win + 2;
        CU(cudaMalloc(&h->dArpsF, (size_t)h->arps_ring * h->vecSize * sizeof(short2)));
        CU(cudaMalloc(&h->dArpsB, (size_t)h->arps_ring * h->vecSize * sizeof(short2)));
        h->tagF.assign(h->arps_ring, pguresvt_handle::ArpsTag());
        h->tagB.assign(h->arps_ring, pguresvt_handle::ArpsTag());
    }
    CU(cudaMalloc(&h->dIds, (size_t)h->P * sizeof(int)));
    CU(cudaMemcpy(h->dIds, ids.data(), (size_t)h->P * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->dCnt, wtot * sizeof(unsigned)));
    CU(cudaMalloc(&h->dAccScale, 4 * sizeof(double)));
    CU(cudaMemset(h->dAccScale, 0, 4 * sizeof(double)));
    for (int k = 0; k < h->nobj; k++)
    {
        const int o = h->objs[k];
        if (o == 0 || (o == 1 && h->eps1f) || !(h->use_fused_eval || h->use_compact))
            CU(cudaMalloc(&h->dAcc[o], wtot * sizeof(double)));
        if (h->use_compact)
        {
            CU(cudaMalloc(&h->dSc[o], (size_t)32 * h->P * sizeof(double)));
            CU(cudaMalloc(&h->dQc[o], (size_t)32 * h->P * sizeof(double)));
            CU(cudaMemset(h->dSc[o], 0, (size_t)32 * h->P * sizeof(double)));
            CU(cudaMemset(h->dQc[o], 0, (size_t)32 * h->P * sizeof(double)));
            continue;
        }
        if (h->lean && o != 0 && !(o == 1 && h->eps1f))
            continue; // perturbed objects leave head entries only; full records on demand (stage_svd)
        CU(cudaMalloc(&h->dFac[o], h->rec * (size_t)h->P * sizeof(double)));
        CU(cudaMemset(h->dFac[o], 0, h->rec * (size_t)h->P * sizeof(double)));
    }
    if (p.optimize_pgure)
    {
        if (h->use_l4)
            CU(cudaMalloc(&h->dUp, wtot * sizeof(double)));
        CU(cudaMalloc(&h->dD1, wtot));
        CU(cudaMalloc(&h->dD2, wtot));
    }
    if (h->use_fused_eval)
    {
        h->eval_blocks = cdiv((long long)h->P * 16, 128);
        if (wtot >= ((size_t)1 << 31))
            return fail(PGS_ERR_UNSUPPORTED, "window of %zu voxels exceeds the fused evaluation kernel's 32-bit indexing", wtot);
        CU(cudaMalloc(&h->dC4, wtot * sizeof(double)));
        CU(cudaMalloc(&h->dPartialE, ((size_t)4 * h->eval_blocks + 64) * sizeof(double)));
        CU(cudaMalloc(&h->dKpart, ((size_t)4 * h->eval_blocks + 64) * sizeof(int)));

        if (!h->lean)
            for (int k = 0; k < 3; k++)
                CU(cudaMalloc(&h->dQ[k], (size_t)16 * h->P * sizeof(double)));
        CU(cudaMalloc(&h->dHead, (size_t)TG_HEAD * h->P * sizeof(double)));
        CU(cudaMemset(h->dHead, 0, (size_t)TG_HEAD * h->P * sizeof(double)));
        if (h->eps1f)
            CU(cudaMalloc(&h->dPartial3, (size_t)RISK_BLOCKS * sizeof(double)));
        if (h->top1)
        {
            CU(cudaMalloc(&h->dUp3, wtot * sizeof(double)));
            CU(cudaMalloc(&h->dLeanList, ((size_t)h->P + 1) * sizeof(int)));
            CU(cudaMalloc(&h->dCrit, 2 * sizeof(double)));
            CU(cudaMallocHost(&h->hLean, 4 * sizeof(double)));
        }
        // the atomics-free gather evaluation (tile_eval.cuh) is exact and deterministic too, but on B200 its irregular gather costs
        // more instructions than the L2 atomic unit costs time: 1.0 ms against 0.55 ms per evaluation at 1024^2 (profiles/r02) —
        // opt-in with PGURESVT_TILE_EVAL=1
        h->use_tile = !h->eps1f && getenv("PGURESVT_TILE_EVAL") && atoi(getenv("PGURESVT_TILE_EVAL")) > 0;
        if (h->use_tile)
        {
            const size_t nbins = h->fsz * h->win;
            h->tile_r = cdiv(h->N, TG_VR), h->tile_c = cdiv(h->N, TG_VC);
            CU(cudaMalloc(&h->dBinCnt, nbins * sizeof(int)));
            CU(cudaMalloc(&h->dBinStart, (nbins + 1) * sizeof(int)));
            CU(cudaMalloc(&h->dScanSums, ((size_t)cdiv(nbins, SCAN_ITEMS) + 1) * sizeof(int)));
            CU(cudaMalloc(&h->dEnt, (size_t)h->P * h->win * sizeof(TgEntry)));
            CU(cudaMalloc(&h->dU0c, (size_t)16 * h->P * sizeof(double)));
            CU(cudaMalloc(&h->dFth, (size_t)h->P * sizeof(double2)));
            CU(cudaMalloc(&h->dTilePart, (size_t)2 * h->tile_r * h->tile_c * h->win * sizeof(double)));
        }
    }
    CU(cudaMalloc(&h->dPartial, (size_t)RISK_BLOCKS * 8 * sizeof(double)));
    CU(cudaMalloc(&h->dOut, 16 * sizeof(double)));
    CU(cudaMemset(h->dOut, 0, 16 * sizeof(double)));
    h->dNeedQ = reinterpret_cast<int *>(h->dOut + 4); // travels home with the four sums of objective_fused
    CU(cudaMalloc(&h->dMaxPartial, (size_t)nres * 64 * 2 * sizeof(double)));
    CU(cudaMalloc(&h->dY, h->fsz * nblk * sizeof(double)));
    CU(cudaMalloc(&h->dEst, (size_t)4 * nblk * sizeof(double)));
    CU(cudaMalloc(&h->dSweeps, 4 * sizeof(int)));
    CU(cudaMalloc(&h->dNcost, sizeof(unsigned long long)));
    CU(cudaMallocHost(&h->hOut, 16 * sizeof(double)));
    if (h->use_compact)
    {
        h->evc_warps = 4 * std::min(cdiv(h->P, 4), h->sm_count * 16);
        CU(cudaMalloc(&h->dC4, wtot * sizeof(double)));
        CU(cudaMalloc(&h->dPartialE, (size_t)h->evc_warps * sizeof(double)));
        CU(cudaMalloc(&h->dKpart, (size_t)h->evc_warps * sizeof(int)));
        CU(cudaMalloc(&h->dOvf, ((size_t)h->P + 1) * sizeof(int)));
        CU(cudaMemset(h->dOvf, 0, sizeof(int)));
        CU(cudaMallocHost(&h->hOvf, sizeof(int)));
        // the leading-triplet cache is sized last: explicit rank_cache, else as many ranks (up to all n) as half of the
        // memory still free holds — the other half stays for the overflow scratch, the noise workspace and the caller
        const size_t per_rank = (size_t)h->P * (h->m + 32) * sizeof(double);
        int R = p.rank_cache;
        if (R == 0)
        {
            size_t freeb = 0, totb = 0;
            CU(cudaMemGetInfo(&freeb, &totb));
            R = (int)std::min<size_t>(EVC_RMAX, (freeb / 2) / per_rank);
        }
        h->Rc = std::max(1, std::min(std::min(R, EVC_RMAX), h->n));
        CU(cudaMalloc(&h->dLead, per_rank * h->Rc));
    }
    h->xmax.assign(nres, 0.0);
    h->xmin.assign(nres, 0.0);
    h->zmax.assign(nres, 0.0);
    h->est.assign((size_t)4 * nblk, 0.0);
    return PGS_OK;
}

static pguresvt_handle *create_handle(int dtype, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, const pguresvt_params *p,
                                      uint32_t frame_begin, uint32_t frame_end, int *rc_out)
{
    g_err.clear();
    int dummy;
    int &rc = rc_out ? *rc_out : dummy;
    if (!p)
    {
        rc = fail(PGS_ERR_ARG, "params is NULL");
        return nullptr;
    }
    if (dtype < PGS_U8 || dtype > PGS_F64)
    {
        rc = fail(PGS_ERR_ARG, "unknown dtype %d", dtype);
        return nullptr;
    }
    if (n_rows != n_cols)
    { // the reference assumes square frames throughout; the CLI rejects others
        rc = fail(PGS_ERR_ARG, "frame dimensions are not square, got %ux%u", n_cols, n_rows);
        return nullptr;
    }
    pguresvt_handle *h = new pguresvt_handle();
    h->p = *p;
    h->dtype = dtype;
    h->N = n_rows;
    h->nframes = n_frames;
    h->fb = frame_begin;
    h->fe = frame_end;
    if ((rc = create_impl(h)) != PGS_OK)
    {
        const std::string keep = g_err;
        free_all(h);
        delete h;
        g_err = keep;
        return nullptr;
    }
    return h;
}

extern "C" pguresvt_handle *pguresvt_create(int dtype, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames,
                                            const pguresvt_params *p, uint32_t frame_begin, uint32_t frame_end)
{
    return create_handle(dtype, n_rows, n_cols, n_frames, p, frame_begin, frame_end, nullptr);
}

extern "C" int pguresvt_resident_range(const pguresvt_handle *h, uint32_t *first, uint32_t *last)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    *first = h->r0;
    *last = h->r1;
    return PGS_OK;
}

static void drop_prelaunched_noise(pguresvt_handle *h);
static int invalidate(pguresvt_handle *h);

// n_frames may change too (a cached handle serving another sequence of the same frame size and parameters)
static int retarget_impl(pguresvt_handle *h, uint32_t n_frames, uint32_t frame_begin, uint32_t frame_end)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    if (n_frames < h->win)
        return fail(PGS_ERR_ARG, "sequence has %u frames, fewer than the %u-frame window", n_frames, h->win);
    if (frame_begin >= frame_end || frame_end > n_frames)
        return fail(PGS_ERR_ARG, "invalid frame block [%u, %u) of %u", frame_begin, frame_end, n_frames);
    const uint32_t keep_n = h->nframes;
    h->nframes = n_frames;
    const uint32_t r0 = window_start(h, frame_begin), r1 = window_start(h, frame_end - 1) + h->win;
    if (frame_end - frame_begin > h->cap_blk || r1 - r0 > h->cap_res)
    {
        h->nframes = keep_n;
        return fail(PGS_ERR_ARG, "block [%u, %u) (%u resident frames) exceeds the handle's capacity of %u frames (%u resident)", frame_begin,
                    frame_end, r1 - r0, h->cap_blk, h->cap_res);
    }
    CU(cudaSetDevice(h->p.device));
    CU(cudaStreamSynchronize(h->st));
    h->fb = frame_begin, h->fe = frame_end, h->r0 = r0, h->r1 = r1;
    h->xmax.assign(r1 - r0, 0.0);
    h->xmin.assign(r1 - r0, 0.0);
    h->zmax.assign(r1 - r0, 0.0);
    h->est.assign((size_t)4 * (frame_end - frame_begin), 0.0);
    invalidate(h);
    h->uploaded = false;
    return PGS_OK;
}

extern "C" int pguresvt_retarget(pguresvt_handle *h, uint32_t frame_begin, uint32_t frame_end)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    return retarget_impl(h, h->nframes, frame_begin, frame_end);
}

// helper thread of the output streaming: waits for a ring slot's device->host copy and moves it into the caller's array
static void copier_main(pguresvt_handle *h)
{
    cudaSetDevice(h->p.device);
    for (;;)
    {
        pguresvt_handle::CopyJob j;
        {
            std::unique_lock<std::mutex> lk(h->mx);
            h->cv.wait(lk, [h] { return h->copier_stop || !h->jobs.empty(); });
            if (h->jobs.empty())
                return;
            j = h->jobs.front();
            h->jobs.pop_front();
        }
        cudaEventSynchronize(h->ring_done[j.slot]);
        memcpy(h->sinkY + h->fsz * j.t, h->ring[j.slot], h->fsz * sizeof(double));
        {
            std::lock_guard<std::mutex> lk(h->mx);
            h->ring_busy[j.slot] = false;
            h->jobs_inflight--;
        }
        h->cv.notify_all();
    }
}

extern "C" int pguresvt_stream_output(pguresvt_handle *h, double *Y_full)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->p.device));
    h->sinkY = Y_full;
    h->sink_pinned = false;
    if (!Y_full)
        return PGS_OK;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, Y_full) == cudaSuccess && at.type == cudaMemoryTypeHost)
        h->sink_pinned = true; // page-locked target: frames are copied straight into it
    else
        cudaGetLastError();
    if (!h->sink_pinned && !h->ring[0])
    {
        for (int i = 0; i < pguresvt_handle::RING; i++)
        {
            CU(cudaMallocHost(&h->ring[i], h->fsz * sizeof(double)));
            CU(cudaEventCreateWithFlags(&h->ring_done[i], cudaEventDisableTiming));
        }
        h->copier = std::thread(copier_main, h);
    }
    return PGS_OK;
}

// frame t (local index lt) is final in dY on the main stream: send it home on the copy stream
static int stream_out_frame(pguresvt_handle *h, uint32_t t)
{
    const uint32_t lt = t - h->fb;
    const size_t bytes = h->fsz * sizeof(double);
    CU(cudaEventRecord(h->evFrame, h->st));
    CU(cudaStreamWaitEvent(h->copy_st, h->evFrame, 0));
    if (h->sink_pinned)
    {
        CU(cudaMemcpyAsync(h->sinkY + h->fsz * t, h->dY + h->fsz * lt, bytes, cudaMemcpyDeviceToHost, h->copy_st));
        h->frames_streamed++;
        return PGS_OK;
    }
    const int slot = (int)(h->frames_streamed % pguresvt_handle::RING);
    {
        std::unique_lock<std::mutex> lk(h->mx);
        h->cv.wait(lk, [h, slot] { return !h->ring_busy[slot]; });
        h->ring_busy[slot] = true;
        h->jobs_inflight++;
    }
    cudaError_t ce = cudaMemcpyAsync(h->ring[slot], h->dY + h->fsz * lt, bytes, cudaMemcpyDeviceToHost, h->copy_st);
    if (ce == cudaSuccess)
        ce = cudaEventRecord(h->ring_done[slot], h->copy_st);
    {
        std::lock_guard<std::mutex> lk(h->mx);
        if (ce == cudaSuccess)
            h->jobs.push_back({slot, t});
        else
        {
            h->ring_busy[slot] = false;
            h->jobs_inflight--;
        }
    }
    if (ce != cudaSuccess)
        return fail(PGS_ERR_CUDA, "CUDA error %s while streaming frame %u to the host", cudaGetErrorString(ce), t);
    h->cv.notify_all();
    h->frames_streamed++;
    return PGS_OK;
}

static int stream_out_drain(pguresvt_handle *h)
{
    if (!h->sinkY)
        return PGS_OK;
    if (!h->sink_pinned)
    {
        std::unique_lock<std::mutex> lk(h->mx);
        h->cv.wait(lk, [h] { return h->jobs_inflight == 0; });
    }
    CU(cudaStreamSynchronize(h->copy_st));
    return PGS_OK;
}

static int invalidate(pguresvt_handle *h)
{
    drop_prelaunched_noise(h);
    h->uploaded = true;
    h->noise_ws.cache.clear();
    for (auto &tg : h->tagF)
        tg.frame = -1;
    for (auto &tg : h->tagB)
        tg.frame = -1;
    h->prefiltered = false;
    h->cur_t = -1;
    return PGS_OK;
}

extern "C" int pguresvt_upload(pguresvt_handle *h, const void *X_full)
{
    if (!h || !X_full)
        return fail(PGS_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->p.device));
    const char *src = (const char *)X_full + h->fsz * h->r0 * h->esz;
    CU(cudaMemcpyAsync(h->dX, src, h->fsz * (h->r1 - h->r0) * h->esz, cudaMemcpyHostToDevice, h->st));
    return invalidate(h);
}

extern "C" int pguresvt_upload_device(pguresvt_handle *h, const void *dX_resident)
{
    if (!h || !dX_resident)
        return fail(PGS_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->p.device));
    CU(cudaMemcpyAsync(h->dX, dX_resident, h->fsz * (h->r1 - h->r0) * h->esz, cudaMemcpyDeviceToDevice, h->st));
    return invalidate(h);
}

// ------------------------------------------------------------------------------------------------------
// stage: median prefilter + per-frame maxima for all resident frames (pguresvt.hpp:69-88,116-117)
// ------------------------------------------------------------------------------------------------------
template <typename T>
static int prefilter_t(pguresvt_handle *h)
{
    const uint32_t nres = h->r1 - h->r0;
    const size_t ntot = h->fsz * nres;
    const int bpf = 64;
    std::vector<double> part((size_t)nres * bpf * 2);
    k_frame_max<T><<<dim3(bpf, nres), 256, 0, h->st>>>((const T *)h->dX, h->fsz, h->dMaxPartial, h->dMaxPartial + (size_t)nres * bpf);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(part.data(), h->dMaxPartial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    for (uint32_t f = 0; f < nres; f++)
    {
        h->xmax[f] = *std::max_element(part.begin() + (size_t)f * bpf, part.begin() + (size_t)(f + 1) * bpf);
        h->xmin[f] = *std::min_element(part.begin() + (size_t)(nres + f) * bpf, part.begin() + (size_t)(nres + f + 1) * bpf);
    }
    if (h->p.median_size > 0)
    {
        const uint16_t *src16;
        if (h->dtype == PGS_U16)
            src16 = (const uint16_t *)h->dX;
        else
        {
            k_to_u16<T><<<std::min(cdiv(ntot, 256), 148 * 16), 256, 0, h->st>>>((const T *)h->dX, h->dTmp16, ntot);
            LAUNCHED(h);
            src16 = h->dTmp16;
        }
        const int r = (int)h->p.median_size;
        const int tw = 32 + 2 * r;
        const size_t smem = (size_t)tw * tw * sizeof(uint16_t);
        if (smem > 48 * 1024)
            CU(cudaFuncSetAttribute(k_median_u16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nb = cdiv(h->N, 32);
        k_median_u16<<<dim3(nb, nb, nres), dim3(32, 8), smem, h->st>>>(src16, h->dZ, h->N, h->N, r);
        LAUNCHED(h);
        k_frame_max<uint16_t><<<dim3(bpf, nres), 256, 0, h->st>>>(h->dZ, h->fsz, h->dMaxPartial);
        LAUNCHED(h);
        CU(cudaMemcpyAsync(part.data(), h->dMaxPartial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        for (uint32_t f = 0; f < nres; f++)
            h->zmax[f] = *std::max_element(part.begin() + (size_t)f * bpf, part.begin() + (size_t)(f + 1) * bpf);
    }
    else
        h->zmax = h->xmax; // Z = conv_to<cube>(X) (pguresvt.hpp:86)
    CU(cudaGetLastError());
    return PGS_OK;
}

static int prefilter(pguresvt_handle *h)
{
    if (h->prefiltered)
        return PGS_OK;
    if (!h->uploaded)
        return fail(PGS_ERR_ARG, "no input uploaded");
    int rc;
    switch (h->dtype)
    {
    case PGS_U8:
        rc = prefilter_t<uint8_t>(h);
        break;
    case PGS_U16:
        rc = prefilter_t<uint16_t>(h);
        break;
    case PGS_F32:
        rc = prefilter_t<float>(h);
        break;
    default:
        rc = prefilter_t<double>(h);
        break;
    }
    if (rc == PGS_OK)
        h->prefiltered = true;
    return rc;
}

// Bernoulli perturbations: identical for every frame (same seed, same size: SURVEY Q10) → once per handle
static int perturb(pguresvt_handle *h)
{
    if (h->perturbed || !h->p.optimize_pgure)
        return PGS_OK;
    int64_t seed = h->p.random_seed;
    if (seed < 0)
    { // pgure.hpp:52-55: external entropy; not reproducible by construction
        std::random_device rd;
        seed = (int64_t)((((uint64_t)rd() << 32) | rd()) >> 1);
    }
    h->seed_used = seed;
    const long long n = (long long)(h->fsz * h->win);
    const int chunk = 64;
    const long long nthreads = (2 * n + chunk - 1) / chunk;
    k_perturb<<<cdiv(nthreads, 128), 128, 0, h->st>>>(h->dD1, h->dD2, n, (unsigned long long)seed, h->vP, chunk);
    LAUNCHED(h);
    CU(cudaGetLastError());
    h->perturbed = true;
    return PGS_OK;
}

// ------------------------------------------------------------------------------------------------------
// per-frame stages
// ------------------------------------------------------------------------------------------------------
template <typename T>
static void launch_window(pguresvt_handle *h, const T *src, double *dst, double vmax)
{
    const size_t n = h->fsz * h->win;
    k_window<T><<<std::min(cdiv(n, 256), h->sm_count * 16), 256, 0, h->st>>>(src, dst, n, vmax);
    LAUNCHED(h);
}

static void window_maxima(const pguresvt_handle *h, uint32_t t, uint32_t &a, double &uMax, double &wMax)
{
    a = window_start(h, t);
    const uint32_t la = a - h->r0; // local index of the first window frame
    uMax = h->xmax[la], wMax = h->zmax[la];
    for (uint32_t k = 1; k < h->win; k++)
    {
        uMax = std::max(uMax, h->xmax[la + k]);
        wMax = std::max(wMax, h->zmax[la + k]);
    }
}

// Noise estimate of frame t's window started ahead of time on the noise stream (called while frame t-1's lambda search
// is about to run): the window is normalised into its own buffer, the estimator thread fills noise_val.
static int prelaunch_noise(pguresvt_handle *h, uint32_t t)
{
    const size_t n = h->fsz * h->win;
    if (!h->dUn)
        CU(cudaMalloc(&h->dUn, n * sizeof(double)));
    uint32_t a;
    double uMax, wMax;
    window_maxima(h, t, a, uMax, wMax);
    const size_t off = h->fsz * (a - h->r0);
    const int grid = std::min(cdiv(n, 256), h->sm_count * 16);
    switch (h->dtype)
    {
    case PGS_U8:
        k_window<uint8_t><<<grid, 256, 0, h->noise_st>>>((const uint8_t *)h->dX + off, h->dUn, n, uMax);
        break;
    case PGS_U16:
        k_window<uint16_t><<<grid, 256, 0, h->noise_st>>>((const uint16_t *)h->dX + off, h->dUn, n, uMax);
        break;
    case PGS_F32:
        k_window<float><<<grid, 256, 0, h->noise_st>>>((const float *)h->dX + off, h->dUn, n, uMax);
        break;
    default:
        k_window<double><<<grid, 256, 0, h->noise_st>>>((const double *)h->dX + off, h->dUn, n, uMax);
        break;
    }
    LAUNCHED(h);
    CU(cudaGetLastError());
    const pguresvt_params &p = h->p;
    h->noise_val[0] = (p.alpha_est >= 0.0) ? p.alpha_est : -1.0;
    h->noise_val[1] = (p.mu_est >= 0.0) ? p.mu_est : -1.0;
    h->noise_val[2] = (p.sigma_est >= 0.0) ? p.sigma_est : -1.0;
    h->noise_launches = 0;
    h->noise_pre_t = t;
    h->noise_thread = std::thread([h, a, uMax]() {
        cudaSetDevice(h->p.device);
        h->noise_rc = noise_estimate_window(h->noise_ws, h->dUn, (int)h->N, (int)h->win, (int)h->p.noise_method, h->sm_count, h->noise_st,
                                            h->noise_val[0], h->noise_val[1], h->noise_val[2], &h->noise_launches, h->noise_err,
                                            (long long)a, uMax);
    });
    return PGS_OK;
}

// an estimate started ahead of time that nobody is going to collect (probes, re-uploads, errors)
static void drop_prelaunched_noise(pguresvt_handle *h)
{
    if (h->noise_thread.joinable())
        h->noise_thread.join();
    h->noise_pre_t = -1;
}

static int stage_window(pguresvt_handle *h, uint32_t t)
{
    uint32_t a;
    double uMax, wMax;
    window_maxima(h, t, a, uMax, wMax);
    h->cur_a = a;
    const uint32_t la = a - h->r0;
    h->cur_uMax = uMax;
    h->cur_wMax = wMax;
    {
        double lo = h->xmin[la];
        for (uint32_t k = 1; k < h->win; k++)
            lo = std::min(lo, h->xmin[la + k]);
        const double am = std::fabs(lo / uMax);
        h->cur_absmax = std::isfinite(am) ? std::max(1.0, am) : 1.0;
    }
    const size_t off = h->fsz * la;
    switch (h->dtype)
    {
    case PGS_U8:
        launch_window<uint8_t>(h, (const uint8_t *)h->dX + off, h->dU, uMax);
        break;
    case PGS_U16:
        launch_window<uint16_t>(h, (const uint16_t *)h->dX + off, h->dU, uMax);
        break;
    case PGS_F32:
        launch_window<float>(h, (const float *)h->dX + off, h->dU, uMax);
        break;
    default:
        launch_window<double>(h, (const double *)h->dX + off, h->dU, uMax);
        break;
    }
    if (h->p.motion_estimation)
    {
        if (h->p.median_size > 0)
            launch_window<uint16_t>(h, h->dZ + off, h->dW, wMax);
        else
            switch (h->dtype)
            {
            case PGS_U8:
                launch_window<uint8_t>(h, (const uint8_t *)h->dX + off, h->dW, wMax);
                break;
            case PGS_U16:
                launch_window<uint16_t>(h, (const uint16_t *)h->dX + off, h->dW, wMax);
                break;
            case PGS_F32:
                launch_window<float>(h, (const float *)h->dX + off, h->dW, wMax);
                break;
            default:
                launch_window<double>(h, (const double *)h->dX + off, h->dW, wMax);
                break;
            }
    }
    // reference slice of the trajectories (arps.hpp:54-113) and output slice (pguresvt.hpp:155-166)
    if (t < h->fw)
    {
        h->cur_ref = (int)t;
        h->cur_sl = (int)t;
    }
    else if (t >= h->nframes - h->fw)
    {
        h->cur_ref = (int)(t - (h->nframes - h->win)); // arps.hpp:82 with Nt = A.n_slices
        h->cur_sl = (int)(t - (h->nframes - h->Nt));   // pguresvt.hpp:161 with the driver's Nt
    }
    else
    {
        h->cur_ref = (int)h->fw;
        h->cur_sl = (int)h->fw;
    }
    CU(cudaGetLastError());
    return PGS_OK;
}

static void arps_pair(pguresvt_handle *h, int f1, int f2, const short2 *pred, short2 *out)
{
    const double oobs2 = 1.0 / (double)(h->p.block_size * h->p.block_size);
    static const bool generic_only = getenv("PGURESVT_ARPS_GENERIC") != nullptr;
    if (h->p.block_size == 4 && h->p.motion_window <= 7 && !generic_only)
        k_arps_pair4<<<cdiv(h->vecSize, 128), 128, 0, h->st>>>(h->dW, h->N, h->p.motion_window, f1, f2, pred, out, h->vecSize, nullptr);
    else
        k_arps_pair<<<cdiv(h->vecSize, 128), 128, 0, h->st>>>(h->dW, h->N, h->p.block_size, h->p.motion_window, f1, f2, pred, out,
                                                              h->vecSize, oobs2, h->dNcost);
    LAUNCHED(h);
    h->stats[14] += 1;
}

// Trajectory slice of a zero-predictor pair, shared between windows: the search for (source frame -> target frame)
// is independent of the output frame it is run for, as long as the window normalisation wMax is bit-identical
// (the block cost is evaluated on w = z / wMax, and ties decide vectors).  Forward results are keyed by the target
// frame g (source g-1), backward results by target g (source g+1).
static int flush_slice_copies(pguresvt_handle *h)
{
    if (h->pend.n == 0)
        return PGS_OK;
    k_copy_slices<<<dim3(std::min(cdiv(h->vecSize, 256), 64), h->pend.n), 256, 0, h->st>>>(h->pend, h->vecSize); // short2 = one 32-bit word
    LAUNCHED(h);
    h->pend.n = 0;
    CU(cudaGetLastError());
    return PGS_OK;
}
static int queue_slice_copy(pguresvt_handle *h, const short2 *src, short2 *dst)
{
    if (h->pend.n == 32)
    {
        int rc = flush_slice_copies(h);
        if (rc)
            return rc;
    }
    h->pend.src[h->pend.n] = reinterpret_cast<const int *>(src);
    h->pend.dst[h->pend.n] = reinterpret_cast<int *>(dst);
    h->pend.n++;
    return PGS_OK;
}

static int arps_cached_pair(pguresvt_handle *h, bool forward, int src_local, int tgt_local)
{
    const long long g = (long long)h->cur_a + tgt_local;
    const int slot = (int)(g % h->arps_ring);
    pguresvt_handle::ArpsTag &tag = forward ? h->tagF[slot] : h->tagB[slot];
    short2 *cache = (forward ? h->dArpsF : h->dArpsB) + (size_t)slot * h->vecSize;
    short2 *dst = h->dPos + (size_t)tgt_local * h->vecSize;
    if (tag.frame == g && tag.wmax == h->cur_wMax)
    {
        h->stats[15] += 1;
        return queue_slice_copy(h, cache, dst);
    }
    arps_pair(h, src_local, tgt_local, nullptr, dst); // (zero predictor: reads no trajectory slice, so queued copies may wait)
    tag.frame = g;
    tag.wmax = h->cur_wMax;
    return queue_slice_copy(h, dst, cache); // (stream order: the copy kernel is launched after this pair's kernel)
}

static int stage_motion(pguresvt_handle *h, uint32_t t) // MotionEstimator::Estimate, arps.hpp:52-134
{
    const int tw = (int)h->fw, Ntw = (int)h->win, nImages = (int)h->nframes, ti = (int)t;
    const int ref = h->cur_ref;
    h->pend.n = 0;
    CU(cudaMemsetAsync(h->dNcost, 0, sizeof(unsigned long long), h->st));
    if (!h->p.motion_estimation)
    { // only the reference slice is populated; every other slice stays (0,0)
        CU(cudaMemsetAsync(h->dPos, 0, (size_t)h->win * h->vecSize * sizeof(short2), h->st));
        k_seed_pos<<<cdiv(h->vecSize, 256), 256, 0, h->st>>>(h->dPos, h->vecSize, h->M1, ref);
        LAUNCHED(h);
        CU(cudaGetLastError());
        return PGS_OK;
    }
    k_seed_pos<<<cdiv(h->vecSize, 256), 256, 0, h->st>>>(h->dPos, h->vecSize, h->M1, ref);
    LAUNCHED(h);
    // schedule of arps.hpp:54-133: nfwd forward pairs (ref+i -> ref+i+1), nbwd backward pairs (ref-i -> ref-i-1)
    int nfwd, nbwd;
    bool special = false; // last frame: the backward pairs read motion slots nobody wrote (arps.hpp:101-104)
    if (ti < tw)
    {
        nfwd = Ntw - ti - 1;
        nbwd = ti;
    }
    else if (ti >= nImages - tw)
    {
        const int endFrame = ti - (nImages - Ntw);
        nfwd = 2 * tw - endFrame;
        nbwd = endFrame;
        special = (2 * tw == endFrame);
    }
    else
    {
        nfwd = tw;
        nbwd = tw;
    }
    int rc;
    for (int i = 0; i < nfwd; i++)
        if ((rc = arps_cached_pair(h, true, ref + i, ref + i + 1)))
            return rc;
    for (int i = 0; i < nbwd; i++)
    {
        if (i == 0 && nfwd > 0 && !special)
        { // predictor = the vector the first forward pair found for the same block; specific to this frame.
          // It reads trajectory slice ref + 1: the queued copies must have landed.
            if ((rc = flush_slice_copies(h)))
                return rc;
            arps_pair(h, ref, ref - 1, h->dPos + (size_t)(ref + 1) * h->vecSize, h->dPos + (size_t)(ref - 1) * h->vecSize);
        }
        else if ((rc = arps_cached_pair(h, false, ref - i, ref - i - 1)))
            return rc;
    }
    if ((rc = flush_slice_copies(h)))
        return rc;
    // slices no pair targets stay (0,0) like the zero-initialised icube (arps.hpp:42); with a full schedule none is left
    for (int k = 0; k < Ntw; k++)
        if (k > ref + nfwd || k < ref - nbwd)
            CU(cudaMemsetAsync(h->dPos + (size_t)k * h->vecSize, 0, (size_t)h->vecSize * sizeof(short2), h->st));
    CU(cudaGetLastError());
    return PGS_OK;
}

// SVD of the patches listed in `ids` for shapes other than 16x15: warp-per-matrix register Jacobi for 64 x n, the
// shared-memory Jacobi otherwise; mode 0 writes full records (o.fac), mode 1 the compact cache (o.S, o.Q, o.lead).
static int launch_svd_generic(pguresvt_handle *h, const Perturb &pt, const int *ids, int np, const SvdOut &o, int mode)
{
    const int max_sweeps = 30;
    const double tol = 1e-15, tol2 = tol * tol;
    if (np <= 0)
        return PGS_OK;
    if (h->use_warp_svd)
    {
        const double big = 1e-6, big2 = big * big;
        const int smem = 4 * SVDW_WARP_DOUBLES(2) * (int)sizeof(double);
        if (!h->attr_svdw)
        {
            CU(cudaFuncSetAttribute(k_svd_warp<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CU(cudaFuncSetAttribute(k_svd_warp<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CU(cudaFuncSetAttribute(k_svd_warp<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            h->attr_svdw = true;
        }
        if (mode == 2)
            k_svd_warp<2, 2><<<cdiv(np, 4), 128, smem, h->st>>>(h->dU, pt, h->dPos, ids, np, h->vecSize, h->N, h->p.block_size, h->n, o,
                                                                 max_sweeps, tol2, big2, h->dSweeps);
        else if (mode == 1)
            k_svd_warp<2, 1><<<cdiv(np, 4), 128, smem, h->st>>>(h->dU, pt, h->dPos, ids, np, h->vecSize, h->N, h->p.block_size, h->n, o,
                                                                 max_sweeps, tol2, big2, h->dSweeps);
        else
            k_svd_warp<2, 0><<<cdiv(np, 4), 128, smem, h->st>>>(h->dU, pt, h->dPos, ids, np, h->vecSize, h->N, h->p.block_size, h->n, o,
                                                                 max_sweeps, tol2, big2, h->dSweeps);
    }
    else
    {
        const size_t per_warp = ((size_t)h->m * h->n + (size_t)h->n * h->n + h->n) * sizeof(double);
        int wpb = (int)std::min<size_t>(8, (size_t)(96 * 1024) / per_warp);
        if (wpb < 1)
            wpb = 1;
        const size_t smem = per_warp * wpb;
        if (smem > 200 * 1024)
            return fail(PGS_ERR_UNSUPPORTED, "Casorati matrix %dx%d too large for the shared-memory SVD kernel", h->m, h->n);
        if (mode == 1)
        {
            if (smem > 48 * 1024)
                CU(cudaFuncSetAttribute(k_svd_smem_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_svd_smem_c<<<cdiv(np, wpb), wpb * 32, smem, h->st>>>(h->dU, pt, h->dPos, ids, np, h->vecSize, h->N, h->p.block_size, h->n, o,
                                                                    max_sweeps, tol2, h->dSweeps);
        }
        else
        {
            if (smem > 48 * 1024)
                CU(cudaFuncSetAttribute(k_svd_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_svd_smem<<<cdiv(np, wpb), wpb * 32, smem, h->st>>>(h->dU, pt, h->dPos, ids, np, h->vecSize, h->N, h->p.block_size, h->n,
                                                                  o.ldv, o.fac, o.rec, max_sweeps, tol2, h->dSweeps);
        }
    }
    LAUNCHED(h);
    CU(cudaGetLastError());
    return PGS_OK;
}

static int stage_svd(pguresvt_handle *h, int obj) // SVT::Decompose, svt.hpp:58-118 (+ pgure.hpp:80-82)
{
    Perturb pt;
    pt.d1 = h->dD1;
    pt.d2neg = h->dD2;
    pt.mode = obj;
    const double eps1 = 1.0 * 0.0001; // U.max() * 1e-4 with U max-normalised (pgure.hpp:72)
    pt.eps = (obj == 1) ? eps1 : 100 * eps1;
    pt.dNeg = h->d2Neg;
    pt.dPos = h->d2Pos;
    const int max_sweeps = 30;
    const double tol = 1e-15, tol2 = tol * tol;
    if (h->use_l4)
    {
        const long long nthreads = (long long)h->P * 4;
        const double big = 1e-6, big2 = big * big;
        // 0: tracked pair norms; 3: legacy variant that recomputes the pair norms every round
        // (unrolling the 15 rounds removes the register moves of the round-robin permutation but the loop then outgrows
        //  the instruction cache: measured 1.15x (3 rounds) to 1.7x (15 rounds) slower on B200)
        // 4: tracked norms with full rotations; 0 (default): tracked norms with fast (scaled) rotations
        const int variant = h->p.svd_kernel == 3 ? 0 : h->p.svd_kernel == 4 ? 1 : 2;
        // epilogue: lean mode leaves head entries (+ the full record of object U); see k_svd16_l4
        const int epi = (!h->lean || obj == 1) ? 0 : (obj == 0 ? 2 : (h->full_mode ? 0 : 1)); // (object U1: always the full record)
        auto cold = epi == 2 ? (variant == 2 ? k_svd16_l4<0, 2, 2> : variant == 1 ? k_svd16_l4<0, 1, 2> : k_svd16_l4<0, 0, 2>)
                             : (variant == 2 ? k_svd16_l4<0, 2, 0> : variant == 1 ? k_svd16_l4<0, 1, 0> : k_svd16_l4<0, 0, 0>);
        auto warm = epi == 1 ? (variant == 2 ? k_svd16_l4<1, 2, 1> : variant == 1 ? k_svd16_l4<1, 1, 1> : k_svd16_l4<1, 0, 1>)
                             : (variant == 2 ? k_svd16_l4<1, 2, 0> : variant == 1 ? k_svd16_l4<1, 1, 0> : k_svd16_l4<1, 0, 0>);
        const int part = obj == 0 ? 0 : obj == 2 ? 1 : 2;
        if (epi == 0 && obj != 0 && !h->dFac[obj])
        { // full records of a perturbed object in lean mode: allocated on first use (4 GB per object at 1024^2)
            CU(cudaMalloc(&h->dFac[obj], h->rec * (size_t)h->P * sizeof(double)));
            CU(cudaMemsetAsync(h->dFac[obj], 0, h->rec * (size_t)h->P * sizeof(double), h->st));
        }
        const double *usrc = h->dU;
        if (obj != 0)
        { // U + eps*delta written out once (pgure.hpp:80-82); the SVD kernel then gathers plain doubles
            const size_t wtot = h->fsz * h->win;
            double *dst = (h->top1 && obj == 3) ? h->dUp3 : h->dUp; // top1 keeps both perturbed windows for the exact fixes
            if (!(h->top1 && h->full_mode)) // (already there when the frame falls back to full records)
            {
                k_perturb_window<<<std::min(cdiv(wtot, 256), h->sm_count * 16), 256, 0, h->st>>>(h->dU, pt, wtot, dst);
                LAUNCHED(h);
            }
            usrc = dst;
        }
        if (h->top1 && !h->full_mode && (obj >= 2 || (obj == 0 && h->top1_all)))
        { // dominant triplet + bound instead of the full decomposition (k_top1_l4)
            auto ktop = obj == 0 ? k_top1_l4<2> : h->top1_all ? k_top1_l4<1> : k_top1_l4<0>;
            ktop<<<cdiv(nthreads, 128), 128, 0, h->st>>>(usrc, h->dPos, h->dIds, h->P, h->vecSize, h->N, h->dFac[0], h->dD2, pt.eps, h->d2Neg,
                                                         h->d2Pos, h->dC4, h->dHead, part, 40, h->dSweeps);
            LAUNCHED(h);
            h->stats[1] += h->P;
            CU(cudaGetLastError());
            return PGS_OK;
        }
        // dynamic shared memory: V of object 0 for the warm start, re-used for the hand-over of z in the V rebuild (both variants)
        const int smem_svd = 32 * SVD16_V0_STRIDE * (int)sizeof(double);
        // (the attribute is per function: set it for the instance about to be launched)
        CU(cudaFuncSetAttribute(obj == 0 ? cold : warm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_svd));
        if (obj == 0)
            cold<<<cdiv(nthreads, 128), 128, smem_svd, h->st>>>(usrc, h->dPos, h->dIds, h->P, h->vecSize, h->N, h->dFac[obj], nullptr,
                                                                max_sweeps, tol2, big2, h->dSweeps, h->dC4, h->dHead, part, nullptr);
        else // perturbed objects start from the V of object 0 (computed first for this frame)
            warm<<<cdiv(nthreads, 128), 128, smem_svd, h->st>>>(usrc, h->dPos, h->dIds, h->P, h->vecSize, h->N, h->dFac[obj], h->dFac[0],
                                                                max_sweeps, tol2, big2, h->dSweeps, h->dC4, h->dHead, part, nullptr);
    }
    else if (h->use_reg_svd)
    {
        const long long nthreads = (long long)h->P * 8;
        k_svd_16x15<<<cdiv(nthreads, 128), 128, 0, h->st>>>(h->dU, pt, h->dPos, h->dIds, h->P, h->vecSize, h->N, h->dFac[obj],
                                                            max_sweeps, tol2, h->dSweeps);
    }
    else if (h->use_warp_svd || h->use_compact)
    {
        SvdOut o{};
        if (h->use_compact)
        {
            if (h->use_warp3)
            { // the three objects of a patch back to back in one launch (perturbed ones warm-started), issued for object 0
                if (obj != 0)
                    return fail(PGS_ERR_ARG, "objects 2 and 3 are decomposed together with object 0");
                o.S[0] = h->dSc[0], o.S[1] = h->dSc[2], o.S[2] = h->dSc[3];
                o.Q[0] = h->dQc[0], o.Q[1] = h->dQc[2], o.Q[2] = h->dQc[3];
                pt.eps = 100 * eps1;
            }
            else
                o.S[0] = h->dSc[obj], o.Q[0] = h->dQc[obj];
            o.c4 = h->dC4;
            o.lead = (obj == 0) ? h->dLead : nullptr;
            o.R = (obj == 0) ? h->Rc : 0;
        }
        else
            o.fac = h->dFac[obj], o.rec = h->rec, o.ldv = h->ldv;
        int rc = launch_svd_generic(h, pt, h->dIds, h->P, o, h->use_compact ? (h->use_warp3 ? 2 : 1) : 0);
        if (rc)
            return rc;
        h->launches--; // counted once below
    }
    else
    {
        const size_t per_warp = ((size_t)h->m * h->n + (size_t)h->n * h->n + h->n) * sizeof(double);
        int wpb = (int)std::min<size_t>(8, (size_t)(96 * 1024) / per_warp);
        if (wpb < 1)
            wpb = 1;
        const size_t smem = per_warp * wpb;
        if (smem > 200 * 1024)
            return fail(PGS_ERR_UNSUPPORTED, "Casorati matrix %dx%d too large for the shared-memory SVD kernel", h->m, h->n);
        if (smem > 48 * 1024)
            CU(cudaFuncSetAttribute(k_svd_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_svd_smem<<<cdiv(h->P, wpb), wpb * 32, smem, h->st>>>(h->dU, pt, h->dPos, h->dIds, h->P, h->vecSize, h->N,
                                                               h->p.block_size, h->n, h->ldv, h->dFac[obj], h->rec, max_sweeps,
                                                               tol2, h->dSweeps);
    }
    LAUNCHED(h);
    h->stats[1] += h->P;
    CU(cudaGetLastError());
    return PGS_OK;
}

#define QFORM_LAZY_K 2
static int launch_qform(pguresvt_handle *h, int kmax)
{
    if (kmax == QFORM_LAZY_K)
        k_qform3_lead<QFORM_LAZY_K><<<h->eval_blocks, 128, 0, h->st>>>(h->dFac[0], h->dFac[2], h->dFac[3], h->dPos,
                                                                       h->P == h->vecSize ? nullptr : h->dIds, h->P, h->vecSize, h->N,
                                                                       h->dC4, h->dQ[0], h->dQ[1], h->dQ[2]);
    else
    {
        const int smem = 8 * 3 * 480 * (int)sizeof(double);
        if (!h->attr_qform)
        {
            CU(cudaFuncSetAttribute(k_qform3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            h->attr_qform = true;
        }
        k_qform3<<<h->eval_blocks, 128, smem, h->st>>>(h->dFac[0], h->dFac[2], h->dFac[3], h->dPos, h->dIds, h->P, h->vecSize, h->N, h->dC4,
                                                       h->dQ[0], h->dQ[1], h->dQ[2], kmax);
    }
    LAUNCHED(h);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(h->dNeedQ, 0, sizeof(double), h->st)); // dOut[4] is shared with the five-sum objective
    h->q_k = kmax;
    return PGS_OK;
}

static int stage_count(pguresvt_handle *h, int only_k)
{
    const size_t wtot = h->fsz * h->win;
    if (only_k >= 0)
        CU(cudaMemsetAsync(h->dCnt + h->fsz * only_k, 0, h->fsz * sizeof(unsigned), h->st));
    else
        CU(cudaMemsetAsync(h->dCnt, 0, wtot * sizeof(unsigned), h->st));
    const long long nt = (long long)h->P * (only_k >= 0 ? 1 : h->win);
    k_count<<<cdiv(nt, 256), 256, 0, h->st>>>(h->dPos, h->dIds, h->P, h->vecSize, h->N, h->p.block_size, h->win, only_k,
                                              h->dCnt);
    LAUNCHED(h);
    { // scale of the fixed-point accumulators: |block entry| <= ||A||_F <= sqrt(m n) max|a| (perturbed objects: + eps2 |delta2|)
        unsigned *dMaxCnt = reinterpret_cast<unsigned *>(h->dAccScale + 2);
        CU(cudaMemsetAsync(dMaxCnt, 0, sizeof(unsigned), h->st));
        const size_t off = only_k >= 0 ? h->fsz * only_k : 0, nv = only_k >= 0 ? h->fsz : wtot;
        k_cnt_max<<<std::min(cdiv(nv, 1024), h->sm_count * 8), 256, 0, h->st>>>(h->dCnt, off, nv, dMaxCnt);
        LAUNCHED(h);
        const double amax = std::max(1.0, h->cur_absmax) + 0.02;
        k_acc_scale<<<1, 1, 0, h->st>>>(dMaxCnt, std::sqrt((double)h->m * h->n) * amax, h->dAccScale);
        LAUNCHED(h);
    }
    CU(cudaGetLastError());
    return PGS_OK;
}

static int launch_recon_generic(pguresvt_handle *h, const double *fac, const int *ids, int np, double lambda, int only_k, double *acc)
{
    const int G = (h->m <= 16) ? 16 : 32;
    const int threads = 128, gpb = threads / G;
    const int NMAX = (h->n <= 16) ? 16 : 32;
    const size_t smem = ((size_t)h->ldv * h->n + 2 * NMAX) * sizeof(double) * gpb;
    if (NMAX == 16)
        k_recon<16><<<cdiv(np, gpb), threads, smem, h->st>>>(fac, h->rec, h->m, h->n, h->ldv, h->p.block_size, h->dPos, ids, np, h->vecSize,
                                                             h->N, lambda, h->p.exp_weighting, only_k, acc, h->dAccScale, G);
    else
        k_recon<32><<<cdiv(np, gpb), threads, smem, h->st>>>(fac, h->rec, h->m, h->n, h->ldv, h->p.block_size, h->dPos, ids, np, h->vecSize,
                                                             h->N, lambda, h->p.exp_weighting, only_k, acc, h->dAccScale, G);
    LAUNCHED(h);
    CU(cudaGetLastError());
    return PGS_OK;
}

// Compact cache: thresholds + s4 partials + overlap-add of object U's block from the leading Rc triplets (k_eval_c).
// Patches with more than Rc survivors at this lambda are decomposed again (object U only) in chunks into a scratch
// buffer of full records and reconstructed by k_recon — the result is exact for any lambda.  The accumulator must be
// clear (whole window, or slice only_k) on entry.
static int compact_accumulate(pguresvt_handle *h, double lambda, int only_k, int want_s4)
{
    CU(cudaMemsetAsync(h->dOvf, 0, sizeof(int), h->st));
    k_eval_c<<<h->evc_warps / 4, 128, 0, h->st>>>(h->dSc[0], h->dSc[2], h->dSc[3], h->dQc[0], h->dQc[2], h->dQc[3], h->dLead, h->Rc, h->m, h->n,
                                                   h->p.block_size, h->dPos, h->dIds, h->P, h->vecSize, h->N, lambda, h->p.exp_weighting, only_k,
                                                   want_s4, h->dAcc[0], h->dAccScale, h->dPartialE, h->dKpart, h->dOvf);
    LAUNCHED(h);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->hOvf, h->dOvf, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    const int novf = *h->hOvf;
    if (novf == 0)
        return PGS_OK;
    h->stats[18] += 1;
    h->stats[20] += novf;
    if (!h->dFacScratch)
    {
        const size_t budget = (size_t)2 << 30;
        h->scratch_patches = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->P, budget / (h->rec * sizeof(double))));
        CU(cudaMalloc(&h->dFacScratch, h->rec * (size_t)h->scratch_patches * sizeof(double)));
        CU(cudaMemset(h->dFacScratch, 0, h->rec * (size_t)h->scratch_patches * sizeof(double)));
    }
    Perturb pt{};
    pt.d1 = h->dD1, pt.d2neg = h->dD2, pt.mode = 0, pt.eps = 0.0, pt.dNeg = h->d2Neg, pt.dPos = h->d2Pos;
    SvdOut o{};
    o.fac = h->dFacScratch, o.rec = h->rec, o.ldv = h->ldv;
    for (int off = 0; off < novf; off += h->scratch_patches)
    {
        const int np = std::min(h->scratch_patches, novf - off);
        int rc = launch_svd_generic(h, pt, h->dOvf + 1 + off, np, o, 0);
        if (rc)
            return rc;
        h->stats[1] += np;
        if ((rc = launch_recon_generic(h, h->dFacScratch, h->dOvf + 1 + off, np, lambda, only_k, h->dAcc[0])))
            return rc;
    }
    return PGS_OK;
}

static int objective_compact(pguresvt_handle *h, double lambda, double alpha, double mu, double sigma, double *value, double *terms)
{
    const size_t wtot = h->fsz * h->win;
    if (!h->acc0_clean)
    { // afterwards every evaluation's voxel pass leaves the accumulator cleared
        CU(cudaMemsetAsync(h->dAcc[0], 0, wtot * sizeof(double), h->st));
        h->acc0_clean = true;
    }
    int rc = compact_accumulate(h, lambda, -1, 1);
    if (rc)
        return rc;
    k_risk_uhat<0><<<RISK_BLOCKS, 256, 0, h->st>>>(h->dU, h->dCnt, h->dAcc[0], wtot, h->dAccScale, h->dPartialE, h->dKpart, h->evc_warps,
                                                h->dPartial);
    LAUNCHED(h);
    k_reduce_partials<<<1, 256, 0, h->st>>>(h->dPartial, RISK_BLOCKS, 4, h->dOut);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(h->hOut, h->dOut, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->stats[16] += h->hOut[3];
    const double s1 = h->hOut[0], s5 = h->hOut[1], s4 = h->hOut[2], s2 = h->cur_sumU, s3 = 0.0;
    const double sigmasq = sigma * sigma;
    const double eps1 = 1.0 * 0.0001, eps2 = 100 * eps1;
    const double OoN = 1.0 / ((double)h->N * h->N * h->win);
    *value = OoN * (s1 - (alpha + mu) * s2 + (2 / eps1 * s3) - (2 * sigmasq * alpha / (eps2 * eps2) * s4) + (2 * mu * s5) + mu) -
             sigmasq;
    if (terms)
    {
        terms[0] = s1, terms[1] = s2, terms[2] = s3, terms[3] = s4, terms[4] = s5;
    }
    h->stats[2] += 1;
    return PGS_OK;
}

static int launch_recon(pguresvt_handle *h, int obj, double lambda, int only_k) // SVT::Reconstruct, svt.hpp:121-160
{
    if (obj == 0)
        h->acc0_clean = false;
    if (obj == 1)
        h->acc1_clean = false;
    if (h->use_compact)
    {
        if (obj != 0)
            return fail(PGS_ERR_ARG, "the compact factor cache reconstructs object U only");
        if (only_k >= 0)
            CU(cudaMemsetAsync(h->dAcc[0] + h->fsz * only_k, 0, h->fsz * sizeof(double), h->st));
        else
            CU(cudaMemsetAsync(h->dAcc[0], 0, h->fsz * h->win * sizeof(double), h->st));
        return compact_accumulate(h, lambda, only_k, 0);
    }
    const size_t wtot = h->fsz * h->win;
    if (only_k >= 0)
        CU(cudaMemsetAsync(h->dAcc[obj] + h->fsz * only_k, 0, h->fsz * sizeof(double), h->st));
    else
        CU(cudaMemsetAsync(h->dAcc[obj], 0, wtot * sizeof(double), h->st));
    if (h->use_l4 && obj == 0 && only_k >= 0)
    { // output slice of the register-SVD configuration: rank-adaptive single-slice kernel
        k_final16<<<cdiv((long long)h->P * 16, 128), 128, 0, h->st>>>(h->dFac[0], h->dPos, h->P == h->vecSize ? nullptr : h->dIds, h->P,
                                                                      h->vecSize, h->N, lambda, h->p.exp_weighting, only_k, h->dAcc[0], h->dAccScale);
        LAUNCHED(h);
        CU(cudaGetLastError());
        return PGS_OK;
    }
    return launch_recon_generic(h, h->dFac[obj], h->dIds, h->P, lambda, only_k, h->dAcc[obj]);
}

// top1 mode, per frame: the critical lambda of the bounds the perturbed objects were given (k_lean_crit)
static int lean_crit(pguresvt_handle *h)
{
    h->hLean[0] = INFINITY, h->hLean[1] = 0.0;
    CU(cudaMemcpyAsync(h->dCrit, h->hLean, 2 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    k_lean_crit<<<cdiv(h->P, 256), 256, 0, h->st>>>(h->dHead, h->P, h->dCrit);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(h->hLean, h->dCrit, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->crit_lambda = h->hLean[0];
    h->crit_bound = h->hLean[1];
    return PGS_OK;
}

// top1 mode, before an evaluation at `lambda`: every patch whose bound would survive the threshold there is decomposed
// exactly (both perturbed objects, warm-started Jacobi over the list), so the evaluation sees exact head entries wherever
// they matter.  Probes on the safe side of the frame's critical lambda skip even the check.
static int lean_fix(pguresvt_handle *h, double lambda)
{
    if (!h->top1 || h->frame_full)
        return PGS_OK;
    const bool maybe = h->p.exp_weighting ? !(lambda <= h->crit_lambda) : !(lambda >= h->crit_bound);
    if (!maybe)
        return PGS_OK;
    CU(cudaMemsetAsync(h->dLeanList, 0, sizeof(int), h->st));
    k_lean_check<<<cdiv(h->P, 256), 256, 0, h->st>>>(h->dHead, h->P, lambda, h->p.exp_weighting, h->dLeanList);
    LAUNCHED(h);
    int *hcnt = reinterpret_cast<int *>(h->hLean + 2);
    CU(cudaMemcpyAsync(hcnt, h->dLeanList, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    const int n = *hcnt;
    h->stats[22] += 1;
    // every patch still carrying a bound is safe at this lambda once the listed ones are exact: the safe side grows
    if (h->p.exp_weighting)
        h->crit_lambda = std::max(h->crit_lambda, lambda);
    else
        h->crit_bound = std::min(h->crit_bound, lambda);
    if (n == 0)
        return PGS_OK;
    h->stats[20] += n;
    const int variant = h->p.svd_kernel == 3 ? 0 : h->p.svd_kernel == 4 ? 1 : 2;
    auto warm = variant == 2 ? k_svd16_l4<1, 2, 1> : variant == 1 ? k_svd16_l4<1, 1, 1> : k_svd16_l4<1, 0, 1>;
    const int smem_svd = 32 * SVD16_V0_STRIDE * (int)sizeof(double);
    CU(cudaFuncSetAttribute(warm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_svd));
    const double big = 1e-6, tol = 1e-15;
    if (h->top1_all)
    { // object U first (cold, full record + head entries): the perturbed objects start from its V
        auto cold = variant == 2 ? k_svd16_l4<0, 2, 2> : variant == 1 ? k_svd16_l4<0, 1, 2> : k_svd16_l4<0, 0, 2>;
        CU(cudaFuncSetAttribute(cold, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_svd));
        cold<<<cdiv((long long)n * 4, 128), 128, smem_svd, h->st>>>(h->dU, h->dPos, h->dIds, n, h->vecSize, h->N, h->dFac[0], nullptr, 30, tol * tol,
                                                                    big * big, nullptr, h->dC4, h->dHead, 0, h->dLeanList + 1);
        LAUNCHED(h);
        h->stats[23] += n;
    }
    for (int obj = 2; obj <= 3; obj++)
    {
        warm<<<cdiv((long long)n * 4, 128), 128, smem_svd, h->st>>>(obj == 2 ? h->dUp : h->dUp3, h->dPos, h->dIds, n, h->vecSize, h->N, nullptr,
                                                                    h->dFac[0], 30, tol * tol, big * big, nullptr, h->dC4, h->dHead, obj - 1,
                                                                    h->dLeanList + 1);
        LAUNCHED(h);
        h->stats[23] += n;
    }
    CU(cudaGetLastError());
    return PGS_OK;
}

// Lean mode, exact fallback: a third singular triplet of some object survived at the probed lambda (or a probe asks for the
// factors of a perturbed object).  The perturbed objects are decomposed again with full records, all q-forms are prepared
// and the rest of the frame runs through the general evaluation.
static int ensure_full(pguresvt_handle *h)
{
    if (!h->lean || h->frame_full)
        return PGS_OK;
    int rc;
    h->full_mode = true;
    for (int obj = h->top1_all ? 0 : 2; obj <= 3; obj += (obj == 0 ? 2 : 1))
        if ((rc = stage_svd(h, obj)))
        {
            h->full_mode = false;
            return rc;
        }
    h->full_mode = false;
    for (int k = 0; k < 3; k++)
        if (!h->dQ[k])
            CU(cudaMalloc(&h->dQ[k], (size_t)16 * h->P * sizeof(double)));
    if ((rc = launch_qform(h, SVD16_N)))
        return rc;
    h->frame_full = true;
    h->stats[18] += 1;
    return PGS_OK;
}

// One evaluation of PGURE::CalculatePGURE (pgure.hpp:120-137).  (alpha, mu, sigma) are the PGURE object's
// members, i.e. AFTER the sigma/mu swap of pguresvt.hpp:133.
static int objective_fused(pguresvt_handle *h, double lambda, double alpha, double mu, double sigma, double *value, double *terms)
{
    const size_t wtot = h->fsz * h->win;
    if (!h->acc0_clean)
    { // afterwards every evaluation's voxel pass leaves the accumulator cleared
        CU(cudaMemsetAsync(h->dAcc[0], 0, wtot * sizeof(double), h->st));
        h->acc0_clean = true;
    }
    if (h->eps1f && !h->acc1_clean)
    {
        CU(cudaMemsetAsync(h->dAcc[1], 0, wtot * sizeof(double), h->st));
        h->acc1_clean = true;
    }
    {
        int rcf = lean_fix(h, lambda);
        if (rcf)
            return rcf;
    }
    const double sigmasq_e = sigma * sigma;
    for (int attempt = 0; attempt < 2; attempt++)
    {
        static const int ppg = getenv("PGURESVT_EVAL_PPG") ? atoi(getenv("PGURESVT_EVAL_PPG")) : EVAL_PPG;
        const bool lean_now = h->lean && !h->frame_full;
        static const bool tiled_off = getenv("PGURESVT_ACC_TILED") && atoi(getenv("PGURESVT_ACC_TILED")) == 0;
        const int tiled = (h->N % 2 == 0 && !tiled_off) ? 1 : 0;
        auto kev = lean_now ? k_eval3<6, 8, 1>
                            : (ppg == 1) ? k_eval3<6, 1, 0> : (ppg == 4) ? k_eval3<6, 4, 0> : (ppg == 16) ? k_eval3<6, 16, 0> : (ppg == 32) ? k_eval3<6, 32, 0> : k_eval3<6, 8, 0>;
        const int ppg_eff = lean_now ? 8 : (ppg == 1 || ppg == 4 || ppg == 16 || ppg == 32) ? ppg : 8;
        // (lean: the head records travel in the q0 argument; fac2/fac3/q2/q3 are not read)
        kev<<<cdiv(h->P, 8 * ppg_eff), 128, 0, h->st>>>(h->dFac[0], h->dFac[2], h->dFac[3], lean_now ? h->dHead : h->dQ[0], h->dQ[1], h->dQ[2],
                                                        h->dPos, h->P == h->vecSize ? nullptr : h->dIds, h->P, h->vecSize, h->N, lambda,
                                                        h->p.exp_weighting, h->dAcc[0], h->dAccScale, h->dPartialE, h->dKpart, h->q_k, h->dNeedQ, tiled);
        LAUNCHED(h);
        if (h->eps1f)
        { // object U1 = U + eps1*delta1: accumulate-only pass into its own accumulator, s3 in the voxel pass
            k_eval3<6, 8, 2><<<cdiv(h->P, 64), 128, 0, h->st>>>(h->dFac[1], nullptr, nullptr, nullptr, nullptr, nullptr, h->dPos,
                                                                 h->P == h->vecSize ? nullptr : h->dIds, h->P, h->vecSize, h->N, lambda,
                                                                 h->p.exp_weighting, h->dAcc[1], h->dAccScale, nullptr, nullptr, SVD16_N, nullptr,
                                                                 tiled);
            LAUNCHED(h);
            k_risk_uhat<1><<<RISK_BLOCKS, 256, 0, h->st>>>(h->dU, h->dCnt, h->dAcc[0], wtot, h->dAccScale, h->dPartialE, h->dKpart,
                                                        4 * cdiv(h->P, 8 * ppg_eff), h->dPartial, tiled ? (int)h->N : 0, h->dAcc[1], h->dD1,
                                                        alpha, sigmasq_e - alpha * mu, h->dPartial3);
            LAUNCHED(h);
            k_reduce_partials<<<1, 256, 0, h->st>>>(h->dPartial3, RISK_BLOCKS, 1, h->dOut + 5);
            LAUNCHED(h);
        }
        else
        {
            k_risk_uhat<0><<<RISK_BLOCKS, 256, 0, h->st>>>(h->dU, h->dCnt, h->dAcc[0], wtot, h->dAccScale, h->dPartialE, h->dKpart,
                                                        4 * cdiv(h->P, 8 * ppg_eff), h->dPartial, tiled ? (int)h->N : 0);
            LAUNCHED(h);
        }
        k_reduce_partials<<<1, 256, 0, h->st>>>(h->dPartial, RISK_BLOCKS, 4, h->dOut);
        LAUNCHED(h);
        CU(cudaMemcpyAsync(h->hOut, h->dOut, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        h->stats[16] += h->hOut[3]; // singular triplets streamed by this pass (algorithmic-bytes accounting for bench.py)
        if (!*reinterpret_cast<const int *>(h->hOut + 4))
            break;
        // a triplet beyond the prepared q-forms survived at this lambda: prepare all of them and redo the pass
        int rc;
        if (h->lean)
        {
            if ((rc = ensure_full(h)))
                return rc;
        }
        else
        {
            if ((rc = launch_qform(h, SVD16_N)))
                return rc;
            h->stats[18] += 1;
        }
    }
    const double s1 = h->hOut[0], s5 = h->hOut[1], s4 = h->hOut[2], s2 = h->cur_sumU, s3 = h->eps1f ? h->hOut[5] : 0.0;
    const double sigmasq = sigma * sigma;
    const double eps1 = 1.0 * 0.0001, eps2 = 100 * eps1;
    const double OoN = 1.0 / ((double)h->N * h->N * h->win);
    *value = OoN * (s1 - (alpha + mu) * s2 + (2 / eps1 * s3) - (2 * sigmasq * alpha / (eps2 * eps2) * s4) + (2 * mu * s5) + mu) -
             sigmasq;
    if (terms)
    {
        terms[0] = s1, terms[1] = s2, terms[2] = s3, terms[3] = s4, terms[4] = s5;
    }
    h->stats[2] += 1;
    return PGS_OK;
}

// Per frame, after the SVDs and the q-forms: head records and the per-slice CSR "patches by destination" (tile_eval.cuh)
static int tile_prepare(pguresvt_handle *h)
{
    const size_t nbins = h->fsz * h->win;
    const int nb = cdiv(nbins, SCAN_ITEMS);
    const int *ids = h->P == h->vecSize ? nullptr : h->dIds;
    const int grid = std::min(cdiv((long long)h->P * h->win, 256), h->sm_count * 32);
    CU(cudaMemsetAsync(h->dBinCnt, 0, nbins * sizeof(int), h->st));
    k_bin_count<<<grid, 256, 0, h->st>>>(h->dPos, ids, h->P, h->vecSize, h->N, h->win, h->dBinCnt);
    LAUNCHED(h);
    k_scan_sums<<<nb, SCAN_THREADS, 0, h->st>>>(h->dBinCnt, nbins, h->dScanSums);
    LAUNCHED(h);
    k_scan_block_sums<<<1, 1024, 0, h->st>>>(h->dScanSums, nb);
    LAUNCHED(h);
    k_scan_apply<<<nb, SCAN_THREADS, 0, h->st>>>(h->dBinCnt, nbins, h->dScanSums, nb, h->dBinStart);
    LAUNCHED(h);
    CU(cudaMemsetAsync(h->dBinCnt, 0, nbins * sizeof(int), h->st)); // reused as the fill cursors
    k_bin_fill<<<grid, 256, 0, h->st>>>(h->dPos, ids, h->P, h->vecSize, h->N, h->win, h->dBinStart, h->dBinCnt, h->dFac[0], h->dEnt);
    LAUNCHED(h);
    k_bin_sort<<<std::min(cdiv(nbins, 256), h->sm_count * 32), 256, 0, h->st>>>(h->dBinStart, nbins, h->dEnt);
    LAUNCHED(h);
    k_head_pack<<<cdiv((long long)h->P * 4, 256), 256, 0, h->st>>>(h->dFac[0], h->dFac[2], h->dFac[3], h->dQ[0], h->dQ[1], h->dQ[2], h->P,
                                                                     h->lean ? nullptr : h->dHead, h->dU0c);
    LAUNCHED(h);
    CU(cudaGetLastError());
    h->frame_fallback = false;
    return PGS_OK;
}

static int objective_fused(pguresvt_handle *h, double lambda, double alpha, double mu, double sigma, double *value, double *terms);

// One evaluation of PGURE::CalculatePGURE through the gather path: thresholds + second-difference sum per patch, then one
// CTA per output tile and slice (no atomics, no accumulator cube), then the fixed-order reduction.
static int objective_tile(pguresvt_handle *h, double lambda, double alpha, double mu, double sigma, double *value, double *terms)
{
    if (h->frame_fallback)
        return objective_fused(h, lambda, alpha, mu, sigma, value, terms);
    {
        int rcf = lean_fix(h, lambda);
        if (rcf)
            return rcf;
    }
    const int nw = cdiv(h->P, 32), ntile = h->tile_r * h->tile_c * (int)h->win;
    k_thresh<<<cdiv(h->P, 256), 256, 0, h->st>>>(h->dHead, h->P, lambda, h->p.exp_weighting, h->dFth, h->dPartialE, h->dKpart, h->dNeedQ);
    LAUNCHED(h);
    k_tile_eval<0><<<dim3(h->win, h->tile_r, h->tile_c), 128, 0, h->st>>>(h->dBinStart, h->dEnt, h->dFac[0], h->dU0c, h->dFth, h->dU, h->dCnt, h->N,
                                                                          0, 1.0, nullptr, h->dTilePart);
    LAUNCHED(h);
    k_reduce_eval<<<1, 1024, 0, h->st>>>(h->dTilePart, ntile, h->dPartialE, h->dKpart, nw, h->dOut);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(h->hOut, h->dOut, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (*reinterpret_cast<const int *>(h->hOut + 4))
    { // a third triplet of some object survived at this lambda: all q-forms, general evaluation for the rest of the frame
        int rc;
        if (h->lean)
            rc = ensure_full(h);
        else
        {
            rc = launch_qform(h, SVD16_N);
            h->stats[18] += 1;
        }
        if (rc)
            return rc;
        h->frame_fallback = true;
        return objective_fused(h, lambda, alpha, mu, sigma, value, terms);
    }
    h->stats[16] += h->hOut[3];
    const double s1 = h->hOut[0], s5 = h->hOut[1], s4 = h->hOut[2], s2 = h->cur_sumU, s3 = 0.0;
    const double sigmasq = sigma * sigma;
    const double eps1 = 1.0 * 0.0001, eps2 = 100 * eps1;
    const double OoN = 1.0 / ((double)h->N * h->N * h->win);
    *value = OoN * (s1 - (alpha + mu) * s2 + (2 / eps1 * s3) - (2 * sigmasq * alpha / (eps2 * eps2) * s4) + (2 * mu * s5) + mu) -
             sigmasq;
    if (terms)
    {
        terms[0] = s1, terms[1] = s2, terms[2] = s3, terms[3] = s4, terms[4] = s5;
    }
    h->stats[2] += 1;
    return PGS_OK;
}

static int objective(pguresvt_handle *h, double lambda, double alpha, double mu, double sigma, double *value, double *terms)
{
    if (h->use_tile)
        return objective_tile(h, lambda, alpha, mu, sigma, value, terms);
    if (h->use_fused_eval)
        return objective_fused(h, lambda, alpha, mu, sigma, value, terms);
    if (h->use_compact)
        return objective_compact(h, lambda, alpha, mu, sigma, value, terms);
    for (int k = 0; k < h->nobj; k++)
    {
        int rc = launch_recon(h, h->objs[k], lambda, -1);
        if (rc)
            return rc;
    }
    const size_t wtot = h->fsz * h->win;
    const double sigmasq = sigma * sigma;
    k_risk<<<RISK_BLOCKS, 256, 0, h->st>>>(h->dU, h->dD1, h->dD2, h->dCnt, h->dAcc[0], h->dAcc[1], h->dAcc[2], h->dAcc[3], h->dAccScale, wtot,
                                           alpha, mu, sigmasq, h->d2Neg, h->d2Pos, h->dPartial);
    LAUNCHED(h);
    k_reduce_partials<<<1, 256, 0, h->st>>>(h->dPartial, RISK_BLOCKS, 5, h->dOut);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(h->hOut, h->dOut, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    const double s1 = h->hOut[0], s2 = h->hOut[1], s3 = h->hOut[2], s4 = h->hOut[3], s5 = h->hOut[4];
    const double eps1 = 1.0 * 0.0001, eps2 = 100 * eps1;
    const double OoN = 1.0 / ((double)h->N * h->N * h->win); // pgure.hpp:49 with Nt = U.n_slices
    *value = OoN * (s1 - (alpha + mu) * s2 + (2 / eps1 * s3) - (2 * sigmasq * alpha / (eps2 * eps2) * s4) + (2 * mu * s5) + mu) -
             sigmasq;
    if (terms)
        for (int q = 0; q < 5; q++)
            terms[q] = h->hOut[q];
    h->stats[2] += 1;
    return PGS_OK;
}

// accu(u) of the current window in Armadillo's order (two sequential accumulators), started right after the window is
// normalised and collected when the SVD launches are in flight
static int sum_u_launch(pguresvt_handle *h)
{
    const size_t wtot = h->fsz * h->win;
    CU(cudaEventRecord(h->evU, h->st));
    CU(cudaStreamWaitEvent(h->sum_st, h->evU, 0));
    const int smem_as = AS_THREADS * (AS_E + 1) * (int)sizeof(double);
    CU(cudaFuncSetAttribute(k_accu_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_as));
    k_accu_seq<<<2, AS_THREADS, smem_as, h->sum_st>>>(h->dU, wtot, h->dSum2);
    LAUNCHED(h);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->hSum2, h->dSum2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->sum_st));
    CU(cudaEventRecord(h->evSum, h->sum_st));
    h->sum_pending = true;
    return PGS_OK;
}
static int sum_u_wait(pguresvt_handle *h)
{
    if (!h->sum_pending)
        return PGS_OK;
    CU(cudaEventSynchronize(h->evSum));
    h->cur_sumU = h->hSum2[0] + h->hSum2[1]; // acc1 + acc2
    h->sum_pending = false;
    return PGS_OK;
}

static cudaEvent_t timer_event(pguresvt_handle *h)
{
    if (h->evused == h->evpool.size())
    {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        h->evpool.push_back(e);
    }
    return h->evpool[h->evused++];
}
// adds the elapsed time of every recorded stage to its stats slot; the stream must have been synchronised
static void resolve_timers(pguresvt_handle *h, bool keep)
{
    if (keep)
        for (const auto &r : h->trecs)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess)
                h->stats[r.slot] += ms;
        }
    h->trecs.clear();
    h->evused = 0;
}
// Stage time by a pair of events recorded on the handle's stream; nothing waits on them until resolve_timers (the first
// version synchronised on every stage: ~6 round trips per frame that production runs paid for nothing).
struct StageTimer
{
    pguresvt_handle *h;
    int slot;
    cudaEvent_t a;
    StageTimer(pguresvt_handle *h_, int slot_) : h(h_), slot(slot_), a(timer_event(h_)) { cudaEventRecord(a, h->st); }
    ~StageTimer()
    {
        cudaEvent_t b = timer_event(h);
        cudaEventRecord(b, h->st);
        h->trecs.push_back({slot, a, b});
        if (h->trecs.size() > 8192)
        { // probes never resolve: keep the pool bounded
            cudaStreamSynchronize(h->st);
            resolve_timers(h, false);
        }
    }
};

// window + trajectories + SVD factors (+ weights for the lambda search) of frame t, cached per handle
static int prepare_frame(pguresvt_handle *h, uint32_t t)
{
    int rc;
    if ((rc = prefilter(h)))
        return rc;
    if ((rc = perturb(h)))
        return rc;
    if (h->cur_t == (long long)t)
        return PGS_OK;
    h->cur_t = -1;
    if ((rc = stage_window(h, t)))
        return rc;
    if (h->p.optimize_pgure && (rc = sum_u_launch(h)))
        return rc;
    if (h->noise_req)
    { // the estimator reads only the window cube: let it run beside ARPS and the SVDs on its own high-priority stream
        CU(cudaEventRecord(h->evWin, h->st));
        h->noise_started = true;
        h->noise_launches = 0;
        h->noise_thread = std::thread([h]() {
            cudaSetDevice(h->p.device);
            cudaStreamWaitEvent(h->noise_st, h->evWin, 0);
            h->noise_rc = noise_estimate_window(h->noise_ws, h->dU, (int)h->N, (int)h->win, (int)h->p.noise_method, h->sm_count, h->noise_st,
                                                h->noise_val[0], h->noise_val[1], h->noise_val[2], &h->noise_launches, h->noise_err,
                                                (long long)h->cur_a, h->cur_uMax);
        });
    }
    {
        StageTimer tm(h, 4);
        if ((rc = stage_motion(h, t)))
            return rc;
    }
    if (h->p.optimize_pgure)
    {
        StageTimer tm(h, 17);
        if ((rc = stage_count(h, -1)))
            return rc;
        if (h->use_fused_eval || h->use_compact)
        { // per-voxel multiplier delta2/weights for the q-forms of this frame
            const size_t wtot = h->fsz * h->win;
            k_c4<<<std::min(cdiv(wtot, 256), h->sm_count * 16), 256, 0, h->st>>>(h->dCnt, h->dD2, h->d2Neg, h->d2Pos, wtot, h->dC4);
            LAUNCHED(h);
        }
    }
    {
        StageTimer tm(h, 5);
        for (int k = 0; k < h->nobj; k++)
        {
            if (h->use_warp3 && h->objs[k] != 0)
            { // decomposed together with object 0
                h->stats[1] += h->P;
                continue;
            }
            CU(cudaMemsetAsync(h->dSweeps, 0, 4 * sizeof(int), h->st));
            if ((rc = stage_svd(h, h->objs[k])))
                return rc;
            int sw[4] = {0, 0, 0, 0};
            CU(cudaMemcpyAsync(sw, h->dSweeps, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            h->stats[10] = std::max(h->stats[10], (double)sw[0]);
            // mean sweeps per warp of object 0 ([12]) and of the warm-started objects ([13]) of the last frame
            const double nwarps = std::ceil((double)h->P * (h->use_l4 ? 4 : 8) / 32.0);
            if (h->use_reg_svd && h->use_l4)
                h->stats[h->objs[k] == 0 ? 12 : 13] = sw[1] / nwarps;
            if (h->use_warp_svd && h->use_compact)
            { // one warp per matrix: [1] sums object U (or the only object of the launch), [2] the warm-started ones
                if (h->use_warp3 || h->objs[k] == 0)
                    h->stats[12] = sw[1] / (double)h->P;
                if (h->use_warp3)
                    h->stats[13] = sw[2] / (2.0 * h->P);
            }
        }
    }
    if (h->use_fused_eval)
    {
        StageTimer tm(h, 17);
        h->frame_full = false;
        if (h->lean)
        { // the SVD kernels left the head entries (S, leading q-forms) behind: nothing to prepare
            CU(cudaMemsetAsync(h->dNeedQ, 0, sizeof(double), h->st));
            if (h->top1 && (rc = lean_crit(h)))
                return rc;
        }
        else if ((rc = launch_qform(h, QFORM_LAZY_K))) // q = u^T C4 v of the leading triplets (the rest lazily, see objective_fused)
            return rc;
        if (h->use_tile && (rc = tile_prepare(h)))
            return rc;
    }
    if ((rc = sum_u_wait(h)))
        return rc;
    h->cur_t = t;
    return PGS_OK;
}

static int estimate_noise(pguresvt_handle *h, double &alpha, double &mu, double &sigma)
{
    // Q9: the reference runs the estimator even when all three are user-supplied and then discards the
    // result; skipping it in that case is bit-identical.
    if (alpha >= 0. && mu >= 0. && sigma >= 0.)
        return PGS_OK;
    StageTimer tm(h, 8);
    return noise_estimate_window(h->noise_ws, h->dU, (int)h->N, (int)h->win, (int)h->p.noise_method, h->sm_count, h->st, alpha, mu,
                                 sigma, &h->launches, g_err, (long long)h->cur_a, h->cur_uMax);
}

static int process_frame(pguresvt_handle *h, uint32_t t) // pgureFunc, pguresvt.hpp:90-167
{
    int rc;
    const pguresvt_params &p = h->p;
    double lambda = (p.lambda_est >= 0.0) ? p.lambda_est : -1.0;
    double alpha = (p.alpha_est >= 0.0) ? p.alpha_est : -1.0;
    double mu = (p.mu_est >= 0.0) ? p.mu_est : -1.0;
    double sigma = (p.sigma_est >= 0.0) ? p.sigma_est : -1.0;
    // Q9: the reference runs the estimator even when all three are user-supplied and then discards the result
    const bool want_noise = p.optimize_pgure && !(alpha >= 0. && mu >= 0. && sigma >= 0.);
    if (h->noise_pre_t >= 0 && (h->noise_pre_t != (long long)t || !want_noise))
        drop_prelaunched_noise(h);
    const bool pre = h->noise_pre_t == (long long)t; // this frame's estimate has been running since the previous lambda search
    h->noise_pre_t = -1;
    h->noise_req = want_noise && !pre;
    h->noise_started = pre;
    if (!pre)
        h->noise_val[0] = alpha, h->noise_val[1] = mu, h->noise_val[2] = sigma;
    rc = prepare_frame(h, t);
    h->noise_req = false;
    if (h->noise_started)
    { // collect the concurrent estimate; [8] is the part of it still exposed after the SVD stage
        const auto w0 = std::chrono::steady_clock::now();
        h->noise_thread.join();
        h->stats[8] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        h->launches += h->noise_launches;
        if (!rc && h->noise_rc)
        {
            g_err = h->noise_err;
            rc = h->noise_rc;
        }
        alpha = h->noise_val[0], mu = h->noise_val[1], sigma = h->noise_val[2];
    }
    if (rc)
        return rc;
    if (p.optimize_pgure)
    {
        if (!h->noise_started && (rc = estimate_noise(h, alpha, mu, sigma)))
            return rc;
        // Optional (PGURESVT_NOISE_PRELAUNCH=1): next frame's estimate beside this frame's lambda search instead of beside its
        // own SVDs.  Measured: the SVD stage gets its 55 ms per 32 frames back, the search loses them (23.60 vs 23.54
        // frames/s) — the estimator costs what it costs wherever it runs, so the simpler schedule stays the default.
        static const bool pre_on = getenv("PGURESVT_NOISE_PRELAUNCH") != nullptr && atoi(getenv("PGURESVT_NOISE_PRELAUNCH")) > 0;
        if (want_noise && pre_on && t + 1 < h->fe && (rc = prelaunch_noise(h, t + 1)))
            return rc;
        StageTimer tm(h, 6);
        CU(cudaMemsetAsync(h->dNcost, 0, sizeof(unsigned long long), h->st));
        const double OoNxNyNt = 1.0 / ((double)h->N * h->N * h->Nt); // pguresvt.hpp:60 (the driver's Nt)
        double start = (lambda >= 0.0) ? lambda : h->cur_sumU * OoNxNyNt;
        start = std::max(0.0, start);
        const double ub = std::max(100.0, start);
        Sbplx1D opt;
        opt.lb = 0.0;
        opt.ub = ub;
        opt.ftol_rel = p.tol;
        opt.xtol_abs = 1E-12;
        opt.maxeval = (int)p.max_iter;
        double last = start;
        int err = PGS_OK;
        // NB the PGURE object receives (alpha, sigma, mu) for its (alpha, mu, sigma) — SURVEY Q1
        // The 1-D subplex search re-probes points it has already visited (restarts re-evaluate the best vertex, shrink
        // steps land on earlier reflections): ~20 % of a frame's probes repeat an earlier lambda bit for bit.  The
        // reference's objective is deterministic, so a repeat returns the identical value there; serving repeats from
        // a per-frame memo reproduces exactly that (and saves the device pass).
        std::vector<std::pair<double, double>> memo;
        auto f = [&](double x) -> double {
            double v = 0;
            last = x; // PGURE::lambda is overwritten by every probe, repeated or not (pgure.hpp:128)
            for (const auto &m : memo)
                if (m.first == x)
                {
                    h->stats[19] += 1;
                    return m.second;
                }
            const int e = objective(h, x, alpha, sigma, mu, &v, nullptr);
            if (e && !err)
                err = e;
            memo.emplace_back(x, v);
            return v;
        };
        const int st = opt.minimize(f, start, std::sqrt(start));
        if (err)
            return err;
        if (st == SBPLX_INVALID_ARGS)
            return fail(PGS_ERR_OPT,
                        "lambda search cannot start from %g (zero initial step; the reference's NLopt call throws here)", start);
        lambda = last; // the LAST evaluated lambda, not the optimum (pgure.hpp:128,236)
    }
    {
        StageTimer tm(h, 7);
        if (!p.optimize_pgure)
            if ((rc = stage_count(h, h->cur_sl)))
                return rc;
        const uint32_t lt = t - h->fb;
        if (h->use_tile && p.optimize_pgure && !h->frame_fallback)
        { // output slice straight from the gather kernel (thresholds of the final lambda; no probe of this frame needed more
          // than two triplets, and the final lambda is one of the probes)
            k_thresh<<<cdiv(h->P, 256), 256, 0, h->st>>>(h->dHead, h->P, lambda, p.exp_weighting, h->dFth, h->dPartialE, h->dKpart, h->dNeedQ);
            LAUNCHED(h);
            k_tile_eval<1><<<dim3(1, h->tile_r, h->tile_c), 128, 0, h->st>>>(h->dBinStart, h->dEnt, h->dFac[0], h->dU0c, h->dFth, h->dU, h->dCnt,
                                                                             h->N, h->cur_sl, h->cur_uMax, h->dY + h->fsz * lt, nullptr);
            LAUNCHED(h);
        }
        else
        {
            if ((rc = launch_recon(h, 0, lambda, h->cur_sl)))
                return rc;
            k_finalize<<<std::min(cdiv(h->fsz, 256), h->sm_count * 8), 256, 0, h->st>>>(h->dAcc[0], h->dAccScale, h->dCnt, h->fsz * h->cur_sl,
                                                                                        h->fsz, h->cur_uMax, h->dY + h->fsz * lt);
            LAUNCHED(h);
        }
        const uint32_t nblk = h->fe - h->fb;
        h->est[lt + (size_t)nblk * 0] = lambda;
        h->est[lt + (size_t)nblk * 1] = alpha;
        h->est[lt + (size_t)nblk * 2] = mu;
        h->est[lt + (size_t)nblk * 3] = sigma;
        if (h->sinkY && (rc = stream_out_frame(h, t)))
            return rc;
    }
    CU(cudaGetLastError());
    return PGS_OK;
}

extern "C" int pguresvt_process(pguresvt_handle *h)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    g_err.clear();
    CU(cudaSetDevice(h->p.device));
    if (!h->uploaded)
        return fail(PGS_ERR_ARG, "pguresvt_process: no input uploaded");
    drop_prelaunched_noise(h);
    CU(cudaStreamSynchronize(h->st));
    resolve_timers(h, false); // whatever earlier probes recorded
    for (int i = 0; i < PGS_NSTATS; i++)
        h->stats[i] = 0;
    h->launches = 0;
    h->prefiltered = false;
    h->cur_t = -1;
    int rc = PGS_OK;
    {
        StageTimer total(h, 9);
        {
            StageTimer tm(h, 3);
            rc = prefilter(h);
        }
        for (uint32_t t = h->fb; t < h->fe && !rc; t++)
            rc = process_frame(h, t);
        if (!rc && cudaMemcpyAsync(h->dEst, h->est.data(), h->est.size() * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess)
            rc = fail(PGS_ERR_CUDA, "CUDA error copying the estimates to the device");
    }
    const std::string keep = g_err;
    const cudaError_t se = cudaStreamSynchronize(h->st);
    const int drc = stream_out_drain(h); // every streamed frame has landed in the caller's array
    resolve_timers(h, rc == PGS_OK);
    if (rc)
    {
        g_err = keep;
        return rc;
    }
    if (se != cudaSuccess)
        return fail(PGS_ERR_CUDA, "CUDA error %s at the end of pguresvt_process", cudaGetErrorString(se));
    if (drc)
        return drc;
    h->stats[0] = (double)h->launches;
    h->stats[11] = h->use_compact ? (double)((size_t)h->P * sizeof(double) * (64 * (size_t)h->nobj + (size_t)h->Rc * (h->m + 32)))
                                  : (double)(h->rec * (size_t)h->P * sizeof(double) * h->nobj);
    h->stats[21] = h->Rc;
    return PGS_OK;
}

extern "C" double *pguresvt_device_output(pguresvt_handle *h) { return h ? h->dY : nullptr; }
extern "C" double *pguresvt_device_estimates(pguresvt_handle *h) { return h ? h->dEst : nullptr; }

extern "C" int pguresvt_download(pguresvt_handle *h, double *Y_full, double *estimates_full)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->p.device));
    const uint32_t nblk = h->fe - h->fb;
    if (Y_full && Y_full != h->sinkY) // (a streamed block is already there)
        CU(cudaMemcpyAsync(Y_full + h->fsz * h->fb, h->dY, h->fsz * nblk * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (estimates_full) // (n_frames, 4) column-major
        for (int q = 0; q < 4; q++)
            for (uint32_t i = 0; i < nblk; i++)
                estimates_full[(h->fb + i) + (size_t)h->nframes * q] = h->est[i + (size_t)nblk * q];
    return PGS_OK;
}

extern "C" int pguresvt_get_stats(const pguresvt_handle *h, double *stats)
{
    if (!h || !stats)
        return fail(PGS_ERR_ARG, "null argument");
    for (int i = 0; i < PGS_NSTATS; i++)
        stats[i] = h->stats[i];
    return PGS_OK;
}

// ------------------------------------------------------------------------------------------------------
// one-shot entry points
// ------------------------------------------------------------------------------------------------------
// One-shot entry points = PGURESVT<T1,T2>() (pguresvt.hpp:17-172) including its fan-out (pguresvt.hpp:169 -> utils.hpp:108-168):
//  * the frames are partitioned into contiguous blocks exactly like pguresvt::parallel (tasksPerThread = ceil(n / workers)), one
//    block per CUDA device, each driven by its own host thread with its own handle; nothing is exchanged between devices (every
//    frame is an independent job, pguresvt.hpp:90-167) and every device writes its frames straight into the caller's Y;
//  * a device whose block does not fit in a quarter of its free HBM (frames + medians + outputs on top of the per-window
//    buffers) streams it in sub-blocks through ONE handle (pguresvt_retarget): while sub-block b is processed, a helper thread
//    stages sub-block b+1 into page-locked memory, and the denoised frames leave through pguresvt_stream_output while the
//    following frames' SVDs run (the reference keeps the whole sequence and its output in host RAM,
//    pguresvt.hpp:44-67).  PGURESVT_BLOCK_FRAMES forces a sub-block length (tests).
static void frame_block(uint32_t n_frames, int parts, int part, uint32_t &b, uint32_t &e)
{
    const uint64_t per = ((uint64_t)n_frames + parts - 1) / parts;
    b = (uint32_t)std::min<uint64_t>(per * part, n_frames);
    e = (uint32_t)std::min<uint64_t>((uint64_t)b + per, n_frames);
}

static int plan_gpus(const pguresvt_params *p, uint32_t n_frames, int n_visible)
{
    const int avail = std::max(1, n_visible - std::max(0, p->device));
    int n = p->n_gpus;
    if (n <= 0) // automatic: every visible device, but no device for fewer than 8 frames (context + window buffers cost more)
        n = std::min<int>(avail, std::max<uint32_t>(1, n_frames / 8));
    n = std::min(n, avail);
    n = std::min<int>(n, std::max<uint32_t>(1, n_frames));
    return std::max(1, n);
}

extern "C" int pguresvt_host_plan_gpus(const pguresvt_params *p, uint32_t n_frames, int n_visible)
{
    if (!p)
        return -1;
    if (n_visible < 0 && cudaGetDeviceCount(&n_visible) != cudaSuccess)
        n_visible = 0;
    return plan_gpus(p, n_frames, n_visible);
}

extern "C" int pguresvt_host_frame_block(uint32_t n_frames, int parts, int part, uint32_t *begin, uint32_t *end)
{
    if (parts < 1 || part < 0 || part >= parts || !begin || !end)
        return fail(PGS_ERR_ARG, "invalid partition %d of %d", part, parts);
    frame_block(n_frames, parts, part, *begin, *end);
    return PGS_OK;
}

// One handle per device is kept between one-shot calls (like a plan cache): creating a handle allocates and clears ~13 GB at
// 1024^2 and page-locks the staging buffers (~110 ms create + destroy, 8 % of a 32-frame call); a later call with the same
// frame size, dtype and parameters re-targets it instead.  pguresvt_release_cached() frees them; PGURESVT_NO_CACHE=1 disables.
static std::mutex g_cache_mx;
static std::vector<pguresvt_handle *> g_cache;

static bool same_config(const pguresvt_handle *h, int dtype, uint32_t N, const pguresvt_params &p)
{
    const pguresvt_params &q = h->p;
    return h->dtype == dtype && h->N == N && q.device == p.device && q.traj_length == p.traj_length && q.block_size == p.block_size &&
           q.block_overlap == p.block_overlap && q.motion_window == p.motion_window && q.median_size == p.median_size &&
           q.noise_method == p.noise_method && q.max_iter == p.max_iter && q.random_seed == p.random_seed &&
           q.optimize_pgure == p.optimize_pgure && q.exp_weighting == p.exp_weighting && q.motion_estimation == p.motion_estimation &&
           q.lambda_est == p.lambda_est && q.alpha_est == p.alpha_est && q.mu_est == p.mu_est && q.sigma_est == p.sigma_est &&
           q.tol == p.tol && q.eps1_mode == p.eps1_mode && q.svd_kernel == p.svd_kernel && q.rank_cache == p.rank_cache;
}
static pguresvt_handle *cache_take(int dtype, uint32_t N, const pguresvt_params &p)
{
    std::lock_guard<std::mutex> lk(g_cache_mx);
    for (size_t i = 0; i < g_cache.size(); i++)
        if (same_config(g_cache[i], dtype, N, p))
        {
            pguresvt_handle *h = g_cache[i];
            g_cache.erase(g_cache.begin() + i);
            return h;
        }
    return nullptr;
}
static void cache_put(pguresvt_handle *h)
{
    pguresvt_handle *old = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mx);
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i]->p.device == h->p.device)
            { // one handle per device
                old = g_cache[i];
                g_cache.erase(g_cache.begin() + i);
                break;
            }
        g_cache.push_back(h);
    }
    if (old)
        pguresvt_destroy(old);
}
extern "C" void pguresvt_release_cached(void)
{
    std::vector<pguresvt_handle *> all;
    {
        std::lock_guard<std::mutex> lk(g_cache_mx);
        all.swap(g_cache);
    }
    for (pguresvt_handle *h : all)
        pguresvt_destroy(h);
}

static bool is_pinned(const void *ptr)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost)
        return true;
    cudaGetLastError();
    return false;
}

// frames [fb, fe) of the sequence on device p.device; called on the device's own host thread
static int run_device(int dtype, const void *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, pguresvt_params p, uint32_t fb,
                      uint32_t fe, double *Y, double *estimates)
{
    if (fb >= fe && n_frames > 0)
        return PGS_OK; // more devices than blocks
    const size_t fsz = (size_t)n_rows * n_cols, esz = dtype_size(dtype);
    uint32_t block = std::max<uint32_t>(1, fe - fb);
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && p.device >= 0 && p.device < ndev && cudaSetDevice(p.device) == cudaSuccess)
        {
            size_t freeb = 0, totb = 0;
            const size_t per_frame = fsz * (esz + sizeof(uint16_t) + sizeof(double));
            if (cudaMemGetInfo(&freeb, &totb) == cudaSuccess && per_frame > 0 && (size_t)block * per_frame > freeb / 4)
                block = (uint32_t)std::max<size_t>(1, (freeb / 4) / per_frame);
        }
        if (const char *e = getenv("PGURESVT_BLOCK_FRAMES"))
            if (atoi(e) > 0)
                block = std::min<uint32_t>(block, (uint32_t)atoi(e));
    }
    int rc = PGS_OK;
    static const bool trace = getenv("PGURESVT_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
    const auto t_begin = now();
    const uint32_t fe0 = (uint32_t)std::min<uint64_t>((uint64_t)fb + block, std::max(fe, fb));
    static const bool use_cache = !(getenv("PGURESVT_NO_CACHE") && atoi(getenv("PGURESVT_NO_CACHE")) > 0);
    pguresvt_handle *h = (use_cache && n_rows == n_cols) ? cache_take(dtype, n_rows, p) : nullptr;
    if (h)
    {
        if (p.random_seed < 0)
            h->perturbed = false; // fresh entropy per call (pgure.hpp:52-55)
        if (retarget_impl(h, n_frames, fb, fe0) != PGS_OK)
        { // too small for this call (or an invalid request: create reports it)
            pguresvt_destroy(h);
            h = nullptr;
        }
    }
    // (an empty sequence still goes through create once for its error message)
    if (!h)
        h = create_handle(dtype, n_rows, n_cols, n_frames, &p, fb, n_frames ? fe0 : 0, &rc);
    if (!h)
        return rc ? rc : PGS_ERR_ARG;
    if (trace)
        fprintf(stderr, "[pguresvt] device %d frames [%u,%u) sub-block %u: create %.1f ms\n", p.device, fb, fe, block, ms_since(t_begin));
    const bool x_pinned = is_pinned(X);
    // page-locked staging of the input sub-blocks (two buffers: one in use by the device, one being filled)
    void **stg = h->stg;
    const size_t stg_bytes = fsz * esz * h->cap_res;
    std::thread stager;
    auto stage = [&](int buf, uint32_t r0, uint32_t r1) { memcpy(stg[buf], (const char *)X + fsz * esz * r0, fsz * esz * (r1 - r0)); };
    auto resident = [&](uint32_t b0, uint32_t b1, uint32_t &r0, uint32_t &r1) {
        r0 = window_start(h, b0);
        r1 = window_start(h, b1 - 1) + h->win;
    };
    if (!x_pinned && h->stg_bytes < stg_bytes)
    {
        for (int i = 0; i < 2 && !rc; i++)
        {
            if (stg[i])
                cudaFreeHost(stg[i]);
            stg[i] = nullptr;
            if (cudaMallocHost(&stg[i], stg_bytes) != cudaSuccess)
                rc = fail(PGS_ERR_CUDA, "cannot allocate %zu bytes of page-locked staging memory", stg_bytes);
        }
        h->stg_bytes = rc ? 0 : stg_bytes;
    }
    if (!rc)
        rc = pguresvt_stream_output(h, Y);
    int buf = 0;
    if (!rc && !x_pinned)
    {
        uint32_t r0, r1;
        resident(fb, fe0, r0, r1);
        stage(0, r0, r1);
    }
    for (uint32_t b0 = fb; b0 < fe && !rc; b0 += block, buf ^= 1)
    {
        const uint32_t b1 = (uint32_t)std::min<uint64_t>((uint64_t)b0 + block, fe);
        if (b0 != fb)
            rc = pguresvt_retarget(h, b0, b1);
        if (stager.joinable())
            stager.join(); // this sub-block's frames are staged
        if (rc)
            break;
        const uint32_t n0 = b1, n1 = (uint32_t)std::min<uint64_t>((uint64_t)b1 + block, fe);
        if (!x_pinned && n0 < fe)
        { // stage the next sub-block while this one is processed (its buffer was last read by the upload before the previous one)
            uint32_t r0, r1;
            resident(n0, n1, r0, r1);
            stager = std::thread(stage, buf ^ 1, r0, r1);
        }
        const auto t_blk = now();
        if (x_pinned)
            rc = pguresvt_upload(h, X);
        else // pguresvt_upload takes the address frame 0 would have
            rc = pguresvt_upload(h, (const char *)stg[buf] - fsz * esz * h->r0);
        if (!rc)
            rc = pguresvt_process(h);
        if (!rc)
            rc = pguresvt_download(h, nullptr, estimates);
        if (trace)
            fprintf(stderr, "[pguresvt] device %d sub-block [%u,%u): %.1f ms (device timeline %.1f ms)\n", p.device, b0, b1, ms_since(t_blk),
                    h->stats[9]);
    }
    if (stager.joinable())
        stager.join();
    const std::string keep = g_err;
    const auto t_end = now();
    pguresvt_stream_output(h, nullptr);
    if (use_cache && rc == PGS_OK)
        cache_put(h);
    else
        pguresvt_destroy(h);
    if (trace)
        fprintf(stderr, "[pguresvt] device %d: destroy %.1f ms, total %.1f ms\n", p.device, ms_since(t_end), ms_since(t_begin));
    g_err = keep;
    return rc;
}

static int run_any(int dtype, const void *X, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, const pguresvt_params *p,
                   double *Y, double *estimates)
{
    if (!X || !Y || !estimates || !p)
        return fail(PGS_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PGS_ERR_CUDA, "no CUDA device available (the PGURE-SVT hot path has no CPU fallback)");
    const int ng = plan_gpus(p, n_frames, ndev);
    if (ng == 1)
        return run_device(dtype, X, n_rows, n_cols, n_frames, *p, 0, n_frames, Y, estimates);
    std::vector<int> rcs(ng, PGS_OK);
    std::vector<std::string> errs(ng);
    std::vector<std::thread> th;
    for (int g = 0; g < ng; g++)
        th.emplace_back([&, g]() {
            pguresvt_params pg = *p;
            pg.device = p->device + g;
            uint32_t b, e;
            frame_block(n_frames, ng, g, b, e);
            rcs[g] = run_device(dtype, X, n_rows, n_cols, n_frames, pg, b, e, Y, estimates);
            if (rcs[g])
                errs[g] = g_err; // (thread-local)
        });
    for (auto &t : th)
        t.join();
    for (int g = 0; g < ng; g++)
        if (rcs[g])
        {
            g_err = "device " + std::to_string(p->device + g) + ": " + errs[g];
            return rcs[g];
        }
    return PGS_OK;
}
extern "C" int pguresvt_run_u8(const uint8_t *X, uint32_t r, uint32_t c, uint32_t f, const pguresvt_params *p, double *Y, double *e)
{
    return run_any(PGS_U8, X, r, c, f, p, Y, e);
}
extern "C" int pguresvt_run_u16(const uint16_t *X, uint32_t r, uint32_t c, uint32_t f, const pguresvt_params *p, double *Y,
                                double *e)
{
    return run_any(PGS_U16, X, r, c, f, p, Y, e);
}
extern "C" int pguresvt_run_f32(const float *X, uint32_t r, uint32_t c, uint32_t f, const pguresvt_params *p, double *Y, double *e)
{
    return run_any(PGS_F32, X, r, c, f, p, Y, e);
}
extern "C" int pguresvt_run_f64(const double *X, uint32_t r, uint32_t c, uint32_t f, const pguresvt_params *p, double *Y, double *e)
{
    return run_any(PGS_F64, X, r, c, f, p, Y, e);
}

// ------------------------------------------------------------------------------------------------------
// stage probes
// ------------------------------------------------------------------------------------------------------
#define CHECK_T(h, t)                                                                        \
    if (!(h))                                                                                \
        return fail(PGS_ERR_ARG, "null handle");                                             \
    if ((t) < (h)->fb || (t) >= (h)->fe)                                                     \
        return fail(PGS_ERR_ARG, "frame %u outside the handle's block [%u,%u)", (t), (h)->fb, (h)->fe); \
    CU(cudaSetDevice((h)->p.device));

extern "C" int pguresvt_probe_median(pguresvt_handle *h, uint32_t t, uint16_t *Z)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    if (t < h->r0 || t >= h->r1)
        return fail(PGS_ERR_ARG, "frame %u not resident", t);
    if (h->p.median_size <= 0)
        return fail(PGS_ERR_ARG, "median prefilter is disabled");
    CU(cudaSetDevice(h->p.device));
    int rc = prefilter(h);
    if (rc)
        return rc;
    CU(cudaMemcpy(Z, h->dZ + h->fsz * (t - h->r0), h->fsz * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    return PGS_OK;
}

extern "C" int pguresvt_probe_arps(pguresvt_handle *h, uint32_t t, int32_t *patches)
{
    CHECK_T(h, t);
    int rc = prepare_frame(h, t);
    if (rc)
        return rc;
    std::vector<short2> hp((size_t)h->win * h->vecSize);
    CU(cudaMemcpy(hp.data(), h->dPos, hp.size() * sizeof(short2), cudaMemcpyDeviceToHost));
    for (uint32_t k = 0; k < h->win; k++)
        for (int it = 0; it < h->vecSize; it++)
        {
            const short2 q = hp[(size_t)k * h->vecSize + it];
            patches[0 + 2 * ((size_t)it + (size_t)h->vecSize * k)] = q.x;
            patches[1 + 2 * ((size_t)it + (size_t)h->vecSize * k)] = q.y;
        }
    return PGS_OK;
}

extern "C" int pguresvt_probe_singular_values(pguresvt_handle *h, uint32_t t, int obj, double *S, int64_t *n_patches)
{
    CHECK_T(h, t);
    const bool lean_obj = (h->lean && (obj == 2 || obj == 3)) || (h->top1_all && obj == 0);
    if (obj < 0 || obj > 3 || !(h->dFac[obj] || h->dSc[obj] || lean_obj))
        return fail(PGS_ERR_ARG, "SVT object %d not present in this configuration", obj);
    int rc = prepare_frame(h, t);
    if (rc)
        return rc;
    if (lean_obj && (rc = ensure_full(h))) // the lean mode keeps no factors of the perturbed objects: decompose them in full
        return rc;
    if (n_patches)
        *n_patches = h->P;
    if (S && h->use_compact)
        CU(cudaMemcpy2D(S, (size_t)h->n * sizeof(double), h->dSc[obj], 32 * sizeof(double), (size_t)h->n * sizeof(double), h->P,
                        cudaMemcpyDeviceToHost));
    else if (S)
    {
        const size_t soff = (size_t)h->m * h->n + (size_t)h->ldv * h->n;
        CU(cudaMemcpy2D(S, (size_t)h->n * sizeof(double), h->dFac[obj] + soff, h->rec * sizeof(double), (size_t)h->n * sizeof(double),
                        h->P, cudaMemcpyDeviceToHost));
        for (int q = 0; q < h->P; q++) // the 4-lane kernel keeps slot order; report LAPACK's descending order
            std::sort(S + (size_t)q * h->n, S + (size_t)(q + 1) * h->n, [](double a, double b) { return a > b; });
    }
    return PGS_OK;
}

extern "C" int pguresvt_probe_pgure(pguresvt_handle *h, uint32_t t, double alpha, double mu, double sigma, int n,
                                    const double *lambdas, double *values, double *terms)
{
    CHECK_T(h, t);
    if (!h->p.optimize_pgure)
        return fail(PGS_ERR_ARG, "handle was created with optimize_pgure = false");
    int rc = prepare_frame(h, t);
    if (rc)
        return rc;
    for (int i = 0; i < n; i++)
        if ((rc = objective(h, lambdas[i], alpha, sigma, mu, &values[i], terms ? terms + 5 * i : nullptr)))
            return rc;
    return PGS_OK;
}

extern "C" int pguresvt_probe_reconstruct(pguresvt_handle *h, uint32_t t, double lambda, double *v)
{
    CHECK_T(h, t);
    int rc = prepare_frame(h, t);
    if (rc)
        return rc;
    const size_t wtot = h->fsz * h->win;
    if (!h->p.optimize_pgure)
        if ((rc = stage_count(h, -1)))
            return rc;
    if (h->top1_all && (rc = ensure_full(h))) // the whole-window reconstruction reads complete records of object U
        return rc;
    if ((rc = launch_recon(h, 0, lambda, -1)))
        return rc;
    if (!h->dV)
        CU(cudaMalloc(&h->dV, wtot * sizeof(double)));
    k_finalize<<<std::min(cdiv(wtot, 256), h->sm_count * 8), 256, 0, h->st>>>(h->dAcc[0], h->dAccScale, h->dCnt, 0, wtot, 1.0, h->dV);
    LAUNCHED(h);
    CU(cudaMemcpyAsync(v, h->dV, wtot * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return PGS_OK;
}

extern "C" int pguresvt_probe_perturbations(pguresvt_handle *h, int8_t *delta1, int8_t *delta2neg)
{
    if (!h)
        return fail(PGS_ERR_ARG, "null handle");
    if (!h->p.optimize_pgure)
        return fail(PGS_ERR_ARG, "handle was created with optimize_pgure = false");
    CU(cudaSetDevice(h->p.device));
    int rc = perturb(h);
    if (rc)
        return rc;
    const size_t wtot = h->fsz * h->win;
    CU(cudaStreamSynchronize(h->st));
    CU(cudaMemcpy(delta1, h->dD1, wtot, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(delta2neg, h->dD2, wtot, cudaMemcpyDeviceToHost));
    return PGS_OK;
}

extern "C" int pguresvt_probe_noise(pguresvt_handle *h, uint32_t t, double *alpha, double *mu, double *sigma)
{
    CHECK_T(h, t);
    drop_prelaunched_noise(h);
    int rc = prefilter(h);
    if (rc)
        return rc;
    h->cur_t = -1;
    if ((rc = stage_window(h, t)))
        return rc;
    return noise_estimate_window(h->noise_ws, h->dU, (int)h->N, (int)h->win, (int)h->p.noise_method, h->sm_count, h->st, *alpha, *mu,
                                 *sigma, &h->launches, g_err, (long long)h->cur_a, h->cur_uMax);
}

extern "C" int pguresvt_probe_window_sum(pguresvt_handle *h, uint32_t t, double *sum)
{
    CHECK_T(h, t);
    if (!sum)
        return fail(PGS_ERR_ARG, "null argument");
    drop_prelaunched_noise(h);
    int rc = prefilter(h);
    if (rc)
        return rc;
    h->cur_t = -1;
    if ((rc = stage_window(h, t)) || (rc = sum_u_launch(h)) || (rc = sum_u_wait(h)))
        return rc;
    *sum = h->cur_sumU;
    return PGS_OK;
}

extern "C" int pguresvt_device_info(int device, char *name, int len)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    {
        fail(PGS_ERR_CUDA, "no CUDA device %d", device);
        return -1;
    }
    if (name && len > 0)
    {
        strncpy(name, prop.name, (size_t)len - 1);
        name[len - 1] = 0;
    }
    return prop.multiProcessorCount;
}

// Host optimiser exposed for the CPU-only tests (the same object the lambda search uses).
extern "C" int pguresvt_host_sbplx(double (*f)(double, void *), void *data, double x0, double lb, double ub, double step,
                                   double ftol_rel, double xtol_abs, int maxeval, double *xbest, double *fbest, int *nevals)
{
    Sbplx1D opt;
    opt.lb = lb;
    opt.ub = ub;
    opt.ftol_rel = ftol_rel;
    opt.xtol_abs = xtol_abs;
    opt.maxeval = maxeval;
    auto fn = [&](double x) { return f(x, data); };
    const int st = opt.minimize(fn, x0, step);
    if (xbest)
        *xbest = opt.xbest;
    if (fbest)
        *fbest = opt.fbest;
    if (nevals)
        *nevals = opt.nevals;
    return st;
}

// out[r][c][f] = in[f][c][r]: the axis reversal svt.py:329 applies to the bridge's (frames, cols, rows) result
// (`np.transpose(X, (2, 1, 0))` made contiguous), cache-blocked and spread over host threads — numpy's strided copy moves
// ~0.3 GB/s, which made SVT.denoise a third slower than the call underneath it.
extern "C" int pguresvt_host_transpose_f64(const double *in, uint32_t nf, uint32_t nc, uint32_t nr, double *out, int n_threads)
{
    if (!in || !out)
        return fail(PGS_ERR_ARG, "null argument");
    if (n_threads <= 0)
        n_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const uint32_t B = 32;
    auto work = [&](uint32_t c0, uint32_t c1) {
        for (uint32_t c = c0; c < c1; c++)
            for (uint32_t fb = 0; fb < nf; fb += B)
                for (uint32_t rb = 0; rb < nr; rb += B)
                {
                    const uint32_t fe = std::min(nf, fb + B), re = std::min(nr, rb + B);
                    for (uint32_t r = rb; r < re; r++)
                    {
                        double *o = out + ((size_t)r * nc + c) * nf;
                        const double *i0 = in + (size_t)c * nr + r;
                        for (uint32_t f = fb; f < fe; f++)
                            o[f] = i0[(size_t)f * nc * nr];
                    }
                }
    };
    n_threads = (int)std::min<uint32_t>((uint32_t)n_threads, std::max(1u, nc));
    std::vector<std::thread> th;
    const uint32_t per = (nc + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++)
    {
        const uint32_t c0 = std::min(nc, (uint32_t)t * per), c1 = std::min(nc, c0 + per);
        if (c0 < c1)
            th.emplace_back(work, c0, c1);
    }
    for (auto &t : th)
        t.join();
    return PGS_OK;
}

// number of patches and the sorted patch-id set of SVT::Decompose (svt.hpp:61-97) — host logic, no GPU needed
extern "C" int64_t pguresvt_host_patch_ids(uint32_t N, uint32_t bs, uint32_t bo, int32_t *out, int64_t cap)
{
    std::vector<int> ids = patch_ids((int)N, (int)bs, (int)bo);
    if (out)
        for (int64_t i = 0; i < (int64_t)ids.size() && i < cap; i++)
            out[i] = ids[i];
    return (int64_t)ids.size();
}

// Roofline denominator measured in the same job as the bench: FP64 DFMA throughput of the vector pipe, 8
// independent chains per thread, sm_count * 16 CTAs of 256 threads; burst = best of 5 launches, sustained = ~1 s back to back.
__global__ void k_dfma_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++)
    {
        a0 = fma(a0, b, c), a1 = fma(a1, b, c), a2 = fma(a2, b, c), a3 = fma(a3, b, c);
        a4 = fma(a4, b, c), a5 = fma(a5, b, c), a6 = fma(a6, b, c), a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
extern "C" int pguresvt_bench_dfma(int device, double *tflops_burst, double *tflops_sustained)
{
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 16, iters = 20000;
    double *out = nullptr;
    CU(cudaMalloc(&out, (size_t)grid * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const double flop = 2.0 * 8 * iters * (double)grid * 256;
    double best = 0, sus = 0;
    for (int rep = 0; rep < 6; rep++)
    {
        cudaEventRecord(e0);
        k_dfma_peak<<<grid, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0)
            best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    {
        const int n = 30;
        cudaEventRecord(e0);
        for (int i = 0; i < n; i++)
            k_dfma_peak<<<grid, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0)
            sus = flop * n / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CU(cudaGetLastError());
    if (tflops_burst)
        *tflops_burst = best;
    if (tflops_sustained)
        *tflops_sustained = sus;
    return PGS_OK;
}

#include "hotpixel.cuh"
extern "C" int pguresvt_hotpixel_u16(uint16_t *seq, uint32_t n_rows, uint32_t n_cols, uint32_t n_frames, double threshold,
                                     int device)
{
    if (!seq)
        return fail(PGS_ERR_ARG, "null argument");
    return hotpixel_filter_u16(seq, n_rows, n_cols, n_frames, threshold, device, g_err);
}
