"""ctypes bridge to libpguresvt_b200.so — the role of the reference's Cython module
pguresvt/_pguresvt.pyx:151-384: four dtype entry points `pguresvt_u8/u16/f/d(input_images, …)` that
return `(X, estimates, result)` with X shaped (frames, cols, rows) C-order and estimates (4, frames).

Differences from the reference bridge, by design: outputs are allocated by numpy and filled through the
C ABI (the reference steals Armadillo's buffers, _pguresvt.pyx:98-138); errors come back as return codes
and are raised as RuntimeError instead of terminating the process.
"""
import ctypes as C
import os

import numpy as np

_LIBNAME = "libpguresvt_b200.so"
_lib = None


class Params(C.Structure):
    """struct pguresvt_params (include/pguresvt_b200.h)."""

    _fields_ = [
        ("traj_length", C.c_uint32),
        ("block_size", C.c_uint32),
        ("block_overlap", C.c_uint32),
        ("motion_window", C.c_uint32),
        ("median_size", C.c_int64),
        ("noise_method", C.c_uint32),
        ("max_iter", C.c_uint32),
        ("n_jobs", C.c_int64),
        ("random_seed", C.c_int64),
        ("optimize_pgure", C.c_int32),
        ("exp_weighting", C.c_int32),
        ("motion_estimation", C.c_int32),
        ("lambda_est", C.c_double),
        ("alpha_est", C.c_double),
        ("mu_est", C.c_double),
        ("sigma_est", C.c_double),
        ("tol", C.c_double),
        ("device", C.c_int32),
        ("eps1_mode", C.c_int32),
        ("svd_kernel", C.c_int32),
        ("rank_cache", C.c_int32),
        ("n_gpus", C.c_int32),
    ]


DTYPES = {np.dtype("uint8"): 0, np.dtype("uint16"): 1, np.dtype("float32"): 2, np.dtype("float64"): 3}
NSTATS = 24
STAT_NAMES = ["launches", "svds", "evals", "ms_median", "ms_arps", "ms_svd", "ms_search", "ms_final", "ms_noise",
              "ms_total", "svd_sweeps", "factor_bytes", "sweeps_obj0", "sweeps_warm", "arps_pairs_computed", "arps_pairs_reused", "eval_triplets", "ms_search_prep", "eval_redone", "evals_memoized",
              "overflow_patches", "rank_cache", "lean_checks", "lean_exact_svds"]


def lib_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), _LIBNAME)


def load():
    """Load the CUDA library; raise loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{_LIBNAME} not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The PGURE-SVT hot path has no CPU fallback.")
    L = C.CDLL(path)
    vp, u32, dp = C.c_void_p, C.c_uint32, C.POINTER(C.c_double)
    pp = C.POINTER(Params)
    for name in ("pguresvt_run_u8", "pguresvt_run_u16", "pguresvt_run_f32", "pguresvt_run_f64"):
        fn = getattr(L, name)
        fn.argtypes = [vp, u32, u32, u32, pp, dp, dp]
        fn.restype = C.c_int
    L.pguresvt_last_error.restype = C.c_char_p
    L.pguresvt_release_cached.argtypes = []
    L.pguresvt_release_cached.restype = None
    L.pguresvt_create.argtypes = [C.c_int, u32, u32, u32, pp, u32, u32]
    L.pguresvt_create.restype = vp
    L.pguresvt_destroy.argtypes = [vp]
    L.pguresvt_destroy.restype = None
    L.pguresvt_resident_range.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.pguresvt_upload.argtypes = [vp, vp]
    L.pguresvt_upload_device.argtypes = [vp, vp]
    L.pguresvt_process.argtypes = [vp]
    L.pguresvt_retarget.argtypes = [vp, u32, u32]
    L.pguresvt_stream_output.argtypes = [vp, dp]
    L.pguresvt_host_plan_gpus.argtypes = [pp, u32, C.c_int]
    L.pguresvt_host_frame_block.argtypes = [u32, C.c_int, C.c_int, C.POINTER(u32), C.POINTER(u32)]
    L.pguresvt_device_output.argtypes = [vp]
    L.pguresvt_device_output.restype = vp
    L.pguresvt_device_estimates.argtypes = [vp]
    L.pguresvt_device_estimates.restype = vp
    L.pguresvt_download.argtypes = [vp, dp, dp]
    L.pguresvt_get_stats.argtypes = [vp, dp]
    L.pguresvt_probe_median.argtypes = [vp, u32, C.POINTER(C.c_uint16)]
    L.pguresvt_probe_arps.argtypes = [vp, u32, C.POINTER(C.c_int32)]
    L.pguresvt_probe_singular_values.argtypes = [vp, u32, C.c_int, dp, C.POINTER(C.c_int64)]
    L.pguresvt_probe_pgure.argtypes = [vp, u32, C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, dp]
    L.pguresvt_probe_reconstruct.argtypes = [vp, u32, C.c_double, dp]
    L.pguresvt_probe_perturbations.argtypes = [vp, C.POINTER(C.c_int8), C.POINTER(C.c_int8)]
    L.pguresvt_probe_noise.argtypes = [vp, u32, dp, dp, dp]
    L.pguresvt_probe_window_sum.argtypes = [vp, u32, dp]
    L.pguresvt_hotpixel_u16.argtypes = [C.POINTER(C.c_uint16), u32, u32, u32, C.c_double, C.c_int]
    L.pguresvt_bench_dfma.argtypes = [C.c_int, dp, dp]
    L.pguresvt_host_transpose_f64.argtypes = [dp, u32, u32, u32, dp, C.c_int]
    L.pguresvt_device_info.argtypes = [C.c_int, C.c_char_p, C.c_int]
    L.pguresvt_host_patch_ids.argtypes = [u32, u32, u32, C.POINTER(C.c_int32), C.c_int64]
    L.pguresvt_host_patch_ids.restype = C.c_int64
    _lib = L
    return L


def plan_gpus(n_frames, n_visible=-1, **kw):
    """Number of devices a one-shot call with these parameters fans out over (host logic, no GPU needed with n_visible >= 0)."""
    p = make_params(**kw)
    return load().pguresvt_host_plan_gpus(C.byref(p), int(n_frames), int(n_visible))


def frame_block(n_frames, parts, part):
    b, e = C.c_uint32(0), C.c_uint32(0)
    check(load().pguresvt_host_frame_block(int(n_frames), int(parts), int(part), C.byref(b), C.byref(e)), "frame_block")
    return b.value, e.value


def reversed_axes_copy(X):
    """C-contiguous copy of `np.transpose(X, (2, 1, 0))` for a C-contiguous float64 (frames, cols, rows) array — what svt.py:329
    does with the bridge's result, through the library's threaded blocked transpose."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    nf, nc, nr = X.shape
    out = np.empty((nr, nc, nf), dtype=np.float64)
    dp = C.POINTER(C.c_double)
    check(load().pguresvt_host_transpose_f64(X.ctypes.data_as(dp), nf, nc, nr, out.ctypes.data_as(dp), 0), "transpose")
    return out


def release_cached():
    """Free the per-device handles the one-shot entry points keep between calls."""
    load().pguresvt_release_cached()


def last_error():
    return load().pguresvt_last_error().decode()


def check(rc, what="pguresvt"):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def make_params(trajectory_length=15, patch_size=4, patch_overlap=1, motion_window=7, motion_filter=5, noise_method=4,
                max_iter=500, n_jobs=-1, random_seed=-1, optimize_pgure=True, exponential_weighting=True,
                motion_estimation=True, lambda1=0.0, noise_alpha=-1.0, noise_mu=-1.0, noise_sigma=-1.0, tol=1e-7,
                device=None, eps1_mode=None, svd_kernel=None, rank_cache=None, n_gpus=None):
    if n_gpus is None:
        # one-shot calls fan out over every visible GPU (the role of n_jobs = -1 in the reference), except inside a
        # one-process-per-GPU launch (torchrun sets LOCAL_RANK), where every rank keeps to its own device
        n_gpus = int(os.environ.get("PGURESVT_NGPUS", "1" if "LOCAL_RANK" in os.environ else "0"))
    if device is None:
        device = int(os.environ.get("PGURESVT_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if eps1_mode is None:
        eps1_mode = int(os.environ.get("PGURESVT_EPS1_MODE", "0"))
    if svd_kernel is None:
        svd_kernel = int(os.environ.get("PGURESVT_SVD_KERNEL", "0"))
    if rank_cache is None:
        rank_cache = int(os.environ.get("PGURESVT_RANK_CACHE", "0"))
    return Params(int(trajectory_length), int(patch_size), int(patch_overlap), int(motion_window), int(motion_filter),
                  int(noise_method), int(max_iter), int(n_jobs), int(random_seed), int(bool(optimize_pgure)),
                  int(bool(exponential_weighting)), int(bool(motion_estimation)), float(lambda1), float(noise_alpha),
                  float(noise_mu), float(noise_sigma), float(tol), int(device), int(eps1_mode), int(svd_kernel), int(rank_cache),
                  int(n_gpus))


def _run(entry, dtype, input_images, **kw):
    L = load()
    X = input_images
    if X.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % X.ndim)
    if X.dtype != dtype:
        raise ValueError(f"Buffer dtype mismatch, expected '{np.dtype(dtype).name}' but got '{X.dtype.name}'")
    # the reference wraps X.data as an arma::Cube(shape[0], shape[1], shape[2]) without copying, i.e. it reads
    # the buffer as column-major (_pguresvt.pyx:72-93); svt.py always passes a Fortran-ordered array
    Xf = np.asfortranarray(X)
    nr, nc, nf = Xf.shape
    p = make_params(**kw)
    Y = np.empty((nf, nc, nr), dtype=np.float64)  # == column-major (rows, cols, frames)
    est = np.zeros((4, nf), dtype=np.float64)  # == column-major (frames, 4)
    rc = getattr(L, entry)(Xf.ctypes.data_as(C.c_void_p), nr, nc, nf, C.byref(p),
                           Y.ctypes.data_as(C.POINTER(C.c_double)), est.ctypes.data_as(C.POINTER(C.c_double)))
    check(rc, entry)
    return Y, est, rc


def pguresvt_u8(input_images, **kw):
    return _run("pguresvt_run_u8", np.uint8, input_images, **kw)


def pguresvt_u16(input_images, **kw):
    return _run("pguresvt_run_u16", np.uint16, input_images, **kw)


def pguresvt_f(input_images, **kw):
    return _run("pguresvt_run_f32", np.float32, input_images, **kw)


def pguresvt_d(input_images, **kw):
    return _run("pguresvt_run_f64", np.float64, input_images, **kw)


class Handle:
    """Thin RAII wrapper over the handle API (upload / process / download / probes)."""

    def __init__(self, X=None, shape=None, dtype=None, frame_begin=0, frame_end=None, **kw):
        L = load()
        self.L = L
        if X is not None:
            X = np.asfortranarray(X)
            shape, dtype = X.shape, X.dtype
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nr, nc, nf = self.shape
        self.params = make_params(**kw)
        fe = nf if frame_end is None else frame_end
        self.fb, self.fe = frame_begin, fe
        self.h = L.pguresvt_create(DTYPES[self.dtype], nr, nc, nf, C.byref(self.params), frame_begin, fe)
        if not self.h:
            raise RuntimeError("pguresvt_create failed: " + last_error())
        bs = self.params.block_size
        nt = bs * bs - 1 if bs * bs < self.params.traj_length else self.params.traj_length
        self.win = 2 * (nt // 2) + 1
        if X is not None:
            self.upload(X)

    def upload(self, X):
        X = np.asfortranarray(X)
        assert X.shape == self.shape and X.dtype == self.dtype
        self._X = X
        check(self.L.pguresvt_upload(self.h, X.ctypes.data_as(C.c_void_p)), "upload")

    def resident_range(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        check(self.L.pguresvt_resident_range(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def upload_device(self, ptr):
        check(self.L.pguresvt_upload_device(self.h, C.c_void_p(ptr)), "upload_device")

    def process(self):
        check(self.L.pguresvt_process(self.h), "process")

    def retarget(self, frame_begin, frame_end):
        check(self.L.pguresvt_retarget(self.h, frame_begin, frame_end), "retarget")
        self.fb, self.fe = frame_begin, frame_end

    def stream_output(self, Y):
        """Y: whole-sequence (rows, cols, frames) float64 F-order array the frames are streamed into during process()."""
        self._Ysink = Y
        ptr = None if Y is None else Y.ctypes.data_as(C.POINTER(C.c_double))
        check(self.L.pguresvt_stream_output(self.h, ptr), "stream_output")

    def device_output(self):
        return self.L.pguresvt_device_output(self.h)

    def device_estimates(self):
        return self.L.pguresvt_device_estimates(self.h)

    def download(self, Y=None, est=None):
        nr, nc, nf = self.shape
        if Y is None:
            Y = np.zeros((nr, nc, nf), dtype=np.float64, order="F")
        if est is None:
            est = np.zeros((nf, 4), dtype=np.float64, order="F")
        check(self.L.pguresvt_download(self.h, Y.ctypes.data_as(C.POINTER(C.c_double)),
                                       est.ctypes.data_as(C.POINTER(C.c_double))), "download")
        return Y, est

    def stats(self):
        s = np.zeros(NSTATS)
        check(self.L.pguresvt_get_stats(self.h, s.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(zip(STAT_NAMES, s))

    # ---- stage probes ----
    def probe_median(self, t):
        nr, nc, _ = self.shape
        Z = np.zeros((nr, nc), dtype=np.uint16, order="F")
        check(self.L.pguresvt_probe_median(self.h, t, Z.ctypes.data_as(C.POINTER(C.c_uint16))), "probe_median")
        return Z

    def probe_arps(self, t):
        nr = self.shape[0]
        vs = (nr - self.params.block_size + 1) ** 2
        p = np.zeros((2, vs, self.win), dtype=np.int32, order="F")
        check(self.L.pguresvt_probe_arps(self.h, t, p.ctypes.data_as(C.POINTER(C.c_int32))), "probe_arps")
        return p

    def probe_singular_values(self, t, obj=0):
        n = C.c_int64(0)
        check(self.L.pguresvt_probe_singular_values(self.h, t, obj, None, C.byref(n)), "probe_sv")
        S = np.zeros((n.value, self.win), dtype=np.float64)
        check(self.L.pguresvt_probe_singular_values(self.h, t, obj, S.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)),
              "probe_sv")
        return S

    def probe_pgure(self, t, alpha, mu, sigma, lambdas):
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        vals = np.zeros(len(lam))
        terms = np.zeros((len(lam), 5))
        dp = C.POINTER(C.c_double)
        check(self.L.pguresvt_probe_pgure(self.h, t, alpha, mu, sigma, len(lam), lam.ctypes.data_as(dp),
                                          vals.ctypes.data_as(dp), terms.ctypes.data_as(dp)), "probe_pgure")
        return vals, terms

    def probe_reconstruct(self, t, lam):
        nr, nc, _ = self.shape
        v = np.zeros((nr, nc, self.win), dtype=np.float64, order="F")
        check(self.L.pguresvt_probe_reconstruct(self.h, t, lam, v.ctypes.data_as(C.POINTER(C.c_double))), "probe_recon")
        return v

    def probe_perturbations(self):
        nr, nc, _ = self.shape
        n = nr * nc * self.win
        d1 = np.zeros(n, dtype=np.int8)
        d2 = np.zeros(n, dtype=np.int8)
        check(self.L.pguresvt_probe_perturbations(self.h, d1.ctypes.data_as(C.POINTER(C.c_int8)),
                                                  d2.ctypes.data_as(C.POINTER(C.c_int8))), "probe_perturbations")
        return d1, d2

    def probe_noise(self, t, alpha=-1.0, mu=-1.0, sigma=-1.0):
        a, m, s = C.c_double(alpha), C.c_double(mu), C.c_double(sigma)
        check(self.L.pguresvt_probe_noise(self.h, t, C.byref(a), C.byref(m), C.byref(s)), "probe_noise")
        return a.value, m.value, s.value

    def probe_window_sum(self, t):
        s = C.c_double(0)
        check(self.L.pguresvt_probe_window_sum(self.h, t, C.byref(s)), "probe_window_sum")
        return s.value

    def close(self):
        if getattr(self, "h", None):
            self.L.pguresvt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
