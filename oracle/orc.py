"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes loader for oracle/liboracle.so and oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package never does.  See oracle/oracle.cpp for the restatement and its pinning status.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)


class OrcParams(C.Structure):
    _fields_ = [
        ("trajLength", C.c_uint32),
        ("blockSize", C.c_uint32),
        ("blockOverlap", C.c_uint32),
        ("motionWindow", C.c_uint32),
        ("medianSize", C.c_int64),
        ("noiseMethod", C.c_uint32),
        ("maxIter", C.c_uint32),
        ("nJobs", C.c_int64),
        ("randomSeed", C.c_int64),
        ("optimizePGURE", C.c_int32),
        ("expWeighting", C.c_int32),
        ("motionEstimation", C.c_int32),
        ("lambdaEst", C.c_double),
        ("alphaEst", C.c_double),
        ("muEst", C.c_double),
        ("sigmaEst", C.c_double),
        ("tol", C.c_double),
    ]


def build(ref=None):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _find_lapack():
    try:
        import scipy

        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        cands = [c for c in cands if "64_" not in os.path.basename(c)]
        if cands:
            return cands[0]
    except Exception:
        pass
    return None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_set_lapack.argtypes = [C.c_char_p]
    L.orc_set_lapack.restype = C.c_int
    L.orc_get_svd_backend.restype = C.c_int
    L.orc_arps.restype = C.c_longlong
    L.orc_svt_new.restype = C.c_void_p
    L.orc_svt_npatches.restype = C.c_int64
    L.orc_svt_npatches.argtypes = [C.c_void_p]
    L.orc_pgure_new.restype = C.c_void_p
    L.orc_pgure_calc.restype = C.c_double
    L.orc_pgure_optimize.restype = C.c_double
    L.orc_sbplx_1d.restype = C.c_int
    for n in ("u8", "u16", "f32", "f64"):
        getattr(L, "orc_pguresvt_" + n).restype = C.c_uint32
    lp = _find_lapack()
    if lp is not None:
        L.orc_set_lapack(lp.encode())
    _LIB = L
    return L


def ref():
    """oracle/_ref/libpguresvt_ref.so (reference's own pcg + CTMF), or None if not built."""
    global _REF
    if _REF is not None:
        return _REF
    path = os.path.join(_HERE, "_ref", "libpguresvt_ref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            build(ref=True)
        else:
            return None
    _REF = C.CDLL(path)
    return _REF


def svd_backend():
    return "dgesdd" if lib().orc_get_svd_backend() == 1 else "jacobi"


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def pcg64_raw(seed, n, which="oracle"):
    out = np.zeros(n, dtype=np.uint64)
    L = lib() if which == "oracle" else ref()
    fn = L.orc_pcg64_raw if which == "oracle" else L.ref_pcg64_raw
    fn(C.c_int64(seed), _p(out, C.c_uint64), C.c_int(n))
    return out


def perturbations(seed, n, which="oracle"):
    d1 = np.zeros(n, dtype=np.int64)
    d2 = np.zeros(n, dtype=np.float64)
    L = lib() if which == "oracle" else ref()
    fn = L.orc_perturbations if which == "oracle" else L.ref_perturbations
    fn(C.c_int64(seed), C.c_int64(n), _p(d1, C.c_int64), _p(d2))
    return d1, d2


def median_u16(img, r, which="oracle"):
    """img: (rows, cols) uint16; returns (2r+1)^2 clamp-to-edge median."""
    src = np.asfortranarray(img, dtype=np.uint16)
    dst = np.zeros_like(src, order="F")
    if which == "oracle":
        lib().orc_median_u16(_p(src, C.c_uint16), _p(dst, C.c_uint16), C.c_int(src.shape[0]), C.c_int(src.shape[1]),
                             C.c_int(r))
    else:
        # reference call convention (pguresvt.hpp:75-76): width = Nx = n_cols, height = Ny = n_rows on col-major memory
        ref().ref_ctmf(_p(src, C.c_uint16), _p(dst, C.c_uint16), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_int(r))
    return dst


def arps(w, bs, time_iter, time_window, motion_window, n_images, estimate=True):
    """w: (N,N,Nt) float64 normalised window.  Returns patches int64 (2, vecSize, Nt), motions, n_cost_evals."""
    w = np.asfortranarray(w, dtype=np.float64)
    N, _, Nt = w.shape
    vs = (N - bs + 1) ** 2
    patches = np.zeros((2, vs, Nt), dtype=np.int64, order="F")
    motions = np.zeros((2, vs, Nt - 1), dtype=np.int64, order="F")
    nc = lib().orc_arps(_p(w), C.c_int(N), C.c_int(Nt), C.c_int(bs), C.c_int(time_iter), C.c_int(time_window),
                        C.c_int(motion_window), C.c_int(n_images), C.c_int(int(estimate)), _p(patches, C.c_int64),
                        _p(motions, C.c_int64))
    return patches, motions, nc


class SVTObj:
    def __init__(self, patches, N, Nt, bs, bo, expw):
        self.patches = np.asfortranarray(patches, dtype=np.int64)
        self.N, self.Nt, self.bs = N, Nt, bs
        self.h = C.c_void_p(lib().orc_svt_new(_p(self.patches, C.c_int64), C.c_int(N), C.c_int(Nt), C.c_int(bs),
                                              C.c_int(bo), C.c_int(int(expw))))

    def npatches(self):
        return int(lib().orc_svt_npatches(self.h))

    def patch_ids(self):
        out = np.zeros(self.npatches(), dtype=np.int64)
        lib().orc_svt_patch_ids(self.h, _p(out, C.c_int64))
        return out

    def decompose(self, u):
        self.u = np.asfortranarray(u, dtype=np.float64)
        lib().orc_svt_decompose(self.h, _p(self.u))

    def singular_values(self):
        S = np.zeros((self.npatches(), self.Nt), dtype=np.float64)
        lib().orc_svt_singular_values(self.h, _p(S))
        return S

    def reconstruct(self, lam):
        v = np.zeros((self.N, self.N, self.Nt), dtype=np.float64, order="F")
        lib().orc_svt_reconstruct(self.h, C.c_double(lam), _p(v))
        return v

    def __del__(self):
        try:
            lib().orc_svt_free(self.h)
        except Exception:
            pass


class PGUREObj:
    def __init__(self, U, patches, alpha, mu, sigma, bs, bo, seed, expw, opt=True):
        self.U = np.asfortranarray(U, dtype=np.float64)
        self.patches = np.asfortranarray(patches, dtype=np.int64)
        N, _, Nt = self.U.shape
        self.N, self.Nt = N, Nt
        self.h = C.c_void_p(lib().orc_pgure_new(_p(self.U), _p(self.patches, C.c_int64), C.c_int(N), C.c_int(Nt),
                                                C.c_double(alpha), C.c_double(mu), C.c_double(sigma), C.c_int(bs),
                                                C.c_int(bo), C.c_int64(seed), C.c_int(int(expw)), C.c_int(int(opt))))

    def calc(self, lam):
        terms = np.zeros(5)
        f = lib().orc_pgure_calc(self.h, C.c_double(lam), _p(terms))
        return float(f), terms

    def optimize(self, tol, start, bound, maxeval):
        ne, st = C.c_int(0), C.c_int(0)
        lam = lib().orc_pgure_optimize(self.h, C.c_double(tol), C.c_double(start), C.c_double(bound), C.c_int(maxeval),
                                       C.byref(ne), C.byref(st))
        return float(lam), ne.value, st.value

    def reconstruct(self, lam):
        v = np.zeros((self.N, self.N, self.Nt), dtype=np.float64, order="F")
        lib().orc_pgure_reconstruct(self.h, C.c_double(lam), _p(v))
        return v

    def __del__(self):
        try:
            lib().orc_pgure_free(self.h)
        except Exception:
            pass


def noise_estimate(u, method=4, alpha=-1.0, mu=-1.0, sigma=-1.0):
    u = np.asfortranarray(u, dtype=np.float64)
    N, _, T = u.shape
    a, m, s = C.c_double(alpha), C.c_double(mu), C.c_double(sigma)
    stats = np.zeros(3, dtype=np.int64)
    lib().orc_noise_estimate(_p(u), C.c_int(N), C.c_int(T), C.c_int(method), C.byref(a), C.byref(m), C.byref(s),
                             _p(stats, C.c_longlong))
    return a.value, m.value, s.value, stats


def quadtree_counts(N, mode=0):
    c, k = C.c_int(0), C.c_int(0)
    lib().orc_quadtree_counts(C.c_int(N), C.c_int(mode), C.byref(c), C.byref(k))
    return c.value, k.value


def hotpixel_u16(seq, threshold):
    s = np.asfortranarray(seq, dtype=np.uint16).copy(order="F")
    lib().orc_hotpixel_u16(_p(s, C.c_uint16), C.c_int(s.shape[0]), C.c_int(s.shape[1]), C.c_int(s.shape[2]),
                           C.c_double(threshold))
    return s


_OBJ = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def sbplx_1d(f, x0, lb, ub, step, ftol_rel=1e-7, xtol_abs=1e-12, maxeval=500):
    trace = []

    def _f(x, _):
        v = float(f(x))
        trace.append((x, v))
        return v

    cb = _OBJ(_f)
    xo, fo, ne = C.c_double(0), C.c_double(0), C.c_int(0)
    st = lib().orc_sbplx_1d(cb, None, C.c_double(x0), C.c_double(lb), C.c_double(ub), C.c_double(step),
                            C.c_double(ftol_rel), C.c_double(xtol_abs), C.c_int(maxeval), C.byref(xo), C.byref(fo),
                            C.byref(ne))
    return dict(x=xo.value, minf=fo.value, nevals=ne.value, status=st, trace=trace)


_ENTRY = {np.dtype("uint8"): ("orc_pguresvt_u8", C.c_uint8), np.dtype("uint16"): ("orc_pguresvt_u16", C.c_uint16),
          np.dtype("float32"): ("orc_pguresvt_f32", C.c_float), np.dtype("float64"): ("orc_pguresvt_f64", C.c_double)}


def pguresvt(X, trajectory_length=15, patch_size=4, patch_overlap=1, motion_window=7, motion_filter=5, noise_method=4,
             max_iter=500, n_jobs=-1, random_seed=-1, optimize_pgure=True, exponential_weighting=True,
             motion_estimation=True, lambda1=0.0, noise_alpha=-1.0, noise_mu=-1.0, noise_sigma=-1.0, tol=1e-7,
             frame_begin=0, frame_end=None):
    """Full-pipeline oracle.  X: (rows, cols, frames).  Returns Y (rows, cols, frames) F-order, estimates (frames,4)."""
    X = np.asfortranarray(X)
    name, ct = _ENTRY[X.dtype]
    nr, nc, nf = X.shape
    p = OrcParams(trajectory_length, patch_size, patch_overlap, motion_window, motion_filter, noise_method, max_iter,
                  n_jobs, random_seed, int(optimize_pgure), int(exponential_weighting), int(motion_estimation),
                  lambda1, noise_alpha, noise_mu, noise_sigma, tol)
    Y = np.zeros((nr, nc, nf), dtype=np.float64, order="F")
    est = np.zeros((nf, 4), dtype=np.float64, order="F")
    if frame_end is None:
        frame_end = nf
    getattr(lib(), name)(_p(X, ct), C.c_uint32(nr), C.c_uint32(nc), C.c_uint32(nf), C.byref(p), _p(Y), _p(est),
                         C.c_uint32(frame_begin), C.c_uint32(frame_end))
    return Y, est


def stage_times(reset=True):
    out = np.zeros(8)
    lib().orc_stage_times(_p(out), C.c_int(int(reset)))
    return dict(zip(["median", "noise", "arps", "svd", "optimize", "reconstruct"], out[:6]))
