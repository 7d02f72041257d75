"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol include/*.h
declares, host logic (patch-id set, 1-D subplex) agrees with the oracle, and the Python mirror keeps the
reference's argument checks (pguresvt/tests/test_svt.py:109-164).  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import orc
from pguresvt import SVT, mixed_noise_model
from pguresvt import _pguresvt as bridge


def test_abi_exports_every_declared_symbol():
    L = bridge.load()
    hdr = open(os.path.join(ROOT, "include", "pguresvt_b200.h")).read()
    names = set(re.findall(r"\b(pguresvt_[a-z0-9_]+)\s*\(", hdr))
    names -= {"pguresvt_params", "pguresvt_handle"}
    assert len(names) >= 26
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_params_struct_layout_matches_header():
    # 4*u32, i64, 2*u32, 2*i64, 3*i32 (+pad), 5*f64, 4*i32
    assert C.sizeof(bridge.Params) == 16 + 8 + 8 + 16 + 16 + 40 + 16 + 8  # + n_gpus and tail padding
    assert bridge.Params.lambda_est.offset == 64


def test_no_gpu_fails_loudly_or_creates():
    L = bridge.load()
    p = bridge.make_params()
    h = L.pguresvt_create(1, 32, 32, 16, C.byref(p), 0, 16)
    if not h:
        assert "CUDA" in bridge.last_error()
    else:
        L.pguresvt_destroy(h)
    # argument errors are reported, not thrown
    assert not L.pguresvt_create(1, 32, 48, 16, C.byref(p), 0, 16)
    assert "square" in bridge.last_error()


@pytest.mark.parametrize("cfg", [(32, 4, 1), (32, 4, 2), (128, 16, 2), (256, 4, 2), (32, 4, 3), (64, 8, 3)])
def test_patch_ids_match_oracle(cfg):
    N, bs, bo = cfg
    L = bridge.load()
    n = L.pguresvt_host_patch_ids(N, bs, bo, None, 0)
    out = np.zeros(n, dtype=np.int32)
    L.pguresvt_host_patch_ids(N, bs, bo, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
    s = orc.SVTObj(np.zeros((2, (N - bs + 1) ** 2, 3), dtype=np.int64), N, 3, bs, bo, True)
    assert np.array_equal(out, s.patch_ids())


_OBJ = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def _host_sbplx(f, x0, lb, ub, step, ftol=1e-7, xtol=1e-12, maxeval=500):
    trace = []

    def g(x, _):
        v = float(f(x))
        trace.append((x, v))
        return v

    cb = _OBJ(g)
    L = bridge.load()
    L.pguresvt_host_sbplx.restype = C.c_int
    xb, fb, ne = C.c_double(0), C.c_double(0), C.c_int(0)
    st = L.pguresvt_host_sbplx(cb, None, C.c_double(x0), C.c_double(lb), C.c_double(ub), C.c_double(step),
                               C.c_double(ftol), C.c_double(xtol), C.c_int(maxeval), C.byref(xb), C.byref(fb), C.byref(ne))
    return dict(x=xb.value, minf=fb.value, nevals=ne.value, status=st, trace=trace)


@pytest.mark.parametrize("fn", [
    lambda x: (x - 0.3) ** 2 + 1.0,
    lambda x: 3.0 * x - 100.0,
    lambda x: -x,
    lambda x: np.cos(3 * x) + 0.1 * x,
    lambda x: 0.0045 + 1e-6 * (np.log1p(x) - 3.4) ** 2,
    lambda x: abs(x - 50.0),
])
@pytest.mark.parametrize("x0", [0.05, 5.0, 99.0])
def test_host_sbplx_follows_oracle_sequence(fn, x0):
    """The product's optimiser and the oracle's restatement of NLopt SBPLX probe the same points in the same
    order (lambda parity is path dependent: SURVEY H1)."""
    a = _host_sbplx(fn, x0, 0.0, 100.0, np.sqrt(x0))
    b = orc.sbplx_1d(fn, x0, 0.0, 100.0, np.sqrt(x0))
    assert a["status"] == b["status"] and a["nevals"] == b["nevals"]
    assert a["trace"] == b["trace"]
    assert a["x"] == b["x"] and a["minf"] == b["minf"]


def test_host_sbplx_maxeval_and_invalid():
    a = _host_sbplx(lambda x: (x - 3) ** 2, 1.0, 0.0, 100.0, 1.0, maxeval=7)
    b = orc.sbplx_1d(lambda x: (x - 3) ** 2, 1.0, 0.0, 100.0, 1.0, maxeval=7)
    assert a["status"] == b["status"] == 5 and a["nevals"] == b["nevals"] == 7 and a["trace"] == b["trace"]
    assert _host_sbplx(lambda x: x, 0.0, 0.0, 100.0, 0.0)["status"] == -2


# ---- the reference's TestErrors, verbatim in behaviour (test_svt.py:109-164) ----
class TestErrors:
    def setup_method(self, method):
        self.X = np.ones((32, 32, 16))

    def test_negative(self):
        with pytest.raises(ValueError, match="Negative values found in data"):
            SVT().denoise(-1 * self.X)

    def test_overlap(self):
        with pytest.raises(ValueError, match="Invalid patch_overlap parameter"):
            SVT(patch_size=4, patch_overlap=5).denoise(self.X)

    @pytest.mark.parametrize("tl", [0, 2, -1])
    def test_trajectory(self, tl):
        with pytest.raises(ValueError, match="Invalid trajectory_length parameter"):
            SVT(trajectory_length=tl).denoise(self.X)

    @pytest.mark.parametrize("mw", [1, 4])
    def test_motion_window(self, mw):
        with pytest.raises(ValueError, match="Invalid motion_window parameter"):
            SVT(motion_window=mw).denoise(self.X)

    def test_motion_filter(self):
        with pytest.raises(ValueError, match="Invalid motion_filter parameter"):
            SVT(motion_filter=1.5).denoise(self.X)

    @pytest.mark.parametrize("lam", [None, -1.0])
    def test_lambda(self, lam):
        with pytest.raises(ValueError, match="Invalid lambda1 parameter"):
            SVT(optimize_pgure=False, lambda1=lam).denoise(self.X)

    def test_square(self):
        with pytest.raises(ValueError, match="requires square images"):
            SVT().denoise(np.ones((32, 64, 16)))

    def test_power_of_two(self):
        with pytest.raises(ValueError, match="requires image dimensions 2"):
            SVT().denoise(np.ones((48, 48, 16)))

    def test_dtype(self):
        with pytest.raises(TypeError, match="Invalid dtype"):
            SVT(optimize_pgure=False, lambda1=1.0).denoise(self.X.astype(np.int32))


def test_mixed_noise_model_errors_and_seeds():
    X = np.random.RandomState(0).uniform(0, 255, size=(8, 8, 4))
    with pytest.raises(ValueError, match="alpha should be in range"):
        mixed_noise_model(X, alpha=-1.0)
    with pytest.raises(ValueError, match="sigma should be"):
        mixed_noise_model(X, sigma=-1.0)
    for rs in (None, 101, np.random.RandomState(101)):
        assert mixed_noise_model(X, random_state=rs).shape == X.shape
    assert np.array_equal(mixed_noise_model(X, random_state=5), mixed_noise_model(X, random_state=5))


def test_odd_even_jacobi_schedule_meets_every_pair_once():
    """Schedule of k_svd_warp (compact.cuh): round E pairs slots (2p, 2p+1), round O pairs (2p+1, 2p+2), and the two
    columns of a pair are written back exchanged.  In 32 rounds (16 E + 16 O) every pair of the 32 columns must meet
    exactly once and the column order must end up reversed (so two sweeps restore it) — the property that lets the
    kernel run a cyclic Jacobi sweep with two static round bodies and no register moves."""
    ns = 32
    slots = list(range(ns))
    met = {}
    for rnd in range(ns):
        odd = rnd & 1
        for p in range(ns // 2 - odd):
            lo, hi = 2 * p + odd, 2 * p + odd + 1
            a, b = slots[lo], slots[hi]
            key = (min(a, b), max(a, b))
            met[key] = met.get(key, 0) + 1
            slots[lo], slots[hi] = b, a
    assert len(met) == ns * (ns - 1) // 2 and set(met.values()) == {1}
    assert slots == list(range(ns))[::-1]


def test_odd_even_jacobi_with_tracked_norms_converges_like_lapack():
    """numpy emulation of the kernel's iteration (odd-even ordering, tracked squared norms refreshed once per sweep, exit
    after 32 consecutive rounds without a rotation above 1e-6) on 64x31 matrices of the kind the path sees (nearly rank one
    plus noise, a duplicated column, an all-zero matrix): singular values equal LAPACK's to 1e-13."""
    rng = np.random.default_rng(0)

    def svd_oe(A0, tol=1e-15, big=1e-6, max_sweeps=30):
        m, n = A0.shape
        a = np.zeros((m, 32))
        a[:, :n] = A0
        quiet, sweeps = 0, 0
        while sweeps < max_sweeps:
            nrm = (a * a).sum(0)
            for rp in range(16):
                if quiet >= 32:
                    break
                for odd in (0, 1):
                    anybig = False
                    for p in range(16 - odd):
                        lo, hi = 2 * p + odd, 2 * p + odd + 1
                        x, y = a[:, lo].copy(), a[:, hi].copy()
                        G, A, B = x @ y, nrm[lo], nrm[hi]
                        g2, ab = G * G, A * B
                        rot = g2 > tol * tol * ab
                        anybig |= g2 > big * big * ab
                        d = B - A
                        q = d * d + 4 * g2
                        t = ((2 * G if d >= 0 else -2 * G) / (abs(d) + np.sqrt(q))) if q > 0 else 0.0
                        cc = 1 / np.sqrt(1 + t * t)
                        c, s = (cc, cc * t) if rot else (1.0, 0.0)
                        w = s * (s * d - c * 2 * G)
                        a[:, lo], a[:, hi] = s * x + c * y, c * x - s * y
                        nrm[lo], nrm[hi] = B - w, A + w
                    quiet = 0 if anybig else quiet + 1
            sweeps += 1
            if quiet >= 32:
                break
        return np.sort(np.sqrt((a * a).sum(0)))[::-1][:n], sweeps

    for trial in range(4):
        A = rng.random((64, 1)) @ np.ones((1, 31)) + 0.05 * rng.standard_normal((64, 31))
        if trial == 2:
            A[:, 5] = A[:, 6]
        if trial == 3:
            A[:] = 0
        s, sweeps = svd_oe(A)
        ref = np.linalg.svd(A, compute_uv=False)
        assert np.abs(s - ref).max() <= 1e-13 * max(ref.max(), 1.0)
        assert sweeps <= 12


def test_qform_identity_behind_the_compact_cache():
    """The identity the truncated factor cache (compact.cuh) and k_eval3 rest on, in plain numpy: the second-difference
    term of the risk, sum_voxels delta2 * (overlap-add of blocks / weights) (pgure.hpp:136 with svt.hpp:148-164), is linear
    in the blocks, so for ANY threshold it equals sum_patches sum_k f_k(lambda) * q_k with q_k = u_k^T (delta2/weights)_patch v_k
    — the perturbed objects need only their singular values and q-forms, never U and V."""
    rng = np.random.default_rng(3)
    N, bs, T = 12, 4, 5
    M1 = N - bs + 1
    # trajectories: every patch jitters by up to one pixel per slice (clamped), like ARPS output
    base = np.array([(r, c) for c in range(M1) for r in range(M1)])
    pos = np.stack([np.clip(base + rng.integers(-1, 2, base.shape), 0, M1 - 1) for _ in range(T)])  # (T, P, 2)
    P = base.shape[0]
    u = rng.random((N, N, T))
    delta2 = np.where(rng.random((N, N, T)) < 0.72, -0.618, 1.618)
    weights = np.zeros((N, N, T))
    mats = np.zeros((P, bs * bs, T))
    for p in range(P):
        for k in range(T):
            r, c = pos[k, p]
            weights[r:r + bs, c:c + bs, k] += 1
            mats[p, :, k] = u[r:r + bs, c:c + bs, k].flatten(order="F")
    c4 = np.divide(delta2, weights, out=np.zeros_like(delta2), where=weights > 0)
    for lam in (0.0, 0.3, 1.5):
        direct_acc = np.zeros((N, N, T))
        via_q = 0.0
        for p in range(P):
            U, S, Vt = np.linalg.svd(mats[p], full_matrices=False)
            f = np.maximum(S - lam, 0.0)
            block = (U * f) @ Vt
            C = np.zeros((bs * bs, T))
            for k in range(T):
                r, c = pos[k, p]
                direct_acc[r:r + bs, c:c + bs, k] += block[:, k].reshape(bs, bs, order="F")
                C[:, k] = c4[r:r + bs, c:c + bs, k].flatten(order="F")
            q = np.einsum("ek,ej,jk->k", U, C, Vt.T)  # q_k = u_k^T C v_k
            via_q += float(f @ q)
        vhat = np.divide(direct_acc, weights, out=np.zeros_like(direct_acc), where=weights > 0)
        direct = float((delta2 * vhat).sum())
        assert abs(direct - via_q) <= 1e-10 * max(1.0, abs(direct))


def test_plan_gpus_and_frame_blocks_match_reference_partition():
    """The one-shot entry's fan-out (pguresvt_params.n_gpus) is host logic: partition of utils.hpp:150-166."""
    from pguresvt.distributed import frame_block

    assert bridge.plan_gpus(1000, n_visible=8, n_gpus=0) == 8
    assert bridge.plan_gpus(1000, n_visible=8, n_gpus=2) == 2
    assert bridge.plan_gpus(16, n_visible=8, n_gpus=0) == 2      # automatic: at least 8 frames per device
    assert bridge.plan_gpus(7, n_visible=8, n_gpus=0) == 1
    assert bridge.plan_gpus(1000, n_visible=8, n_gpus=0, device=6) == 2
    assert bridge.plan_gpus(1000, n_visible=1, n_gpus=4) == 1
    assert bridge.plan_gpus(3, n_visible=8, n_gpus=8) == 3
    for n, w in [(16, 2), (17, 2), (1000, 8), (5, 8), (23, 4)]:
        for r in range(w):
            assert bridge.frame_block(n, w, r) == frame_block(r, w, n)


def test_threaded_transpose_equals_numpy():
    """svt.py:329 reverses the axes of the bridge's result; the library's blocked, threaded copy must equal numpy's."""
    rng = np.random.RandomState(0)
    for shape in [(16, 32, 32), (17, 33, 21), (1, 5, 7), (40, 64, 48)]:
        X = rng.rand(*shape)
        out = bridge.reversed_axes_copy(X)
        assert out.flags.c_contiguous and np.array_equal(out, np.transpose(X, (2, 1, 0)))
