"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs
and against the committed golden fixtures (tests/golden/oracle_small.npz, produced here by the oracle).

Bars (BASELINE.json north_star): ARPS trajectories bit-exact; median and Bernoulli perturbations bit-exact
(integer work); per-frame lambda within 1e-3 relative; denoised pixels within 1e-6 relative
(max|dY| / max|Y_ref| per frame).  PGURE objective values: 1e-9 relative (FP64 reduction order differs)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, nsed, synthetic_sequence
from oracle import orc
from pguresvt import SVT
from pguresvt import _pguresvt as bridge

pytestmark = pytest.mark.gpu

PIX_TOL = 1e-6
LAM_TOL = 1e-3


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def per_frame_rel_err(Y, R):
    return max(np.abs(Y[:, :, t] - R[:, :, t]).max() / np.abs(R[:, :, t]).max() for t in range(R.shape[2]))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "oracle_small.npz"))


# ------------------------------------------------------------------ stage parity (integer / bit-exact)
def test_median_bit_exact(golden):
    X = golden["X"]
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15)
    for t in (0, 7, 15):
        assert np.array_equal(h.probe_median(t), golden["Z"][:, :, t])
    h.close()


@pytest.mark.parametrize("N,r", [(64, 1), (64, 3), (96, 5), (40, 7)])
def test_median_random_images_vs_oracle(N, r):
    rng = np.random.RandomState(N + r)
    X = np.asfortranarray(rng.randint(0, 65535, size=(N, N, 15)).astype(np.uint16))
    X[:, :, 1] = rng.poisson(20, size=(N, N))  # heavy ties
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, motion_filter=r)
    for t in (0, 1, 14):
        assert np.array_equal(h.probe_median(t), orc.median_u16(X[:, :, t], r))
    h.close()


def test_perturbations_bit_exact(golden):
    X = golden["X"]
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
    d1, d2 = h.probe_perturbations()
    assert np.array_equal(d1, golden["delta1"]) and np.array_equal(d2, golden["delta2neg"])
    h.close()
    # other seeds and a bigger cube straight against the oracle generator
    X2 = np.zeros((64, 64, 15), dtype=np.uint16, order="F")
    X2[0, 0, :] = 1
    for seed in (0, 101, 123456789):
        h = bridge.Handle(X2, optimize_pgure=True, noise_alpha=0.1, noise_mu=0.1, noise_sigma=0.1, random_seed=seed)
        d1, d2 = h.probe_perturbations()
        o1, o2 = orc.perturbations(seed, 64 * 64 * 15)
        assert np.array_equal(d1, o1.astype(np.int8)) and np.array_equal(d2, (o2 < 0).astype(np.int8))
        h.close()


def test_arps_bit_exact_golden(golden):
    X = golden["X"]
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15)
    p = h.probe_arps(8)
    assert np.array_equal(p.astype(np.int16), golden["patches8"])
    h.close()


@pytest.mark.parametrize("t", [0, 3, 7, 8, 12, 15])
def test_arps_bit_exact_all_window_cases(ref_test_cube, t):
    """First-fw, middle and last-fw frames take different schedules (arps.hpp:54-133)."""
    _, Y = ref_test_cube
    fw, F = 7, Y.shape[2]
    h = bridge.Handle(Y, optimize_pgure=False, lambda1=5.0)
    a = 0 if t < fw else (F - 2 * fw - 1 if t >= F - fw else t - fw)
    Z = np.stack([orc.median_u16(Y[:, :, i], 5) for i in range(a, a + 2 * fw + 1)], axis=2).astype(np.float64)
    w = Z / Z.max()
    want, _, _ = orc.arps(w, 4, t, fw, 7, F, True)
    got = h.probe_arps(t)
    assert np.array_equal(got, want.astype(np.int32)), f"{(got != want).sum()} trajectory entries differ"
    h.close()


def test_arps_motion_estimation_off_leaves_zero_positions(ref_test_cube):
    _, Y = ref_test_cube
    h = bridge.Handle(Y, optimize_pgure=False, lambda1=5.0, motion_estimation=False)
    p = h.probe_arps(8)
    want, _, _ = orc.arps(np.ones((32, 32, 15)), 4, 8, 7, 7, 16, False)
    assert np.array_equal(p, want.astype(np.int32))
    assert (p[:, :, [0, 14]] == 0).all() and p[:, :, 7].max() == 28
    h.close()


# ------------------------------------------------------------------ SVD / reconstruct / risk
@pytest.mark.parametrize("svd_kernel", [0, 1, 2, 3, 4])
def test_singular_values_vs_lapack(golden, svd_kernel):
    X = golden["X"]
    t, fw = 8, 7
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, svd_kernel=svd_kernel)
    S = h.probe_singular_values(t, 0)
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    o = orc.SVTObj(golden["patches8"].astype(np.int64), 32, 15, 4, 1, True)
    o.decompose(u)
    So = o.singular_values()
    assert S.shape == So.shape
    assert np.abs(S - So).max() / So.max() < 1e-12
    h.close()


@pytest.mark.parametrize("obj", [2, 3])
def test_singular_values_of_perturbed_objects_warm_start(golden, obj):
    """SVT objects U +- eps2*delta2 (pgure.hpp:81-82) are decomposed with a warm start from object 0's V; their
    singular values must still be LAPACK's."""
    X = golden["X"]
    t, fw = 8, 7
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
    S = h.probe_singular_values(t, obj)
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    _, d2 = orc.perturbations(1, u.size)
    d2 = d2.reshape(u.shape, order="F")
    up = u + (d2 * 0.01) if obj == 2 else u - (d2 * 0.01)
    o = orc.SVTObj(golden["patches8"].astype(np.int64), 32, 15, 4, 1, True)
    o.decompose(up)
    So = o.singular_values()
    assert np.abs(S - So).max() / So.max() < 1e-12
    h.close()


@pytest.mark.parametrize("svd_kernel", [1, 2, 3, 4])
def test_pgure_objective_other_svd_kernels(golden, svd_kernel):
    """The generic (shared-memory) and 8-lane register SVD kernels feed the unfused evaluation path."""
    X = golden["X"]
    alpha, mu, sigma = golden["pgure_params"]
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma, random_seed=1,
                      svd_kernel=svd_kernel, rank_cache=-1)
    vals, terms = h.probe_pgure(8, alpha, mu, sigma, golden["pgure_lambdas"])
    assert np.abs(vals - golden["pgure_values"]).max() <= 1e-9 * np.abs(golden["pgure_values"]).max()
    h.close()


@pytest.mark.parametrize("lam", [0.0, 0.15, 2.0, 50.0])
@pytest.mark.parametrize("expw", [True, False])
def test_reconstruct_window_vs_oracle(golden, lam, expw):
    X = golden["X"]
    t, fw = 8, 7
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, exponential_weighting=expw)
    v = h.probe_reconstruct(t, lam)
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    o = orc.SVTObj(golden["patches8"].astype(np.int64), 32, 15, 4, 1, expw)
    o.decompose(u)
    vo = o.reconstruct(lam)
    assert np.abs(v - vo).max() <= 1e-9 * max(1.0, np.abs(vo).max())
    h.close()


def test_pgure_objective_golden(golden):
    X = golden["X"]
    alpha, mu, sigma = golden["pgure_params"]
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma, random_seed=1)
    vals, terms = h.probe_pgure(8, alpha, mu, sigma, golden["pgure_lambdas"])
    assert np.allclose(terms, golden["pgure_terms"], rtol=1e-7, atol=1e-9)
    assert np.abs(vals - golden["pgure_values"]).max() <= 1e-9 * np.abs(golden["pgure_values"]).max()
    h.close()


def test_pgure_objective_intended_eps1_mode(golden):
    """eps1_mode=1 (four SVT objects, first-order term alive) against the oracle in the same mode."""
    X = golden["X"]
    t, fw = 8, 7
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    lams = np.array([0.05, 0.3, 3.0])
    orc.lib().orc_set_eps1_mode(1)
    try:
        P = orc.PGUREObj(u, golden["patches8"].astype(np.int64), 0.05, 0.03, 0.03, 4, 1, 1, True, True)
        want = np.array([P.calc(l)[0] for l in lams])
    finally:
        orc.lib().orc_set_eps1_mode(0)
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1, eps1_mode=1)
    vals, _ = h.probe_pgure(t, 0.05, 0.03, 0.03, lams)
    assert np.abs(vals - want).max() <= 1e-7 * np.abs(want).max()
    h.close()


# ------------------------------------------------------------------ whole pipeline through the public API
def test_fixed_lambda_golden(golden):
    X = golden["X"]
    s = SVT(optimize_pgure=False, lambda1=0.15, random_seed=1).denoise(X)
    assert per_frame_rel_err(s.Y_, golden["Y_fixed"]) < PIX_TOL
    s = SVT(optimize_pgure=False, lambda1=0.15, random_seed=1, motion_estimation=False).denoise(X)
    assert per_frame_rel_err(s.Y_, golden["Y_fixed_nome"]) < PIX_TOL
    assert np.all(s.lambda1s_ == 0.15)


def test_pgure_lambda_golden(golden):
    X = golden["X"]
    alpha, mu, sigma = golden["pgure_params"]
    s = SVT(noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma, random_seed=1).denoise(X)
    est = golden["est_pgure"]
    assert np.abs(s.lambda1s_ - est[:, 0]).max() / np.abs(est[:, 0]).max() < LAM_TOL
    assert per_frame_rel_err(s.Y_, golden["Y_pgure"]) < PIX_TOL
    assert np.array_equal(s.noise_alphas_, est[:, 1]) and np.array_equal(s.noise_sigmas_, est[:, 3])


def test_reference_test_cases_on_gpu(ref_test_cube):
    """The reference's own integration tests (test_svt.py:75-106) with known noise, same thresholds,
    plus exact parity with the oracle."""
    X, Y = ref_test_cube
    s = SVT(lambda1=5.0, optimize_pgure=False, random_seed=101).denoise(np.asfortranarray(Y))
    assert nsed(X, s.Y_) < 0.025
    ref, _ = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, random_seed=101)
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL
    s = SVT(noise_alpha=0.0109, noise_mu=100.0, noise_sigma=100.0, random_seed=101).denoise(Y)
    ref, est = orc.pguresvt(Y, lambda1=-1.0, noise_alpha=0.0109, noise_mu=100.0, noise_sigma=100.0, random_seed=101)
    assert nsed(X, s.Y_) < 0.3
    # raw-unit noise parameters (mu = sigma = 100 on max-normalised data, as the reference's test passes them) put a -1e4
    # offset on the risk.  Round 1 accepted 80 % of the frames here; the divergence came from the START POINT of the search
    # (a tree-ordered sum of u instead of Armadillo's sequential accu): with the exact sum every frame follows the oracle.
    rel = np.abs(s.lambda1s_ - est[:, 0]) / np.abs(est[:, 0])
    assert rel.max() < LAM_TOL, rel
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.float64])
def test_dtypes_fixed_lambda(dtype):
    X, _ = synthetic_sequence(32, 16, seed=5, dtype=np.uint16)
    if dtype == np.uint8:
        X = np.asfortranarray((X.astype(np.float64) / X.max() * 255).astype(np.uint8))
    else:
        X = np.asfortranarray(X.astype(dtype))
    s = SVT(optimize_pgure=False, lambda1=0.2, random_seed=1).denoise(X)
    ref, _ = orc.pguresvt(X, optimize_pgure=False, lambda1=0.2, random_seed=1)
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


def test_config1_like_patch16_overlap2():
    """BASELINE config 1 shape class: 256x15 Casorati matrices, patch_overlap=2 (skewed patch set, Q5),
    exponential weighting, ARPS on, median radius 5 — on a 64x64 crop-sized synthetic for the oracle's sake."""
    X, _ = synthetic_sequence(64, 17, seed=9)
    kw = dict(patch_size=16, patch_overlap=2, trajectory_length=15, optimize_pgure=False, lambda1=0.15, random_seed=1)
    s = SVT(**kw).denoise(X)
    ref, _ = orc.pguresvt(X, **kw)
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


def test_plain_threshold_and_no_median():
    X, _ = synthetic_sequence(32, 16, seed=11)
    kw = dict(optimize_pgure=False, lambda1=0.4, exponential_weighting=False, motion_filter=None, random_seed=1)
    s = SVT(**kw).denoise(X)
    kwo = dict(kw)
    kwo["motion_filter"] = -1
    ref, _ = orc.pguresvt(X, **kwo)
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


def test_frame_block_equals_full_run():
    """A block [fb, fe) processed through the handle API (what one GPU of a frame-sharded run does) equals the
    same frames of the full run (to FP64 rounding), including first/last-window frames."""
    X, _ = synthetic_sequence(32, 24, seed=13)
    full = bridge.Handle(X, optimize_pgure=False, lambda1=0.15)
    full.process()
    Yf, ef = full.download()
    full.close()
    for fb, fe in [(0, 5), (5, 17), (17, 24)]:
        h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, frame_begin=fb, frame_end=fe)
        h.process()
        Y, e = h.download()
        # the overlap-add uses FP64 atomics: summation order, hence the last bits, may differ between runs
        assert np.allclose(Y[:, :, fb:fe], Yf[:, :, fb:fe], rtol=1e-12, atol=1e-9)
        assert np.all(Y[:, :, :fb] == 0) and np.all(Y[:, :, fe:] == 0)
        h.close()


def test_one_shot_entry_streams_blocks(monkeypatch):
    """The one-shot C entry streams a long sequence in blocks of frames: forcing 5-frame blocks must
    give the result of the single-handle run (each block carries its halo frames and applies the global edge rules)."""
    X, _ = synthetic_sequence(32, 23, seed=9)
    kw = dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
    Y0, e0, _ = bridge.pguresvt_u16(X, **kw)
    monkeypatch.setenv("PGURESVT_BLOCK_FRAMES", "5")
    Y1, e1, _ = bridge.pguresvt_u16(X, **kw)
    assert np.abs(e1 - e0).max() <= 1e-9 * np.abs(e0).max()
    assert np.abs(Y1 - Y0).max() <= 1e-9 * np.abs(Y0).max()


def test_idempotent_and_deterministic_fixed_lambda():
    X, _ = synthetic_sequence(32, 16, seed=17)
    a = SVT(optimize_pgure=False, lambda1=0.15).denoise(X).Y_
    b = SVT(optimize_pgure=False, lambda1=0.15).denoise(X).Y_
    assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
    # lambda = 0 with plain thresholding reproduces the input exactly where every voxel is covered
    c = SVT(optimize_pgure=False, lambda1=0.0, exponential_weighting=False, motion_estimation=False).denoise(X).Y_
    assert np.abs(c - X).max() <= 1e-9 * X.max()


def test_full_size_properties_512():
    """BASELINE config 3 frame size (512^2, patch 4, trajectory 15, fixed lambda): size-independent properties —
    lambda=0/plain reproduces the input (reconstruct∘decompose = identity, overlap-add weights correct) and the
    exp-weighted output is finite and within the input's range scale."""
    X, _ = synthetic_sequence(512, 15, seed=3)
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.0, exponential_weighting=False, motion_estimation=False,
                      frame_begin=7, frame_end=8)
    h.process()
    Y, _ = h.download()
    assert np.abs(Y[:, :, 7] - X[:, :, 7]).max() <= 1e-9 * X.max()
    st = h.stats()
    assert st["svds"] == 509 * 509 and st["launches"] > 0
    h.close()
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, frame_begin=7, frame_end=8)
    h.process()
    Y, _ = h.download()
    assert np.isfinite(Y).all() and Y[:, :, 7].max() < 2.0 * X.max() and Y[:, :, 7].std() > 0
    h.close()


def test_hyperspy_wrapper_with_duck_typed_signal():
    from pguresvt.hspy import HSPYSVT

    X, _ = synthetic_sequence(32, 16, seed=19)

    class _Meta:
        class General:
            title = "stack"

    class _Axes:
        signal_dimension = 2

    class FakeSignal:
        """Image stack (frames, rows, cols) exposing the handful of methods hspy.py uses."""

        def __init__(self, data):
            self.data = data
            self.metadata = _Meta()
            self.axes_manager = _Axes()

        def unfold_navigation_space(self):
            pass

        def fold(self):
            pass

        def as_signal1D(self, spectral_axis=0):
            outer = self

            class S1:
                _data_aligned_with_axes = np.transpose(outer.data, (1, 2, 0))

            return S1()

        def _deepcopy_with_new_data(self, data):
            return FakeSignal(data)

    sig = FakeSignal(np.ascontiguousarray(np.transpose(X, (2, 0, 1))))
    out = HSPYSVT(optimize_pgure=False, lambda1=0.15).denoise(sig)
    ref = SVT(optimize_pgure=False, lambda1=0.15).denoise(X).Y_
    assert out.data.shape == sig.data.shape and out.metadata.General.title == "Denoised stack"
    assert np.allclose(out.data, np.transpose(ref, (2, 0, 1)), rtol=1e-12, atol=1e-9)


# ------------------------------------------------------------------ noise estimation (noise.hpp) on the GPU
def test_noise_estimate_golden(golden):
    X = golden["X"]
    h = bridge.Handle(X, optimize_pgure=True, random_seed=1)
    a, m, s = h.probe_noise(8)
    assert np.allclose([a, m, s], golden["noise8"], rtol=1e-6, atol=0), (a, m, s, golden["noise8"])
    # user-supplied values are kept, the others estimated (noise.hpp:113,141-145)
    a2, m2, s2 = h.probe_noise(8, alpha=0.07, mu=-1.0, sigma=-1.0)
    assert a2 == 0.07 and np.isclose(m2, golden["noise8"][1], rtol=1e-6)
    h.close()


@pytest.mark.parametrize("N", [64, 128, 256])
def test_noise_estimate_vs_oracle(N):
    X, _ = synthetic_sequence(N, 15, seed=N)
    h = bridge.Handle(X, optimize_pgure=True, random_seed=1)
    got = h.probe_noise(7)
    u = X.astype(np.float64)
    u /= u.max()
    want = orc.noise_estimate(u, 4)[:3]
    assert np.allclose(got, want, rtol=1e-6, atol=0), (got, want)
    h.close()


@pytest.mark.parametrize("method", [1, 2, 3])
def test_noise_methods_1_to_3(method):
    """The mode-based variants of noise.hpp:115-137 (non-default, 'currently undocumented' in the reference)."""
    X, _ = synthetic_sequence(64, 15, seed=64)
    h = bridge.Handle(X, optimize_pgure=True, random_seed=1, noise_method=method)
    got = h.probe_noise(7)
    u = X.astype(np.float64)
    u /= u.max()
    want = orc.noise_estimate(u, method)[:3]
    assert np.allclose(got, want, rtol=1e-6, atol=0), (got, want)
    h.close()


def test_default_api_estimates_noise(golden):
    """SVT() with the reference's defaults: noise parameters unknown -> estimated per frame on the GPU."""
    X = golden["X"]
    s = SVT(random_seed=1).denoise(X)
    ref, est = orc.pguresvt(X, optimize_pgure=True, lambda1=-1.0, random_seed=1)
    assert np.allclose(s.noise_alphas_, est[:, 1], rtol=1e-6) and np.allclose(s.noise_mus_, est[:, 2], rtol=1e-6)
    assert np.allclose(s.noise_sigmas_, est[:, 3], rtol=1e-6)
    rel = np.abs(s.lambda1s_ - est[:, 0]) / np.abs(est[:, 0])
    assert rel.max() < LAM_TOL, rel
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


def test_reference_default_test_on_gpu(ref_test_cube):
    """test_svt.py:75-90 (default SVT, nsed < 0.025) through the GPU path."""
    X, Y = ref_test_cube
    s = SVT(n_jobs=1, random_seed=101).denoise(Y)
    for a in ("Y_", "lambda1s_", "noise_alphas_", "noise_mus_", "noise_sigmas_"):
        assert hasattr(s, a)
    assert nsed(X, s.Y_) < 0.025


# ------------------------------------------------------------------ hot-pixel prefilter + CLI (configs 1-2)
def test_hotpixel_vs_oracle():
    import ctypes as C

    # Under the uint16 modular arithmetic of SURVEY Q22 every pixel below the frame median wraps to a huge deviation,
    # so the filter only fires when the median is the frame minimum: a flat background with sparse structure on top.
    rng = np.random.RandomState(4)
    X = np.full((64, 64, 6), 100, dtype=np.uint16, order="F")
    for t in range(6):
        sel = rng.rand(64, 64) < 0.3
        X[:, :, t][sel] = rng.randint(101, 200, sel.sum())
        rr, cc = rng.randint(0, 64, 40), rng.randint(0, 64, 40)
        X[rr, cc, t] = 60000  # hot pixels, some adjacent, some on the edge (order-dependent replacement)
        X[10, 10, t] = X[11, 10, t] = X[10, 11, t] = 65000
        X[0, 5, t] = X[63, 63, t] = 64000
    want = orc.hotpixel_u16(X, 10.0)
    got = X.copy(order="F")
    L = bridge.load()
    bridge.check(L.pguresvt_hotpixel_u16(got.ctypes.data_as(C.POINTER(C.c_uint16)), 64, 64, 6, C.c_double(10.0), 0), "hotpixel")
    assert (want != X).sum() > 100
    assert np.array_equal(got, want)


def test_cli_config1_example_tif(tmp_path):
    """BASELINE configs[0]: examples/example.tif with param_example.svt through the PGURE-SVT CLI (frames 1-17 here to
    keep the CPU oracle short): 128x128 uint16, patch 16 / overlap 2 (256x15 Casorati), fixed lambda 0.15,
    exponential weighting, ARPS on, median radius 5."""
    import shutil
    import subprocess

    import cv2

    exe = os.path.join(os.path.dirname(bridge.lib_path()), "PGURE-SVT")
    shutil.copy(os.path.join(GOLDEN, "example.tif"), tmp_path / "example.tif")
    par = open(os.path.join(GOLDEN, "param_example.svt")).read().replace("end_frame   : 25", "end_frame   : 17")
    (tmp_path / "p.svt").write_text(par)
    r = subprocess.run([exe, "p.svt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PGURE-SVT:" in r.stdout and "TIFF import:" in r.stdout
    ok, pages = cv2.imreadmulti(str(tmp_path / "example-CLEANED.tif"), flags=cv2.IMREAD_UNCHANGED)
    assert ok and len(pages) == 17 and pages[0].dtype == np.uint16
    got = np.stack(pages, axis=2)  # (y, x, t) == cube(row, col, frame)
    ok, inp = cv2.imreadmulti(os.path.join(GOLDEN, "example.tif"), flags=cv2.IMREAD_UNCHANGED)
    X = np.asfortranarray(np.stack(inp[:17], axis=2))
    ref, _ = orc.pguresvt(X, trajectory_length=15, patch_size=16, patch_overlap=2, motion_window=7, motion_filter=5,
                          optimize_pgure=False, lambda1=0.15, exponential_weighting=True, motion_estimation=True,
                          noise_alpha=0.1, noise_mu=0.1, noise_sigma=0.1, random_seed=1, max_iter=1000, n_jobs=-1)
    want = np.where(ref < 0, 0, ref).astype(np.int64).astype(np.uint16)  # conv_to<uint16>: truncate, negatives -> 0
    # truncation makes +-1 count differences possible where the double result sits within 1e-6 of an integer
    diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-4


# ------------------------------------------------------------------ edge cases of the driver logic
@pytest.mark.parametrize("kw,N,F", [
    (dict(trajectory_length=15), 32, 15),                      # sequence exactly one window long: every frame shares it
    (dict(trajectory_length=3), 32, 5),                        # 16x3 Casorati matrices (generic SVD kernel)
    (dict(trajectory_length=1), 32, 3),                        # no temporal dimension at all
    (dict(trajectory_length=15, patch_size=3), 30, 12),        # bs^2 < T -> Nt = 8, 9-slice windows
    (dict(trajectory_length=7, patch_size=5, patch_overlap=3), 40, 9),   # skewed patch set with uncovered pixels (Q5, Q17)
    (dict(trajectory_length=15, motion_window=3, motion_filter=1), 48, 16),  # small search window, 3x3 median, non-2^N frame
    (dict(trajectory_length=9, patch_size=8, patch_overlap=4), 64, 11),   # 64x9 Casorati matrices
    (dict(trajectory_length=31, patch_size=8), 40, 33),        # BASELINE configs[4] shape: 64x31 Casorati matrices, ARPS on
])
def test_edge_case_configurations_fixed_lambda(kw, N, F):
    X, _ = synthetic_sequence(N, F, seed=N + F)
    # noise parameters given so that the reference's 2^N check (svt.py:262-271) does not apply to odd frame sizes
    args = dict(optimize_pgure=False, lambda1=0.2, random_seed=1, noise_alpha=0.1, noise_mu=0.1, noise_sigma=0.1)
    args.update(kw)
    s = SVT(**args).denoise(X)
    ref, _ = orc.pguresvt(X, **args)
    assert np.isfinite(s.Y_).all()
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


def test_edge_case_pgure_small_windows():
    """PGURE search on a configuration that takes the generic kernels (patch 5, trajectory 7)."""
    X, _ = synthetic_sequence(40, 9, seed=77)
    args = dict(trajectory_length=7, patch_size=5, patch_overlap=2, optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05,
                noise_mu=0.03, noise_sigma=0.03, random_seed=3)
    s = SVT(**{k: v for k, v in args.items() if k != "lambda1"}).denoise(X)
    ref, est = orc.pguresvt(X, **args)
    rel = np.abs(s.lambda1s_ - est[:, 0]) / np.abs(est[:, 0])
    assert rel.max() < LAM_TOL, rel
    assert per_frame_rel_err(s.Y_, ref) < PIX_TOL


@pytest.mark.parametrize("rank_cache", [0, 1, -1])
def test_config5_shape_pgure_small(rank_cache):
    """BASELINE configs[4] shape at a size the oracle finishes in seconds: patch 8, trajectory 31 (64x31 Casorati
    matrices), PGURE lambda search, ARPS on.  rank_cache 0: warp-per-matrix register SVD + truncated factor cache
    (compact.cuh); 1: the same with a single cached triplet, so that probes overflow into the exact chunked fallback;
    -1: full factor cache + unfused evaluation kernels."""
    X, _ = synthetic_sequence(32, 33, seed=5)
    args = dict(trajectory_length=31, patch_size=8, optimize_pgure=True, lambda1=-1.0, noise_alpha=0.1, noise_mu=0.05,
                noise_sigma=0.05, random_seed=1)  # interior optimum (lambda ~ 36.9) rather than the upper bound
    t = 16
    h = bridge.Handle(X, frame_begin=t, frame_end=t + 1, rank_cache=rank_cache, **args)
    h.process()
    Yh, eh = h.download()
    st = h.stats()
    h.close()
    ref, est = orc.pguresvt(X, frame_begin=t, frame_end=t + 1, **args)
    assert abs(eh[t, 0] - est[t, 0]) / abs(est[t, 0]) < LAM_TOL
    assert np.abs(Yh[:, :, t] - ref[:, :, t]).max() / np.abs(ref[:, :, t]).max() < PIX_TOL
    if rank_cache == 1:
        assert st["overflow_patches"] > 0 and st["rank_cache"] == 1
    if rank_cache == -1:
        assert st["rank_cache"] == 0


@pytest.mark.parametrize("kw", [dict(trajectory_length=15, patch_size=8), dict(trajectory_length=15, patch_size=8, patch_overlap=3),
                                dict(trajectory_length=31, patch_size=8, exponential_weighting=False),
                                dict(trajectory_length=9, patch_size=6, patch_overlap=2)])
def test_warp_svd_other_shapes_pgure(kw):
    """64 x 15 (zero-padded column slots), overlapping patch sets, plain thresholding and a 36 x 9 shape (shared-memory
    kernel + compact cache) through the whole PGURE pipeline against the oracle."""
    traj = kw["trajectory_length"]
    X, _ = synthetic_sequence(32, traj + 3, seed=21)
    args = dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.1, noise_mu=0.05, noise_sigma=0.05, random_seed=2, **kw)
    t = traj // 2 + 1
    h = bridge.Handle(X, frame_begin=t, frame_end=t + 2, **args)
    h.process()
    Yh, eh = h.download()
    ref, est = orc.pguresvt(X, frame_begin=t, frame_end=t + 2, **args)
    for f in (t, t + 1):
        # (round 1 accepted the plain-thresholding case on the objective value only: its search ended elsewhere in a flat
        #  basin.  The cause was the start point of the search — see k_accu_seq — and is gone: lambda agrees on every case.)
        assert abs(eh[f, 0] - est[f, 0]) / abs(est[f, 0]) < LAM_TOL
        pix_tol = PIX_TOL
        assert np.abs(Yh[:, :, f] - ref[:, :, f]).max() / np.abs(ref[:, :, f]).max() < pix_tol
    h.close()


@pytest.mark.parametrize("svd_kernel", [0, 1])
@pytest.mark.parametrize("obj", [0, 2, 3])
def test_config5_shape_singular_values_vs_lapack(svd_kernel, obj):
    """64x31 Casorati matrices: warp-per-matrix register Jacobi (svd_kernel 0) and shared-memory Jacobi (1), both with the
    compact epilogue, against LAPACK on the same (perturbed) window."""
    X, _ = synthetic_sequence(32, 33, seed=5)
    t, fw = 16, 15
    h = bridge.Handle(X, trajectory_length=31, patch_size=8, optimize_pgure=True, noise_alpha=0.1, noise_mu=0.05,
                      noise_sigma=0.05, random_seed=1, svd_kernel=svd_kernel, motion_estimation=False)
    S = h.probe_singular_values(t, obj)
    h.close()
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    if obj:
        _, d2 = orc.perturbations(1, u.size)
        d2 = d2.reshape(u.shape, order="F")
        u = u + (d2 * 0.01) if obj == 2 else u - (d2 * 0.01)
    patches, _, _ = orc.arps(u, 8, t, fw, 7, 33, False)
    o = orc.SVTObj(patches.astype(np.int64), 32, 31, 8, 1, True)
    o.decompose(u)
    So = o.singular_values()
    assert S.shape == So.shape
    assert np.abs(S - So).max() / So.max() < 1e-12


@pytest.mark.parametrize("svd_kernel,rank_cache", [(0, 0), (0, 2), (1, 0), (1, 1), (0, -1)])
def test_config5_shape_objective_and_window_vs_oracle(svd_kernel, rank_cache):
    """PGURE objective (pgure.hpp:120-137) and the reconstructed window (svt.hpp:121-167) of the 64x31 shape at lambdas
    from 'nothing survives' to 'every triplet survives' — the latter only through the overflow fallback."""
    X, _ = synthetic_sequence(32, 33, seed=5)
    t, fw = 16, 15
    alpha, mu, sigma = 0.1, 0.05, 0.05
    h = bridge.Handle(X, trajectory_length=31, patch_size=8, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu,
                      noise_sigma=sigma, random_seed=1, svd_kernel=svd_kernel, rank_cache=rank_cache, motion_estimation=False)
    lams = np.array([0.0, 0.5, 5.0, 36.9, 99.0])
    vals, terms = h.probe_pgure(t, alpha, mu, sigma, lams)
    v = h.probe_reconstruct(t, 36.9)
    v_all = h.probe_reconstruct(t, 99.0)
    st = h.stats()
    h.close()
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    patches, _, _ = orc.arps(u, 8, t, fw, 7, 33, False)
    Pg = orc.PGUREObj(u, patches.astype(np.int64), alpha, mu, sigma, 8, 1, 1, True, True)  # mu == sigma: the swap of SURVEY Q1 is invisible
    want = np.array([Pg.calc(l)[0] for l in lams])
    assert np.abs(vals - want).max() <= 1e-8 * np.abs(want).max(), (vals, want)
    o = orc.SVTObj(patches.astype(np.int64), 32, 31, 8, 1, True)
    o.decompose(u)
    for got, lam in ((v, 36.9), (v_all, 99.0)):
        vo = o.reconstruct(lam)
        assert np.abs(got - vo).max() <= 1e-9 * max(1.0, np.abs(vo).max())


@pytest.mark.parametrize("patch,traj,rank_cache", [(4, 15, 0), (8, 31, 0), (8, 31, 2)])
def test_zero_region_rank_deficient_patches(patch, traj, rank_cache):
    """A region of exact zeros (clipped background) gives rank-deficient — all-zero — Casorati matrices for object U while
    the perturbed objects U +- eps2*delta2 are NOT zero there: the warm start of the perturbed SVDs (V of object U) must
    fall back to a cold start for such patches, otherwise their singular values vanish from the risk."""
    X, _ = synthetic_sequence(32, traj + 2, seed=5)
    X[3:19, 8:24, :] = 0
    fw = traj // 2
    t = fw + 1
    alpha, mu, sigma = 0.1, 0.05, 0.05
    h = bridge.Handle(X, trajectory_length=traj, patch_size=patch, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu,
                      noise_sigma=sigma, random_seed=1, rank_cache=rank_cache, motion_estimation=False)
    lams = np.array([0.5, 5.0, 40.0, 99.0])
    vals, terms = h.probe_pgure(t, alpha, mu, sigma, lams)
    S2 = h.probe_singular_values(t, 2)
    h.close()
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    patches, _, _ = orc.arps(u, patch, t, fw, 7, traj + 2, False)
    Pg = orc.PGUREObj(u, patches.astype(np.int64), alpha, mu, sigma, patch, 1, 1, True, True)
    want = np.array([Pg.calc(l) for l in lams], dtype=object)
    wv = np.array([w[0] for w in want])
    wt = np.array([w[1] for w in want])
    _, d2 = orc.perturbations(1, u.size)
    o = orc.SVTObj(patches.astype(np.int64), 32, traj, patch, 1, True)
    o.decompose(u + (d2.reshape(u.shape, order="F") * 0.01))
    So = o.singular_values()
    assert np.abs(S2 - So).max() / So.max() < 1e-12
    assert np.allclose(terms, wt, rtol=1e-7, atol=1e-9)
    assert np.abs(vals - wv).max() <= 1e-8 * np.abs(wv).max(), (vals, wv)


@pytest.mark.parametrize("rank_cache", [0, 1, 3])
def test_compact_cache_16x15_matches_golden_objective(golden, rank_cache):
    """The truncated factor cache on the headline shape (forced through svd_kernel=1): same objective values as the
    golden vectors, whatever the number of cached triplets."""
    X = golden["X"]
    alpha, mu, sigma = golden["pgure_params"]
    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma, random_seed=1,
                      svd_kernel=1, rank_cache=rank_cache)
    vals, terms = h.probe_pgure(8, alpha, mu, sigma, golden["pgure_lambdas"])
    h.close()
    assert np.allclose(terms, golden["pgure_terms"], rtol=1e-7, atol=1e-9)
    assert np.abs(vals - golden["pgure_values"]).max() <= 1e-9 * np.abs(golden["pgure_values"]).max()


def test_full_size_properties_1024_pgure():
    """BASELINE configs[3] frame size (1024^2, patch 4, trajectory 15, PGURE): the oracle cannot run this size, so the
    production path (4-lane register SVD + fused three-object evaluation with q-forms) is checked against the
    independent generic path (shared-memory SVD + per-object reconstruction + five-sum risk kernel), which the small
    tests pin to the oracle: same objective values, same per-frame lambda and pixels within the parity tolerances."""
    X, _ = synthetic_sequence(1024, 15, seed=11)
    kw = dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1,
              frame_begin=7, frame_end=8)
    lams = [0.0, 0.05, 0.3, 2.0]
    out = {}
    for k in (0, 1):
        h = bridge.Handle(X, svd_kernel=k, rank_cache=-k, **kw)  # k = 1: full factor cache, unfused kernels
        vals, _ = h.probe_pgure(7, 0.05, 0.03, 0.03, lams)
        h.process()
        Y, e = h.download()
        out[k] = (vals, Y[:, :, 7].copy(), e[7, 0], h.stats())
        h.close()
    assert out[0][3]["svds"] == 3 * 1021 * 1021
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-9 * np.abs(out[1][0]).max()
    assert abs(out[0][2] - out[1][2]) / abs(out[1][2]) < LAM_TOL
    assert np.abs(out[0][1] - out[1][1]).max() / np.abs(out[1][1]).max() < PIX_TOL


def test_too_short_sequence_and_bad_arguments_are_errors_not_crashes():
    X, _ = synthetic_sequence(32, 10, seed=1)
    with pytest.raises(RuntimeError, match="fewer than"):
        SVT(optimize_pgure=False, lambda1=0.1).denoise(X)          # 10 frames < 15-frame window
    with pytest.raises(RuntimeError, match="square"):
        bridge.pguresvt_u16(np.zeros((32, 48, 16), dtype=np.uint16, order="F"), optimize_pgure=False, lambda1=0.1)
    with pytest.raises(RuntimeError, match="initial step|cannot start"):
        bridge.pguresvt_u16(np.zeros((32, 32, 16), dtype=np.uint16, order="F") + 5, optimize_pgure=True, lambda1=0.0,
                            noise_alpha=0.1, noise_mu=0.1, noise_sigma=0.1)  # the CLI's lambda = 0.0 start
