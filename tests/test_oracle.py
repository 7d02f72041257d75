"""CPU tests of the oracle (restatement of the reference) against every golden vector the survey gathered
(SURVEY §8c), against the reference's own vendored pieces compiled into oracle/_ref when present, and against
the committed fixtures under tests/golden/."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, nsed, synthetic_sequence
from oracle import orc

PCG_GOLDEN = {
    0: "01070196e695f8f1 703ec840c59f4493 e54954914b3a44fa 96130ff204b9285e",
    1: "e175e32ed3507bfa c0bf922a0b283109 140bfa21e68785bb c5ec8bcc4fe35830",
    101: "31b84a71188fb148 6ccfc8fe8ba62a69 8cd656d3f723d117 cd5df7cf5e7c6d9a",
}
DELTA1_GOLDEN = {
    0: [-1, -1, 1, 1, -1, -1, 1, 1, 1, -1, -1, -1, 1, -1, 1, -1],
    1: [1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, -1, -1, 1, -1],
    101: [-1, -1, 1, 1, -1, -1, 1, -1, 1, 1, -1, 1, -1, 1, 1, 1],
}


@pytest.mark.parametrize("seed", [0, 1, 101])
def test_pcg64_golden(seed):
    raw = orc.pcg64_raw(seed, 4)
    assert " ".join("%016x" % x for x in raw) == PCG_GOLDEN[seed]


@pytest.mark.parametrize("seed", [0, 1, 101])
def test_bernoulli_golden(seed):
    d1, d2 = orc.perturbations(seed, 16)
    assert d1.tolist() == DELTA1_GOLDEN[seed]
    assert set(np.round(np.unique(d2), 12)) <= {-0.618033988750, 1.618033988750}


def test_perturbations_vs_reference_pcg():
    if orc.ref() is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    for seed in (0, 1, 7, 101):
        a = orc.perturbations(seed, 100000)
        b = orc.perturbations(seed, 100000, "ref")
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        assert (orc.pcg64_raw(seed, 64) == orc.pcg64_raw(seed, 64, "ref")).all()


def test_median_vs_reference_ctmf():
    if orc.ref() is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.RandomState(3)
    for N, r in [(32, 1), (32, 2), (32, 3), (32, 5), (128, 5), (64, 7)]:
        img = rng.randint(0, 65535, size=(N, N)).astype(np.uint16)
        assert (orc.median_u16(img, r) == orc.median_u16(img, r, "ref")).all()
        img = (rng.poisson(30, size=(N, N))).astype(np.uint16)  # many ties
        assert (orc.median_u16(img, r) == orc.median_u16(img, r, "ref")).all()


def test_median_brute_force():
    rng = np.random.RandomState(5)
    img = rng.randint(0, 1000, size=(24, 24)).astype(np.uint16)
    r = 2
    pad = np.pad(img, r, mode="edge")
    want = np.zeros_like(img)
    for i in range(24):
        for j in range(24):
            want[i, j] = np.sort(pad[i:i + 2 * r + 1, j:j + 2 * r + 1].ravel())[(2 * r + 1) ** 2 // 2]
    assert (orc.median_u16(img, r) == want).all()


@pytest.mark.parametrize("cfg,want", [((32, 4, 1), 841), ((32, 4, 2), 252), ((128, 16, 2), 3360), ((256, 4, 2), 16380),
                                      ((32, 4, 3), 120)])
def test_patch_set_sizes(cfg, want):
    N, bs, bo = cfg
    s = orc.SVTObj(np.zeros((2, (N - bs + 1) ** 2, 3), dtype=np.int64), N, 3, bs, bo, True)
    assert s.npatches() == want
    if cfg == (32, 4, 2):
        ids = s.patch_ids()[:17]
        M1 = 29
        rc = [(int(i % M1), int(i // M1)) for i in ids]
        assert rc[:3] == [(0, 0), (2, 0), (4, 0)] and rc[14] == (28, 0) and rc[15] == (27, 1) and rc[16] == (0, 2)


def test_quadtree_golden_counts():
    assert orc.quadtree_counts(32) == (41, 33)
    assert orc.quadtree_counts(64) == (169, 129)
    assert orc.quadtree_counts(128) == (681, 513)
    assert orc.quadtree_counts(32, 1) == (9, 9)


def test_jacobi_matches_dgesdd():
    if orc.svd_backend() != "dgesdd":
        pytest.skip("LAPACK not found")
    import ctypes as C
    rng = np.random.RandomState(0)
    for m, n in [(16, 15), (64, 31), (256, 15), (9, 9)]:
        A = np.asfortranarray(rng.rand(m, n))
        out = []
        for backend in (1, 0):
            orc.lib().orc_set_svd_backend(backend)
            U = np.zeros((m, n), order="F"); S = np.zeros(n); V = np.zeros((n, n), order="F")
            orc.lib().orc_svd(C.c_int(m), C.c_int(n), orc._p(A), orc._p(U), orc._p(S), orc._p(V))
            out.append((U, S, V))
        orc.lib().orc_set_svd_backend(1)
        assert np.allclose(out[0][1], out[1][1], rtol=1e-12, atol=1e-13)
        for U, S, V in out:
            assert np.allclose(U @ np.diag(S) @ V.T, A, atol=1e-12)


def test_sbplx_quadratic_and_monotone():
    r = orc.sbplx_1d(lambda x: (x - 0.3) ** 2 + 1.0, 0.05, 0.0, 100.0, np.sqrt(0.05), ftol_rel=1e-7)
    assert abs(r["x"] - 0.3) < 1e-3 and r["status"] == 3
    assert 15 < r["nevals"] < 80
    # monotone increasing objective: converges to the lower bound by repeated step shrinking
    r = orc.sbplx_1d(lambda x: 3.0 * x - 100.0, 0.05, 0.0, 100.0, np.sqrt(0.05), ftol_rel=1e-7)
    assert r["x"] == 0.0 and r["trace"][-1][0] < 1e-5
    # zero initial step is rejected like nlopt_set_initial_step(0)
    r = orc.sbplx_1d(lambda x: x, 0.0, 0.0, 100.0, 0.0)
    assert r["status"] == -2 and r["nevals"] == 0


def test_mixed_noise_model_reference_goldens():
    """The only numbers the reference's own tests pin (test_svt.py:30,34,45)."""
    from pguresvt import mixed_noise_model

    rng = np.random.RandomState(101)
    X = rng.uniform(low=0, high=255, size=(64, 64, 32))
    np.testing.assert_allclose(nsed(X, mixed_noise_model(X, random_state=rng)), 0.3753506, rtol=1e-6)
    rng = np.random.RandomState(101)
    X = rng.uniform(low=0, high=255, size=(64, 64, 32))
    np.testing.assert_allclose(nsed(X, mixed_noise_model(X, alpha=1e-5, random_state=rng)), 1.4986839425e-05, rtol=1e-6)
    rng = np.random.RandomState(101)
    X = rng.uniform(low=0, high=255, size=(64, 64, 32))
    np.testing.assert_allclose(nsed(X, mixed_noise_model(X, alpha=1e-5, sigma=0.1, random_state=rng)), 0.02836968,
                               rtol=1e-6)


def test_reference_test_cube_fixed_lambda(ref_test_cube):
    """Survey-time cross-check values (restatement-derived): lambda=5, frames {0,3,8,12,15}."""
    X, Y = ref_test_cube
    fr = [0, 3, 8, 12, 15]
    D, est = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, motion_estimation=True, random_seed=101)
    assert abs(nsed(X[:, :, fr], D[:, :, fr]) - 0.008845) < 2e-6
    D, est = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, motion_estimation=False, random_seed=101)
    assert abs(nsed(X[:, :, fr], D[:, :, fr]) - 0.019262) < 2e-6
    assert np.all(est[:, 0] == 5.0)


def test_reference_own_thresholds(ref_test_cube):
    """The reference's TestGaussianNoise assertions (test_svt.py:75-106) hold for the oracle — including
    test_known_noise (< 0.3), which only holds with the integer-truncated eps1."""
    X, Y = ref_test_cube
    D, est = orc.pguresvt(Y, optimize_pgure=True, lambda1=-1.0, random_seed=101, n_jobs=1)
    assert nsed(X, D) < 0.025
    D, est = orc.pguresvt(Y, optimize_pgure=True, lambda1=-1.0, noise_mu=100.0, noise_sigma=100.0, random_seed=101)
    assert nsed(X, D) < 0.3
    orc.lib().orc_set_eps1_mode(1)
    try:
        D, est = orc.pguresvt(Y, optimize_pgure=True, lambda1=-1.0, noise_mu=100.0, noise_sigma=100.0, random_seed=101)
        assert nsed(X, D) > 0.3  # the "intended" maths would FAIL the reference's own test
    finally:
        orc.lib().orc_set_eps1_mode(0)


def test_threads_do_not_change_results(ref_test_cube):
    X, Y = ref_test_cube
    a = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, n_jobs=1)
    b = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, n_jobs=4)
    assert np.array_equal(a[0], b[0])
    c = orc.pguresvt(Y, optimize_pgure=False, lambda1=5.0, n_jobs=2, frame_begin=5, frame_end=9)
    assert np.array_equal(a[0][:, :, 5:9], c[0][:, :, 5:9]) and np.all(c[0][:, :, :5] == 0)


def test_golden_fixture_roundtrip():
    """tests/golden/oracle_small.npz was produced by tests/golden/make_golden.py with this oracle; the oracle
    must keep reproducing it bit for bit (guards against accidental changes of the checker)."""
    path = os.path.join(GOLDEN, "oracle_small.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = np.load(path)
    D, est = orc.pguresvt(g["X"], optimize_pgure=False, lambda1=0.15, random_seed=1)
    assert np.allclose(D, g["Y_fixed"], rtol=1e-12, atol=1e-9)
    p, m, nc = orc.arps(g["w8"], 4, 8, 7, 7, 16, True)
    assert np.array_equal(p.astype(np.int16), g["patches8"])
