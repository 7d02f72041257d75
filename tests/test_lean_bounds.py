"""CPU tests (plain numpy) of the two facts the lean PGURE path of the device rests on (k_top1_l4 / k_lean_check,
pgure-svt_b200/csrc/kernels.cuh): (1) the Gram bound  sigma_2 <= (||A^T A||_F^2 - sigma_1^4)^(1/4)  is rigorous and tight
enough on noise-dominated 16x15 Casorati matrices, the Weyl bound  sigma_k(A + E) <= sigma_2(A) + ||E||_F  likewise; (2) the
soft threshold of svt.hpp:135-143 / utils.hpp:96-106 is monotone in the singular value, so a bound that does not survive
proves that nothing below it does — which is what lets the device skip the full SVD and still answer every probe exactly."""
import numpy as np


def casorati(rng, signal=1.0, noise=0.05):
    u = np.abs(rng.randn(16)) + 1.0
    v = np.abs(rng.randn(15)) + 1.0
    A = signal * np.outer(u / np.linalg.norm(u), v / np.linalg.norm(v)) * rng.uniform(0.5, 4.0)
    return A + noise * rng.randn(16, 15)


def soft_f(s, smax, lam, expw):
    w = np.abs(smax * np.exp(-0.5 * lam * s * s)) if expw else lam
    return np.maximum(s - w, 0.0)


def test_gram_and_weyl_bounds_are_rigorous_and_tight():
    rng = np.random.RandomState(3)
    ratios = []
    for _ in range(400):
        A = casorati(rng, noise=rng.uniform(0.01, 0.3))
        S = np.linalg.svd(A, compute_uv=False)
        G = A.T @ A
        g2 = max((G * G).sum() - S[0] ** 4, 0.0) + 1e-13 * S[0] ** 4
        B = g2 ** 0.25 * (1 + 1e-9)
        assert B >= S[1]
        assert B <= 1.0001 * ((S[1:] ** 4).sum() + 1e-13 * S[0] ** 4) ** 0.25
        ratios.append(B / S[1])
        # Weyl: perturbed object against the unperturbed sigma_2
        delta = np.where(rng.rand(16, 15) < 0.7236, -0.6180339887, 1.6180339887)
        E = 0.01 * delta
        for sgn in (1.0, -1.0):
            Sp = np.linalg.svd(A + sgn * E, compute_uv=False)
            assert Sp[1] <= (S[1] + np.linalg.norm(E)) * (1 + 1e-12)
    assert np.median(ratios) < 1.35  # (Frobenius norm of the residual: 1.9)


def test_soft_threshold_is_monotone_in_the_singular_value():
    rng = np.random.RandomState(4)
    for expw in (True, False):
        for _ in range(200):
            smax = rng.uniform(0.5, 8.0)
            lam = 10.0 ** rng.uniform(-3, 2)
            s = np.sort(rng.uniform(0.0, smax, 64))
            f = soft_f(s, smax, lam, expw)
            assert np.all(np.diff(f) >= 0.0)
            # hence: a bound B >= s whose thresholded value is zero proves f(s) = 0
            B = s[40]
            if soft_f(B, smax, lam, expw) == 0.0:
                assert np.all(f[:41] == 0.0)


def test_critical_lambda_of_a_bound():
    """k_lean_crit: with exponential weighting a bound B < s1 survives iff lambda > 2 ln(s1 / B) / B^2."""
    rng = np.random.RandomState(5)
    for _ in range(200):
        s1 = rng.uniform(0.5, 5.0)
        B = rng.uniform(0.05, 0.99) * s1
        lc = 2.0 * np.log(s1 / B) / (B * B)
        assert soft_f(B, s1, lc * (1 - 1e-6), True) == 0.0
        assert soft_f(B, s1, lc * (1 + 1e-6), True) > 0.0
