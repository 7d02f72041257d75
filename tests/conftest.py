import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pgure-svt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from pguresvt import _pguresvt as b

        return b.load().pguresvt_device_info(0, None, 0) > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def nsed(A, B):
    A_m = A - A.mean()
    B_m = B - B.mean()
    return 0.5 * np.linalg.norm(A_m - B_m) ** 2 / (np.linalg.norm(A_m) ** 2 + np.linalg.norm(B_m) ** 2)


def synthetic_sequence(N, F, seed=123, alpha=0.1, mu=0.1, sigma=0.1, dtype=np.uint16, pitch=16, blob=2.5):
    """Drifting lattice of Gaussian blobs in [0,1] * 4095 with the reference's own Poisson-Gaussian generator
    (SURVEY §8d).  Returns (noisy uint16 F-order (N,N,F), clean float)."""
    from pguresvt import mixed_noise_model

    yy, xx = np.mgrid[0:N, 0:N].astype(np.float64)
    frames = []
    for t in range(F):
        dy, dx = 0.3 * t, 0.2 * t
        py = (yy - dy) % pitch - pitch / 2
        px = (xx - dx) % pitch - pitch / 2
        img = np.exp(-(py ** 2 + px ** 2) / (2 * blob ** 2)) * (1.0 + 0.1 * np.sin(0.2 * t))
        frames.append(img)
    clean = np.stack(frames, axis=2)
    clean /= clean.max()
    noisy = mixed_noise_model(clean * 4095, alpha=alpha, mu=mu, sigma=sigma, random_state=seed)
    noisy[noisy < 0] = 0
    if np.issubdtype(np.dtype(dtype), np.integer):
        noisy = np.minimum(noisy, np.iinfo(dtype).max)
    return np.asfortranarray(noisy.astype(dtype)), clean


@pytest.fixture(scope="session")
def ref_test_cube():
    """The reference's own test cube (pguresvt/tests/data.npz) + the noise its TestGaussianNoise adds."""
    X = np.load(os.path.join(GOLDEN, "ref_test_data.npz"))["a"]
    rng = np.random.RandomState(101)
    Y = X + 100.0 + 100.0 * rng.randn(*X.shape)
    Y[Y < 0.0] = 0.0
    return X, Y.astype(np.uint16)
