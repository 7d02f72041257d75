import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
traj = 31
X, _ = synthetic_sequence(32, traj + 3, seed=21)
t = traj // 2 + 1
alpha, mu, sigma = 0.1, 0.05, 0.05
lams = np.array([0.0630, 0.0641552, 0.0642, 0.0643669, 0.0650])
for rc, sk in ((0, 0), (-1, 0), (0, 1), (-1, 1)):
    h = b.Handle(X, trajectory_length=traj, patch_size=8, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma,
                 random_seed=2, rank_cache=rc, svd_kernel=sk, lambda1=-1.0, exponential_weighting=False, frame_begin=t, frame_end=t + 1)
    vals, terms = h.probe_pgure(t, alpha, mu, sigma, lams)
    h.process(); Y, e = h.download()
    print("rank_cache", rc, "svd_kernel", sk, "lambda %.16g" % e[t, 0], "evals", h.stats()["evals"], "values", " ".join("%.15e" % v for v in vals))
    h.close()
