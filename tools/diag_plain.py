"""Diagnostic: PGURE objective of the 64x31 shape with plain thresholding, compact vs full cache vs oracle (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import synthetic_sequence
from oracle import orc
from pguresvt import _pguresvt as b
traj = 31
X, _ = synthetic_sequence(32, traj + 3, seed=21)
t, fw = traj // 2 + 1, traj // 2
alpha, mu, sigma = 0.1, 0.05, 0.05
lams = np.array([0.01, 0.03, 0.06, 0.0641, 0.0642, 0.0643, 0.0644, 0.07, 0.1, 0.3, 1.0])
u = X[:, :, t - fw:t + fw + 1].astype(np.float64); u /= u.max()
w = u.copy()
res = {}
for rc in (0, -1):
    for me in (False,):
        h = b.Handle(X, trajectory_length=traj, patch_size=8, optimize_pgure=True, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma,
                     random_seed=2, rank_cache=rc, motion_estimation=me, lambda1=-1.0, exponential_weighting=False, frame_begin=t, frame_end=t + 1)
        vals, terms = h.probe_pgure(t, alpha, mu, sigma, lams)
        h.process(); Y, e = h.download()
        res[rc] = (vals, terms, e[t, 0])
        print("rank_cache", rc, "lambda", e[t, 0], h.stats()["evals"])
        h.close()
patches, _, _ = orc.arps(u, 8, t, fw, 7, traj + 3, False)
Pg = orc.PGUREObj(u, patches.astype(np.int64), alpha, mu, sigma, 8, 1, 2, False, True)
want = [Pg.calc(l) for l in lams]
wv = np.array([x[0] for x in want]); wt = np.array([x[1] for x in want])
for rc in (0, -1):
    print("rc", rc, "rel err values", np.abs(res[rc][0] - wv) / np.abs(wv))
    print("rc", rc, "rel err terms", (np.abs(res[rc][1] - wt) / np.maximum(np.abs(wt), 1e-300)).max(0))
print("oracle values", wv)
ref, est = orc.pguresvt(X, frame_begin=t, frame_end=t + 1, trajectory_length=traj, patch_size=8, optimize_pgure=True, lambda1=-1.0,
                        noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma, random_seed=2, exponential_weighting=False, motion_estimation=False)
print("oracle lambda (ME off)", est[t, 0])
