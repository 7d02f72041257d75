"""Generates tests/golden/oracle_small.npz with the CPU oracle (oracle/oracle.cpp).  Run here (build container);
the GPU box only reads the committed fixture.  `python tests/golden/make_golden.py`"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_sequence  # noqa: E402
from oracle import orc  # noqa: E402

X, clean = synthetic_sequence(32, 16, seed=123)
out = {"X": X, "svd_backend": orc.svd_backend()}
out["Y_fixed"], _ = orc.pguresvt(X, optimize_pgure=False, lambda1=0.15, random_seed=1)
out["Y_fixed_nome"], _ = orc.pguresvt(X, optimize_pgure=False, lambda1=0.15, random_seed=1, motion_estimation=False)
# stage-level goldens for frame 8 (window [1,15], reference slice 7)
t, fw = 8, 7
Z = np.stack([orc.median_u16(X[:, :, i], 5) for i in range(16)], axis=2)
out["Z"] = Z
w = Z[:, :, t - fw:t + fw + 1].astype(np.float64)
w /= w.max()
u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
u /= u.max()
p, m, nc = orc.arps(w, 4, t, fw, 7, 16, True)
out["w8"] = w
out["patches8"] = p.astype(np.int16)
out["ncost8"] = nc
d1, d2 = orc.perturbations(1, 32 * 32 * 15)
out["delta1"] = d1.astype(np.int8)
out["delta2neg"] = (d2 < 0).astype(np.int8)
alpha, mu, sigma = 0.05, 0.03, 0.03
P = orc.PGUREObj(u, p, alpha, sigma, mu, 4, 1, 1, True, True)  # NB PGURE(alpha, sigma, mu): pguresvt.hpp:133
lams = np.array([0.0, 0.01, 0.05, 0.1, 0.3, 1.0, 3.0, 10.0, 30.0, 100.0])
vals, terms = zip(*[P.calc(l) for l in lams])
out["pgure_lambdas"] = lams
out["pgure_values"] = np.array(vals)
out["pgure_terms"] = np.array(terms)
out["pgure_params"] = np.array([alpha, mu, sigma])
Yp, est = orc.pguresvt(X, optimize_pgure=True, lambda1=-1.0, noise_alpha=alpha, noise_mu=mu, noise_sigma=sigma,
                       random_seed=1)
out["Y_pgure"] = Yp
out["est_pgure"] = est
a, m_, s, stats = orc.noise_estimate(u, 4)
out["noise8"] = np.array([a, m_, s])
np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
print("wrote oracle_small.npz; backend", orc.svd_backend(), "lambda[8]", est[8, 0], "noise", a, m_, s)
