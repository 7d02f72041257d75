"""Generates the FULL-SIZE oracle fixtures tests/golden/oracle_full_<name>.npz (BASELINE.json configs 3, 4 and a
config-5-shaped frame) with the CPU oracle (oracle/oracle.cpp).  Run in the build container (one 1024^2 PGURE frame is
about ten CPU-minutes); the GPU box only reads the committed fixtures.

    python tests/golden/make_golden_full.py [c3] [c4] [c5]

A fixture holds what a parity test needs without the multi-megabyte frames: the per-frame estimates (lambda, alpha,
mu, sigma), a strided sample of the denoised frame, its 16x16 block sums, the SHA-256 and a strided sample of the ARPS
trajectories, the PGURE objective at a few lambdas, and the generator arguments of the input (conftest.synthetic_sequence
is seeded, so the input is regenerated, not stored)."""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_sequence  # noqa: E402
from oracle import orc  # noqa: E402

CASES = {
    # BASELINE.json configs[3]: 1024^2, patch 4, trajectory 15, per-frame PGURE lambda, ARPS, median 5, noise estimated
    "c4": dict(N=1024, F=17, t=8, seed=123, kw=dict(trajectory_length=15, patch_size=4, patch_overlap=1, motion_window=7,
                                                    motion_filter=5, noise_method=4, max_iter=500, random_seed=1,
                                                    optimize_pgure=True, exponential_weighting=True, motion_estimation=True,
                                                    lambda1=-1.0, tol=1e-7)),
    # BASELINE.json configs[2]: 512^2, fixed lambda (pure SVT path)
    "c3": dict(N=512, F=17, t=8, seed=123, kw=dict(trajectory_length=15, patch_size=4, patch_overlap=1, motion_window=7,
                                                   motion_filter=5, noise_method=4, max_iter=500, random_seed=1,
                                                   optimize_pgure=False, exponential_weighting=True, motion_estimation=True,
                                                   lambda1=0.15, tol=1e-7)),
    # BASELINE.json configs[4] shape (64x31 Casorati, PGURE + ARPS) on a 256^2 frame; on this data lambda ends at the upper
    # bound region like the full-size run (profiles/r01/config5_*)
    "c5": dict(N=256, F=33, t=16, seed=123, kw=dict(trajectory_length=31, patch_size=8, patch_overlap=1, motion_window=7,
                                                    motion_filter=5, noise_method=4, max_iter=500, random_seed=1,
                                                    optimize_pgure=True, exponential_weighting=True, motion_estimation=True,
                                                    lambda1=-1.0, tol=1e-7)),
}


def block_sums(a, b=16):
    n = a.shape[0] // b
    return a[: n * b, : n * b].reshape(n, b, n, b).sum(axis=(1, 3))


def make(name):
    c = CASES[name]
    N, F, t, kw = c["N"], c["F"], c["t"], c["kw"]
    fw = kw["trajectory_length"] // 2
    bs = kw["patch_size"]
    X, _ = synthetic_sequence(N, F, seed=c["seed"])
    out = dict(N=N, F=F, t=t, seed=c["seed"], svd_backend=orc.svd_backend(), kw_keys=np.array(list(kw.keys())),
               kw_vals=np.array([float(v) for v in kw.values()]))
    t0 = time.time()
    Y, est = orc.pguresvt(X, n_jobs=1, frame_begin=t, frame_end=t + 1, **kw)
    print(name, "pipeline", round(time.time() - t0, 1), "s; est", est[t], flush=True)
    y = np.ascontiguousarray(Y[:, :, t])
    out["est"] = est[t].copy()
    out["y_stride"] = 8
    out["y_sample"] = y[::8, ::8].copy()
    out["y_blocksum"] = block_sums(y)
    out["y_max"] = np.abs(y).max()
    out["y_sha256"] = hashlib.sha256(y.tobytes()).hexdigest()
    # stage-level: trajectories and objective of the same frame
    Z = np.stack([orc.median_u16(X[:, :, i], kw["motion_filter"]) for i in range(t - fw, t + fw + 1)], axis=2)
    w = Z.astype(np.float64)
    w /= w.max()
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    p, _, nc = orc.arps(w, bs, t, fw, kw["motion_window"], F, True)
    p16 = np.ascontiguousarray(p.astype(np.int16))  # (2, vecSize, win) F-order in the oracle; C-order bytes hashed here
    out["arps_sha256"] = hashlib.sha256(p16.tobytes()).hexdigest()
    out["arps_stride"] = 97
    out["arps_sample"] = p16[:, ::97, :].copy()
    out["arps_ncost"] = nc
    if kw["optimize_pgure"]:
        a, m, s = est[t, 1], est[t, 2], est[t, 3]
        na, nm, ns, _ = orc.noise_estimate(u, kw["noise_method"])
        out["noise"] = np.array([na, nm, ns])
        t0 = time.time()
        P = orc.PGUREObj(u, p, a, s, m, bs, 1, kw["random_seed"], True, True)  # PGURE(alpha, sigma, mu): pguresvt.hpp:133
        lam = est[t, 0]
        lams = np.array([0.5 * lam, 0.9 * lam, lam, 1.1 * lam, 2.0 * lam])
        vals, terms = zip(*[P.calc(float(x)) for x in lams])
        out["pgure_lambdas"] = lams
        out["pgure_values"] = np.array(vals)
        out["pgure_terms"] = np.array(terms)
        print(name, "objective", round(time.time() - t0, 1), "s", vals, flush=True)
        del P
    np.savez_compressed(os.path.join(HERE, f"oracle_full_{name}.npz"), **out)
    print("wrote", name, flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["c3", "c5", "c4"]):
        make(nm)
