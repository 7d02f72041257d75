"""Diagnostic (not a pytest): prints stage-by-stage parity of the CUDA path vs the oracle and first timings.
Run on the GPU box: python tests/diag_gpu.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_sequence  # noqa: E402
from oracle import orc  # noqa: E402
from pguresvt import _pguresvt as bridge  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_small.npz"))
    X = g["X"]
    import ctypes
    name = ctypes.create_string_buffer(64)
    print("device:", bridge.load().pguresvt_device_info(0, name, 64), name.value.decode())

    def step(msg, fn):
        try:
            t = time.time()
            r = fn()
            print(f"[ok ] {msg}: {r}  ({time.time() - t:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[ERR] {msg}: {type(e).__name__}: {e}", flush=True)

    h = bridge.Handle(X, optimize_pgure=True, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
    step("median mismatches", lambda: int((h.probe_median(8) != g["Z"][:, :, 8]).sum()))
    step("perturbation mismatches", lambda: [int((a != b).sum()) for a, b in zip(h.probe_perturbations(), (g["delta1"], g["delta2neg"]))])
    step("arps mismatches", lambda: int((h.probe_arps(8).astype(np.int16) != g["patches8"]).sum()))
    u = X[:, :, 1:16].astype(np.float64)
    u /= u.max()
    o = orc.SVTObj(g["patches8"].astype(np.int64), 32, 15, 4, 1, True)
    o.decompose(u)
    So = o.singular_values()
    step("singular values rel err (register kernel)", lambda: float(np.abs(h.probe_singular_values(8, 0) - So).max() / So.max()))
    for lam in (0.0, 0.15, 50.0):
        vo = o.reconstruct(lam)
        step(f"reconstruct lam={lam} abs err", lambda: float(np.abs(h.probe_reconstruct(8, lam) - vo).max()))

    def pg():
        vals, terms = h.probe_pgure(8, 0.05, 0.03, 0.03, g["pgure_lambdas"])
        return float(np.abs(vals - g["pgure_values"]).max() / np.abs(g["pgure_values"]).max()), \
            float(np.abs(terms - g["pgure_terms"]).max())
    step("pgure objective rel err / terms abs err", pg)
    h.close()
    h = bridge.Handle(X, optimize_pgure=False, lambda1=0.15, svd_kernel=1)
    step("singular values rel err (smem kernel)", lambda: float(np.abs(h.probe_singular_values(8, 0) - So).max() / So.max()))
    h.close()

    def full(**kw):
        hh = bridge.Handle(X, **kw)
        hh.process()
        Y, e = hh.download()
        st = hh.stats()
        hh.close()
        return Y, e, st
    def fixed():
        Y, e, st = full(optimize_pgure=False, lambda1=0.15, random_seed=1)
        return float(np.abs(Y - g["Y_fixed"]).max() / np.abs(g["Y_fixed"]).max()), {k: round(v, 2) for k, v in st.items()}
    step("fixed-lambda full pipeline rel err", fixed)
    def pgure():
        Y, e, st = full(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
        eo = g["est_pgure"]
        return float(np.abs(e[:, 0] - eo[:, 0]).max() / eo[:, 0].max()), float(np.abs(Y - g["Y_pgure"]).max() / np.abs(g["Y_pgure"]).max()), \
            {k: round(v, 2) for k, v in st.items()}
    step("PGURE full pipeline: lambda rel err, pixel rel err", pgure)

    # first timings at real sizes
    for N, F, kw in [(512, 17, dict(optimize_pgure=False, lambda1=0.15)),
                     (1024, 17, dict(optimize_pgure=False, lambda1=0.15)),
                     (1024, 17, dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.1, noise_mu=0.02, noise_sigma=0.02, random_seed=1)),
                     (1024, 17, dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.1, noise_mu=0.02, noise_sigma=0.02, random_seed=1, motion_estimation=False)),
                     (512, 17, dict(optimize_pgure=False, lambda1=0.15, svd_kernel=1))]:
        def timing():
            Xb, _ = synthetic_sequence(N, F, seed=1)
            hh = bridge.Handle(Xb, frame_begin=7, frame_end=9, **kw)
            hh.process()
            hh.process()
            st = hh.stats()
            Y, e = hh.download()
            hh.close()
            return {k: round(v, 2) for k, v in st.items()}, "lambda", e[7:9, 0].tolist()
        step(f"timing N={N} {kw}", timing)


if __name__ == "__main__":
    main()
