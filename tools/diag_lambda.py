"""Diagnostic (GPU box): where does the lambda search of the device path leave the oracle's on the bench's CPU sample?
For every sampled frame: the oracle's probe sequence (its SBPLX restatement driving its own objective), the device
objective at the same probes, and the device's own sequence; prints the first probe at which the two searches differ
and the objective disagreement there.  python tools/diag_lambda.py [crop] [nframes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import synthetic_sequence  # noqa: E402
from oracle import orc  # noqa: E402
from pguresvt import _pguresvt as b  # noqa: E402

crop = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fw = 7
X, _ = synthetic_sequence(crop, nfr + 2 * fw, seed=123)
kw = dict(trajectory_length=15, patch_size=4, patch_overlap=1, motion_window=7, motion_filter=5, noise_method=4, max_iter=500,
          random_seed=1, exponential_weighting=True, motion_estimation=True, tol=1e-7, optimize_pgure=True, lambda1=-1.0)
Yo, eo = orc.pguresvt(X, n_jobs=os.cpu_count(), frame_begin=fw, frame_end=fw + nfr, **kw)
h = b.Handle(X, frame_begin=fw, frame_end=fw + nfr, **kw)
h.process()
Y, e = h.download()
rel = np.abs(e[fw:fw + nfr, 0] - eo[fw:fw + nfr, 0]) / eo[fw:fw + nfr, 0]
print("lambda rel err per frame:", " ".join(f"{r:.1e}" for r in rel))
for t in range(fw, fw + nfr):
    a, m, s = eo[t, 1:]
    Z = np.stack([orc.median_u16(X[:, :, i], 5) for i in range(t - fw, t + fw + 1)], axis=2).astype(np.float64)
    u = X[:, :, t - fw:t + fw + 1].astype(np.float64)
    u /= u.max()
    p, _, _ = orc.arps(Z / Z.max(), 4, t, fw, 7, X.shape[2], True)
    P = orc.PGUREObj(u, p, a, s, m, 4, 1, 1, True, True)
    start = u.sum() / (crop * crop * 15)
    tr_o = orc.sbplx_1d(lambda x: P.calc(x)[0], start, 0.0, max(100.0, start), np.sqrt(start), 1e-7, 1e-12, 500)["trace"]
    xs = np.array([x for x, _ in tr_o])
    vo = np.array([v for _, v in tr_o])
    vd, _ = h.probe_pgure(t, a, m, s, xs)
    dis = np.abs(vd - vo) / np.abs(vo)
    tr_d = orc.sbplx_1d(lambda x: float(h.probe_pgure(t, a, m, s, [x])[0][0]), start, 0.0, max(100.0, start), np.sqrt(start), 1e-7,
                        1e-12, 500)["trace"]
    n = min(len(tr_o), len(tr_d))
    first = next((i for i in range(n) if tr_o[i][0] != tr_d[i][0]), None)
    print(f"frame {t}: oracle {len(tr_o)} probes -> {tr_o[-1][0]:.9g}; device {len(tr_d)} probes -> {tr_d[-1][0]:.9g}; "
          f"pipeline {e[t, 0]:.9g} / {eo[t, 0]:.9g}; objective rel disagreement max {dis.max():.2e} median {np.median(dis):.2e}")
    if first is not None:
        lo = max(0, first - 3)
        print("   first differing probe", first, "of", n)
        for i in range(lo, min(n, first + 2)):
            print(f"     {i}: oracle x={tr_o[i][0]:.12g} f={tr_o[i][1]:.15e} | device x={tr_d[i][0]:.12g} f={tr_d[i][1]:.15e} "
                  f"| device f at oracle x={vd[i]:.15e}")
h.close()
