"""Experiment: two handles (two contiguous half blocks) driven by two host threads on one GPU vs one handle."""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
X, _ = synthetic_sequence(1024, 14 + nf + 2, seed=1)
kw = dict(optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
fb = 8
h = b.Handle(X, frame_begin=fb, frame_end=fb + nf, **kw)
h.process()
t0 = time.time(); h.process(); t1 = time.time()
Y1, e1 = h.download(); h.close()
print("one handle:", nf, "frames", round((t1 - t0) * 1e3, 1), "ms ->", round(nf / (t1 - t0), 2), "fps")
for nsplit in (2, 3):
    bounds = [fb + (nf * i) // nsplit for i in range(nsplit + 1)]
    hs = [b.Handle(X, frame_begin=bounds[i], frame_end=bounds[i + 1], **kw) for i in range(nsplit)]
    def run(hh): hh.process()
    for rep in range(2):
        th = [threading.Thread(target=run, args=(hh,)) for hh in hs]
        t0 = time.time()
        # stagger the second thread so that the phases interleave
        for i, t in enumerate(th):
            t.start()
            if rep == 1 and i + 1 < len(th):
                time.sleep(0.03)
        for t in th: t.join()
        t1 = time.time()
        print(nsplit, "handles (rep", rep, "):", round((t1 - t0) * 1e3, 1), "ms ->", round(nf / (t1 - t0), 2), "fps")
    Y2 = np.zeros_like(Y1); e2 = np.zeros_like(e1)
    for hh in hs:
        hh.download(Y2, e2); hh.close()
    print("   max rel diff vs one handle", np.abs(Y2 - Y1).max() / np.abs(Y1).max(), np.abs(e2 - e1).max())
