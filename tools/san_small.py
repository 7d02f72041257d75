"""Small end-to-end runs of every evaluation path for compute-sanitizer (memcheck / racecheck) on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import synthetic_sequence  # noqa: E402
from pguresvt import _pguresvt as b  # noqa: E402

X, _ = synthetic_sequence(64, 18, seed=123)
base = dict(optimize_pgure=True, lambda1=-1.0, random_seed=1)
for name, env, kw in [("lean top1 (default)", {}, {}), ("tile eval", {"PGURESVT_TILE_EVAL": "1"}, {}),
                      ("jacobi lean (top1 off)", {"PGURESVT_TOP1": "0"}, {}), ("eps1 mode 1", {}, {"eps1_mode": 1}),
                      ("compact 64x15", {}, {"patch_size": 8})]:
    for k, v in env.items():
        os.environ[k] = v
    h = b.Handle(X, frame_begin=8, frame_end=10, **base, **kw)
    h.process()
    Y, e = h.download()
    st = h.stats()
    print(name, "lambda", e[8:10, 0], "evals", st["evals"], "exact", st["lean_exact_svds"], flush=True)
    h.close()
    for k in env:
        os.environ.pop(k)
Y1, e1, _ = b.pguresvt_u16(X, n_gpus=1, **base)
print("one-shot", e1[0, :3])
