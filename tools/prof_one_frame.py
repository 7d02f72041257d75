"""One 1024^2 config-4 frame through the handle API (for ncu captures).  python tools/prof_one_frame.py [size] [nframes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import synthetic_sequence  # noqa: E402
from pguresvt import _pguresvt as b  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, _ = synthetic_sequence(size, 15 + nfr - 1, seed=123)
h = b.Handle(X, frame_begin=7, frame_end=7 + nfr, optimize_pgure=True, lambda1=-1.0, random_seed=1)
h.process()
print(h.stats())
h.close()
