import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
N = 512
X, _ = synthetic_sequence(N, 16, seed=1)
h = b.Handle(X, frame_begin=8, frame_end=9, optimize_pgure=False, lambda1=0.15)
p = h.probe_arps(8)  # (2, vs, 15)
M1 = N - 3
py = p[0].reshape(M1, M1, 15, order="F")  # [row i, col j, k]
px = p[1].reshape(M1, M1, 15, order="F")
ii, jj = np.mgrid[0:M1, 0:M1]
my = py - ii[:, :, None]; mx = px - jj[:, :, None]
print("nonzero motion fraction", ((my != 0) | (mx != 0)).mean())
same_v = ((my[1:] == my[:-1]) & (mx[1:] == mx[:-1])).mean()
print("same motion as vertical neighbour", same_v)
# runs of 8 vertical neighbours all same
blk = (M1 // 8) * 8
a = my[:blk].reshape(blk // 8, 8, M1, 15); c = mx[:blk].reshape(blk // 8, 8, M1, 15)
print("8-run uniform", ((a == a[:, :1]).all(1) & (c == c[:, :1]).all(1)).mean())
print("|m| histogram", np.bincount(np.maximum(np.abs(my), np.abs(mx)).ravel(), minlength=8) / my.size)
