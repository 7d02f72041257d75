import sys, os, time
sys.path.insert(0,'.'); sys.path.insert(0,'pgure-svt_b200'); sys.path.insert(0,'tests')
import numpy as np
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
X,_ = synthetic_sequence(1024, 32, seed=77)
kw = dict(optimize_pgure=True, lambda1=-1.0, random_seed=1, n_gpus=1)
for i in range(3):
    t0=time.perf_counter(); Y,e,_ = b.pguresvt_u16(X, **kw); print("call", i, time.perf_counter()-t0, file=sys.stderr)
