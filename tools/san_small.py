import sys
sys.path.insert(0,'.'); sys.path.insert(0,'pgure-svt_b200'); sys.path.insert(0,'tests')
import numpy as np
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
X,_ = synthetic_sequence(32, 16, seed=123)
h = b.Handle(X, frame_begin=8, frame_end=9, optimize_pgure=True, lambda1=-1.0, noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03, random_seed=1)
h.process(); print(h.stats()); h.close()
