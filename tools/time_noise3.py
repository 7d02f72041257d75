import sys,time,os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
X,_=synthetic_sequence(1024,24,seed=1)
h=b.Handle(X,frame_begin=8,frame_end=14,optimize_pgure=True,lambda1=-1.0,random_seed=1)
for rep in range(3):
    h.upload(X)   # clears the caches
    t0=time.time(); r=h.probe_noise(8); print("cold window", round((time.time()-t0)*1e3,2),"ms")
    t0=time.time(); r=h.probe_noise(9); print("next window", round((time.time()-t0)*1e3,2),"ms")
