import sys,time; sys.path.insert(0,"pgure-svt_b200"); sys.path.insert(0,"tests")
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
X,_=synthetic_sequence(1024,24,seed=1)
h=b.Handle(X,frame_begin=8,frame_end=14,optimize_pgure=True,lambda1=-1.0,random_seed=1)
h.process(); st=h.stats(); print({k:round(v,2) for k,v in st.items()})
Y,e=h.download(); print(e[8:14])
