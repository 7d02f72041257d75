"""Time one PGURE frame of the config-5 shape (patch 8, trajectory 31 -> 64x31 Casorati) at a given frame size and
print per-stage stats (developer tool, GPU box)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_sequence
from pguresvt import _pguresvt as b
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 1
X, _ = synthetic_sequence(N, 31 + nf + 1, seed=1)
kw = dict(trajectory_length=31, patch_size=8, optimize_pgure=True, lambda1=-1.0, random_seed=1)  # noise estimated
if len(sys.argv) > 3 and sys.argv[3] == "known":
    kw.update(noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03)
if len(sys.argv) > 4:
    kw.update(rank_cache=int(sys.argv[4]))
if len(sys.argv) > 3 and sys.argv[3] == "fixed":
    kw = dict(trajectory_length=31, patch_size=8, optimize_pgure=False, lambda1=0.15)
h = b.Handle(X, frame_begin=16, frame_end=16 + nf, **kw)
h.process()
t0 = time.time(); h.process(); wall = time.time() - t0
st = h.stats()
print("N", N, "wall_s", round(wall, 3), {k: round(v, 2) for k, v in st.items()}, "ms/eval",
      round(st["ms_search"] / max(st["evals"], 1), 3))
Y, e = h.download(); print("lambda, alpha, mu, sigma", e[16:16 + nf, :])
