"""BASELINE configs[4] at FULL frame size on one B200: one 4096x4096 frame, patch 8, trajectory 31 (64x31 Casorati,
16.7 M patches per SVT object), PGURE lambda search + ARPS, noise estimated.  The 31-frame window is generated on the
GPU (Poisson-Gaussian, same model as mixed_noise_model) because numpy needs minutes for 5.5e8 Poisson draws.
Prints per-stage stats and peak device memory (developer tool, GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pgure-svt_b200"))
import numpy as np
import torch
from pguresvt import _pguresvt as b

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NF = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # output frames (steady state: ARPS pairs and noise slices are reused)
F = 32 + NF
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(123)
yy, xx = torch.meshgrid(torch.arange(N, device=dev, dtype=torch.float32), torch.arange(N, device=dev, dtype=torch.float32), indexing="ij")
X = np.empty((N, N, F), dtype=np.uint16, order="F")
for t in range(F):
    py = (yy - 0.3 * t) % 16 - 8; px = (xx - 0.2 * t) % 16 - 8
    clean = torch.exp(-(py * py + px * px) / (2 * 2.5 ** 2)) * (1.0 + 0.1 * np.sin(0.2 * t)) / 1.1 * 4095.0
    noisy = 0.1 * torch.poisson(clean / 0.1, generator=g) + 0.1 + 0.1 * torch.randn(N, N, device=dev, generator=g)
    X[:, :, t] = noisy.clamp_(0, 65535).to(torch.int32).cpu().numpy().astype(np.uint16)
del yy, xx, clean, noisy, py, px
torch.cuda.empty_cache()
kw = dict(trajectory_length=31, patch_size=8, optimize_pgure=True, lambda1=-1.0, random_seed=1)
if len(sys.argv) > 2 and sys.argv[2] == "known":
    kw.update(noise_alpha=0.05, noise_mu=0.03, noise_sigma=0.03)
t0 = time.time()
h = b.Handle(X, frame_begin=16, frame_end=16 + NF, **kw)
free0, tot = torch.cuda.mem_get_info()
t1 = time.time(); h.process(); wall = time.time() - t1
st = h.stats()
print("N", N, "frames", NF, "create_s", round(t1 - t0, 2), "process_wall_s", round(wall, 2), "device_GB_in_use", round((tot - free0) / 2 ** 30, 1),
      {k: round(float(v), 2) for k, v in st.items()})
Y, e = h.download(); print("lambda, alpha, mu, sigma", e[16:16 + NF, :], "finite", bool(np.isfinite(Y[:, :, 16]).all()),
                           "output range", float(Y[:, :, 16].min()), float(Y[:, :, 16].max()))
