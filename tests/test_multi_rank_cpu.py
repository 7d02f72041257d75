"""world_size-2 gloo test (CPU) of the frame-sharded path: block partition (utils.hpp:150-166), first/last-window
rules under sharding, and the all-gather assembly.  The compute per block is the CPU oracle (injected), so no GPU
is needed; the GPU flavour of the same function is exercised by bench.py under torchrun."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, synthetic_sequence
from oracle import orc


def test_frame_block_matches_reference_partition():
    from pguresvt.distributed import frame_block

    for n, w in [(16, 2), (17, 2), (1000, 8), (5, 8), (23, 4)]:
        blocks = [frame_block(r, w, n) for r in range(w)]
        per = (n + w - 1) // w
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for r, (b, e) in enumerate(blocks):
            assert b == min(r * per, n) and e == min(b + per, n)
        assert sum(e - b for b, e in blocks) == n


def test_two_rank_gloo_equals_single_process(tmp_path):
    X, _ = synthetic_sequence(32, 19, seed=31)
    np.save(tmp_path / "X.npy", X)
    script = textwrap.dedent(f"""
        import sys, numpy as np, torch.distributed as dist
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'pgure-svt_b200')!r})
        from oracle import orc
        from pguresvt.distributed import denoise_sharded
        dist.init_process_group("gloo")
        X = np.load({str(tmp_path / 'X.npy')!r})
        def block(X, b, e, **kw):
            Y, est = orc.pguresvt(X, frame_begin=b, frame_end=e, n_jobs=1, **kw)
            return Y[:, :, b:e], est[b:e]
        Y, est = denoise_sharded(X, compute_block=block, optimize_pgure=False, lambda1=0.15)
        np.save({str(tmp_path)!r} + f"/Y{{dist.get_rank()}}.npy", Y)
        np.save({str(tmp_path)!r} + f"/e{{dist.get_rank()}}.npy", est)
        dist.destroy_process_group()
    """)
    (tmp_path / "run.py").write_text(script)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", str(tmp_path / "run.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    want, west = orc.pguresvt(X, optimize_pgure=False, lambda1=0.15, n_jobs=2)
    for rank in (0, 1):
        Y = np.load(tmp_path / f"Y{rank}.npy")
        e = np.load(tmp_path / f"e{rank}.npy")
        assert np.array_equal(Y, want) and np.array_equal(e, west)
