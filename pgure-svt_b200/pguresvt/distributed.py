"""Frame-sharded denoising over several GPUs of one node: one process per GPU, torch.distributed for the plumbing.

The reference fans frames out over threads with a static contiguous partition (`pguresvt::parallel`,
src/utils.hpp:150-166: `tasksPerThread = ceil(n / workers)`, worker w takes `[w*tpt, min((w+1)*tpt, n))`).  Frames are
independent jobs (pguresvt.hpp:90-169), so the same partition maps onto GPUs with no data-path exchange: every rank
denoises its block (the C ABI loads the block plus the fw halo frames its windows need and applies the first/last
window rules with the GLOBAL frame count), then one all-gather assembles the sequence on every rank.
"""
import numpy as np


def frame_block(rank, world, n_frames):
    """Contiguous block [begin, end) of rank `rank` — the partition of utils.hpp:150-166."""
    per = (n_frames + world - 1) // world
    begin = min(rank * per, n_frames)
    return begin, min(begin + per, n_frames)


def _gpu_block(X, begin, end, **kw):
    from . import _pguresvt as bridge

    h = bridge.Handle(X, frame_begin=begin, frame_end=end, **kw)
    try:
        h.process()
        Y, est = h.download()
    finally:
        h.close()
    return Y[:, :, begin:end], est[begin:end]


def denoise_sharded(X, compute_block=None, group=None, **kw):
    """Denoise X (rows, cols, frames) with the frames sharded over the ranks of `group`.

    Every rank must pass the same X and kwargs (the kwargs of the `_pguresvt` entry points).  Returns
    (Y (rows, cols, frames) float64 F-order, estimates (frames, 4)) on every rank.  `compute_block(X, begin, end,
    **kw) -> (Y_block, est_block)` defaults to the CUDA path; tests inject a CPU stand-in to exercise the sharding
    and gather logic without a GPU.
    """
    import torch
    import torch.distributed as dist

    X = np.asfortranarray(X)
    nr, nc, nf = X.shape
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, end = frame_block(rank, world, nf)
    fn = compute_block or _gpu_block
    per = (nf + world - 1) // world
    Yb = np.zeros((per, nc, nr), dtype=np.float64)  # frame-major so that the gather concatenates frames
    eb = np.zeros((per, 4), dtype=np.float64)
    if end > begin:
        Yblk, est = fn(X, begin, end, **kw)
        Yb[: end - begin] = np.transpose(Yblk, (2, 1, 0))
        eb[: end - begin] = est
    if world == 1:
        Yall, eall = Yb, eb
    else:
        use_cuda = dist.get_backend(group) == "nccl"
        if use_cuda:
            # gather on the SAME device the C ABI computes on (make_params: PGURESVT_DEVICE / LOCAL_RANK), not on whatever
            # torch's current device happens to be: every rank on cuda:0 is an NCCL duplicate-GPU error
            from ._pguresvt import make_params

            dev = torch.device("cuda", int(kw.get("device", make_params().device)))
        else:
            dev = torch.device("cpu")
        ty, te = torch.from_numpy(Yb).to(dev), torch.from_numpy(eb).to(dev)
        gy = torch.empty((world * per, nc, nr), dtype=ty.dtype, device=dev)
        ge = torch.empty((world * per, 4), dtype=te.dtype, device=dev)
        dist.all_gather_into_tensor(gy, ty, group=group)
        dist.all_gather_into_tensor(ge, te, group=group)
        Yall = gy.reshape(world * per, nc, nr).cpu().numpy()
        eall = ge.reshape(world * per, 4).cpu().numpy()
    Y = np.asfortranarray(np.transpose(Yall[:nf], (2, 1, 0)))
    return Y, np.asfortranarray(eall[:nf])
