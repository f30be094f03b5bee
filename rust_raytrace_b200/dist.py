"""One-process-per-GPU plumbing (torchrun / torch.distributed) around the C ABI.

The reference shards a frame by image row over CPU threads (raytrace.rs:1179-1194).  Here the unit is a GPU:

* 1-spp frames: 8-row bands, band b -> rank b % world (`band_rows`, the host mirror of rtb_partition_rows).  Bands are
  disjoint, every rank renders its own into its own buffer: no data-path collective.
* multi-sample frames: samples [sample_range(spp, rank, world)) per rank over the FULL frame with
  RTB_FLAG_SUM_ONLY, then ONE reduce(sum) of the f32 accumulation buffers to rank 0 (NCCL over NVLink on GPUs,
  gloo in the CPU tests) and walk_ray_set's final `* (1/spp)` (raytrace.rs:1426) on the root.

Only host logic lives here; pixels are produced by librtb.so (or, in the CPU tests, by the oracle standing in).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def band_rows(height: int, rank: int, world: int) -> np.ndarray:
    """Image rows rendered by `rank` (ascending).  Same answer as the library's own partition."""
    rows = np.zeros(height, np.uint32)
    n = _lib.lib().rtb_partition_rows(int(height), int(rank), int(world), rows.ctypes.data, int(height))
    if n < 0:
        raise ValueError("rtb_partition_rows: bad rank/world")
    return rows[:n]


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous share [begin, end) of `spp` samples for `rank`; shares differ by at most one sample and a rank
    may get none when world > spp."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (spp * rank) // world, (spp * (rank + 1)) // world


def sample_view(view: _lib.RtbView, rank: int, world: int) -> _lib.RtbView | None:
    """The view `rank` renders in a sample-partitioned frame (None when its share is empty)."""
    b, e = sample_range(view.spp, rank, world)
    if b == e:
        return None
    v = _lib.RtbView.from_buffer_copy(view)
    v.sample_begin, v.sample_end = b, e
    v.flags |= _lib.RTB_FLAG_SUM_ONLY
    return v


def reduce_samples(acc, spp: int, dst: int = 0, group=None, scale=None):
    """Combine the per-rank sample sums: reduce(sum) to `dst`, then `* (1/spp)` on `dst`.

    acc   : torch tensor [H, W, 4] f32 holding this rank's un-normalised sum (zeros for an empty share)
    scale : callable(tensor, spp) applying the final scale in place; default = rtb_scale_device on CUDA tensors
            (the library's kernel), a plain f32 multiply on CPU tensors (gloo tests).
    Returns the tensor on `dst` (the normalised frame), None elsewhere."""
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if dist.get_rank(group) != dst:
            return None
    if scale is not None:
        scale(acc, spp)
    elif acc.is_cuda:
        st = torch.cuda.current_stream(acc.device).cuda_stream
        _lib.check(_lib.lib().rtb_scale_device(acc.data_ptr(), acc.numel() // 4, int(spp), 0, C.c_void_p(st)),
                   "rtb_scale_device")
    else:
        acc.mul_(np.float32(1.0) / np.float32(spp))
        acc[..., 3] = 0.0
    return acc
