"""world_size-2 CPU test (gloo) of the N>1 host logic: the band partition the GPU ranks use
(rtb_partition_rows: 8-row band b -> rank b % world), the per-rank ray counting, and the gather of the
disjoint bands into one frame.  Each rank renders ITS rows with the CPU oracle (the checker standing in for
the GPU, which does not exist here); the summed frame must equal the single-process oracle frame bit for bit
and the ray counts must add up."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_raytrace_b200 as R
    from oracle import oracle as O
    from rust_raytrace_b200 import _lib

    rows = np.zeros(H, np.uint32)
    n = _lib.lib().rtb_partition_rows(H, rank, world, rows.ctypes.data, H)
    rows = rows[:n]
    scene = R.main_scene(deterministic=False)
    osc = O.Scene(scene.tris.view(O.TRI_DTYPE), O.ACCEL_BVH)
    ov = O.main_viewport(W, H, 5, 1)
    img = np.zeros((H, W, 4), np.float32)
    rays = 0
    # bands are contiguous runs of <= 8 rows
    start = 0
    while start < len(rows):
        end = start
        while end + 1 < len(rows) and rows[end + 1] == rows[end] + 1:
            end += 1
        rgba, _, _, st = osc.render(ov, seed=5, threads=1, rows=(int(rows[start]), int(rows[end]) + 1), want_ids=False)
        img[rows[start]:rows[end] + 1] = rgba[rows[start]:rows[end] + 1]
        rays += st.rays
        start = end + 1
    t = torch.from_numpy(img)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)            # disjoint bands: the sum is the gather
    cnt = torch.tensor([rays, n], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "frame.npy"), t.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), cnt.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_band_partition_gather_gloo(tmp_path, world, O, R):
    import torch.multiprocessing as mp

    W, H = 96, 52   # 6.5 bands: ragged last band
    port = _free_port()
    mp.spawn(_worker, args=(world, port, W, H, str(tmp_path)), nprocs=world, join=True)
    frame = np.load(tmp_path / "frame.npy")
    counts = np.load(tmp_path / "counts.npy")
    scene = R.main_scene(deterministic=False)
    ref, _, _, st = O.Scene(scene.tris.view(O.TRI_DTYPE), O.ACCEL_BVH).render(O.main_viewport(W, H, 5, 1), seed=5)
    assert counts[1] == H
    assert counts[0] == st.rays
    assert np.array_equal(frame.view(np.uint32), ref.view(np.uint32))
