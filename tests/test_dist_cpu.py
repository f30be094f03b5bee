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


def _sample_worker(rank, world, port, W, H, spp, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_raytrace_b200 as R
    from oracle import oracle as O
    from rust_raytrace_b200 import dist as RD

    scene = R.main_scene(deterministic=False)
    osc = O.Scene(scene.tris.view(O.TRI_DTYPE), O.ACCEL_BVH)
    ov = O.main_viewport(W, H, 5, spp)
    view = R.main_viewport(W, H, 5, spp)
    sv = RD.sample_view(view, rank, world)
    if sv is None:
        acc, rays = np.zeros((H, W, 4), np.float32), 0
    else:
        assert sv.flags & 1                      # RTB_FLAG_SUM_ONLY
        acc, st = osc.render_samples(ov, sv.sample_begin, sv.sample_end, seed=11, threads=1, sum_only=True)
        rays = st.rays
    t = torch.from_numpy(acc)
    out = RD.reduce_samples(t, spp, dst=0)
    cnt = torch.tensor([rays], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "frame.npy"), out.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), cnt.numpy())
    else:
        assert out is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 5), (3, 2)])
def test_sample_partition_reduce_gloo(tmp_path, world, spp, O, R):
    """Multi-sample frame, samples partitioned over ranks, one reduce(sum) + 1/spp on the root (dist.py).  The f32
    summation order differs from the single-process loop (raytrace.rs:1418-1426), so the frame agrees to rounding
    (checked tightly) rather than bit for bit; the ray counts must add up exactly."""
    import torch.multiprocessing as mp

    W, H = 64, 40
    port = _free_port()
    mp.spawn(_sample_worker, args=(world, port, W, H, spp, str(tmp_path)), nprocs=world, join=True)
    frame = np.load(tmp_path / "frame.npy")
    counts = np.load(tmp_path / "counts.npy")
    scene = R.main_scene(deterministic=False)
    ref, _, _, st = O.Scene(scene.tris.view(O.TRI_DTYPE), O.ACCEL_BVH).render(O.main_viewport(W, H, 5, spp), seed=11)
    assert counts[0] == st.rays
    assert np.all(frame[..., 3] == 0)
    assert np.max(np.abs(frame - ref)) < 1e-6


def test_sample_range_covers_all_samples(R):
    from rust_raytrace_b200 import dist as RD
    for spp in (1, 2, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [RD.sample_range(spp, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == spp
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    for h in (1, 8, 52, 2160):
        for world in (1, 2, 8):
            rows = np.concatenate([RD.band_rows(h, r, world) for r in range(world)])
            assert np.array_equal(np.sort(rows), np.arange(h))
