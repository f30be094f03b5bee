#!/usr/bin/env python
"""Times the 4K benchmark frame (device-resident output) for a list of env-var settings, one subprocess each.

    python tools/tune.py "RTB_WF_DESCEND=4 RTB_WF_REFILL=16" "RTB_WF_DESCEND=8" ...
    python tools/tune.py --world 8 ""          # per-rank times of an 8-way band split, all on one GPU

Prints ms/frame (mean of 10 after 3 warm-ups, L2 flushed between) and Mrays/s per setting."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(world):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import rust_raytrace_b200 as R
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    _lib.check(L.rtb_init(1, None), "init")
    scene = R.main_scene(False)
    h = scene.upload()
    v = R.main_viewport(3840, 2160, 5, 1)
    v.seed = 7
    d = torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    out = []
    for rank in range(world):
        s = _lib.RtbStats()
        _lib.check(L.rtb_render_device(h, C.byref(v), 0, rank, world, d.data_ptr(), None, None, C.c_void_p(st.cuda_stream), C.byref(s)), "r")
        rays = int(s.rays)
        for _ in range(3):
            flush.fill_(1)
            L.rtb_render_device(h, C.byref(v), 0, rank, world, d.data_ptr(), None, None, C.c_void_p(st.cuda_stream), None)
        ms = []
        for _ in range(10):
            flush.fill_(0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            L.rtb_render_device(h, C.byref(v), 0, rank, world, d.data_ptr(), None, None, C.c_void_p(st.cuda_stream), None)
            b.record(st)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        out.append((rank, rays, float(np.mean(ms)), float(np.min(ms))))
    for rank, rays, m, mn in out:
        print(f"  rank {rank}/{world}: {rays} rays  mean {m:.3f} ms  min {mn:.3f} ms  {rays / m / 1e3:.0f} Mrays/s", flush=True)
    if world > 1:
        worst = max(o[2] for o in out)
        tot = sum(o[1] for o in out)
        print(f"  => frame {worst:.3f} ms (slowest rank), {tot / worst / 1e3:.0f} Mrays/s aggregate", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        sys.exit(0)
    args = sys.argv[1:]
    world = 1
    if args and args[0] == "--world":
        world = int(args[1]); args = args[2:]
    for setting in (args or [""]):
        env = dict(os.environ)
        for kv in setting.split():
            k, v = kv.split("=", 1)
            env[k] = v
        print(f"[{setting or 'defaults'}]", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(world)], env=env, check=False)
