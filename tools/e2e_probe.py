"""Breaks the end-to-end frame time of rtb_render down: PCIe floor, device time, host wall time."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib

W, H = 3840, 2160
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = _lib.lib()
_lib.check(L.rtb_init(n_gpus, None), "init")
scene = R.main_scene(False)
h = scene.upload()
v = R.main_viewport(W, H, 5, 1)
v.seed = 7
host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
dev = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
torch.cuda.synchronize()
for _ in range(3):
    host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"PCIe floor: D2H of {W*H*16/1e6:.0f} MB pinned: {dt*1e3:.3f} ms = {W*H*16/dt/1e9:.1f} GB/s")
st = _lib.RtbStats()
data = host.numpy()
for _ in range(3):
    _lib.check(L.rtb_render(h, C.byref(v), data.ctypes.data, None, None, C.byref(st)), "render")
ts, dev_ms, tot_ms = [], [], []
for _ in range(10):
    t0 = time.perf_counter()
    _lib.check(L.rtb_render(h, C.byref(v), data.ctypes.data, None, None, C.byref(st)), "render")
    ts.append(time.perf_counter() - t0); dev_ms.append(st.ms_render); tot_ms.append(st.ms_total)
print(f"rtb_render x{n_gpus} GPU: wall {np.mean(ts)*1e3:.3f} ms (min {np.min(ts)*1e3:.3f}), device compute {np.mean(dev_ms):.3f} ms, "
      f"lib total {np.mean(tot_ms):.3f} ms, launches {st.kernel_launches}, rays {st.rays}")
