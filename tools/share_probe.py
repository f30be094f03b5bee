"""Per-stage device times (RTB_FLAG_TIMING) of a 1/world band share of the 4K benchmark frame, all on one GPU."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L = _lib.lib()
_lib.check(L.rtb_init(1, None), "init")
scene = R.main_scene(False); h = scene.upload()
MD = int(os.environ.get("MAXDEPTH", "5"))
v = R.main_viewport(3840, 2160, MD, 1); v.seed = 7; v.flags = _lib.RTB_FLAG_TIMING | int(os.environ.get("FLAGS", "0"))      # FLAGS=64: RTB_FLAG_FUSED
d = torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for world in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    acc = np.zeros(4); rays = 0
    for it in range(8):
        flush.fill_(it)
        s = _lib.RtbStats()
        _lib.check(L.rtb_render_device(h, C.byref(v), 0, 0, world, d.data_ptr(), None, None, C.c_void_p(st.cuda_stream), C.byref(s)), "r")
        if it >= 3: acc += np.array(s.ms_stage[:]); rays = int(s.rays)
    acc /= 5
    print(f"world {world}: rays {rays}  trace {acc[1]:.3f}  shade {acc[2]:.3f}  bounce {acc[3]:.3f}  sum {acc.sum():.3f} ms   (x{world}: {acc.sum()*world:.3f})", flush=True)
