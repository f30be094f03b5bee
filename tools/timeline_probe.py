"""Per-warp timeline of the bounce launch (librtb built with -DWF_TIMELINE prints `TL cta warp t_start t_exhausted t_end rays`):
when the launch's warps start, when each sees the queue run dry and when it exits.  usage: timeline_probe.py WORLD"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L = _lib.lib()
_lib.check(L.rtb_init(1, None), "init")
scene = R.main_scene(False); h = scene.upload()
v = R.main_viewport(3840, 2160, 5, 1); v.seed = 7; v.flags = _lib.RTB_FLAG_TIMING
d = torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
world = int(sys.argv[1])
for it in range(4):
    flush.fill_(it)
    s = _lib.RtbStats()
    _lib.check(L.rtb_render_device(h, C.byref(v), 0, 0, world, d.data_ptr(), None, None, C.c_void_p(st.cuda_stream), C.byref(s)), "r")
    torch.cuda.synchronize()
    sys.stdout.flush()
    os.write(1, f"FRAME {it} world {world} stages {list(s.ms_stage)[:4]}\n".encode())
