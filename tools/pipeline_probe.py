"""How much of the path kernel's tail would the NEXT frame hide?  K frames of a 1/world band share of the 4K benchmark frame on
one GPU: (a) one after the other on one stream (what bench.py times), (b) alternating between two scene handles / workspaces /
streams / output frames, so that frame i+1's primary launch starts in the SM slots frame i's bounce launch frees as its warps
exit.  Not used by bench.py: the metric is one frame at a time."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L = _lib.lib()
_lib.check(L.rtb_init(1, None), "init")
scenes = [R.main_scene(False) for _ in range(2)]      # kept alive: a Scene releases its handle when it is collected
hs = [sc.upload() for sc in scenes]
v = R.main_viewport(3840, 2160, 5, 1); v.seed = 7
frames = [torch.zeros((2160, 3840, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
K = 40
def run(world, two):
    for rep in range(2):                       # first pass warms up (workspaces, L2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(K):
            j = (i & 1) if two else 0
            _lib.check(L.rtb_render_device(hs[j], C.byref(v), 0, 0, world, frames[j].data_ptr(), None, None,
                                           C.c_void_p(streams[j].cuda_stream), None), "render")
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    return dt * 1e3
for world in [int(x) for x in (sys.argv[1:] or ["1", "8"])]:
    a = run(world, False); b = run(world, True)
    same = torch.equal(frames[0], frames[1])
    print(f"world {world}: one frame at a time {a:.3f} ms/frame, two frames in flight {b:.3f} ms/frame ({a / b:.2f}x), frames identical: {same}", flush=True)
