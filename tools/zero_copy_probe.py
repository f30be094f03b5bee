"""Zero-copy output: the path kernel stores finished pixels straight into the caller's pinned HOST frame (mapped memory, posted
PCIe writes from finish_path) instead of into a device frame that is copied home in pieces.  Wall time per frame, against
rtb_render (pieces + cudaMemcpy2DAsync).  usage: zero_copy_probe.py"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
W, H = 3840, 2160
L = _lib.lib()
_lib.check(L.rtb_init(1, None), "init")
scene = R.main_scene(False); h = scene.upload()
v = R.main_viewport(W, H, 5, 1); v.seed = 7
host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
host2 = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
dev = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
s = _lib.RtbStats()
# reference frame through the normal API
_lib.check(L.rtb_render(h, C.byref(v), host2.numpy().ctypes.data, None, None, C.byref(s)), "render")
ref = host2.numpy().copy()
def run(name, flags, target_ptr, check):
    v.flags = flags
    ts = []
    for it in range(10):
        flush.fill_(it); torch.cuda.synchronize()
        if check is not None: check.zero_()
        t0 = time.perf_counter()
        _lib.check(L.rtb_render_device(h, C.byref(v), 0, 0, 1, target_ptr, None, None, C.c_void_p(st.cuda_stream), C.byref(s)), "r")
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ok = "" if check is None else f" frame == rtb_render's: {np.array_equal(check.numpy().view(np.uint32), ref.view(np.uint32))}"
    print(f"{name}: wall {np.mean(ts[3:])*1e3:.3f} ms (min {np.min(ts[3:])*1e3:.3f}){ok}", flush=True)
run("device frame, two launches      ", 0, dev.data_ptr(), None)
run("device frame, fused launch      ", _lib.RTB_FLAG_FUSED, dev.data_ptr(), None)
run("pinned host frame, two launches ", 0, host.data_ptr(), host)
run("pinned host frame, fused launch ", _lib.RTB_FLAG_FUSED, host.data_ptr(), host)
ts = []
for it in range(10):
    flush.fill_(it); torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.rtb_render(h, C.byref(v), host2.numpy().ctypes.data, None, None, C.byref(s)), "render")
    ts.append(time.perf_counter() - t0)
print(f"rtb_render (pieces + D2H copies)  : wall {np.mean(ts[3:])*1e3:.3f} ms (min {np.min(ts[3:])*1e3:.3f})")
