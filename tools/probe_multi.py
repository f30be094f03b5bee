"""Probe: one process driving 1..N GPUs — rtb_render (bands) and rtb_render_progressive (samples): device vs wall time."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L = _lib.lib()
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H, SPP = 3840, 2160, 16
for n in [1, nmax]:
    _lib.check(L.rtb_init(n, None), "init")
    sc = R.main_scene(False)
    h = sc.upload()
    data = np.zeros((H, W, 4), np.float32)
    _lib.check(L.rtb_host_register(data.ctypes.data, data.nbytes), "reg")
    st = _lib.RtbStats()
    v1 = R.main_viewport(W, H, 5, 1); v1.seed = 7
    vp = R.main_viewport(W, H, 5, SPP); vp.seed = 7
    for name, fn in (("bands 1spp", lambda: L.rtb_render(h, C.byref(v1), data.ctypes.data, None, None, C.byref(st))),
                     (f"samples {SPP}spp", lambda: L.rtb_render_progressive(h, C.byref(vp), data.ctypes.data, C.byref(st)))):
        for _ in range(3): _lib.check(fn(), name)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); _lib.check(fn(), name); ts.append((time.perf_counter() - t0) * 1e3)
        print(f"n_gpus={n} {name}: wall {np.mean(ts):.2f} ms  ms_render(max over GPUs) {st.ms_render:.2f}  ms_total {st.ms_total:.2f}  rays {st.rays}", flush=True)
    L.rtb_host_unregister(data.ctypes.data)
    sc.release()
