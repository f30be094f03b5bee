#!/bin/bash
# round 2, run H: extension A/B on a deep scene (teapot + light, 4K, maxdepth 5) and on circles at maxdepth 2..5
python - <<'PY'
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L=_lib.lib(); _lib.check(L.rtb_init(1,None),"init")
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
st=torch.cuda.Stream(); torch.cuda.set_stream(st)
def run(s,W,H,md,tag):
    h=s.upload()
    d=torch.zeros((H,W,4),dtype=torch.float32,device="cuda")
    for name,fl in (("wavefront EXT",0),("k_trace_ext",_lib.RTB_FLAG_MEGAKERNEL)):
        v=R.main_viewport(W,H,md,1); v.seed=7; v.flags=fl
        ms=[]
        for it in range(8):
            flush.fill_(it); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            a.record(st); _lib.check(L.rtb_render_device(h,C.byref(v),0,0,1,d.data_ptr(),None,None,C.c_void_p(st.cuda_stream),None),"r"); b.record(st)
            torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        print(tag, name, "ms/frame %.3f"%np.mean(ms[3:]), flush=True)
    s.release()
s=R.main_scene(False); s.set_light((6.0,-2.0,0.0),0.5); run(s,3840,2160,5,"teapot+light 4K md5")
for md in (2,3,5):
    run(R.circles_scene(),2560,1440,md,"circles 2K md%d"%md)
PY
