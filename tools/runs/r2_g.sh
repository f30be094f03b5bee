#!/bin/bash
# round 2, run G: the extension on the wavefront renderer (EXT variants): tests, circles2k bench, A/B against k_trace_ext
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "extension or shadow or circles or sphere or golden or 640" > gpurun_out/r2_g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_g_tests.log
tail -6 gpurun_out/r2_g_tests.log
timeout 600 python bench.py --workload circles2k --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_g_circles.json 2> gpurun_out/r2_g_circles.err
echo "circles rc=$?"; tail -3 gpurun_out/r2_g_circles.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2_g_circles.json") if l.startswith("{")][-1]
print("circles2k", d["value"], d["ms_per_step"], d["stages_ms"], "e2e", d["e2e"]["ms_per_frame"], d["parity"], d["roofline"]["frac"], d["roofline"]["kernel"])
PY
python - <<'PY'
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, ".")
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L=_lib.lib(); _lib.check(L.rtb_init(1,None),"init")
s=R.circles_scene(); h=s.upload()
d=torch.zeros((1440,2560,4),dtype=torch.float32,device="cuda")
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
st=torch.cuda.Stream(); torch.cuda.set_stream(st)
for name,fl in (("wavefront EXT",0),("k_trace_ext",_lib.RTB_FLAG_MEGAKERNEL)):
    v=R.main_viewport(2560,1440,2,1); v.seed=7; v.flags=fl
    ms=[]
    for it in range(8):
        flush.fill_(it); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(st); _lib.check(L.rtb_render_device(h,C.byref(v),0,0,1,d.data_ptr(),None,None,C.c_void_p(st.cuda_stream),None),"r"); b.record(st)
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    print(name, "ms/frame", np.mean(ms[3:]))
PY
