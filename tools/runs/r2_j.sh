#!/bin/bash
# round 2, run J (8 GPUs): scaling of the default workload with the frame assembled on GPU 0, progressive8k, multi-GPU tests
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_j_topo.txt 2>&1
run() {  # n workload steps tag port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $5 \
     bench.py --gpus $1 --steps $3 --warmup 3 --workload $2 > gpurun_out/r2_j_$4.json 2> gpurun_out/r2_j_$4.err
  echo "$4 rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_j_$4.err | tail -3
}
run 8 teapot4k 20 n8_teapot 29521
run 4 teapot4k 20 n4_teapot 29522
run 8 progressive8k 3 n8_prog 29523
timeout 300 python -m pytest tests -x -q -m gpu --timeout 200 -k "multi_gpu or progressive_psnr" > gpurun_out/r2_j_tests.log 2>&1
tail -3 gpurun_out/r2_j_tests.log
python - <<PY
import json
for f in ("n8_teapot","n4_teapot","n8_prog"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2_j_{f}.json") if l.startswith("{")][-1]
        print(f, "value %.0f ms %.4f"%(d["value"], d["ms_per_step"]), "stages", {k:round(v,3) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["ms_per_frame"],3), "floor", d["e2e"].get("d2h_floor_ms"), "rgb8", d["e2e"].get("rgb8",{}).get("ms_per_frame"), "ms_reduce", d["e2e"].get("ms_reduce"))
        print("  parity", {k:v for k,v in d["parity"].items() if k not in ("golden","frame_on")})
    except Exception as e: print(f, "no json", e)
PY
