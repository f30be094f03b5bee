#!/bin/bash
# round 2, last run (8 GPUs): final tree — default workload at N = 2, 4, 8 (frame on GPU 0, parity), progressive8k at N = 2, 4, 8
mkdir -p gpurun_out
run() {  # n workload steps tag port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $5 \
     bench.py --gpus $1 --steps $3 --warmup 3 --workload $2 > gpurun_out/r2_last_$4.json 2> gpurun_out/r2_last_$4.err
  echo "$4 rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_last_$4.err | tail -2
}
run 8 teapot4k 20 n8_teapot 29531
run 4 teapot4k 20 n4_teapot 29532
run 2 teapot4k 20 n2_teapot 29533
run 8 progressive8k 3 n8_prog 29534
run 4 progressive8k 3 n4_prog 29535
run 2 progressive8k 3 n2_prog 29536
timeout 600 python -m pytest tests -q -m gpu -k 'multi_gpu or progressive or ipc or peer' > gpurun_out/r2_last_tests_8gpu.log 2>&1; echo "multi-GPU tests rc=$?"; tail -3 gpurun_out/r2_last_tests_8gpu.log
python - <<PY
import json
for f in ("n2_teapot","n4_teapot","n8_teapot","n2_prog","n4_prog","n8_prog"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2_last_{f}.json") if l.startswith("{")][-1]
        print(f, "value %.0f ms %.4f"%(d["value"], d["ms_per_step"]), "e2e", round(d["e2e"]["ms_per_frame"],3), "floor", d["e2e"].get("d2h_floor_ms"), "rgb8", d["e2e"].get("rgb8",{}).get("ms_per_frame"), "ms_reduce", d["e2e"].get("ms_reduce"))
        print("  parity", {k:v for k,v in d["parity"].items() if k not in ("golden","frame_on","frame_sha","golden_sha")})
    except Exception as e: print(f, "no json", e)
PY
