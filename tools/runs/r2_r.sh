#!/bin/bash
# round 2, run R (2 GPUs): bench.py's fallback for boxes without peer access (forced), and the normal path again
mkdir -p gpurun_out
for mode in ipc noipc; do
  [ $mode = noipc ] && export RTB_BENCH_NO_IPC=1 || unset RTB_BENCH_NO_IPC
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((RANDOM%9)) \
     bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_r_$mode.json 2> gpurun_out/r2_r_$mode.err
  echo "$mode rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_r_$mode.err | tail -2
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2_r_$mode.json") if l.startswith("{")][-1]
print("$mode", d["value"], d["ms_per_step"], d["parity"])
PY
done
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 -k "multi_gpu or ipc or progressive_psnr" 2>&1 | tail -2
