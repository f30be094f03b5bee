#!/bin/bash
# round 2, run A: first contact of the fused path kernel with the GPU (parity subset, then fused / split / in-place timings)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "golden or split or repeatable or 640 or ragged or maxdepth or multisample or pool or megakernel or brute or tiny or cull" \
  > gpurun_out/r2_a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_a_tests.log
tail -5 gpurun_out/r2_a_tests.log
for mode in fused split inplace; do
  case $mode in
    fused) envs="" ;;
    split) envs="RTB_WF_SPLIT=1" ;;
    inplace) envs="RTB_WF_INPLACE=1" ;;
  esac
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_a_bench_$mode.json 2> gpurun_out/r2_a_bench_$mode.err
  echo "$mode rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_a_bench_$mode.json"))
    print("$mode", d["value"], d["ms_per_step"], d["stages_ms"], d["e2e"]["ms_per_frame"])
except Exception as e:
    print("$mode: no json", e)
PY
done
