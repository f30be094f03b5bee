#!/bin/bash
# round 2, run E: full GPU suite with the reworked API (locks, progressive, camera check), bench N=1, BVH8 stats
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 --durations=8 > gpurun_out/r2_e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_e_tests.log
tail -16 gpurun_out/r2_e_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_e_bench.json 2> gpurun_out/r2_e_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_e_bench.err
RTB_BVH8=1 RTB_BVH=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_e_bench_bvh8.json 2> gpurun_out/r2_e_bench_bvh8.err
echo "bench bvh8 rc=$?"; tail -3 gpurun_out/r2_e_bench_bvh8.err
python - <<PY
import json
for f in ("r2_e_bench","r2_e_bench_bvh8"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["stages_ms"], "e2e", d["e2e"]["ms_per_frame"], d["e2e"].get("d2h_floor_ms"), d["e2e"]["rgb8"]["ms_per_frame"])
        print("  parity", d["parity"]); print("  roofline", d["roofline"]["frac"], d["roofline"]["per_ray"], "cpu", d["cpu_baseline"])
    except Exception as e: print(f, "no json", e)
PY
