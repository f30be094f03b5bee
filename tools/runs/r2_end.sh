#!/bin/bash
# round 2, last check of the final tree, the way the driver runs it: GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_end_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_end_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_end_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_end_smoke.log
timeout 600 python bench.py > gpurun_out/r2_end_bench.json 2> gpurun_out/r2_end_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2_end_bench.json") if l.startswith("{")][-1]
print("value %.1f ms %.4f e2e %.1f (%.3f ms)"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_frame"]), d["parity"]["equals_golden"], d["roofline"]["frac"], d["roofline"]["traffic_source"], d["gpu_launches"], d["clocks"])
PY
