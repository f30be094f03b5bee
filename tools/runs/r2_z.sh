#!/bin/bash
# round 2, run Z: primary phase with the partial sort; full parity suite, share probe, default bench
mkdir -p gpurun_out
timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_z_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_z_tests.log
timeout 600 python bench.py > gpurun_out/r2_z_bench.json 2> gpurun_out/r2_z_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2_z_bench.json") if l.startswith("{")][-1]
print("value %.1f ms %.4f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["stages_ms"], d["parity"]["equals_golden"], d["roofline"]["frac"])
PY
