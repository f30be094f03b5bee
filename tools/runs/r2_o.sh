#!/bin/bash
# round 2, run O: re-verification of the final tree on one GPU (full suite, smoke, default bench + circles2k)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 > gpurun_out/r2_o_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_o_tests.log; tail -4 gpurun_out/r2_o_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_o_bench.json 2> gpurun_out/r2_o_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload circles2k --steps 20 --warmup 3 > gpurun_out/r2_o_circles2k.json 2> gpurun_out/r2_o_circles2k.err; echo "circles rc=$?"
python - <<PY
import json
for f in ("bench","circles2k"):
    d=[json.loads(l) for l in open(f"gpurun_out/r2_o_{f}.json") if l.startswith("{")][-1]
    e=d["e2e"]
    print(f, "value %.1f ms %.4f"%(d["value"], d["ms_per_step"]), d["stages_ms"], "e2e", e["ms_per_frame"], "floor", e.get("d2h_floor_ms"), "rgb8", e["rgb8"]["ms_per_frame"], "parity", {k:v for k,v in d["parity"].items() if k.startswith("equals") or k.startswith("e2e")}, "roof", d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["traffic"], d["gpu_launches"])
PY
