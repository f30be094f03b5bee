#!/bin/bash
# round 2, run U: BVH8 vs BVH4 on the 986 K-triangle field (nodes and references no longer L1/L2 resident)
mkdir -p gpurun_out
for b in 4 8; do
  RTB_BVH8=1 RTB_BVH=$b timeout 600 python bench.py --workload field1m --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_u_field_bvh$b.json 2> gpurun_out/r2_u_field_bvh$b.err
  echo "bvh$b rc=$?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2_u_field_bvh$b.json") if l.startswith("{")][-1]
print("bvh$b", "value %.1f ms %.4f"%(d["value"], d["ms_per_step"]), d["stages_ms"], "tests/ray", d["details"]["node_tests_per_ray"], d["details"]["tri_tests_per_ray"], "build ms", d["details"]["bvh"]["ms_build"], d["parity"]["equals_golden"])
PY
done
