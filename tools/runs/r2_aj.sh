#!/bin/bash
# round 2, run AJ: BVH4 child boxes as (centre, half extent) — slab planes by FFMA instead of per-axis FMNMX — vs (lo, hi)
mkdir -p gpurun_out
V=rust_raytrace_b200/csrc/build/variants
echo "== (centre, half extent)"; timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
echo "== (lo, hi)"; RTB_LIB=$PWD/$V/librtb_lohi.so timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_aj_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_aj_tests.log
for w in field1m; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]; print('$w', d['value'], d['ms_per_step'], d['stages_ms'], d['parity']['equals_golden'], d['details']['node_tests_per_ray'], d['details']['tri_tests_per_ray'])"; done
