#!/bin/bash
# round 2, last measurements on one GPU: full GPU suite, smoke, every workload, the reference arm, launch list + ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/r2_last_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_last_tests.log; tail -4 gpurun_out/r2_last_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_last_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_last_smoke.log
timeout 300 python tools/share_probe.py 1 2 4 8 2>&1 | tail -4 | tee gpurun_out/r2_last_share_probe.txt
bash tools/profile_gpu.sh r2_last
for w in field1m circles2k progressive8k; do
  steps=10; [ $w = progressive8k ] && steps=3
  timeout 900 python bench.py --workload $w --steps $steps --warmup 3 > gpurun_out/r2_last_$w.json 2> gpurun_out/r2_last_$w.err; echo "$w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_last_reference.json 2> gpurun_out/r2_last_reference.err; echo "reference rc=$?"
python - <<PY
import json
for f in ("bench","field1m","circles2k","progressive8k","reference"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r2_last_{f}.json") if l.startswith("{")][-1]
        e=d.get("e2e") or {}
        print(f, "value %.1f ms %.4f"%(d["value"], d["ms_per_step"]), "e2e", e.get("ms_per_frame"), "floor", e.get("d2h_floor_ms"), "rgb8", (e.get("rgb8") or {}).get("ms_per_frame"), "parity", {k:v for k,v in (d.get("parity") or {}).items() if k.startswith("equals") or k.startswith("e2e")}, "roof", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("one_thread") or {}).get("value"), d.get("stages_ms"))
    except Exception as ex: print(f, "no json", ex)
PY
