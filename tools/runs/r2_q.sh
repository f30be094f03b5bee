#!/bin/bash
# round 2, run Q: IPC two-process test; leaf-phase load experiments (eager edge records, L1::evict_last hot halves)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "ipc or golden or 640" > gpurun_out/r2_q_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_q_tests.log; tail -4 gpurun_out/r2_q_tests.log
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; }
echo "== default"; probe
for v in eager hint eagerhint; do echo "== $v"; RTB_LIB=$PWD/$V/librtb_$v.so probe; done
