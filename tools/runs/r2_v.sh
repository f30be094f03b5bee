#!/bin/bash
# round 2, run V: tail launches of the bounce phase (RTB_WF_TAIL = 0 / 1 / 2 / 3) on a full frame and a 1/8 band share
mkdir -p gpurun_out
for t in 0 1 2 3; do
  echo "== RTB_WF_TAIL=$t"
  RTB_WF_TAIL=$t timeout 300 python tools/share_probe.py 1 2 4 8 2>&1 | tail -4
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_v_tests.log
