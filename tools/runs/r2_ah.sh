#!/bin/bash
# round 2, run AH: per-frame constants moved to the host (pixel deltas, first RNG mixing step, magic-number divisions), 1-spp paths seeded at their first hit
mkdir -p gpurun_out
timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_ah_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_ah_tests.log
