#!/bin/bash
# round 2, run N: hot / cold split of the triangle records
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "golden or 640 or ragged or bvh8 or extension or shadow or circles or megakernel or brute or debug or tiny or field" > gpurun_out/r2_n_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_n_tests.log; tail -5 gpurun_out/r2_n_tests.log
timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
