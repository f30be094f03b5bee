#!/bin/bash
# round 2, run S: randomised differential tests + full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 --durations=5 > gpurun_out/r2_s_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_s_tests.log; tail -10 gpurun_out/r2_s_tests.log
