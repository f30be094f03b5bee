#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_z_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_z_tests.log
