#!/bin/bash
# round 2, run AL: far-camera parity test (BVH4 centre / half-extent boxes at the camera range limit)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "far_camera or rotated or bruteforce" > gpurun_out/r2_al_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_al_tests.log
