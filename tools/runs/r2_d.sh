#!/bin/bash
# round 2, run D: first contact of the BVH8 (compressed 8-wide) traversal
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 \
  -k "golden or bvh8 or 640 or ragged or tiny or cull or brute or multisample or fused_launch or rotated or sphere_scene" > gpurun_out/r2_d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_d_tests.log
tail -15 gpurun_out/r2_d_tests.log
V=rust_raytrace_b200/csrc/build/variants
for cfg in "bvh8:0::" "bvh4:128::" "bvh8_w7:0:$V/librtb_w7.so:" "bvh8_w6:0:$V/librtb_w6.so:" "bvh8_w5:0:$V/librtb_w5.so:" \
           "bvh8_d2:0::RTB_WF_DESCEND8_P=2 RTB_WF_DESCEND8_B=2" "bvh8_d1:0::RTB_WF_DESCEND8_P=1 RTB_WF_DESCEND8_B=1" "bvh8_d5:0::RTB_WF_DESCEND8_P=5 RTB_WF_DESCEND8_B=5"; do
  IFS=: read name flags lib envs <<< "$cfg"
  echo "== $name"
  if [ -n "$lib" ]; then export RTB_LIB=$PWD/$lib; else unset RTB_LIB; fi
  env $envs FLAGS=$flags timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
done > gpurun_out/r2_d_share.log 2>&1
unset RTB_LIB
cat gpurun_out/r2_d_share.log
