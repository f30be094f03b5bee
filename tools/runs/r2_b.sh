#!/bin/bash
# round 2, run B: self-validating 64-byte queue entries (no flags, no fences): parity subset, fused vs split per band share
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "golden or split or repeatable or 640 or ragged or maxdepth or multisample or megakernel or brute or tiny or cull or 4k" \
  > gpurun_out/r2_b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_b_tests.log
tail -5 gpurun_out/r2_b_tests.log
V=rust_raytrace_b200/csrc/build/variants
for cfg in "fused:0:" "split:64:" "fused_b7:0:$V/librtb_b7.so" "fused_b6:0:$V/librtb_b6.so" "split_b7:64:$V/librtb_b7.so"; do
  IFS=: read name flags lib <<< "$cfg"
  echo "== $name"
  if [ -n "$lib" ]; then export RTB_LIB=$PWD/$lib; else unset RTB_LIB; fi
  FLAGS=$flags timeout 300 python tools/share_probe.py 1 2 4 8 2>&1 | tail -4
done > gpurun_out/r2_b_share.log 2>&1
cat gpurun_out/r2_b_share.log
