#!/bin/bash
# round 2, run C: padded counters + deferred primary_done + stream-ordered workspace clear; ncu fused vs split
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "golden or split or repeatable or 640 or multisample or megakernel or 4k or rgb8" > gpurun_out/r2_c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c_tests.log
tail -5 gpurun_out/r2_c_tests.log
V=rust_raytrace_b200/csrc/build/variants
for cfg in "fused:0:" "split:64:" "fused_b7:0:$V/librtb_b7.so"; do
  IFS=: read name flags lib <<< "$cfg"
  echo "== $name"
  if [ -n "$lib" ]; then export RTB_LIB=$PWD/$lib; else unset RTB_LIB; fi
  FLAGS=$flags timeout 300 python tools/share_probe.py 1 2 4 8 2>&1 | tail -4
done > gpurun_out/r2_c_share.log 2>&1
unset RTB_LIB
cat gpurun_out/r2_c_share.log
FLAGS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wf_path -s 4 -c 1 -f -o gpurun_out/r2_c_fused \
   python tools/share_probe.py 1 > gpurun_out/r2_c_ncu_fused.log 2>&1
FLAGS=64 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wf_path -s 8 -c 2 -f -o gpurun_out/r2_c_split \
   python tools/share_probe.py 1 > gpurun_out/r2_c_ncu_split.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
