#!/bin/bash
# round 2, run W: per-warp timeline of the bounce launch (start, queue exhausted, exit) on a full frame and a 1/8 share
mkdir -p gpurun_out
for w in 1 8; do
  RTB_LIB=$PWD/rust_raytrace_b200/csrc/build/variants/librtb_timeline.so timeout 300 python tools/timeline_probe.py $w > gpurun_out/r2_w_timeline_w$w.txt 2>&1
  echo "w$w rc=$? lines $(wc -l < gpurun_out/r2_w_timeline_w$w.txt)"; grep FRAME gpurun_out/r2_w_timeline_w$w.txt
done
gzip -f gpurun_out/r2_w_timeline_w*.txt
