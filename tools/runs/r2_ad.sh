#!/bin/bash
# round 2, run AD: per-warp timeline of the bounce launch + one tail launch (patch tools/experiments/tail_launches_requeue.patch)
mkdir -p gpurun_out
for w in 1 8; do
  RTB_WF_TAIL=1 RTB_LIB=$PWD/rust_raytrace_b200/csrc/build/variants/librtb_tail_tl.so timeout 300 python tools/timeline_probe.py $w > gpurun_out/r2_ad_timeline_w$w.txt 2>&1
  echo "w$w rc=$?"; grep FRAME gpurun_out/r2_ad_timeline_w$w.txt | tail -2
done
gzip -f gpurun_out/r2_ad_timeline_w*.txt
