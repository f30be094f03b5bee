#!/bin/bash
# round 2, run AN: rtb_render piece count re-swept with the faster kernels (f32 frame home and rgb8)
for cfg in "X=1" "RTB_PIECES=4" "RTB_PIECES=5" "RTB_PIECES=6" "RTB_PIECES=10" "RTB_PIECES=12"; do
  echo "== $cfg"; env $cfg timeout 120 python tools/e2e_probe.py 2>&1 | grep "rtb_render\|floor"
done
