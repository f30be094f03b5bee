#!/bin/bash
# round 2, run I: resident CTAs per SM vs band share (tail experiment)
for c in 8 6 5 4 3; do
  echo "== RTB_WF_CTAS=$c"
  RTB_WF_CTAS=$c timeout 300 python tools/share_probe.py 1 4 8 2>&1 | tail -3
done
