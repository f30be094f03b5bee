#!/bin/bash
# round 2, run L: how much is coherence in the bounce queue worth?  (sorted consumption order, sort cost not counted: stage "shade")
for m in 0 1 2 3; do
  echo "== RTB_WF_SORTQ=$m"
  RTB_WF_SORTQ=$m timeout 300 python tools/share_probe.py 1 2>&1 | tail -1
done
