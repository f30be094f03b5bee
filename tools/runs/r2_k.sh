#!/bin/bash
# round 2, run K: service threshold in the tail (queue exhausted)
for t in 20 12 8 4 1; do
  echo "== RTB_WF_REFILL_TAIL=$t"
  RTB_WF_REFILL_TAIL=$t timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2
done
