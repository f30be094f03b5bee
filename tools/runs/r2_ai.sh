#!/bin/bash
# round 2, run AI: service thresholds re-swept with the final kernel
for p in 16 20 24 28; do echo "== RTB_WF_REFILL_P=$p"; RTB_WF_REFILL_P=$p timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; done
for b in 16 18 22 24; do echo "== RTB_WF_REFILL=$b"; RTB_WF_REFILL=$b RTB_WF_REFILL_TAIL=$b timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; done
