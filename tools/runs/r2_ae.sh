#!/bin/bash
# round 2, run AE: bounded shared-memory stack (more L1) on a 1/8 share; CTAs per SM with it
for s in 0 4 8 12; do echo "== RTB_WF_STACK=$s"; RTB_WF_STACK=$s timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2; done
