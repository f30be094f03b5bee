#!/bin/bash
# round 2, run AA: node visits per round 1 / 2 / 3 (default) / 4 with the final kernel
for d in 1 2 3 4; do echo "== RTB_WF_DESCEND=$d"; RTB_WF_DESCEND=$d timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; done
