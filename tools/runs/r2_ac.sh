#!/bin/bash
# round 2, run AC: bounce queue filled from both ends by first-hit material (RTB_WF_LPT: 0 queue order, 1 Matte first, 2 Reflective first)
for l in 0 1 2; do echo "== RTB_WF_LPT=$l"; RTB_WF_LPT=$l timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2; done
RTB_WF_LPT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or octree or fused or band" 2>&1 | tail -2
