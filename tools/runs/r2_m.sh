#!/bin/bash
# round 2, run M: leaf formation knobs (the leaf phase holds ~47 % of the bounce kernel's stall samples)
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; }
echo "== default"; probe
for s in 25 50 200 400; do echo "== RTB_SAH_LEAVES=$s"; RTB_SAH_LEAVES=$s probe; done
for c in 60 200 300; do echo "== RTB_COLLAPSE_CT=$c"; RTB_COLLAPSE_CT=$c probe; done
for v in 1 2 3 6; do echo "== RTB_LEAF_MAX=$v"; RTB_LIB=$PWD/$V/librtb_leaf$v.so probe; done
echo "== RTB_LEAF_MAX=2 RTB_SAH_LEAVES=50"; RTB_SAH_LEAVES=50 RTB_LIB=$PWD/$V/librtb_leaf2.so probe
echo "== RTB_SPLIT_DIV=32"; RTB_SPLIT_DIV=32 probe
echo "== RTB_SPLIT_DIV=64"; RTB_SPLIT_DIV=64 probe
