#!/bin/bash
# round 2, run X: partial child sort (3 compare-exchanges), 96-byte triangle records read with 256-bit loads
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2; }
echo "== default"; probe
for v in sort3 tri96 tri96sort3; do echo "== $v"; RTB_LIB=$PWD/$V/librtb_$v.so probe; done
echo "== default again"; probe
