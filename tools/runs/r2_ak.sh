#!/bin/bash
# round 2, run AK: one-select key masking (PTX predicate chain), no software pipelining of the leaf records
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; }
echo "== default"; probe
for v in keyasm nopipe keynopipe; do echo "== $v"; RTB_LIB=$PWD/$V/librtb_$v.so probe; done
echo "== default"; probe
