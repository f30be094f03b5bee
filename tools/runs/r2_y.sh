#!/bin/bash
# round 2, run Y: sensitivity of the path kernel to L1 wavefronts and to issue slots: +1 / +2 extra one-word loads of another node's line
# per node visit (+25 / +50 % node wavefronts), +16 / +32 extra FFMA per node visit (+12 / +24 % node-visit instructions)
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; }
echo "== default"; probe
for v in diag_ldg1 diag_ldg2 diag_alu16 diag_alu32; do echo "== $v"; RTB_LIB=$PWD/$V/librtb_$v.so probe; done
