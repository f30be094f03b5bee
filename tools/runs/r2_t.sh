#!/bin/bash
# round 2, run T: 9 / 10 CTAs per SM (56 / 48 registers, compiler spills)
V=rust_raytrace_b200/csrc/build/variants
probe() { timeout 300 python tools/share_probe.py 1 8 2>&1 | tail -2; }
echo "== default (8 CTAs, 64 regs)"; probe
for v in b9 b10; do echo "== $v"; RTB_LIB=$PWD/$V/librtb_$v.so probe; done
