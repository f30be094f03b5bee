#!/bin/bash
# round 2, run AG: adaptive rounds — one node visit or one leaf pass per round, whichever more lanes wait for (RTB_WF_ADAPT = weight of a leaf lane x 4)
export RTB_LIB=$PWD/rust_raytrace_b200/csrc/build/variants/librtb_adapt.so
for a in 0 2 4 6 8 12; do echo "== RTB_WF_ADAPT=$a"; RTB_WF_ADAPT=$a timeout 300 python tools/share_probe.py 1 2>&1 | tail -1; done
echo "== primary too (RTB_WF_ADAPT_P=4, ADAPT=4)"; RTB_WF_ADAPT_P=4 RTB_WF_ADAPT=4 timeout 300 python tools/share_probe.py 1 2>&1 | tail -1
