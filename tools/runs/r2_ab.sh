#!/bin/bash
# round 2, run AB: zero-copy output (pixels stored straight into the pinned host frame) vs pieces + copies
timeout 300 python tools/zero_copy_probe.py 2>&1 | tail -8
