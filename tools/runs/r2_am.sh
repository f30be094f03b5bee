#!/bin/bash
# round 2, run AM: two frames in flight (next frame's primary launch fills the SM slots the bounce launch's tail frees)
timeout 300 python tools/pipeline_probe.py 1 2 4 8 2>&1 | tail -4
