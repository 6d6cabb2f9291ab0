#!/bin/bash
# ncu --set full over every kernel of ONE forward frame (C2) -> gpurun_out/prof_frame.ncu-rep
set -u
mkdir -p gpurun_out
# frame 0 warms up (21 launches); capture the 21 launches of frame 1
timeout 1200 ncu --set full --clock-control none --import-source on -s 21 -c 21 -f \
    -o gpurun_out/prof_frame python tools/profile_frame.py --frames 2 > gpurun_out/ncu_frame.log 2>&1
tail -3 gpurun_out/ncu_frame.log
