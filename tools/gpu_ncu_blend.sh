#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_forward -s 2 -c 1 -f \
    -o gpurun_out/prof_blend_fwd python tools/profile_frame.py --frames 4 > gpurun_out/ncu_blend.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_backward -s 1 -c 1 -f \
    -o gpurun_out/prof_blend_bwd python tools/profile_frame.py --frames 3 --backward > gpurun_out/ncu_blend_bwd.log 2>&1
tail -2 gpurun_out/ncu_blend.log gpurun_out/ncu_blend_bwd.log
