#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'depth_pass|range_partition|plan_kernel|column_hist' -s 8 -c 8 -f \
    -o gpurun_out/prof_bin python tools/profile_frame.py --frames 2 > gpurun_out/ncu_bin.log 2>&1
tail -3 gpurun_out/ncu_bin.log
