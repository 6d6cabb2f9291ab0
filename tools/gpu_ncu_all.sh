#!/bin/bash
# ncu --set full over every library kernel of ONE forward + backward frame (C2 shape) -> gpurun_out/prof_frame.ncu-rep,
# and the per-launch duration list of a bench run -> gpurun_out/launches_bench.csv (never a bench value)
set -u
mkdir -p gpurun_out
KREG='regex:preprocess|depth_|row_count|row_scan|range_partition|column_hist|plan_kernel|blend_'
# frames 0-1 warm up (14 library kernels each: 12 forward + 2 backward); capture the 14 of frame 2
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -s 28 -c 14 -f \
    -o gpurun_out/prof_frame python tools/profile_frame.py --frames 3 --backward > gpurun_out/ncu_frame.log 2>&1
tail -2 gpurun_out/ncu_frame.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 200 -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200
