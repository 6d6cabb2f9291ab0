#!/bin/bash
# quick perf check: bench (b200 arm) + ncu launch list of a few frames
set -u
mkdir -p gpurun_out
( timeout 600 python bench.py --no-cpu-baseline ${BENCH_ARGS:-} 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 100 --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 10 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/bench_b200.json; tail -2 gpurun_out/bench.err
