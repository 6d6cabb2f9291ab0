#!/bin/bash
# blend A/B: single-frame stage times + in-flight throughput (bench, short) and the per-unit timeline
set -u
mkdir -p gpurun_out
( timeout 300 python tools/single_frame.py C2 ) 2> gpurun_out/ab.err | tail -1
( timeout 300 python tools/frontend_cost.py C2 6 ) 2>> gpurun_out/ab.err | tail -1
( timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "golden or fuzz or full_size or reference" 2>&1 | tail -3 )
GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_tl.so timeout 200 python tools/blend_timeline.py | grep -v "^  start" > gpurun_out/blend_timeline_grouped.txt 2>> gpurun_out/ab.err
cat gpurun_out/blend_timeline_grouped.txt
