#!/bin/bash
# grouped hit loop: parity + single-frame / in-flight numbers, then the per-unit timeline
set -u
mkdir -p gpurun_out
timeout 400 python tools/grouped_check.py C2 $MODES > gpurun_out/grouped_check_c2.txt 2> gpurun_out/grouped_check.err || { echo "GROUPED CHECK FAILED"; tail -5 gpurun_out/grouped_check.err; }
cat gpurun_out/grouped_check_c2.txt
GSPLAT_B200_BLEND_LONGSMS=${LONGSMS:-0} GSPLAT_B200_BLEND_HIMODE=${HIMODE:-0} GSPLAT_B200_BLEND_GROUPED=1 GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_tl.so timeout 200 python tools/blend_timeline.py > gpurun_out/blend_timeline_grouped.txt 2>> gpurun_out/grouped_check.err
cat gpurun_out/blend_timeline_grouped.txt
