#!/bin/bash
# team A/B: parity, per-unit timeline, then single-frame / stage times for several hand-over thresholds and variants
set -u
mkdir -p gpurun_out
GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_tl.so timeout 90 python tools/blend_timeline.py > gpurun_out/blend_timeline.txt 2>&1 || { echo "TIMELINE HUNG/FAILED"; tail -5 gpurun_out/blend_timeline.txt; exit 1; }
cat gpurun_out/blend_timeline.txt
( timeout 400 python -m pytest tests -m gpu -x -q --timeout 100 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in $VARIANTS; do
for t in "$@"; do
  lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_$v.so
  [ "$v" = "base" ] && lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200.so
  ( GSPLAT_B200_LIB=$lib GSPLAT_B200_TEAM_AFTER=$t timeout 90 python tools/single_frame.py 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/sf_${v}_$t.json
  echo "variant $v team_after $t: $(cat gpurun_out/sf_${v}_$t.json)"
done
done
tail -3 gpurun_out/bench.err
