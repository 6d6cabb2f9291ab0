#!/bin/bash
# round 2, session a: parity of the tail-mode blend + A/B of variants + per-unit timeline
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_variants.sh base notail q16 q32 q128 q64pg8 2>&1 | tee gpurun_out/variants.log
GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_tl.so timeout 300 python tools/blend_timeline.py > gpurun_out/blend_timeline.txt 2>&1
cat gpurun_out/blend_timeline.txt
