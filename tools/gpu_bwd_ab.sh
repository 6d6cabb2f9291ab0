#!/bin/bash
# A/B of the backward kernel: main library vs libgsplat_b200_bwd1.so (previous kernel), both opacity modes.
mkdir -p gpurun_out
for op in ones uniform; do
  for lib in "" bwd1; do
    if [ -n "$lib" ]; then export GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_$lib.so; else unset GSPLAT_B200_LIB; fi
    echo "== $op ${lib:-main}"
    timeout 300 python tools/bench_backward.py $op 2>&1 | tail -1 | cut -c1-330
  done
done > gpurun_out/bwd_ab.log 2>&1
cat gpurun_out/bwd_ab.log
