#!/bin/bash
# compare library variants: pipelined throughput + blend kernel time
set -u
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200${v:+_$v}.so
  [ "$v" = "base" ] && lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200.so
  ( GSPLAT_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --steps 600 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_v_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v_$v.json"))
print("variant $v value %.1f blend_ms %.3f stage %s" % (d["value"], d["roofline"]["kernel_ms"], {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}))
PY
done
