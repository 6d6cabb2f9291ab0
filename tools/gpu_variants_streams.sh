#!/bin/bash
set -u
mkdir -p gpurun_out
v=$1; shift
lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_$v.so
[ "$v" = "base" ] && lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200.so
for s in "$@"; do
  ( GSPLAT_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --steps 240 --streams $s 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_vs.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_vs.json"))
print("variant $v streams $s value %.1f" % d["value"])
PY
done
