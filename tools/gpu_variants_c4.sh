#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_$v.so
  [ "$v" = "base" ] && lib=$PWD/gaussian-pcloud-render_b200/libgsplat_b200.so
  for wl in C2 C4; do
    steps=240; [ $wl = C4 ] && steps=40
    ( GSPLAT_B200_LIB=$lib timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps $steps --warmup 5 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_v_${v}_$wl.json
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_v_${v}_$wl.json"))
print("variant $v $wl value %.1f stage %s" % (d["value"], {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}))
PY
  done
done
