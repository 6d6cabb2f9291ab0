#!/bin/bash
set -u
mkdir -p gpurun_out
for s in 1 2 3 4 6; do
  ( timeout 300 python bench.py --no-cpu-baseline --streams $s --steps 240 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_s$s.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_s$s.json"))
print("streams $s value %.1f e2e %.1f clocks %s blend_ms %.3f" % (d["value"], d["e2e"]["value"], d["clocks"], d["roofline"]["kernel_ms"]))
PY
done
tail -3 gpurun_out/bench.err
