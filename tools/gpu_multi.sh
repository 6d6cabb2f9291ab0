#!/bin/bash
# multi-GPU check: view-parallel and tile-sharded bench at N = $1
set -u
N=${1:-2}
mkdir -p gpurun_out
for mode in views tiles; do
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 120 --warmup 10 --parallel $mode --no-cpu-baseline 2> gpurun_out/bench_n${N}_$mode.err | tail -1 ) > gpurun_out/bench_n${N}_$mode.json
  tail -2 gpurun_out/bench_n${N}_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$mode.json"))
    print("N=$N $mode value %.1f e2e %.1f resident %.1f scaling %s" % (d["value"], d["e2e"]["value"], d["e2e"]["resident_cloud"]["value"], d["scaling"]))
except Exception as e: print("failed", e)
PY
done
