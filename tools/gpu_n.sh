#!/bin/bash
# the driver's bench command at N GPUs (view-parallel value + tile-row sharded records for C2 and C4)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 120 --warmup 12 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -3 gpurun_out/bench_n${N}.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n${N}.json") if l.startswith("{")][-1])
t=d["tiles"]
print("N=${N} value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
print("tiles C2:", {k: t[k] for k in ("value","ms_per_frame","single_gpu_frame_ms","speedup_vs_single_gpu_frame","bit_identical_to_single_gpu_frame","num_rendered_share_per_rank")})
if "C4" in t: print("tiles C4:", {k: t["C4"][k] for k in ("value","ms_per_frame","single_gpu_frame_ms","speedup_vs_single_gpu_frame","bit_identical_to_single_gpu_frame","num_rendered_share_per_rank")})
PY
