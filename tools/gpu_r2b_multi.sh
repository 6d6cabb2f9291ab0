#!/bin/bash
# round 2, session b, N GPUs: peer stores + sharded backward + e2e fan-out checks, then the driver's bench command
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/peer_check.py > gpurun_out/peer_check_n$N.log 2>&1
grep -v "^W\|^\[W" gpurun_out/peer_check_n$N.log | tail -12
timeout 300 python -m pytest tests/test_gpu.py -q -k "two_devices or peer_store" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 120 --warmup 12 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 2500 gpurun_out/bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err
