#!/bin/bash
# round 2, session b: new full-size parity tests + the restructured bench line (both arms)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 120 --warmup 12 > gpurun_out/bench_b200.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench_b200.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_reference.json; tail -5 gpurun_out/bench_ref.err
