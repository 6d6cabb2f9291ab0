#!/bin/bash
# session r02e: GPU tests, smoke, bench (own arm) after the grouped hit loop became the blend kernel of plain frames
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200.json
( GSPLAT_B200_BLEND_PLAIN=1 timeout 600 python bench.py --no-cpu-baseline 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200_plainloop.json
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_b200.json; echo; cat gpurun_out/bench_b200_plainloop.json
