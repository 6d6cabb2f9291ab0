#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, ncu launch list + one full capture of the blend kernel.
# Everything that should come back is written under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200.json
( timeout 600 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tail -1 ) > gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 120 --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 8 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_forward -s 2 -c 2 -f \
    -o gpurun_out/prof_blend_fwd python tools/profile_frame.py --frames 4 > gpurun_out/ncu_blend.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_b200.json; cat gpurun_out/bench_reference.json
