#!/bin/bash
# Round evidence run: tests, smoke, both bench arms, ncu launch list of the bench command, ncu --set full of the
# dominant kernel.  Results under gpurun_out/ (copied into profiles/ by hand).
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200.json
( timeout 600 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tail -1 ) > gpurun_out/bench_reference.json
# launch list of the bench command itself (first 400 launches after 60 warm-up launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --streams 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_forward -s 2 -c 1 -f \
    -o gpurun_out/prof_blend_fwd python tools/profile_frame.py --frames 4 > gpurun_out/ncu_blend.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench_b200.json; echo; cut -c1-300 gpurun_out/bench_reference.json
# side measurements quoted in DESIGN.md section 5
( timeout 300 python tools/bench_head.py 2>&1 | tail -1 ) > gpurun_out/bench_head.log
( timeout 300 python tools/bench_supersample.py 2>&1 | tail -1 ) > gpurun_out/bench_supersample.log
( timeout 300 python tools/bench_backward.py uniform 2>&1 | tail -1 ) > gpurun_out/bench_backward.log
( timeout 300 python tools/bench_passes.py 2>&1 | tail -1 ) > gpurun_out/bench_passes.log
