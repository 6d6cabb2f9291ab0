#!/bin/bash
# Round evidence run: tests, smoke, both bench arms, ncu launch list of the bench command, ncu --set full of every
# kernel of a frame (forward + backward) and of the dominant kernel alone, the side measurements quoted in DESIGN.md.
# Results under gpurun_out/ (the ones worth keeping are copied into profiles/ by hand).
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_b200.json
( timeout 600 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tail -1 ) > gpurun_out/bench_reference.json
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_b200.json; echo; cut -c1-300 gpurun_out/bench_reference.json
KREG='regex:preprocess|depth_|row_count|row_scan|range_partition|column_hist|plan_kernel|blend_'
# launch list of the bench command itself (never a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 200 -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches.log 2>&1
# every kernel of one forward + backward frame (frames 0-1 warm up: 14 library kernels each)
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -s 28 -c 14 -f \
    -o gpurun_out/prof_frame python tools/profile_frame.py --frames 3 --backward > gpurun_out/ncu_frame.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_forward -s 2 -c 1 -f \
    -o gpurun_out/prof_blend_fwd python tools/profile_frame.py --frames 4 > gpurun_out/ncu_blend.log 2>&1
tail -1 gpurun_out/ncu_frame.log gpurun_out/ncu_blend.log
# side measurements quoted in DESIGN.md section 5
( timeout 300 python tools/bench_head.py 2>&1 | tail -1 ) > gpurun_out/bench_head.log
( timeout 300 python tools/bench_supersample.py 2>&1 | tail -1 ) > gpurun_out/bench_supersample.log
( timeout 300 python tools/bench_backward.py uniform 2>&1 | tail -1 ) > gpurun_out/bench_backward.log
( timeout 300 python tools/bench_passes.py 2>&1 | tail -1 ) > gpurun_out/bench_passes.log
( timeout 300 python tools/bench_render_call.py 2>&1 | tail -1 ) > gpurun_out/bench_render_call.log
( timeout 300 python tools/frontend_cost.py C2 6 2>&1 | tail -1 ) > gpurun_out/frontend_cost.log
( timeout 300 python tools/blend_only.py 2>&1 | tail -2 ) > gpurun_out/blend_only.log
GSPLAT_B200_LIB=$PWD/gaussian-pcloud-render_b200/libgsplat_b200_tl.so timeout 200 python tools/blend_timeline.py > gpurun_out/blend_timeline.txt 2>&1
( timeout 300 python tools/fuzz_vs_reference.py 2>&1 | tail -3 ) > gpurun_out/fuzz.log
tail -n 2 gpurun_out/bench_*.log gpurun_out/frontend_cost.log gpurun_out/blend_only.log gpurun_out/fuzz.log | cut -c1-400
