#!/bin/bash
# other BASELINE configs (C1, C4): both arms, short runs
set -u
mkdir -p gpurun_out
for wl in C1 C4; do
  ( timeout 600 python bench.py --workload $wl --steps 60 --warmup 6 --no-cpu-baseline 2> gpurun_out/bench_$wl.err | tail -1 ) > gpurun_out/bench_${wl}_b200.json
  ( timeout 600 python bench.py --workload $wl --steps 30 --warmup 4 --impl reference 2>> gpurun_out/bench_$wl.err | tail -1 ) > gpurun_out/bench_${wl}_reference.json
  tail -2 gpurun_out/bench_$wl.err
  python - <<PY
import json
for arm in ("b200","reference"):
    try:
        d=json.load(open("gpurun_out/bench_${wl}_%s.json" % arm))
        print("$wl", arm, "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d.get("roofline",{}) and d["roofline"]["stage_ms"])
    except Exception as e: print("$wl", arm, "failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_bwd.csv python tools/profile_frame.py --frames 4 --backward > /dev/null 2>&1
