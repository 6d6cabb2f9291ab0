#!/bin/bash
# round 2, session c: parity after the preprocess-backward re-derivation + free identity sort pass; stage times; bench
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "single frame C2: $(timeout 120 python tools/single_frame.py C2 2>&1 | tail -1)"
echo "single frame C4: $(timeout 200 python tools/single_frame.py C4 2>&1 | tail -1)"
timeout 600 python bench.py --steps 600 --warmup 30 > gpurun_out/bench_b200.json 2> gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_b200.json"))
print("value %.1f single_frame_ms %.3f e2e %.1f padded %.1f serial %.1f stage %s" % (d["value"], d["single_frame_ms"], d["e2e"]["value"], d["e2e"]["padded_layout"]["value"], d["dropin_serial_fps"], {k: round(x,4) for k,x in d["roofline"]["stage_ms"].items()}))
x=d["extra_workloads"]; print("C3 fwd+bwd %.3f ms (bwd kernel %.3f, pre-bwd %.3f)  C4 %.1f fps single %.3f ms" % (x["C3"]["fwd_bwd_ms"], x["C3"]["roofline"]["stage_ms"]["blend_backward"], x["C3"]["roofline"]["stage_ms"]["preprocess_backward"], x["C4"]["value"], x["C4"]["single_frame_ms"]))
PY
tail -3 gpurun_out/bench.err
