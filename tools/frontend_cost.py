"""Developer tool (GPU box): what the stages of a frame cost with frames in flight -- frames/s of the pipeline cut after
preprocess / depth sort / tile lists (developer switch GSPLAT_B200_STOP_AFTER) and of the whole frame.
usage: frontend_cost.py [workload] [depth]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from renderer import FramePipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload(name)
pipe = FramePipeline(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, depth=depth, capacity=40_000_000 if name == "C4" else 24_000_000)
vd = [pipe.upload_view(v) for v in views]


def run(n):
    pipe.begin()
    for i in range(n):
        pipe.enqueue(vd[(i * 7) % len(vd)])
    pipe.end()


out = {}
for stop in (0, 1, 2, 3, 0):
    os.environ["GSPLAT_B200_STOP_AFTER"] = str(stop)
    run(24)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 480
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[["full", "preprocess", "+depth sort", "+tile lists"][stop]] = {"us_per_frame": round(ms * 1e3, 1), "frames_per_s": round(1e3 / ms, 1)}
print(json.dumps(out))
