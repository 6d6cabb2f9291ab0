"""Developer tool (GPU box): GPU cost of the stages of a C2 frame with six frames in flight, WITHOUT the host in the way:
the pipeline cut after preprocess / depth sort / tile lists (developer switch GSPLAT_B200_STOP_AFTER) is captured into
one CUDA graph per lane and replayed, so that a frame costs the host one small copy and one graph launch."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
import scenes  # noqa: E402
from renderer import FramePipeline, ViewBatch  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C2")
vb = ViewBatch(scenes.orbit_c2w(len(views)), 45.0, dev)
out = {}
for stop in (0, 1, 2, 3):
    os.environ["GSPLAT_B200_STOP_AFTER"] = str(stop)
    pipe = FramePipeline(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, depth=6, capacity=24_000_000)
    pipe.capture_graphs((vb.tanfov, vb.tanfov))

    def run(m, off):
        pipe.begin()
        for i in range(m):
            pipe.enqueue_graph(vb.buf[(off + 7 * i) % len(vb)])
        pipe.end()

    run(48, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 960
    e0.record()
    run(n, 5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[["full", "preprocess", "+depth sort", "+tile lists"][stop]] = {"us_per_frame": round(ms * 1e3, 1), "frames_per_s": round(1e3 / ms, 1)}
    del pipe
    torch.cuda.empty_cache()
print(json.dumps(out))
