"""Developer tool (GPU box): frames in flight with and without the per-lane CUDA graph (FrameRenderer.capture_graph),
C1 (small frames: host / launch bound) and C2.  Writes gpurun_out/bench_graph.json."""
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
out = {}
for name, cap, n in (("C1", 12_000_000, 3000), ("C2", 24_000_000, 1200)):
    cloud, views, w = bench.make_workload(name)
    vb = ViewBatch(scenes.orbit_c2w(len(views)), 45.0, dev)
    vd = [vb[k] for k in range(len(vb))]
    res = {}
    for depth in (6,):
        pipe = FramePipeline(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, depth=depth, capacity=cap)
        pipe.capture_graphs((vb.tanfov, vb.tanfov))

        def run(graph, m, off):
            pipe.begin()
            for i in range(m):
                k = (off + i) % len(vd)
                if graph:
                    pipe.enqueue_graph(vb.buf[k])
                else:
                    pipe.enqueue(vd[k], slot=i)
            pipe.end()

        for graph in (False, True, False, True):
            run(graph, 60, 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(graph, n, 60)
            e1.record()
            torch.cuda.synchronize()
            res.setdefault("graph" if graph else "plain", []).append(n / (e0.elapsed_time(e1) / 1e3))
    out[name] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_graph.json"), "w"), indent=1)
print(json.dumps(out))
