"""Developer tool (GPU box): the reference caller's four raster passes per view (position / RGB / hit-map / normal,
simple_raw_render.py:411-522) at C2: four independent frames vs one frame + three colour passes that reuse the
geometry (gs_forward_recolor).  Writes gpurun_out/bench_passes.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from renderer import FrameRenderer  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C2")
W, H, P = w["W"], w["H"], cloud["means3D"].shape[0]
fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=24_000_000)
vd = [fr.upload_view(v) for v in views]
normals = torch.randn(P, 3, device=dev)
ones = torch.ones(P, 3, device=dev)
xyz = fr.means3D
outs = [torch.empty((3, H, W), device=dev) for _ in range(4)]


def fused(i):
    v = vd[i % len(vd)]
    fr.enqueue(v, out_color=outs[1])                      # RGB pass (SH) = the full frame
    fr.enqueue_pass(v, outs[0], colors_precomp=xyz)       # position pass
    fr.enqueue_pass(v, outs[2], colors_precomp=ones)      # hit map
    fr.enqueue_pass(v, outs[3], colors_precomp=normals)   # normals


clouds = [dict(cloud), dict(cloud), dict(cloud), dict(cloud)]
rs = [fr] + [FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=24_000_000, share=fr) for _ in range(0)]


def independent(i):
    v = vd[i % len(vd)]
    for k in range(4):  # four full frames (same kernels as four calls of the drop-in module, without its host sync)
        fr.enqueue(v, out_color=outs[k])


def one_walk(i):
    v = vd[i % len(vd)]
    fr.enqueue(v, out_color=outs[1], extra_passes=[(xyz, outs[0]), (ones, outs[2]), (normals, outs[3])])


def timeit(fn, n=60):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(5 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"workload": "C2, four raster passes per view (one stream)", "independent_ms_per_view": timeit(independent),
       "recolor_ms_per_view": timeit(fused), "one_walk_ms_per_view": timeit(one_walk)}
out["speedup_recolor"] = out["independent_ms_per_view"] / out["recolor_ms_per_view"]
out["speedup_one_walk"] = out["independent_ms_per_view"] / out["one_walk_ms_per_view"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_passes.json"), "w"), indent=1)
print(json.dumps(out))
