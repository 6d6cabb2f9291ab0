"""Developer tool (GPU box): pipelined throughput of blend-only frames (gs_forward_recolor on 4 lanes) vs full frames,
to see how much of the frames-in-flight throughput the front end (preprocess, sort, tile lists) costs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from renderer import FramePipeline  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C2")
pipe = FramePipeline(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, depth=4, capacity=24_000_000)
vd = [pipe.upload_view(v) for v in views]
cols = torch.rand(cloud["means3D"].shape[0], 3, device=dev)
outs = [torch.empty((3, w["H"], w["W"]), device=dev) for _ in pipe.lanes]


def run(n, blend_only):
    pipe.begin()
    for i in range(n):
        k = i % pipe.depth
        with torch.cuda.stream(pipe.streams[k]):
            if blend_only:
                pipe.lanes[k].enqueue_pass(vd[k], outs[k], colors_precomp=cols)
            else:
                pipe.lanes[k].enqueue(vd[(i * 7) % len(vd)])
    pipe.end()


for k in range(pipe.depth):  # each lane holds a binned frame of view k
    with torch.cuda.stream(pipe.streams[k]):
        pipe.lanes[k].enqueue(vd[k])
torch.cuda.synchronize()
for mode in (False, True):
    run(20, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(240, mode)
    e1.record()
    torch.cuda.synchronize()
    print("blend-only (recolor + blend)" if mode else "full frames", "%.1f frames/s" % (240 / (e0.elapsed_time(e1) / 1e3)))
    if not mode:
        for k in range(pipe.depth):
            with torch.cuda.stream(pipe.streams[k]):
                pipe.lanes[k].enqueue(vd[k])
        torch.cuda.synchronize()
