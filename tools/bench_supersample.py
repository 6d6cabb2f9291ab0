"""Developer tool (GPU box): config C1 as the reference's caller runs it (simple_raw_render.py:227-288): 200 K points
rasterised at 1024x1024 (super-sample rate 2) and halved to 512x512 with F.interpolate(bilinear).  Compares the fused
epilogue (GsScene.downsample = 2) with frame + F.interpolate, one stream and six frames in flight, and with the
unmodified reference kernels + F.interpolate.  Writes gpurun_out/bench_supersample.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
from oracle.oracle import ReferenceCUDA  # noqa: E402
from renderer import FramePipeline, FrameRenderer  # noqa: E402

dev = torch.device("cuda:0")
cloud, views, w = bench.make_workload("C1")
W, H = w["W"], w["H"]
fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=12_000_000)
fd = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=12_000_000, share=fr, downsample=2)
vd = [fr.upload_view(v) for v in views]


def separate(i):
    return F.interpolate(fr.enqueue(vd[i % len(vd)])[None], size=(H // 2, W // 2), mode="bilinear", align_corners=False)


def fused(i):
    return fd.enqueue(vd[i % len(vd)])


def timeit(fn, n=200):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(10 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def piped(downsample, n=600, depth=6):
    pipe = FramePipeline(cloud, W, H, [1.0, 1.0, 1.0], dev, depth=depth, capacity=12_000_000, downsample=downsample)

    def go(m, off):
        pipe.begin()
        for i in range(m):
            k, out = pipe.enqueue(vd[(off + i) % len(vd)], slot=i)
            if downsample == 1:
                with torch.cuda.stream(pipe.streams[k]):
                    F.interpolate(out[None], size=(H // 2, W // 2), mode="bilinear", align_corners=False)
        pipe.end()

    go(30, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go(n, 30)
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / 1e3)


assert torch.equal(separate(3)[0], fused(3))
out = {"workload": "C1: 221712 pts, raster 1024x1024 -> 512x512 (super-sample rate 2), forward",
       "frame_plus_interpolate_ms": timeit(separate), "fused_epilogue_ms": timeit(fused),
       "frames_per_s_6_in_flight": {"frame_plus_interpolate": piped(1), "fused_epilogue": piped(2)}}
if ReferenceCUDA.available():
    ref = ReferenceCUDA()
    d = {k: cloud[k].to(dev) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    t = lambda a: torch.from_numpy(a).to(dev)
    rv = [(t(v.viewmatrix), t(v.projmatrix), t(v.campos)) for v in views]
    bg = torch.ones(3, device=dev)

    def theirs(i):
        v = views[i % len(views)]
        vm, pm, cp = rv[i % len(views)]
        f = ref.forward(means3D=d["means3D"], opacities=d["opacities"], W=W, H=H, viewmatrix=vm, projmatrix=pm, campos=cp,
                        bg=bg, tanfovx=v.tanfovx, tanfovy=v.tanfovy, sh_degree=cloud["sh_degree"], shs=d["shs"],
                        scales=d["scales"], rotations=d["rotations"])
        return F.interpolate(f[0][None], size=(H // 2, W // 2), mode="bilinear", align_corners=False)

    out["reference_plus_interpolate_ms"] = timeit(theirs, n=60)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_supersample.json"), "w"), indent=1)
print(json.dumps(out))
