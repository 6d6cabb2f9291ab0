"""Developer tool (GPU box): one `render()` of the reference's caller at C1 shape (simple_raw_render.py:291-545 /
SURVEY 3.1): 221712 Gaussians, 12 orbit views, four raster passes per view (position, RGB, hit map, normal) at
1024x1024 (super-sample rate 2) halved to 512x512.
  reference flow : 48 calls of the unmodified reference kernels + F.interpolate + permute + the per-view torch glue
  this library   : renderer.render_passes -- gs_make_views once, per view one frame with three extra colour passes in
                   the same list walk and the 2x2 mean in the blend epilogue
Writes gpurun_out/bench_render_call.json."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
import scenes  # noqa: E402
from oracle.oracle import ReferenceCUDA  # noqa: E402
from renderer import FramePipeline, FrameRenderer, ViewBatch, render_passes  # noqa: E402

dev = torch.device("cuda:0")
cloud, _, w = bench.make_workload("C1")
W, H, P = w["W"], w["H"], cloud["means3D"].shape[0]
c2w = scenes.orbit_c2w(12)
normals = F.normalize(torch.randn(P, 3, generator=torch.Generator().manual_seed(1)), dim=-1).to(dev)
fr = FrameRenderer(cloud, W, H, [1.0, 1.0, 1.0], dev, capacity=12_000_000, downsample=2)


pipe = FramePipeline(cloud, W, H, [1.0, 1.0, 1.0], dev, depth=4, capacity=12_000_000, downsample=2)


def ours(target=fr):
    vb = ViewBatch(c2w, 45.0, dev)  # includes the upload of the 12 camera matrices
    return render_passes(target, vb, normals=normals)


d = {k: cloud[k].to(dev) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
bg = torch.ones(3, device=dev)
ref = ReferenceCUDA() if ReferenceCUDA.available() else None


def theirs(vb=None):
    """vb: take the camera matrices from a ViewBatch (for the bit-exact comparison) instead of building them per view
    on the host as get_rasterize_param_from_camera does (the timed flavour)."""
    nrm = normals
    outs = {k: [] for k in ("xyz_w", "rgb", "hitmap", "normal")}
    for i, c in enumerate(c2w):
        v = scenes.make_view(c, W, H)
        t = lambda a: torch.from_numpy(a).to(dev)
        vm, pm, cp = (t(v.viewmatrix), t(v.projmatrix), t(v.campos)) if vb is None else vb[i][:3]
        camera_dir = d["means3D"] - cp.reshape(1, 1, 3)
        sgn = (torch.sum(camera_dir * nrm, -1, keepdim=True) > 0).float() * 2 - 1
        nrm = nrm * (-1) * sgn[0]
        for name, kw in (("xyz_w", dict(colors_precomp=d["means3D"])), ("rgb", dict(shs=d["shs"], sh_degree=1)),
                         ("hitmap", dict(colors_precomp=torch.ones_like(d["means3D"]))),
                         ("normal", dict(colors_precomp=nrm))):
            img = ref.forward(means3D=d["means3D"], opacities=d["opacities"], W=W, H=H, viewmatrix=vm, projmatrix=pm,
                              campos=cp, bg=bg, tanfovx=v.tanfovx, tanfovy=v.tanfovy, scales=d["scales"],
                              rotations=d["rotations"], **kw)[0]
            outs[name].append(img)
    return {k: F.interpolate(torch.stack(x, 0), size=(H // 2, W // 2), mode="bilinear",
                             align_corners=False).permute(0, 2, 3, 1) for k, x in outs.items()}


def timeit(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"workload": f"one render(): {P} Gaussians, 12 views x 4 passes, raster {W}x{H} -> {W // 2}x{H // 2}",
       "this_library_ms": timeit(ours, 20), "this_library_4_lanes_ms": timeit(lambda: ours(pipe), 20)}
if ref is not None:
    out["reference_flow_ms"] = timeit(theirs, 3)
    a, b = ours(), theirs(ViewBatch(c2w, 45.0, dev))
    torch.cuda.synchronize()
    out["bit_identical_with_the_same_camera_matrices"] = {k: bool(torch.equal(a[k], b[k])) for k in a}
    out["speedup"] = out["reference_flow_ms"] / out["this_library_ms"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_render_call.json"), "w"), indent=1)
print(json.dumps(out))
