"""Developer tool (GPU box): config C3 -- forward + backward (gradients to means/scales/rotations/opacity/SH) at
800K points, 1920x1080, this library vs the unmodified reference CUDA library.  Writes gpurun_out/bench_backward.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import bench  # noqa: E402
import scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C  # noqa: E402
from oracle.oracle import ReferenceCUDA  # noqa: E402

dev = torch.device("cuda:0")
opacity = sys.argv[1] if len(sys.argv) > 1 else "ones"
cloud = scenes.human_cloud(799957, scale_factor=448.0, seed=0, opacity=opacity)
views = [scenes.make_view(c, 1920, 1080) for c in scenes.orbit_c2w(120)]
W, H = 1920, 1080
d = {k: cloud[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
wgt = torch.from_numpy(np.random.default_rng(7).standard_normal((3, H, W)).astype(np.float32)).to(dev)
bg = torch.ones(3, device=dev)
t = lambda a: torch.from_numpy(a).to(dev)
vd = [(t(v.viewmatrix), t(v.projmatrix), t(v.campos)) for v in views]
L = _C.lib()


bwd_events = None


def ours(i, backward=True):
    v = views[i % 120]
    vm, pm, cp = vd[i % 120]
    rs = GaussianRasterizationSettings(H, W, v.tanfovx, v.tanfovy, bg, 1.0, vm, pm, 1, cp, False, False)
    m2 = torch.zeros_like(d["means3D"], requires_grad=True)
    color, _ = GaussianRasterizer(rs)(d["means3D"], m2, d["opacities"], shs=d["shs"], scales=d["scales"],
                                      rotations=d["rotations"])
    if backward:
        for x in d.values():
            x.grad = None
        if bwd_events is not None:
            bwd_events[0].record()
        color.backward(wgt)
        if bwd_events is not None:
            bwd_events[1].record()


ref = ReferenceCUDA() if ReferenceCUDA.available() else None


def theirs(i, backward=True):
    v = views[i % 120]
    vm, pm, cp = vd[i % 120]
    ref.forward(means3D=d["means3D"].detach(), opacities=d["opacities"].detach(), W=W, H=H, viewmatrix=vm,
                projmatrix=pm, campos=cp, bg=bg, tanfovx=v.tanfovx, tanfovy=v.tanfovy, sh_degree=1,
                shs=d["shs"].detach(), scales=d["scales"].detach(), rotations=d["rotations"].detach())
    if backward:
        ref.backward(wgt)


def timeit(fn, n=40, **kw):
    """Median over n views of the per-call device time (the mean is dominated by caching-allocator growth spikes)."""
    for i in range(10):
        fn(i * 7, **kw)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(5 + i * 7, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def time_backward_only(n=40):
    global bwd_events
    ts = []
    for i in range(n):
        bwd_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ours(5 + i * 7)
        torch.cuda.synchronize()
        ts.append(bwd_events[0].elapsed_time(bwd_events[1]))
    bwd_events = None
    return float(np.median(ts))


out = {"workload": f"C3: 800K pts, 1920x1080, forward+backward, opacity={opacity}"}
out["b200_fwd_ms"] = timeit(ours, backward=False)
out["b200_fwd_bwd_ms"] = timeit(ours)
out["b200_bwd_only_ms"] = time_backward_only()
if ref is not None:
    out["reference_fwd_ms"] = timeit(theirs, backward=False)
    out["reference_fwd_bwd_ms"] = timeit(theirs)
    # gradient parity at full size on one view
    ours(3)
    g_ours = {k: x.grad.clone() for k, x in d.items()}
    theirs(3)
    g = ref.backward(wgt)
    names = {"means3D": "dL_dmeans3D", "opacities": "dL_dopacity", "scales": "dL_dscales", "rotations": "dL_drotations",
             "shs": "dL_dsh"}
    out["grad_rel_err"] = {k: float((g_ours[k] - g[n].reshape(g_ours[k].shape)).abs().max() /
                                    (g[n].abs().max() + 1e-30)) for k, n in names.items()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"bench_backward_{opacity}.json"), "w"), indent=1)
print(json.dumps(out))
