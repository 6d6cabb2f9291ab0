"""Developer tool (GPU box): renders a few frames of the headline workload (C2: 800K points, 1920x1080) through
gs_forward_nosync, optionally followed by a backward pass, for use under ncu.  Nothing is timed here.

usage: python tools/profile_frame.py [--frames N] [--backward] [--workload C2|C1|C4] [--view K]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))

import bench  # noqa: E402
from renderer import FrameRenderer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--view", type=int, default=7)
    ap.add_argument("--backward", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cloud, views, w = bench.make_workload(a.workload)
    if a.backward:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        d = {k: cloud[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
        for i in range(a.frames):
            v = views[(a.view + i) % len(views)]
            t = lambda x: torch.from_numpy(x).to(dev)
            rs = GaussianRasterizationSettings(w["H"], w["W"], v.tanfovx, v.tanfovy, torch.ones(3, device=dev), 1.0,
                                               t(v.viewmatrix), t(v.projmatrix), cloud["sh_degree"], t(v.campos), False, False)
            m2 = torch.zeros_like(d["means3D"], requires_grad=True)
            color, _ = GaussianRasterizer(rs)(d["means3D"], m2, d["opacities"], shs=d["shs"], scales=d["scales"],
                                              rotations=d["rotations"])
            color.sum().backward()
        torch.cuda.synchronize()
        return
    fr = FrameRenderer(cloud, w["W"], w["H"], [1.0, 1.0, 1.0], dev, capacity=24_000_000 if a.workload != "C4" else 120_000_000)
    for i in range(a.frames):
        fr.render(fr.upload_view(views[(a.view + i) % len(views)]))
    print("num_rendered", fr.status()[0])


if __name__ == "__main__":
    main()
