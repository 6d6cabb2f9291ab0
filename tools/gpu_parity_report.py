"""Developer tool (GPU box): stage-by-stage comparison of libgsplat_b200 against (a) the UNMODIFIED reference CUDA
library (oracle/_ref/libgs_ref.so) and (b) the CPU oracle, plus rough timings.  Writes gpurun_out/parity_report.json.
Not part of the product; imports oracle/ as a checker only.

usage: python tools/gpu_parity_report.py [--big] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))

import scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C  # noqa: E402
from oracle.oracle import Oracle, ReferenceCUDA  # noqa: E402

dev = torch.device("cuda:0")


def settings_from_view(v, bg, sh_degree, debug=False):
    return GaussianRasterizationSettings(
        image_height=v.image_height, image_width=v.image_width, tanfovx=v.tanfovx, tanfovy=v.tanfovy,
        bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=1.0,
        viewmatrix=torch.from_numpy(v.viewmatrix).to(dev)[None], projmatrix=torch.from_numpy(v.projmatrix).to(dev)[None],
        sh_degree=sh_degree, campos=torch.from_numpy(v.campos).to(dev)[None, None], prefiltered=False, debug=debug)


def run_case(name, cloud, view, bg, use_cpu=True, backward=True, timing_iters=0):
    rep = {"case": name, "P": int(cloud["means3D"].shape[0]), "W": view.image_width, "H": view.image_height}
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in cloud.items()}
    rs = settings_from_view(view, bg, d["sh_degree"])
    means2D = torch.zeros_like(d["means3D"], requires_grad=True)
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    rast = GaussianRasterizer(rs)
    color, radii = rast(leaves["means3D"], means2D, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                        rotations=leaves["rotations"])
    torch.cuda.synchronize()
    rep["visible"] = int((radii > 0).sum())
    gen = torch.Generator(device="cpu").manual_seed(7)
    wgt = torch.randn(color.shape, generator=gen).to(dev)
    if backward:
        (color * wgt).sum().backward()
        torch.cuda.synchronize()

    kw = dict(means3D=d["means3D"], opacities=d["opacities"], W=view.image_width, H=view.image_height,
              viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, campos=rs.campos, bg=rs.bg, tanfovx=view.tanfovx,
              tanfovy=view.tanfovy, sh_degree=d["sh_degree"], shs=d["shs"], scales=d["scales"], rotations=d["rotations"])
    # ---- reference CUDA library ----
    if ReferenceCUDA.available():
        ref = ReferenceCUDA()
        rc, rr, R = ref.forward(**kw)
        torch.cuda.synchronize()
        rep["ref_R"] = int(R)
        rep["ref_color_maxabs"] = float((color.detach() - rc).abs().max())
        rep["ref_color_nbad_1e-4"] = int(((color.detach() - rc).abs() > 1e-4).sum())
        rep["ref_radii_mismatch"] = int((radii != rr).sum())
        if backward:
            g = ref.backward(wgt)
            pairs = {"means3D": "dL_dmeans3D", "opacities": "dL_dopacity", "shs": "dL_dsh", "scales": "dL_dscales",
                     "rotations": "dL_drotations"}
            for k, gk in pairs.items():
                a, b = leaves[k].grad, g[gk].reshape(leaves[k].grad.shape)
                den = float(b.abs().max()) + 1e-20
                rep[f"ref_grad_{k}_relmax"] = float((a - b).abs().max()) / den
            a, b = means2D.grad, g["dL_dmeans2D"]
            rep["ref_grad_means2D_relmax"] = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-20)
        # internal structures: per-tile lists must be identical (same order incl. ties)
        try:
            ref_list = ref.fetch("point_list")
            ref_ranges = ref.fetch("ranges").reshape(-1, 2)
            rep["ref_internal_R"] = int(ref_list.shape[0])
        except Exception as e:  # noqa: BLE001
            rep["ref_fetch_error"] = repr(e)
            ref_list = None
    else:
        rep["ref"] = "unavailable"
        ref_list = None

    # ---- my internals (re-run through _C to get the buffers) ----
    out = _C.rasterize_gaussians(rs.bg, d["means3D"], torch.Tensor([]), d["opacities"], d["scales"], d["rotations"], 1.0,
                                 torch.Tensor([]), rs.viewmatrix, rs.projmatrix, view.tanfovx, view.tanfovy,
                                 view.image_height, view.image_width, d["shs"], d["sh_degree"], rs.campos, False, False)
    R_me, color2, radii2, gb, bb, ib = out
    torch.cuda.synchronize()
    rep["R"] = int(R_me)
    rep["rerun_identical"] = bool(torch.equal(color2, color.detach()))
    keepalive = [t.contiguous() for t in (rs.bg, d["means3D"], d["shs"], d["opacities"], d["scales"], d["rotations"],
                                          rs.viewmatrix, rs.projmatrix, rs.campos)]
    scene = _C.make_scene(P=rep["P"], sh_degree=d["sh_degree"], sh_stride=d["shs"].shape[1], width=view.image_width,
                          height=view.image_height, tan_fovx=view.tanfovx, tan_fovy=view.tanfovy, scale_modifier=1.0,
                          prefiltered=False, debug=False, background=keepalive[0], means3D=keepalive[1],
                          shs=keepalive[2], colors_precomp=None, opacities=keepalive[3], scales=keepalive[4],
                          rotations=keepalive[5], cov3D_precomp=None, viewmatrix=keepalive[6], projmatrix=keepalive[7],
                          campos=keepalive[8])
    my_list = _C.fetch("point_list", scene, gb, bb, ib, R_me).numpy().view(np.uint32)
    my_ranges = _C.fetch("ranges", scene, gb, bb, ib, R_me).numpy().view(np.uint32).reshape(-1, 2)
    if ref_list is not None and ref_list.shape[0] == my_list.shape[0]:
        rep["list_equal_ref"] = bool(np.array_equal(ref_list, my_list))
        ne = ref_ranges[:, 0] != ref_ranges[:, 1]
        rep["ranges_equal_ref_nonempty"] = bool(np.array_equal(ref_ranges[ne], my_ranges[ne]))
        rep["ranges_empty_consistent"] = bool(np.all(my_ranges[~ne, 0] == my_ranges[~ne, 1]))
    # ---- CPU oracle ----
    if use_cpu:
        o = Oracle(32)
        t0 = time.time()
        f = o.forward(**kw)
        rep["oracle_fwd_s"] = time.time() - t0
        rep["oracle_R"] = int(f["num_rendered"])
        diff = np.abs(f["color"] - color.detach().cpu().numpy())
        rep["oracle_color_maxabs"] = float(diff.max())
        rep["oracle_color_nbad_1e-4"] = int((diff > 1e-4).sum())
        rep["oracle_radii_mismatch"] = int((f["radii"] != radii.cpu().numpy()).sum())
        if f["point_list"].shape[0] == my_list.shape[0]:
            rep["list_equal_oracle"] = bool(np.array_equal(f["point_list"], my_list))
        if backward:
            kwb = {k: v for k, v in kw.items() if k != "opacities"}
            t0 = time.time()
            g = o.backward(f, wgt.cpu().numpy(), **kwb)
            rep["oracle_bwd_s"] = time.time() - t0
            pairs = {"means3D": "dL_dmeans3D", "opacities": "dL_dopacity", "shs": "dL_dsh", "scales": "dL_dscales",
                     "rotations": "dL_drotations"}
            for k, gk in pairs.items():
                a, b = leaves[k].grad.cpu().numpy(), g[gk].reshape(leaves[k].grad.shape)
                rep[f"oracle_grad_{k}_relmax"] = float(np.abs(a - b).max() / (np.abs(b).max() + 1e-20))
    # ---- timings ----
    if timing_iters:
        def time_fn(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(timing_iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / timing_iters

        with torch.no_grad():
            rep["ms_forward"] = time_fn(lambda: rast(d["means3D"], means2D, d["opacities"], shs=d["shs"],
                                                     scales=d["scales"], rotations=d["rotations"]))
        if ReferenceCUDA.available():
            rep["ms_forward_ref"] = time_fn(lambda: ref.forward(**kw))
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    orbit = scenes.orbit_c2w(12)
    reports = []
    cases = [
        ("tiny_sh3", scenes.tiny_cloud(3000, seed=1, sh_degree=3), scenes.make_view(orbit[1], 200, 136), [0.2, 0.4, 0.6]),
        ("tiny_ties", scenes.tiny_cloud(5000, seed=2, sh_degree=2, depth_ties=True), scenes.make_view(orbit[0], 256, 256), [1, 1, 1]),
        ("human20k", scenes.human_cloud(20000, scale_factor=256.0, seed=0), scenes.make_view(orbit[3], 512, 512), [1, 1, 1]),
    ]
    for name, cl, v, bg in cases:
        r = run_case(name, cl, v, bg, use_cpu=not args.no_cpu, timing_iters=5)
        print(json.dumps(r), flush=True)
        reports.append(r)
    if args.big:
        cl = scenes.human_cloud(799957, scale_factor=448.0, seed=0)
        v = scenes.make_view(scenes.orbit_c2w(120)[7], 1920, 1080)
        r = run_case("C2_800k_1080p", cl, v, [1, 1, 1], use_cpu=not args.no_cpu, backward=True, timing_iters=20)
        print(json.dumps(r), flush=True)
        reports.append(r)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as fh:
        json.dump(reports, fh, indent=1)


if __name__ == "__main__":
    main()
