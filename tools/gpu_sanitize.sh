#!/bin/bash
# compute-sanitizer over a small forward + backward through the drop-in module and the resident renderer
# (memcheck, racecheck, synccheck; initcheck is too noisy with torch's caching allocator).
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "gaussian-pcloud-render_b200"))
import scenes
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from renderer import FrameRenderer
dev = torch.device("cuda:0")
for (P, W, H, ds) in ((6000, 200, 136, 1), (6000, 208, 144, 2), (40, 33, 17, 1)):
    cl = scenes.human_cloud(P, scale_factor=120.0, seed=4, opacity="uniform")
    v = scenes.make_view(scenes.orbit_c2w(12)[2], W, H)
    t = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(dev)
    rs = GaussianRasterizationSettings(H, W, v.tanfovx, v.tanfovy, t([1.0, 1.0, 1.0]), 1.0, t(v.viewmatrix), t(v.projmatrix),
                                       1, t(v.campos), False, False)
    lv = {k: cl[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    m2 = torch.zeros_like(lv["means3D"], requires_grad=True)
    color, radii = GaussianRasterizer(rs, downsample=ds)(lv["means3D"], m2, lv["opacities"], shs=lv["shs"],
                                                         scales=lv["scales"], rotations=lv["rotations"])
    color.sum().backward()
    fr = FrameRenderer(cl, W, H, [1, 1, 1], dev, capacity=2_000_000, downsample=ds)
    vd = fr.upload_view(v)
    ex = [(torch.rand(P, 3, device=dev), torch.empty_like(fr.color)) for _ in range(3)]
    fr.enqueue(vd, extra_passes=ex)
    fr.enqueue_pass(vd, torch.empty_like(fr.color), colors_precomp=ex[0][0])
    torch.cuda.synchronize()
    print("ok", P, W, H, ds, int((radii > 0).sum()), fr.status())
# camera set-up, head decode and the caller-level pass loop
from diff_gaussian_rasterization import _C
from renderer import ViewBatch, render_passes
vb = ViewBatch(scenes.orbit_c2w(12)[:3], 45.0, dev)
cl = scenes.human_cloud(3000, scale_factor=120.0, seed=5, opacity="uniform")
fr = FrameRenderer(cl, 96, 64, [1, 1, 1], dev, capacity=1_000_000, downsample=2)
out = render_passes(fr, vb, normals=torch.nn.functional.normalize(torch.randn(3000, 3), dim=-1))
for C_, kw in ((8, {}), (26, dict(use_offset=True, use_dc_offset=True, est_normal=True, sh_ac_coeffs=3)), (0, dict(use_rotation=False, use_scale=False, use_opacity=False))):
    hd = _C.decode_head(torch.randn(777, C_, device=dev), torch.rand(777, 3, device=dev), torch.rand(777, 3, device=dev) * 1000,
                        scale_factor=448, xyz_offset=512, **kw)
torch.cuda.synchronize()
print("ok extras", out["rgb"].shape, hd["shs"].shape)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | grep -v "^ok" | tail -15
done > gpurun_out/sanitizer.log 2>&1
cat gpurun_out/sanitizer.log
