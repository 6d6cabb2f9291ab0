"""Developer probe (GPU box): sh_degree 0 with an SH array -- this library vs the live reference library, and the
padded (P,13,3)/degree-1 form of the same colours."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from oracle.oracle import ReferenceCUDA  # noqa: E402

dev = torch.device("cuda:0")
cl = scenes.human_cloud(50000, scale_factor=300.0, seed=13, opacity="uniform")
v = scenes.make_view(scenes.orbit_c2w(12)[7], 640, 400)
t = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(dev)
d = {k: cl[k].to(dev) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
bg = torch.ones(3, device=dev)
ref = ReferenceCUDA()
res = {}
for name, shs, deg in (("padded_deg1", d["shs"], 1), ("packed_deg0", d["shs"][:, :1].contiguous(), 0),
                       ("padded_deg0", d["shs"], 0)):
    rs = GaussianRasterizationSettings(400, 640, v.tanfovx, v.tanfovy, bg, 1.0, t(v.viewmatrix), t(v.projmatrix), deg,
                                       t(v.campos), False, False)
    ours, _ = GaussianRasterizer(rs)(d["means3D"], None, d["opacities"], shs=shs, scales=d["scales"],
                                     rotations=d["rotations"])
    theirs = ref.forward(means3D=d["means3D"], opacities=d["opacities"], W=640, H=400, viewmatrix=t(v.viewmatrix),
                         projmatrix=t(v.projmatrix), campos=t(v.campos), bg=bg, tanfovx=v.tanfovx, tanfovy=v.tanfovy,
                         sh_degree=deg, shs=shs, scales=d["scales"], rotations=d["rotations"])[0]
    res[name] = (ours.clone(), theirs.clone())
    print(name, "ours==reference:", bool(torch.equal(ours, theirs)), "max diff", float((ours - theirs).abs().max()))
for a in ("packed_deg0", "padded_deg0"):
    print(a, "vs padded_deg1: ours", float((res[a][0] - res["padded_deg1"][0]).abs().max()),
          " reference", float((res[a][1] - res["padded_deg1"][1]).abs().max()))
