"""Generates tests/golden/camera_params.npz by EXECUTING the reference's own camera set-up code in the build
container (where /root/reference exists): `getProjectionMatrix` + `get_rasterize_param_from_camera`
(simple_raw_render.py:50-112) and `inv_homogeneous_tensors` (plib/rigid_motion.py:687-703).

simple_raw_render.py cannot be imported here (MinkowskiEngine, open3d ... are absent), so the three function
definitions are cut out of the reference files with `ast` and executed as they are; the only stand-ins are a minimal
camera object (H_c2w, width/height, get_H_w2c -> the reference's inv_homogeneous_tensors) and `Tensor.cuda`, which is
made the identity because this container has no GPU.  Nothing of the reference is copied into the repository: only the
numbers it produces.

    python tests/golden/make_camera_golden.py      # build container only
"""
import ast
import math
import os
import sys
from typing import NamedTuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def cut(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    return "\n\n".join(ast.get_source_segment(src, n) for n in tree.body
                       if isinstance(n, ast.FunctionDef) and n.name in names)


class GaussianRasterizationSettings(NamedTuple):  # field names of dgr/diff_gaussian_rasterization/__init__.py:157-169
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def main():
    ns = {"torch": torch, "np": np, "math": math, "Camera": object,
          "GaussianRasterizationSettings": GaussianRasterizationSettings}
    exec(cut(os.path.join(REF, "plib", "rigid_motion.py"), {"inv_homogeneous_tensors"}), ns)
    exec(cut(os.path.join(REF, "simple_raw_render.py"), {"getProjectionMatrix", "get_rasterize_param_from_camera"}), ns)
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU in the build container

    class Cam:
        def __init__(self, H_c2w, w, h):
            self.H_c2w, self.width_px, self.height_px = H_c2w, w, h

        def get_H_w2c(self):
            return ns["inv_homogeneous_tensors"](self.H_c2w)

    orbit = np.load(os.path.join(HERE, "orbit12_H_c2w.npy")).astype(np.float32).reshape(-1, 4, 4)
    rng = np.random.default_rng(77)
    extra = []
    for _ in range(20):  # random rigid motions
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        H = np.eye(4)
        H[:3, :3] = R
        H[:3, 3] = rng.uniform(-4, 4, 3)
        extra.append(H.astype(np.float32))
    c2w = np.concatenate([orbit, np.stack(extra)], 0)
    out = {"c2w": c2w}
    for fov in (45.0, 40.0):
        view, proj, campos, tan = [], [], [], None
        for H in c2w:
            cam = Cam(torch.from_numpy(H)[None, None], 512, 512)  # (1,1,4,4) as the reference's chunks
            rs = ns["get_rasterize_param_from_camera"](cam, device=torch.device("cpu"), fovX_deg=fov, fovY_deg=fov,
                                                       sh_degree=1, bg=torch.ones(3), super_sample_rate=2)
            view.append(rs.viewmatrix.reshape(4, 4).numpy())
            proj.append(rs.projmatrix.reshape(4, 4).numpy())
            campos.append(rs.campos.reshape(3).numpy())
            tan = (rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width)
        out[f"view_{int(fov)}"] = np.stack(view)
        out[f"proj_{int(fov)}"] = np.stack(proj)
        out[f"campos_{int(fov)}"] = np.stack(campos)
        out[f"scalars_{int(fov)}"] = np.array(tan, np.float64)
    np.savez_compressed(os.path.join(HERE, "camera_params.npz"), **out)
    print("wrote camera_params.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container)")
    main()
