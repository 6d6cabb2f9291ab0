"""Synthetic workloads and camera/settings construction for the rasterizer hot path.

Host-side restatement of what the reference's callers feed the rasterizer (paths relative to /root/reference/):
  * projection matrix           simple_raw_render.py:51-71   (getProjectionMatrix)
  * camera -> raster settings   simple_raw_render.py:79-112  (get_rasterize_param_from_camera; note the
                                full-angle tan(fov) quirk at :101-102 and the transposed matrices at :83-93)
  * per-point attributes        simple_raw_render.py:239-250, models/model_v2.py:292-324,358-365
  * camera orbit                structures.py:3950-4053 (generate_camera_circle_path, d=0, r=3,
                                center_angles=[90,0]); view k must equal validate/temp_state_dict.pt, which is
                                pinned by tests/golden/orbit12_H_c2w.npy.
Everything is numpy + torch on the host; nothing here touches the GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """Perspective matrix with the half-angle tangent (simple_raw_render.py:51-71)."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def orbit_c2w(n: int, r: float = 3.0) -> np.ndarray:
    """(n,4,4) camera-to-world poses of the reference's circle path (d=0, center_angles=[90,0], yz inverted):
    position (-r cos t, 0, r sin t), t = linspace(0, 2pi, n); z axis looks at the origin, y axis = (0,-1,0)."""
    th = np.linspace(0.0, 2.0 * np.pi, n, dtype=np.float32).astype(np.float64)
    out = np.zeros((n, 4, 4), np.float64)
    for k, t in enumerate(th):
        pos = np.array([-r * math.cos(t), 0.0, r * math.sin(t)])
        z = -pos / np.linalg.norm(pos)
        y = np.array([0.0, -1.0, 0.0])
        x = np.cross(y, z)
        out[k, :3, 0], out[k, :3, 1], out[k, :3, 2], out[k, :3, 3] = x, y, z, pos
        out[k, 3, 3] = 1.0
    return out.astype(np.float32)


@dataclass
class View:
    """The per-view fields of GaussianRasterizationSettings, as host float32 arrays."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # (4,4) = H_w2c^T  (column-major for the kernels)
    projmatrix: np.ndarray   # (4,4) = (P @ H_w2c)^T
    campos: np.ndarray       # (3,)


def make_view(H_c2w: np.ndarray, width: int, height: int, fov_deg: float = 45.0, super_sample: int = 1) -> View:
    """simple_raw_render.py:79-112 for one camera."""
    H_c2w = np.asarray(H_c2w, np.float32)
    H_w2c = np.linalg.inv(H_c2w.astype(np.float64)).astype(np.float32)
    fov = math.pi * fov_deg / 180.0
    Pm = projection_matrix(0.01, 100.0, fov, fov)
    view_t = np.ascontiguousarray(H_w2c.T)
    full_t = np.ascontiguousarray((view_t.astype(np.float32) @ Pm.T.astype(np.float32)).astype(np.float32))
    t = math.tan(fov_deg / 180.0 * math.pi)  # FULL angle: the reference's quirk
    return View(height * super_sample, width * super_sample, t, t, view_t, full_t,
                np.ascontiguousarray(H_c2w[:3, 3]))


# ------------------------------------------------------------------------------------------------
def _human_surface(n: int, rng: np.random.Generator) -> np.ndarray:
    """Surface samples of a union of ellipsoids (torso, head, arms, legs) inside the THuman bbox
    x[-.53,.53] y[-1,1] z[-.23,.23]."""
    parts = [  # centre, radii, weight
        ((0.0, 0.25, 0.0), (0.22, 0.36, 0.14), 0.30),
        ((0.0, 0.80, 0.0), (0.12, 0.16, 0.13), 0.10),
        ((-0.38, 0.30, 0.0), (0.15, 0.09, 0.08), 0.10),
        ((0.38, 0.30, 0.0), (0.15, 0.09, 0.08), 0.10),
        ((-0.13, -0.50, 0.0), (0.11, 0.48, 0.12), 0.20),
        ((0.13, -0.50, 0.0), (0.11, 0.48, 0.12), 0.20),
    ]
    w = np.array([p[2] for p in parts])
    counts = np.floor(w / w.sum() * n).astype(int)
    counts[0] += n - counts.sum()
    pts = []
    for (c, rad, _), m in zip(parts, counts):
        v = rng.standard_normal((m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts.append(v * np.array(rad) + np.array(c))
    p = np.concatenate(pts, 0)
    rng.shuffle(p, axis=0)
    return p.astype(np.float32)


def human_cloud(P: int, scale_factor: float = 448.0, seed: int = 0, voxelize: Optional[int] = None,
                opacity: str = "ones", decoded_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """THuman-shaped cloud with the network head's attribute statistics (C1/C2/C3 of BASELINE.md)."""
    rng = np.random.default_rng(seed)
    xyz = _human_surface(P, rng)
    if voxelize:
        xyz = np.unique(np.round(xyz * voxelize), axis=0).astype(np.float32) / voxelize
        rng.shuffle(xyz, axis=0)
    n = xyz.shape[0]
    rot = np.zeros((n, 4), np.float32)
    rot[:, 0] = 1.0
    rot += 0.1 * rng.standard_normal((n, 4)).astype(np.float32)
    radius = math.sqrt(3.0) / scale_factor * 6.0
    scales = np.clip(1.0 + 0.25 * rng.standard_normal((n, 3)), 0.0, None).astype(np.float32) * np.float32(
        radius * decoded_scale)
    if opacity == "ones":
        op = np.ones((n, 1), np.float32)
    else:
        op = rng.uniform(0.0, 1.0, (n, 1)).astype(np.float32)
    ph = rng.uniform(0, 2 * np.pi, 3)
    rgb = 0.5 + 0.45 * np.sin(xyz @ rng.uniform(4.0, 9.0, (3, 3)).astype(np.float32) + ph)
    sh = np.zeros((n, 13, 3), np.float32)  # DC + 12 zero coefficients: model_v2.py:358-365
    sh[:, 0] = (rgb - 0.5) / SH_C0
    return dict(means3D=torch.from_numpy(xyz), rotations=torch.from_numpy(rot), scales=torch.from_numpy(scales),
                opacities=torch.from_numpy(op), shs=torch.from_numpy(sh), sh_degree=1)


def random_cloud(P: int, seed: int = 1, sh_degree: int = 3, extent: float = 1.0, smin: float = 0.002,
                 smax: float = 0.02) -> Dict[str, torch.Tensor]:
    """C4: uniform random Gaussians, log-uniform scales, unit quaternions, opacity U(.05,1), full SH."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-extent, extent, (P, 3)).astype(np.float32)
    scales = np.exp(rng.uniform(math.log(smin), math.log(smax), (P, 3))).astype(np.float32)
    q = rng.standard_normal((P, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    op = rng.uniform(0.05, 1.0, (P, 1)).astype(np.float32)
    M = (sh_degree + 1) ** 2
    sh = (0.3 * rng.standard_normal((P, M, 3))).astype(np.float32)
    return dict(means3D=torch.from_numpy(xyz), rotations=torch.from_numpy(q), scales=torch.from_numpy(scales),
                opacities=torch.from_numpy(op), shs=torch.from_numpy(sh), sh_degree=sh_degree)


def pack_cloud(cloud: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The same cloud without the SH coefficients that are zero for every point, and the SH degree that is left.
    The reference's head stores a DC colour plus 12 zero coefficients per point and renders with degree 1
    (models/model_v2.py:358-365); the rasterizer accepts any coefficient stride >= (degree + 1)^2, and terms with zero
    coefficients add exactly 0, so the packed cloud renders the same frame bit for bit
    (tests/test_gpu.py::test_packed_sh_without_zero_tail_renders_the_same_frame).  This is the layout gs_decode_head
    emits; 56 instead of 200 bytes per point for the pcrender shape."""
    sh = cloud["shs"]
    nz = (sh != 0).flatten(2).any(2).any(0).nonzero()
    used = int(nz.max()) + 1 if nz.numel() else 1
    deg = 0 if used <= 1 else 1 if used <= 4 else 2 if used <= 9 else 3
    deg = min(deg, int(cloud["sh_degree"]))
    out = dict(cloud)
    out["shs"] = sh[:, : (deg + 1) ** 2].contiguous()
    out["sh_degree"] = deg
    return out


def tiny_cloud(P: int, seed: int = 0, sh_degree: int = 3, M: Optional[int] = None, spread: float = 0.6,
               scale: float = 0.05, opacity_lo: float = 0.05, depth_ties: bool = False) -> Dict[str, torch.Tensor]:
    """Small general-purpose test cloud around the origin (anisotropic, un-normalised quats, mixed opacity)."""
    rng = np.random.default_rng(seed)
    xyz = (spread * rng.standard_normal((P, 3))).astype(np.float32)
    if depth_ties:  # voxelised coordinates -> many exact depth ties under axis-aligned views
        xyz = (np.round(xyz * 16) / 16).astype(np.float32)
    scales = (scale * np.exp(rng.uniform(-1.2, 0.8, (P, 3)))).astype(np.float32)
    q = np.zeros((P, 4), np.float32)
    q[:, 0] = 1
    q += 0.4 * rng.standard_normal((P, 4)).astype(np.float32)
    op = rng.uniform(opacity_lo, 1.0, (P, 1)).astype(np.float32)
    M = M or (sh_degree + 1) ** 2
    sh = (0.4 * rng.standard_normal((P, M, 3))).astype(np.float32)
    sh[:, 0] += 0.8
    return dict(means3D=torch.from_numpy(xyz), rotations=torch.from_numpy(q), scales=torch.from_numpy(scales),
                opacities=torch.from_numpy(op), shs=torch.from_numpy(sh), sh_degree=sh_degree)
