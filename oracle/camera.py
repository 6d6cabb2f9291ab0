"""CPU restatement (numpy, fp32) of the reference's per-view camera set-up -- the oracle for gs_make_views.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Pinned against tests/golden/camera_params.npz, which holds the
outputs of the reference's own functions executed in the build container (tests/golden/make_camera_golden.py).

Follows, term for term:
  inv_homogeneous_tensors          plib/rigid_motion.py:687-703   inverse of a rigid motion: [R^T | -1.0 * (R^T t)]
  getProjectionMatrix              simple_raw_render.py:50-69     Python-double arithmetic stored into an fp32 matrix
  get_rasterize_param_from_camera  simple_raw_render.py:79-112    viewmatrix = w2c^T, projmatrix = viewmatrix @ P^T,
                                                                  campos = H_c2w @ [0,0,0,1], tanfov = tan(FULL angle),
                                                                  raster size = camera size * super_sample_rate
"""
import math

import numpy as np


def inv_rigid(H: np.ndarray) -> np.ndarray:
    H = np.asarray(H, np.float32)
    inv = np.zeros_like(H)
    Rt = np.swapaxes(H[..., :3, :3], -2, -1)
    inv[..., :3, :3] = Rt
    inv[..., :3, 3:4] = np.float32(-1.0) * (Rt @ H[..., :3, 3:4])
    inv[..., 3, 3] = 1
    return inv


def projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> np.ndarray:
    top = math.tan(fovY / 2) * znear
    right = math.tan(fovX / 2) * znear
    bottom, left = -top, -right
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = 1.0 * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def raster_params(c2w: np.ndarray, fovX_deg: float, fovY_deg: float, width_px: int = 0, height_px: int = 0,
                  super_sample_rate: int = 2) -> dict:
    """c2w (N,4,4) -> dict(viewmatrix (N,4,4), projmatrix (N,4,4), campos (N,3), tanfovx, tanfovy, image_height/width)."""
    c2w = np.asarray(c2w, np.float32).reshape(-1, 4, 4)
    view = np.ascontiguousarray(np.swapaxes(inv_rigid(c2w), -2, -1))
    P = projection_matrix(0.01, 100, np.pi * fovX_deg / 180, np.pi * fovY_deg / 180)
    proj = (view @ np.ascontiguousarray(P.T)[None]).astype(np.float32)
    campos = (c2w @ np.array([0, 0, 0, 1], np.float32))[..., 0:3]
    return dict(viewmatrix=view, projmatrix=proj, campos=np.ascontiguousarray(campos),
                tanfovx=math.tan(fovX_deg / 180. * math.pi), tanfovy=math.tan(fovY_deg / 180. * math.pi),
                image_height=height_px * super_sample_rate, image_width=width_px * super_sample_rate)
