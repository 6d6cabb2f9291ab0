"""CPU restatement (torch CPU tensors, fp32) of the reference's network-head decode -- the oracle for gs_decode_head.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Pinned against tests/golden/head_decode.npz, the outputs of the
reference's own statements executed in the build container (tests/golden/make_head_golden.py).

Follows models/model_v2.py:287-375 (feature slicing in the order rotation, scale, opacity, offset, dc offset, normal,
SH AC; `+ default_quaternion`, `+ ones`, clamps, RGB2SH = (rgb - 0.5) / C0 of models/sh_utils.py:114-115, zero SH
padding of `2 ** (sh_deg + 1) * 3` coefficients) and the caller's glue in simple_raw_render.py: pcgc_rescale
(:71-75, means = (xyz - offset) / factor), `scales = decoded_s * radius` with radius = sqrt(3) / scale_factor * 6
(:248-249), opacity replaced by ones when `enable_opacity` is false (:243-247, :396-398).

`cuda_scalar_division=True` evaluates the two divisions by a scalar the way torch does on a CUDA tensor (multiplication
by the reciprocal, taken in double and narrowed to fp32; ATen/native/cuda/BinaryDivTrueKernel.cu) -- that is the arithmetic of the reference's real
(GPU) run and what the kernel reproduces; False is torch's CPU arithmetic (a true division), which is what the
golden file, generated without a GPU, contains.  The two differ by at most one ulp.
"""
import numpy as np
import torch

C0 = 0.28209479177387814


def _div(t: torch.Tensor, b: float, cuda_scalar_division: bool) -> torch.Tensor:
    if not cuda_scalar_division:
        return t / b
    return t * float(np.float32(1.0 / float(b)))  # reciprocal in double, narrowed (measured: tools/div_probe.py)


def decode_head(features, dc_rgb, primitives, *, scale_factor, xyz_offset, use_rotation=True, use_scale=True,
                use_opacity=True, use_offset=False, use_dc_offset=False, est_normal=False, normalize_normal=True,
                sh_deg=1, sh_feat_deg=0, enable_opacity=True, cuda_scalar_division=False) -> dict:
    f = torch.as_tensor(features, dtype=torch.float32)
    rgb = torch.as_tensor(dc_rgb, dtype=torch.float32)
    prim = torch.as_tensor(primitives, dtype=torch.float32)
    P = f.shape[0]
    quat = torch.tensor([[1, 0, 0, 0]], dtype=torch.float32)
    used = 0
    if use_rotation:
        r = f[:, 0:4] + quat
        used += 4
    else:
        r = quat.expand(P, 4)
    if use_scale:
        s = torch.clamp(f[:, used:used + 3] + torch.ones_like(f[:, used:used + 3]), min=0.)
        used += 3
    else:
        s = torch.ones_like(f[:, 0:3]) if f.shape[1] >= 3 else torch.ones(P, 3)
    if use_opacity:
        o = torch.clamp(f[:, used:used + 1], min=0., max=1.)
        used += 1
    else:
        o = torch.ones(P, 1)
    offset = None
    if use_offset:
        offset = f[:, used:used + 3]
        used += 3
    dc = _div(rgb - 0.5, C0, cuda_scalar_division)
    if use_dc_offset:
        sh_dc = (f[:, used:used + 3] + dc).unsqueeze(-2)
        used += 3
    else:
        sh_dc = dc.unsqueeze(-2)
    n = None
    if est_normal:
        n = f[:, used:used + 3]
        used += 3
        if normalize_normal:
            n = torch.nn.functional.normalize(n, dim=-1)
    if sh_deg > 0 and sh_feat_deg > 0:
        sh = torch.cat([sh_dc, f[:, used:].reshape(P, -1, 3)], dim=1)
    elif sh_deg > 0 and sh_feat_deg == 0:
        sh = torch.cat([sh_dc, torch.zeros((P, (2 ** (sh_deg + 1)) * 3, 3))], dim=1)
    else:
        sh = sh_dc
    prim_aug = prim + offset if use_offset else prim
    means = _div(prim_aug - xyz_offset, scale_factor, cuda_scalar_division)
    radius = np.sqrt(3) / scale_factor * 6
    if not enable_opacity:
        o = torch.ones_like(o)
    return dict(means3D=means, rotations=r.contiguous(), scales=s * radius, opacities=o, shs=sh, normals=n)
