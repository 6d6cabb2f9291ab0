"""Drop-in `diff_gaussian_rasterization` for huzi96/gaussian-pcloud-render, backed by libgsplat_b200.so.

Mirrors the reference's Python API surface (dgr/diff_gaussian_rasterization/__init__.py):
  GaussianRasterizationSettings   :157-169  same 12 fields, same order
  GaussianRasterizer              :171-220  forward(means3D, means2D, opacities, shs, colors_precomp, scales,
                                            rotations, cov3D_precomp) -> (color (3,H,W), radii (P,) int32),
                                            markVisible(positions) -> bool (P,)
  rasterize_gaussians             :21-42
  _RasterizeGaussians             :44-155   autograd Function; gradient order as :143-153
so `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(simple_raw_render.py:12) keeps working unchanged.  Same exceptions for the exactly-one-of rules, same
`debug=True` behaviour (inputs dumped to snapshot_fw.dump / snapshot_bw.dump when the library raises).

Extensions (defaults = reference behaviour): `GaussianRasterizer(settings, tile_rows=(r0, r1))` restricts binning
and blending to tile rows [r0, r1) (multi-GPU path); with `grad_group=<process group>` the backward of
such a shard sums the per-Gaussian partial gradients over the ranks (one all-reduce) so that every rank returns the
gradients of the whole frame; `GaussianRasterizer(settings, downsample=2)` returns
the (3, H/2, W/2) image the reference's caller computes with `F.interpolate(..., mode="bilinear")` after a
super-sampled render (simple_raw_render.py:281-284), folded into the blend epilogue, differentiable.
"""
from typing import NamedTuple, Optional, Tuple

import torch
import torch.nn as nn

from . import _C


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, tile_rows: Optional[Tuple[int, int]] = None, downsample: int = 1,
                        grad_group=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, tile_rows, downsample, grad_group)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, tile_rows=None, downsample=1, grad_group=None):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh,
                rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        # no input needs a gradient (e.g. the benchmark's torch.no_grad() rendering, simple_benchmark.py:198):
        # backward will never run, so the scratch buffers can be recycled instead of allocated per frame
        reuse = not any(ctx.needs_input_grad)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy before anything can corrupt them
            try:
                out = _C.rasterize_gaussians(*args, tile_rows=tile_rows, reuse_workspace=reuse, downsample=downsample)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _C.rasterize_gaussians(*args, tile_rows=tile_rows, reuse_workspace=reuse, downsample=downsample)
        num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = out
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.tile_rows = tile_rows
        ctx.downsample = downsample
        ctx.grad_group = grad_group
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, sh, rs.sh_degree, rs.campos,
                geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads_c = _C.rasterize_gaussians_backward(*args, tile_rows=ctx.tile_rows, downsample=ctx.downsample,
                                                          grad_group=ctx.grad_group)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads_c = _C.rasterize_gaussians_backward(*args, tile_rows=ctx.tile_rows, downsample=ctx.downsample,
                                                          grad_group=ctx.grad_group)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = grads_c

        def _like(g, ref):  # "not provided" inputs were empty CPU tensors: hand autograd a matching empty gradient
            return g if ref.numel() != 0 else None

        return (grad_means3D, grad_means2D, _like(grad_sh, sh), _like(grad_colors_precomp, colors_precomp),
                grad_opacities, _like(grad_scales, scales),
                _like(grad_rotations, rotations), _like(grad_cov3Ds_precomp, cov3Ds_precomp), None, None, None, None)


class GaussianRasterizationSettings(NamedTuple):
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


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, tile_rows: Optional[Tuple[int, int]] = None, downsample: int = 1,
                 grad_group=None):
        super().__init__()
        self.raster_settings = raster_settings
        self.tile_rows = tile_rows
        self.downsample = downsample
        self.grad_group = grad_group

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            rs = self.raster_settings
            visible = _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings, self.tile_rows, self.downsample, self.grad_group)
