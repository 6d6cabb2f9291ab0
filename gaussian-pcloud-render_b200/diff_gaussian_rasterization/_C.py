"""`_C` -- binding of libgsplat_b200.so (include/gsplat_b200.h) with the call signatures of the reference's pybind
module (dgr/ext.cpp:15-19, dgr/rasterize_points.h:18-66):

    rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                        projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
                        debug) -> (num_rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer)
    rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R,
                        binningBuffer, imageBuffer, debug)
                     -> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    mark_visible(means3D, viewmatrix, projmatrix) -> bool tensor

PyTorch is only used for device memory and the current stream; the library itself sees raw pointers.  There is NO
CPU or eager fallback: if the shared library is missing or the tensors are not CUDA tensors this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GSPLAT_B200_LIB: developer override used by the profiling tools to load an instrumented build of the same library
_SO = os.environ.get("GSPLAT_B200_LIB") or os.path.join(_PKG, "libgsplat_b200.so")

GS_ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "buffer allocation failed",
             -4: "instance capacity exceeded", -5: "unsupported size (more than 65536 tiles or too many Gaussians)",
             -6: "point culled although `prefiltered` is set"}


class GsScene(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_degree", C.c_int32), ("sh_stride", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
                ("prefiltered", C.c_int32), ("debug", C.c_int32), ("tile_row_begin", C.c_int32),
                ("tile_row_end", C.c_int32),
                ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
                ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p),
                ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("num_peers", C.c_int32), ("downsample", C.c_int32), ("peer_out_color", C.c_void_p * 8),
                ("num_extra", C.c_int32), ("team_after", C.c_int32), ("extra_colors", C.c_void_p * 3),
                ("extra_out", C.c_void_p * 3), ("shard_cull", C.c_int32), ("blend_split", C.c_int32)]


class GsHeadLayout(C.Structure):
    _fields_ = [("C", C.c_int32), ("use_rotation", C.c_int32), ("use_scale", C.c_int32), ("use_opacity", C.c_int32),
                ("use_offset", C.c_int32), ("use_dc_offset", C.c_int32), ("est_normal", C.c_int32),
                ("normalize_normal", C.c_int32), ("sh_ac_coeffs", C.c_int32), ("enable_opacity", C.c_int32),
                ("radius", C.c_float), ("xyz_offset", C.c_float), ("xyz_factor", C.c_float)]


RESIZE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class GsBuffer(C.Structure):
    _fields_ = [("fn", RESIZE_FN), ("user", C.c_void_p)]


class GsStatus(C.Structure):
    _fields_ = [("num_rendered", C.c_int64), ("num_visible", C.c_int32), ("code", C.c_int32)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libgsplat_b200.so; fails loudly when it has not been built (`__graft_entry__.build()`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: build it with `make -C {os.path.join(_PKG, 'csrc')}` "
                              "(there is no CPU fallback)")
        L = C.CDLL(_SO)
        L.gs_forward.restype = C.c_int64
        L.gs_forward.argtypes = [C.POINTER(GsScene), GsBuffer, GsBuffer, GsBuffer, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gs_forward_nosync.restype = C.c_int32
        L.gs_forward_nosync.argtypes = [C.POINTER(GsScene), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.gs_forward_recolor.restype = C.c_int32
        L.gs_forward_recolor.argtypes = [C.POINTER(GsScene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gs_read_status.restype = C.c_int32
        L.gs_read_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gs_backward.restype = C.c_int32
        L.gs_backward.argtypes = [C.POINTER(GsScene), C.c_int64] + [C.c_void_p] * 15
        L.gs_backward_stage.restype = C.c_int32
        L.gs_backward_stage.argtypes = [C.POINTER(GsScene), C.c_int64] + [C.c_void_p] * 14 + [C.c_int32, C.c_void_p]
        L.gs_mark_visible.restype = C.c_int32
        L.gs_mark_visible.argtypes = [C.c_int32] + [C.c_void_p] * 5
        L.gs_make_views.restype = C.c_int32
        L.gs_make_views.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        L.gs_decode_head.restype = C.c_int32
        L.gs_decode_head.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(GsHeadLayout)] + \
                                    [C.c_void_p] * 7
        L.gs_fetch.restype = C.c_int64
        L.gs_fetch.argtypes = [C.POINTER(GsScene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p,
                               C.c_void_p, C.c_int64, C.c_void_p]
        L.gs_geometry_bytes.restype = C.c_size_t
        L.gs_geometry_bytes.argtypes = [C.c_int32]
        L.gs_image_bytes.restype = C.c_size_t
        L.gs_image_bytes.argtypes = [C.c_int32, C.c_int32]
        L.gs_binning_bytes.restype = C.c_size_t
        L.gs_binning_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32]
        L.gs_launch_count.restype = C.c_int64
        L.gs_profile_enable.restype = None
        L.gs_profile_enable.argtypes = [C.c_int32]
        L.gs_profile_read.restype = C.c_int32
        L.gs_profile_read.argtypes = [C.c_void_p]
        L.gs_profile_read_backward.restype = C.c_int32
        L.gs_profile_read_backward.argtypes = [C.c_void_p]
        L.gs_last_error.restype = C.c_char_p
        L.gs_abi_version.restype = C.c_int32
        _lib = L
    return _lib


def _check(rc: int, what: str) -> int:
    if rc < 0:
        detail = lib().gs_last_error().decode() if rc == -2 else ""
        raise RuntimeError(f"{what}: {GS_ERRORS.get(int(rc), rc)} {detail}".strip())
    return rc


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous fp32 CUDA tensor; an empty tensor (the reference's "not provided") -> NULL."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError("diff_gaussian_rasterization (B200): all tensors must be CUDA tensors; no CPU path exists")
    return t.data_ptr()


def _prep(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class _Growable:
    """A uint8 CUDA tensor that the library resizes through a C callback (resizeFunctional, rasterize_points.cu:27-33).
    Sizes are rounded up to 1/8-octave buckets so that PyTorch's caching allocator finds a block of the same size
    again on the next frame (the instance count, hence the binning buffer, changes from view to view)."""

    def __init__(self, device, stream=None):
        self.t = torch.empty(0, dtype=torch.uint8, device=device)

        def _resize(_user, nbytes):
            n = int(nbytes)
            if n > self.t.numel():
                step = max(1 << 20, 1 << max(0, n.bit_length() - 4))
                if stream is not None and self.t.numel():
                    self.t.record_stream(stream)  # pooled: earlier frames of this stream may still read the old block
                self.t = torch.empty(((n + step - 1) // step) * step, dtype=torch.uint8, device=self.t.device)
            return self.t.data_ptr()

        self._cb = RESIZE_FN(_resize)
        self.buf = GsBuffer(self._cb, None)


# Workspaces that may be recycled from call to call: only used when the caller states that no backward pass will
# consume the buffers of this forward (see _RasterizeGaussians.forward).  Grow-only, one set per (device, stream):
# frames queued on different streams of one device are in flight at the same time and must not share scratch
# memory; a buffer that is replaced by a larger one is handed back to the caching allocator only after the stream
# that may still be reading it has been told (record_stream).
_WORKSPACE_POOL = {}


def _pooled_workspaces(dev):
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev)
    key = (dev.type, index, stream.cuda_stream)
    ws = _WORKSPACE_POOL.get(key)
    if ws is None:
        ws = _WORKSPACE_POOL[key] = (_Growable(dev, stream), _Growable(dev, stream), _Growable(dev, stream))
    return ws


def make_scene(*, P, sh_degree, sh_stride, width, height, tan_fovx, tan_fovy, scale_modifier, prefiltered, debug,
               background, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, viewmatrix,
               projmatrix, campos, tile_rows: Optional[Tuple[int, int]] = None, peer_out=None,
               extra_passes=None, downsample: int = 1, team_after: int = 0, shard_cull: bool = False, blend_split: int = 0) -> GsScene:
    """blend_split: > 0 = latency mode of the blend (GsScene.blend_split): list walks longer than that many batches are
    finished by a CTA as parallel segments merged associatively (pixels within ~1e-6, NOT bit-identical); 0 = off.
    shard_cull: tile-row shards only -- Gaussians that cannot reach the shard's rows are dropped before the
    per-Gaussian stage (GsScene.shard_cull; pixels unchanged, radii only written for the survivors).
    team_after: blend scheduling hint (GsScene.team_after): > 0 = hand-over threshold in batches (long list walks are
    finished by CTA teams; never changes a result), 0 = library default (off), < 0 = off.
    downsample: 2 = the blend epilogue stores the 2x2 box mean (all output images are (3,H/2,W/2)).
    peer_out: optional list of (peer-mapped) device pointers of (3,H,W) images the blend epilogue writes to.
    extra_passes: optional list of up to three (colors (P,3) CUDA tensor, out (3,H,W) CUDA tensor) pairs blended in
    the same list walk (the caller keeps the tensors alive)."""
    r0, r1 = tile_rows if tile_rows is not None else (0, 0)
    peers = list(peer_out) if peer_out else []
    if len(peers) > 8:
        raise ValueError("at most 8 peer images")
    arr = (C.c_void_p * 8)(*([int(p) for p in peers] + [None] * (8 - len(peers))))
    extras = list(extra_passes) if extra_passes else []
    if len(extras) > 3:
        raise ValueError("at most 3 extra colour passes")
    return GsScene(P, sh_degree, sh_stride, width, height, tan_fovx, tan_fovy, scale_modifier, int(bool(prefiltered)),
                   int(bool(debug)), int(r0), int(r1), _ptr(background), _ptr(means3D), _ptr(shs),
                   _ptr(colors_precomp), _ptr(opacities), _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp),
                   _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), len(peers), int(downsample), arr, len(extras),
                   int(team_after),
                   (C.c_void_p * 3)(*([_ptr(c) for c, _ in extras] + [None] * (3 - len(extras)))),
                   (C.c_void_p * 3)(*([_ptr(o) for _, o in extras] + [None] * (3 - len(extras)))),
                   int(bool(shard_cull)), int(blend_split))


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug, tile_rows: Optional[Tuple[int, int]] = None, out_color=None,
                        reuse_workspace: bool = False, downsample: int = 1):
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("diff_gaussian_rasterization (B200): means3D must be a CUDA tensor; no CPU path exists")
    L = lib()
    dev = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    with torch.cuda.device(dev):
        keep = [_prep(t) for t in (background, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos)]
        bg_, m3_, sh_, col_, op_, sc_, rot_, cov_, view_, proj_, cam_ = keep
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
        if downsample not in (1, 2) or (downsample == 2 and (H % 2 or W % 2)):
            raise ValueError("downsample must be 1 or 2 (2 needs an even raster size)")
        if out_color is None:
            out_color = torch.zeros((3, H // downsample, W // downsample), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        # reuse_workspace: the three scratch buffers come from a grow-only per-device pool instead of being
        # allocated per call; valid only if nothing (no backward) reads them after the next forward on this device
        geom, binning, img = _pooled_workspaces(dev) if reuse_workspace else (_Growable(dev), _Growable(dev), _Growable(dev))
        rendered = 0
        if P != 0:
            scene = make_scene(P=P, sh_degree=int(degree), sh_stride=M, width=W, height=H, tan_fovx=float(tan_fovx),
                               tan_fovy=float(tan_fovy), scale_modifier=float(scale_modifier), prefiltered=prefiltered,
                               debug=debug, background=bg_, means3D=m3_, shs=sh_, colors_precomp=col_, opacities=op_,
                               scales=sc_, rotations=rot_, cov3D_precomp=cov_, viewmatrix=view_, projmatrix=proj_,
                               campos=cam_, tile_rows=tile_rows, downsample=downsample)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rendered = _check(L.gs_forward(C.byref(scene), geom.buf, binning.buf, img.buf, out_color.data_ptr(),
                                           radii.data_ptr(), stream), "rasterize_gaussians")
    return int(rendered), out_color, radii, geom.t, binning.t, img.t


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
                                 geomBuffer, R, binningBuffer, imageBuffer, debug,
                                 tile_rows: Optional[Tuple[int, int]] = None, downsample: int = 1, grad_group=None,
                                 grad_reduce=None):
    """grad_group: tile-row sharded backward -- a torch.distributed process group over which the per-Gaussian
    partial gradients of the blend stage are summed (ONE all-reduce of a (P, 11) buffer) before the per-Gaussian
    stage runs, so that every rank returns the gradients of the WHOLE frame; dL_dout_color is the full (3,H,W)
    gradient on every rank, of which only this rank's tile rows are read.  grad_reduce: the same hook as a callable
    `f(partial)` that sums the flat (11 P) buffer over the shards in place (used instead of grad_group)."""
    L = lib()
    dev = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)) * downsample, int(dL_dout_color.size(2)) * downsample  # raster size
    M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
    with torch.cuda.device(dev):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        # the four arrays the blend stage accumulates into live in one buffer: a sharded backward sums them over the
        # ranks with a single collective
        partial = z(P * 11)
        dL_dmeans2D, dL_dconic = partial[:3 * P].view(P, 3), partial[3 * P:7 * P].view(P, 2, 2)
        dL_dopacity, dL_dcolors = partial[7 * P:8 * P].view(P, 1), partial[8 * P:].view(P, 3)
        dL_dmeans3D, dL_dcov3D = z(P, 3), z(P, 6)
        dL_dsh, dL_dscales, dL_drotations = z(P, M, 3), z(P, 3), z(P, 4)
        if P != 0:
            # note: unlike rasterize_points.cu:169,171 scales/rotations are made contiguous here too
            keep = [_prep(t) for t in (background, means3D, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                       projmatrix, campos, dL_dout_color)]
            bg_, m3_, sh_, col_, sc_, rot_, cov_, view_, proj_, cam_, dpix_ = keep
            radii_ = _prep(radii, torch.int32)
            scene = make_scene(P=P, sh_degree=int(degree), sh_stride=M, width=W, height=H, tan_fovx=float(tan_fovx),
                               tan_fovy=float(tan_fovy), scale_modifier=float(scale_modifier), prefiltered=False,
                               debug=debug, background=bg_, means3D=m3_, shs=sh_, colors_precomp=col_,
                               opacities=None, scales=sc_, rotations=rot_, cov3D_precomp=cov_, viewmatrix=view_,
                               projmatrix=proj_, campos=cam_, tile_rows=tile_rows, downsample=downsample)
            stream = torch.cuda.current_stream(dev).cuda_stream
            args = (C.byref(scene), int(R), _ptr(radii_), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                    _ptr(dpix_), _ptr(dL_dmeans2D), _ptr(dL_dconic), _ptr(dL_dopacity), _ptr(dL_dcolors),
                    _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations))
            if grad_group is None and grad_reduce is None:
                _check(L.gs_backward(*args, stream), "rasterize_gaussians_backward")
            else:
                _check(L.gs_backward_stage(*args, GS_BWD_BLEND, stream), "rasterize_gaussians_backward (blend stage)")
                if grad_reduce is not None:
                    grad_reduce(partial)
                else:
                    import torch.distributed as dist
                    dist.all_reduce(partial, group=grad_group)
                _check(L.gs_backward_stage(*args, GS_BWD_PREPROCESS, stream),
                       "rasterize_gaussians_backward (per-Gaussian stage)")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D, viewmatrix, projmatrix):
    if not means3D.is_cuda:
        raise RuntimeError("diff_gaussian_rasterization (B200): means3D must be a CUDA tensor; no CPU path exists")
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        with torch.cuda.device(means3D.device):
            m, v, p = _prep(means3D), _prep(viewmatrix), _prep(projmatrix)
            _check(lib().gs_mark_visible(P, _ptr(m), _ptr(v), _ptr(p), present.data_ptr(),
                                         torch.cuda.current_stream(means3D.device).cuda_stream), "mark_visible")
    return present


GS_VIEW_STRIDE = 48
GS_BWD_BLEND, GS_BWD_PREPROCESS = 1, 2


def projection_entries(fovx_deg: float, fovy_deg: float, znear: float = 0.01, zfar: float = 100.0):
    """The four non-trivial entries P[0][0], P[1][1], P[2][2], P[2][3] of getProjectionMatrix
    (simple_raw_render.py:50-69), in Python doubles as the reference computes them before storing fp32."""
    import math
    fovx, fovy = math.pi * fovx_deg / 180, math.pi * fovy_deg / 180  # np.pi * fov / 180, simple_raw_render.py:87
    top, right = math.tan(fovy / 2) * znear, math.tan(fovx / 2) * znear
    bottom, left = -top, -right
    return (2.0 * znear / (right - left), 2.0 * znear / (top - bottom), 1.0 * zfar / (zfar - znear),
            -(zfar * znear) / (zfar - znear))


def make_views(c2w: torch.Tensor, fovx_deg: float, fovy_deg: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gs_make_views: (N,4,4) camera-to-world matrices on the device -> (N, 48) fp32 rows holding viewmatrix [0:16],
    projmatrix [16:32] and campos [32:35] of every view, built by one kernel on the current stream."""
    if not c2w.is_cuda or c2w.dtype != torch.float32 or c2w.shape[-2:] != (4, 4):
        raise ValueError("make_views expects a float32 CUDA tensor of shape (N, 4, 4)")
    c2w = c2w.reshape(-1, 4, 4).contiguous()
    N = int(c2w.size(0))
    if out is None:
        out = torch.empty((N, GS_VIEW_STRIDE), dtype=torch.float32, device=c2w.device)
    p4 = (C.c_float * 4)(*projection_entries(fovx_deg, fovy_deg))
    with torch.cuda.device(c2w.device):
        _check(lib().gs_make_views(c2w.data_ptr(), N, p4, out.data_ptr(),
                                   torch.cuda.current_stream(c2w.device).cuda_stream), "make_views")
    return out


def decode_head(features: torch.Tensor, dc_rgb: torch.Tensor, primitives: torch.Tensor, *, scale_factor: float,
                xyz_offset: float, use_rotation=True, use_scale=True, use_opacity=True, use_offset=False,
                use_dc_offset=False, est_normal=False, normalize_normal=True, sh_ac_coeffs: int = 0,
                enable_opacity=True) -> dict:
    """gs_decode_head: the network head's feature rows -> rasterizer inputs in one kernel (models/model_v2.py:287-375,
    simple_raw_render.py:243-250,390-394).  Returns dict(means3D, rotations, scales, opacities, shs (P,1+K,3),
    sh_degree, normals or None); `shs` carries no zero padding -- render it with the returned sh_degree."""
    import math
    if not (features.is_cuda and dc_rgb.is_cuda and primitives.is_cuda):
        raise RuntimeError("decode_head: all tensors must be CUDA tensors; no CPU path exists")
    f32 = lambda t: t.to(torch.float32).contiguous()
    features, dc_rgb, primitives = f32(features), f32(dc_rgb), f32(primitives)
    P, Cc = int(features.shape[0]), int(features.shape[1])
    dev = features.device
    radius = math.sqrt(3) / scale_factor * 6  # np.sqrt(3) / self.scale_factor * 6, simple_raw_render.py:248
    lay = GsHeadLayout(Cc, int(use_rotation), int(use_scale), int(use_opacity), int(use_offset), int(use_dc_offset),
                       int(est_normal), int(normalize_normal), int(sh_ac_coeffs), int(enable_opacity), radius,
                       float(xyz_offset), float(scale_factor))
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = dict(means3D=e(P, 3), rotations=e(P, 4), scales=e(P, 3), opacities=e(P, 1), shs=e(P, 1 + sh_ac_coeffs, 3),
               normals=e(P, 3) if est_normal else None)
    with torch.cuda.device(dev):
        _check(lib().gs_decode_head(features.data_ptr(), dc_rgb.data_ptr(), primitives.data_ptr(), P, C.byref(lay),
                                    out["means3D"].data_ptr(), out["rotations"].data_ptr(), out["scales"].data_ptr(),
                                    out["opacities"].data_ptr(), out["shs"].data_ptr(),
                                    out["normals"].data_ptr() if est_normal else None,
                                    torch.cuda.current_stream(dev).cuda_stream), "decode_head")
    k = 1 + sh_ac_coeffs
    out["sh_degree"] = 0 if k < 4 else 1 if k < 9 else 2 if k < 16 else 3
    return out


_FETCH_DT = {"records": torch.float32, "sorted_idx": torch.int32, "sorted_key": torch.int32,
             "clamped": torch.uint8, "tiles_touched": torch.int32, "point_list": torch.int32, "ranges": torch.int32,
             "n_contrib": torch.int32, "final_T": torch.float32}


def fetch(name: str, scene: GsScene, geomBuffer, binningBuffer, imageBuffer, num_rendered: int) -> torch.Tensor:
    """Debug/test access to the library's internal arrays (host copy)."""
    cap = max(64, 48 * scene.P, 8 * scene.width * scene.height, 4 * num_rendered)
    host = torch.empty(cap, dtype=torch.uint8)
    n = _check(lib().gs_fetch(C.byref(scene), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                              int(num_rendered), name.encode(), host.data_ptr(), cap,
                              torch.cuda.current_stream().cuda_stream), f"fetch({name})")
    return host[:n].view(_FETCH_DT[name]).clone()
