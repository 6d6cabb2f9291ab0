"""ctypes front-end for the CPU oracle (oracle/gs_oracle.c) and the reference-library shim
(oracle/_ref/libgs_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

`Oracle(bits).forward(...)` runs the whole reference pipeline on the CPU in the order of
dgr/cuda_rasterizer/rasterizer_impl.cu:198-336 and returns every intermediate array;
`Oracle(bits).backward(...)` follows rasterizer_impl.cu:340-434.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p


def build(force: bool = False) -> None:
    """Compile the C oracle (and, when /root/reference is present, the reference shim)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in ("liboracle32.so", "liboracle64.so"))
    ref_present = os.path.isdir("/root/reference/diff-gaussian-rasterization/cuda_rasterizer")
    if ref_present and not os.path.exists(os.path.join(_HERE, "_ref", "libgs_ref.so")):
        need = True
    if need:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


class Oracle:
    """CPU restatement.  bits=32 (parity checker / cpu_baseline) or 64 (gradient ground truth)."""

    def __init__(self, bits: int = 32):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, f"liboracle{bits}.so"))
        self.real = np.float32 if bits == 32 else np.float64
        assert self.lib.gso_real_bytes() == np.dtype(self.real).itemsize
        self.lib.gso_inclusive_scan.restype = C.c_uint32
        self.lib.gso_higher_msb.restype = C.c_uint32
        self.threads = int(self.lib.gso_num_threads())

    # -- forward -------------------------------------------------------------------------------------
    def forward(self, *, means3D, opacities, W: int, H: int, viewmatrix, projmatrix, campos, bg,
                tanfovx: float, tanfovy: float, sh_degree: int = 0, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None, scale_modifier: float = 1.0,
                prefiltered: bool = False, stop_after: Optional[str] = None) -> Dict[str, np.ndarray]:
        r = self.real
        means3D = _f32(means3D).reshape(-1, 3) if _f32(means3D) is not None else np.zeros((0, 3), np.float32)
        P = means3D.shape[0]
        shs, colors_precomp = _f32(shs), _f32(colors_precomp)
        scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
        opac = _f32(opacities)
        opac = opac.reshape(-1) if opac is not None else np.zeros((0,), np.float32)
        view, proj = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16)
        cam, bgc = _f32(campos).reshape(3), _f32(bg).reshape(3)
        M = 0 if shs is None else int(shs.reshape(P, -1, 3).shape[1])
        gx, gy = (W + 15) // 16, (H + 15) // 16
        out: Dict[str, np.ndarray] = {}
        out["radii"] = np.zeros(P, np.int32)
        out["means2D"] = np.zeros((P, 2), r)
        out["depths"] = np.zeros(P, r)
        out["cov3D"] = np.zeros((P, 6), r)
        out["rgb"] = np.zeros((P, 3), r)
        out["conic_opacity"] = np.zeros((P, 4), r)
        out["clamped"] = np.zeros((P, 3), np.uint8)
        out["tiles_touched"] = np.zeros(P, np.uint32)
        rc = self.lib.gso_preprocess(
            C.c_int(P), C.c_int(sh_degree), C.c_int(M), _ptr(means3D), _ptr(scales), C.c_float(scale_modifier),
            _ptr(rotations), _ptr(opac), _ptr(shs), _ptr(cov3D_precomp), _ptr(colors_precomp), _ptr(view), _ptr(proj),
            _ptr(cam), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(prefiltered)),
            _ptr(out["radii"]), _ptr(out["means2D"]), _ptr(out["depths"]), _ptr(out["cov3D"]), _ptr(out["rgb"]),
            _ptr(out["conic_opacity"]), _ptr(out["clamped"]), _ptr(out["tiles_touched"]))
        if rc != 0:
            raise RuntimeError("oracle: point culled although prefiltered is set (reference traps here)")
        if stop_after == "preprocess":
            return out
        out["point_offsets"] = np.zeros(P, np.uint32)
        R = int(self.lib.gso_inclusive_scan(C.c_int(P), _ptr(out["tiles_touched"]), _ptr(out["point_offsets"]))) if P else 0
        out["num_rendered"] = R
        keys_u = np.zeros(R, np.uint64)
        vals_u = np.zeros(R, np.uint32)
        self.lib.gso_duplicate_with_keys(C.c_int(P), C.c_int(W), C.c_int(H), _ptr(out["means2D"]), _ptr(out["depths"]),
                                         _ptr(out["point_offsets"]), _ptr(out["radii"]), _ptr(keys_u), _ptr(vals_u))
        bit = int(self.lib.gso_higher_msb(C.c_uint32(gx * gy)))
        out["point_list_keys"] = np.zeros(R, np.uint64)
        out["point_list"] = np.zeros(R, np.uint32)
        self.lib.gso_sort_pairs(C.c_uint32(R), C.c_int(32 + bit), _ptr(keys_u), _ptr(vals_u),
                                _ptr(out["point_list_keys"]), _ptr(out["point_list"]))
        out["ranges"] = np.zeros((gx * gy, 2), np.uint32)
        self.lib.gso_tile_ranges(C.c_uint32(R), C.c_int(gx * gy), _ptr(out["point_list_keys"]), _ptr(out["ranges"]))
        if stop_after == "binning":
            return out
        colors = out["rgb"] if colors_precomp is None else np.ascontiguousarray(colors_precomp.reshape(P, 3), dtype=r)
        out["colors_used"] = colors
        out["color"] = np.zeros((3, H, W), r)
        out["final_T"] = np.zeros((H, W), r)
        out["n_contrib"] = np.zeros((H, W), np.uint32)
        self.lib.gso_blend_forward(C.c_int(W), C.c_int(H), _ptr(out["ranges"]), _ptr(out["point_list"]),
                                   _ptr(out["means2D"]), _ptr(colors), _ptr(out["conic_opacity"]), _ptr(bgc),
                                   _ptr(out["color"]), _ptr(out["final_T"]), _ptr(out["n_contrib"]))
        return out

    # -- backward ------------------------------------------------------------------------------------
    def backward(self, fwd: Dict[str, np.ndarray], dL_dcolor_img, *, means3D, W: int, H: int, viewmatrix, projmatrix,
                 campos, bg, tanfovx: float, tanfovy: float, sh_degree: int = 0, shs=None, colors_precomp=None,
                 scales=None, rotations=None, cov3D_precomp=None, scale_modifier: float = 1.0) -> Dict[str, np.ndarray]:
        r = self.real
        means3D = _f32(means3D).reshape(-1, 3)
        P = means3D.shape[0]
        shs, colors_precomp = _f32(shs), _f32(colors_precomp)
        scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
        view, proj = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16)
        cam, bgc = _f32(campos).reshape(3), _f32(bg).reshape(3)
        M = 0 if shs is None else int(shs.reshape(P, -1, 3).shape[1])
        dpix = np.ascontiguousarray(np.asarray(dL_dcolor_img), dtype=r).reshape(3, H, W)
        g = {
            "dL_dmeans2D": np.zeros((P, 3), r), "dL_dconic": np.zeros((P, 2, 2), r),
            "dL_dopacity": np.zeros((P, 1), r), "dL_dcolors": np.zeros((P, 3), r),
            "dL_dmeans3D": np.zeros((P, 3), r), "dL_dcov3D": np.zeros((P, 6), r),
            "dL_dsh": np.zeros((P, M, 3), r), "dL_dscales": np.zeros((P, 3), r),
            "dL_drotations": np.zeros((P, 4), r),
        }
        self.lib.gso_blend_backward(C.c_int(W), C.c_int(H), _ptr(fwd["ranges"]), _ptr(fwd["point_list"]),
                                    _ptr(fwd["means2D"]), _ptr(fwd["colors_used"]), _ptr(fwd["conic_opacity"]),
                                    _ptr(bgc), _ptr(fwd["final_T"]), _ptr(fwd["n_contrib"]), _ptr(dpix),
                                    _ptr(g["dL_dmeans2D"]), _ptr(g["dL_dconic"]), _ptr(g["dL_dopacity"]),
                                    _ptr(g["dL_dcolors"]))
        cov3D = fwd["cov3D"] if cov3D_precomp is None else np.ascontiguousarray(cov3D_precomp.reshape(P, 6), dtype=r)
        self.lib.gso_preprocess_backward(
            C.c_int(P), C.c_int(sh_degree), C.c_int(M), _ptr(means3D), _ptr(fwd["radii"]), _ptr(shs),
            _ptr(fwd["clamped"]), _ptr(scales), _ptr(rotations), C.c_float(scale_modifier), _ptr(cov3D), _ptr(view),
            _ptr(proj), _ptr(cam), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy),
            _ptr(g["dL_dmeans2D"]), _ptr(g["dL_dconic"]), _ptr(g["dL_dcolors"]), _ptr(g["dL_dmeans3D"]),
            _ptr(g["dL_dcov3D"]), _ptr(g["dL_dsh"]), _ptr(g["dL_dscales"]), _ptr(g["dL_drotations"]))
        return g

    def mark_visible(self, means3D, viewmatrix, projmatrix) -> np.ndarray:
        means3D = _f32(means3D).reshape(-1, 3)
        out = np.zeros(means3D.shape[0], np.uint8)
        self.lib.gso_mark_visible(C.c_int(means3D.shape[0]), _ptr(means3D), _ptr(_f32(viewmatrix).reshape(16)),
                                  _ptr(_f32(projmatrix).reshape(16)), _ptr(out))
        return out.astype(bool)


# ---------------------------------------------------------------------------------------------------
class ReferenceCUDA:
    """The UNMODIFIED reference CUDA rasterizer (oracle/_ref/libgs_ref.so) driven through ref_shim.cu.
    Needs a GPU.  Takes/returns torch CUDA tensors; plays the role of the reference's `_C` module."""

    SO = os.path.join(_HERE, "_ref", "libgs_ref.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.SO)

    def __init__(self):
        import torch  # noqa: F401
        self.lib = C.CDLL(self.SO)
        self.lib.gsref_create.restype = _vp
        self.lib.gsref_fetch.restype = C.c_longlong
        self.state = _vp(self.lib.gsref_create())
        self._meta = None

    def __del__(self):
        try:
            self.lib.gsref_destroy(self.state)
        except Exception:
            pass

    @staticmethod
    def _dp(t):
        if t is None or t.numel() == 0:
            return None
        assert t.is_cuda and t.is_contiguous()
        return _vp(t.data_ptr())

    def forward(self, *, means3D, opacities, W, H, viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy, sh_degree=0,
                shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
                prefiltered=False, debug=False):
        import torch
        P = means3D.shape[0]
        c = lambda t: None if t is None else t.contiguous().float()
        a = dict(means3D=c(means3D), opacities=c(opacities), shs=c(shs), colors_precomp=c(colors_precomp),
                 scales=c(scales), rotations=c(rotations), cov3D_precomp=c(cov3D_precomp), view=c(viewmatrix),
                 proj=c(projmatrix), campos=c(campos), bg=c(bg))
        M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
        color = torch.zeros((3, H, W), dtype=torch.float32, device=means3D.device)
        radii = torch.zeros((P,), dtype=torch.int32, device=means3D.device)
        # The reference launches on the legacy default stream, which is torch's default stream: no synchronisation is
        # needed (the reference's own binding has none, rasterize_points.cu:35-115).  Only a caller on a side stream
        # has to be ordered in front of it.
        if torch.cuda.current_stream(means3D.device) != torch.cuda.default_stream(means3D.device):
            torch.cuda.current_stream(means3D.device).synchronize()
        R = self.lib.gsref_forward(self.state, C.c_int(P), C.c_int(sh_degree), C.c_int(M), self._dp(a["bg"]), C.c_int(W),
                                   C.c_int(H), self._dp(a["means3D"]), self._dp(a["shs"]), self._dp(a["colors_precomp"]),
                                   self._dp(a["opacities"]), self._dp(a["scales"]), C.c_float(scale_modifier),
                                   self._dp(a["rotations"]), self._dp(a["cov3D_precomp"]), self._dp(a["view"]),
                                   self._dp(a["proj"]), self._dp(a["campos"]), C.c_float(tanfovx), C.c_float(tanfovy),
                                   C.c_int(int(prefiltered)), self._dp(color), self._dp(radii), C.c_int(int(debug)))
        if R < 0:
            raise RuntimeError("reference forward failed")
        self._meta = dict(a=a, P=P, M=M, W=W, H=H, D=sh_degree, tanfovx=tanfovx, tanfovy=tanfovy,
                          scale_modifier=scale_modifier, radii=radii)
        return color, radii, R

    def backward(self, dL_dcolor, sync=True):
        import torch
        m = self._meta
        a, P, M = m["a"], m["P"], m["M"]
        dev = a["means3D"].device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        g = dict(dL_dmeans2D=z(P, 3), dL_dconic=z(P, 2, 2), dL_dopacity=z(P, 1), dL_dcolors=z(P, 3),
                 dL_dmeans3D=z(P, 3), dL_dcov3D=z(P, 6), dL_dsh=z(P, M, 3), dL_dscales=z(P, 3), dL_drotations=z(P, 4))
        dpix = dL_dcolor.contiguous().float()
        if torch.cuda.current_stream(dev) != torch.cuda.default_stream(dev):
            torch.cuda.current_stream(dev).synchronize()
        rc = self.lib.gsref_backward(
            self.state, C.c_int(P), C.c_int(m["D"]), C.c_int(M), self._dp(a["bg"]), C.c_int(m["W"]), C.c_int(m["H"]),
            self._dp(a["means3D"]), self._dp(a["shs"]), self._dp(a["colors_precomp"]), self._dp(a["scales"]),
            C.c_float(m["scale_modifier"]), self._dp(a["rotations"]), self._dp(a["cov3D_precomp"]), self._dp(a["view"]),
            self._dp(a["proj"]), self._dp(a["campos"]), C.c_float(m["tanfovx"]), C.c_float(m["tanfovy"]),
            self._dp(m["radii"]), self._dp(dpix), self._dp(g["dL_dmeans2D"]), self._dp(g["dL_dconic"]),
            self._dp(g["dL_dopacity"]), self._dp(g["dL_dcolors"]), self._dp(g["dL_dmeans3D"]), self._dp(g["dL_dcov3D"]),
            self._dp(g["dL_dsh"]), self._dp(g["dL_dscales"]), self._dp(g["dL_drotations"]), C.c_int(0))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        if sync:
            torch.cuda.synchronize()
        return g

    _DT = dict(depths=np.float32, means2D=np.float32, cov3D=np.float32, conic_opacity=np.float32, rgb=np.float32,
               clamped=np.uint8, tiles_touched=np.uint32, point_offsets=np.uint32, point_list=np.uint32,
               point_list_keys=np.uint64, ranges=np.uint32, n_contrib=np.uint32, accum_alpha=np.float32)

    def fetch(self, name: str) -> np.ndarray:
        m = self._meta
        cap = max(64, 8 * max(m["P"] * 6, m["W"] * m["H"], int(self.lib.gsref_num_rendered(self.state))))
        buf = np.zeros(cap, np.uint8)
        n = int(self.lib.gsref_fetch(self.state, name.encode(), _ptr(buf), C.c_longlong(cap)))
        if n < 0:
            raise KeyError(name)
        return buf[:n].view(self._DT[name]).copy()
