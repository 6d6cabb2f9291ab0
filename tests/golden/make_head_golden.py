"""Generates tests/golden/head_decode.npz by EXECUTING the reference's own head-decode statements in the build
container: the body of the network's `forward` from `dc_color_rgb = [...]` to its `return`
(models/model_v2.py:286-375), `RGB2SH` (models/sh_utils.py), `pcgc_rescale` (simple_raw_render.py:71-75) and the
`radius` / `scales` statements of `PCML_Render._rasterize` (simple_raw_render.py:248-249).

models/model_v2.py cannot be imported here (MinkowskiEngine is absent), so the statements are cut out of the file with
`ast` and executed as they are on plain tensors; `self` is a stand-in carrying `args` and `default_quaternion`.
Nothing of the reference is copied into the repository: only the numbers it produces (CPU tensors: this container has
no GPU, so divisions by a scalar are true divisions here; see oracle/head.py).

    python tests/golden/make_head_golden.py      # build container only
"""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CONFIGS = {  # the reference's args; "shipped" = use_rotation/scale/opacity, DC from the input colours, sh_deg 1, no SH features
    "shipped": dict(use_rotation=True, use_scale=True, use_opacity=True, use_offset=False, use_dc_offset=False,
                    est_normal=False, normalize_normal=True, sh_deg=1, sh_feat_deg=0, C=8),
    "all_heads": dict(use_rotation=True, use_scale=True, use_opacity=True, use_offset=True, use_dc_offset=True,
                      est_normal=True, normalize_normal=True, sh_deg=1, sh_feat_deg=1, C=26),
    "bare": dict(use_rotation=False, use_scale=False, use_opacity=False, use_offset=False, use_dc_offset=False,
                 est_normal=False, normalize_normal=True, sh_deg=0, sh_feat_deg=0, C=3),
    "raw_normal": dict(use_rotation=False, use_scale=True, use_opacity=True, use_offset=True, use_dc_offset=False,
                       est_normal=True, normalize_normal=False, sh_deg=1, sh_feat_deg=0, C=10),
}


def inputs(name, P=600):
    cfg = CONFIGS[name]
    rng = np.random.default_rng(sorted(CONFIGS).index(name) + 900)
    feat = (0.6 * rng.standard_normal((P, cfg["C"]))).astype(np.float32)
    feat[::37] = 0.0
    feat[5::41, :] *= 4.0  # some scales clamp at 0, some opacities at 0 / 1
    rgb = rng.random((P, 3)).astype(np.float32)
    prim = rng.integers(0, 1024, (P, 3)).astype(np.float32)
    return feat, rgb, prim


def lines_of(src, stmts):
    """Whole source lines of the statements, dedented (get_source_segment only strips the first line's indentation)."""
    import textwrap
    rows = src.split("\n")
    return textwrap.dedent("\n".join("\n".join(rows[s.lineno - 1:s.end_lineno]) for s in stmts))


def head_statements():
    src = open(os.path.join(REF, "models", "model_v2.py")).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == "forward":
            body = [s for s in node.body if not isinstance(s, ast.Return)]
            start = next((i for i, s in enumerate(body) if isinstance(s, ast.Assign) and isinstance(s.targets[0], ast.Name)
                          and s.targets[0].id == "dc_color_rgb"), None)
            if start is not None and any("default_quaternion" in ast.get_source_segment(src, s) for s in body[start:]):
                return lines_of(src, body[start:])
    raise RuntimeError("head statements not found")


def cut_function(path, name):
    src = open(path).read()
    return next(ast.get_source_segment(src, n) for n in ast.parse(src).body
                if isinstance(n, ast.FunctionDef) and n.name == name)


def rasterize_glue():
    src = open(os.path.join(REF, "simple_raw_render.py")).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == "PCML_Render":
            fn = next(n for n in node.body if isinstance(n, ast.FunctionDef) and n.name == "_rasterize")
            want = [s for s in ast.walk(fn) if isinstance(s, ast.Assign) and isinstance(s.targets[0], ast.Name)
                    and s.targets[0].id in ("radius", "scales")]
            want.sort(key=lambda s: s.lineno)
            return lines_of(src, want)
    raise RuntimeError("_rasterize not found")


def main():
    spec = importlib.util.spec_from_file_location("ref_sh_utils", os.path.join(REF, "models", "sh_utils.py"))
    sh_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh_utils)
    head, glue = head_statements(), rasterize_glue()
    rescale = cut_function(os.path.join(REF, "simple_raw_render.py"), "pcgc_rescale")
    out = {}
    for name, cfg in CONFIGS.items():
        feat, rgb, prim = inputs(name)
        args = SimpleNamespace(**{k: v for k, v in cfg.items() if k != "C"})
        self = SimpleNamespace(args=args, default_quaternion=torch.tensor([[1, 0, 0, 0]], dtype=torch.float32),
                               scale_factor=448, offset=512)
        ns = {"torch": torch, "np": np, "RGB2SH": sh_utils.RGB2SH, "self": self,
              "decoded_color_feature": [torch.from_numpy(feat)], "decoded_primitives": [torch.from_numpy(prim)],
              # the statements start by taking the last three columns of the input features as the DC colours
              "dc_color_rgb": [torch.cat([torch.zeros(len(rgb), 2), torch.from_numpy(rgb)], 1)]}
        exec(head, ns)
        exec(rescale, ns)
        ns["i"] = 0
        ns["decoded_s"] = ns["decoded_s"]
        exec(glue, ns)
        means = ns["pcgc_rescale"](ns["decoded_primitives_aug"][0].float(), self.offset, self.scale_factor)
        out[f"{name}.means3D"] = means.numpy()
        out[f"{name}.rotations"] = ns["decoded_r"][0].contiguous().numpy()
        out[f"{name}.scales"] = ns["scales"].numpy()
        out[f"{name}.opacities"] = ns["decoded_o"][0].numpy()
        out[f"{name}.shs"] = ns["decoded_sh"][0].numpy()
        if ns["decoded_n"] is not None:
            out[f"{name}.normals"] = ns["decoded_n"][0].numpy()
    np.savez_compressed(os.path.join(HERE, "head_decode.npz"), **out)
    print("wrote head_decode.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container)")
    main()
