"""Developer tool (GPU box): one case of tools/fuzz_vs_reference.py in detail -- every gradient of this library and of the
live reference library against the fp64 CPU oracle, and each library against itself on a second run (both add with
atomics in an order that changes from run to run).  usage: fuzz_case.py <case index> [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from oracle.oracle import Oracle, ReferenceCUDA  # noqa: E402

target, seed = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 2024
dev = torch.device("cuda:0")
rng = np.random.default_rng(seed)
orbit = scenes.orbit_c2w(12)
for case in range(target + 1):  # replay the generator of fuzz_vs_reference.run
    P = int(rng.choice([1, 7, 100, 1500, 8000, 30000]))
    W, H = int(rng.integers(8, 700)), int(rng.integers(8, 500))
    D = int(rng.integers(0, 4))
    M = (D + 1) ** 2 + int(rng.choice([0, 0, 1, 5]))
    use_sh = bool(rng.random() < 0.7)
    use_cov = bool(rng.random() < 0.25)
    mod = float(rng.choice([1.0, 1.0, 0.5, 1.7]))
    spread, scale = float(rng.choice([0.4, 0.8, 2.0])), float(rng.choice([0.01, 0.05, 0.2]))
    vi, fov = int(rng.integers(0, 12)), float(rng.choice([30.0, 45.0, 60.0]))
    ysc = float(rng.choice([1.0, 0.75]))
    bg_np = rng.random(3)
    col_np = None if use_sh else rng.random((P, 3))
    A = rng.standard_normal((P, 3, 3)).astype(np.float32) * 0.05 if use_cov else None
cl = scenes.tiny_cloud(P, seed=1000 + target, sh_degree=D, M=M, spread=spread, scale=scale)
v = scenes.make_view(orbit[vi], W, H, fov_deg=fov)
tanx, tany = v.tanfovx, v.tanfovy * ysc
t = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(dev)
bg = t(bg_np)
means, opac = cl["means3D"].to(dev), cl["opacities"].to(dev)
kw = {}
if use_sh:
    kw["shs"] = cl["shs"].to(dev)
else:
    kw["colors_precomp"] = t(col_np)
if use_cov:
    S = A @ A.transpose(0, 2, 1)
    kw["cov3D_precomp"] = t(np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1))
else:
    kw["scales"], kw["rotations"] = cl["scales"].to(dev), cl["rotations"].to(dev)
wgt = torch.from_numpy(np.random.default_rng(target).standard_normal((3, H, W)).astype(np.float32)).to(dev)
names = dict(means3D="dL_dmeans3D", opacities="dL_dopacity", shs="dL_dsh", colors_precomp="dL_dcolors",
             scales="dL_dscales", rotations="dL_drotations", cov3D_precomp="dL_dcov3D")
print(f"case {target}: P={P} {W}x{H} D={D} M={M} sh={use_sh} cov={use_cov} mod={mod} spread={spread} scale={scale}")


def ours():
    leaves = {k: x.clone().requires_grad_(True) for k, x in dict(means3D=means, opacities=opac, **kw).items()}
    rs = GaussianRasterizationSettings(H, W, tanx, tany, bg, mod, t(v.viewmatrix), t(v.projmatrix), D, t(v.campos), False, False)
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, _ = GaussianRasterizer(rs)(leaves["means3D"], m2, leaves["opacities"], shs=leaves.get("shs"),
                                      colors_precomp=leaves.get("colors_precomp"), scales=leaves.get("scales"),
                                      rotations=leaves.get("rotations"), cov3D_precomp=leaves.get("cov3D_precomp"))
    (color * wgt).sum().backward()
    out = {names[k]: x.grad.detach().cpu().numpy().astype(np.float64) for k, x in leaves.items()}
    out["dL_dmeans2D"] = m2.grad.detach().cpu().numpy().astype(np.float64)
    return out


def theirs():
    ref = ReferenceCUDA()
    ref.forward(means3D=means, opacities=opac, W=W, H=H, viewmatrix=t(v.viewmatrix), projmatrix=t(v.projmatrix),
                campos=t(v.campos), bg=bg, tanfovx=tanx, tanfovy=tany, sh_degree=D, scale_modifier=mod, **kw)
    g = ref.backward(wgt)
    return {k: x.detach().cpu().numpy().astype(np.float64) for k, x in g.items()}


o1, o2, r1, r2 = ours(), ours(), theirs(), theirs()
o64 = Oracle(64)
npk = {k: x.cpu().numpy() for k, x in kw.items()}
f = o64.forward(means3D=means.cpu().numpy(), opacities=opac.cpu().numpy(), W=W, H=H, viewmatrix=v.viewmatrix,
                projmatrix=v.projmatrix, campos=v.campos, bg=bg_np.astype(np.float32), tanfovx=tanx, tanfovy=tany,
                sh_degree=D, scale_modifier=mod, **npk)
g64 = o64.backward(f, wgt.cpu().numpy(), means3D=means.cpu().numpy(), W=W, H=H, viewmatrix=v.viewmatrix,
                   projmatrix=v.projmatrix, campos=v.campos, bg=bg_np.astype(np.float32), tanfovx=tanx, tanfovy=tany,
                   sh_degree=D, scale_modifier=mod, **npk)
rel = lambda a, b: float(np.abs(a.reshape(-1) - b.reshape(-1)).max() / (np.abs(b).max() + 1e-30))
print(f"{'gradient':16s} {'ours-fp64':>10s} {'ref-fp64':>10s} {'ours-ref':>10s} {'ours-ours':>10s} {'ref-ref':>10s}")
for k in o1:
    if k in g64 and k in r1:
        print(f"{k:16s} {rel(o1[k], g64[k]):10.2e} {rel(r1[k], g64[k]):10.2e} {rel(o1[k], r1[k]):10.2e} {rel(o1[k], o2[k]):10.2e} {rel(r1[k], r2[k]):10.2e}")
