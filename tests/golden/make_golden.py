"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference CUDA rasterizer
(oracle/_ref/libgs_ref.so = the reference's own forward.cu / backward.cu / rasterizer_impl.cu compiled for
sm_100a, see oracle/Makefile) on a B200.  The reference ships no tests or stored renders, so
these files are what pins both the CPU oracle and the CUDA product to the reference's actual behaviour.

Run on the GPU box:   python tests/golden/make_golden.py          (writes gpurun_out/golden/*.npz)
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  Also (in the build container only)
`python tests/golden/make_golden.py --orbit` extracts the 12-view orbit of the reference's
validate/temp_state_dict.pt into orbit12_H_c2w.npy.

Inputs are NOT stored: they are regenerated from seeds by golden_cases() below (numpy Generator streams are
stable); a checksum of the inputs is stored and verified by the tests.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-pcloud-render_b200"))
import scenes  # noqa: E402


def golden_cases():
    """name -> dict(cloud kwargs for the rasterizer, view, bg).  Small enough for the CPU oracle in < 1 s each."""
    orbit = scenes.orbit_c2w(12)
    cases = {}
    c = scenes.tiny_cloud(3000, seed=1, sh_degree=3)
    cases["tiny_sh3"] = dict(cloud=c, view=scenes.make_view(orbit[1], 200, 136), bg=[0.2, 0.4, 0.6])
    c = scenes.tiny_cloud(5000, seed=2, sh_degree=2, depth_ties=True)
    cases["tiny_ties"] = dict(cloud=c, view=scenes.make_view(orbit[0], 256, 256), bg=[1.0, 1.0, 1.0])
    c = scenes.human_cloud(6000, scale_factor=256.0, seed=3)  # M=13 stride with sh_degree 1, opacity 1 (pcrender shape)
    cases["human_m13"] = dict(cloud=c, view=scenes.make_view(orbit[4], 160, 240), bg=[1.0, 1.0, 1.0])
    # precomputed colours + precomputed 3D covariance branch (forward.cu:208,244)
    c = scenes.tiny_cloud(2000, seed=5, sh_degree=0)
    rng = np.random.default_rng(11)
    A = rng.standard_normal((2000, 3, 3)).astype(np.float32) * 0.04
    S = A @ A.transpose(0, 2, 1)
    cov6 = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1).astype(np.float32)
    c2 = dict(means3D=c["means3D"], opacities=c["opacities"], colors_precomp=torch.from_numpy(
        rng.uniform(0, 1, (2000, 3)).astype(np.float32)), cov3D_precomp=torch.from_numpy(cov6), sh_degree=0)
    cases["precomp"] = dict(cloud=c2, view=scenes.make_view(orbit[7], 128, 96), bg=[0.0, 0.0, 0.0])
    # points behind / near the camera, off-screen splats, zero scales: culling edge cases
    c = scenes.tiny_cloud(4000, seed=9, sh_degree=1, spread=2.5, scale=0.08)
    c["scales"][::7] = 0.0
    cases["cull_edges"] = dict(cloud=c, view=scenes.make_view(orbit[2], 176, 144), bg=[0.5, 0.5, 0.5])
    # SH degree 0 read from a packed (P,1,3) array: the form gs_decode_head emits (forward.cu:30 + `+ 0.5f` only)
    c = scenes.human_cloud(5000, scale_factor=200.0, seed=6, opacity="uniform")
    c = dict(c, shs=c["shs"][:, :1].contiguous(), sh_degree=0)
    cases["sh0_packed"] = dict(cloud=c, view=scenes.make_view(orbit[9], 192, 128), bg=[0.1, 0.9, 0.3])
    return cases


def rast_kwargs(case):
    c, v = case["cloud"], case["view"]
    kw = dict(means3D=c["means3D"], opacities=c["opacities"], W=v.image_width, H=v.image_height,
              viewmatrix=v.viewmatrix, projmatrix=v.projmatrix, campos=v.campos, bg=np.asarray(case["bg"], np.float32),
              tanfovx=v.tanfovx, tanfovy=v.tanfovy, sh_degree=c["sh_degree"])
    for k in ("shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"):
        if k in c:
            kw[k] = c[k]
    return kw


def input_checksum(kw) -> str:
    h = hashlib.sha256()
    for k in sorted(kw):
        v = kw[k]
        if torch.is_tensor(v):
            v = v.numpy()
        h.update(k.encode())
        h.update(np.ascontiguousarray(np.asarray(v, dtype=np.float32)).tobytes())
    return h.hexdigest()


def loss_weights(shape, seed=7):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


def main():
    if "--orbit" in sys.argv:
        d = torch.load("/root/reference/validate/temp_state_dict.pt", weights_only=False)
        np.save(os.path.join(HERE, "orbit12_H_c2w.npy"), d["H_c2w"][0].numpy().astype(np.float32))
        print("wrote orbit12_H_c2w.npy")
        return
    from oracle.oracle import ReferenceCUDA
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    ref = ReferenceCUDA()
    for name, case in golden_cases().items():
        kw = rast_kwargs(case)
        tk = {k: (torch.as_tensor(v).to(dev) if not isinstance(v, (int, float)) else v) for k, v in kw.items()}
        color, radii, R = ref.forward(**tk)
        w = loss_weights(tuple(color.shape))
        g = ref.backward(torch.from_numpy(w).to(dev))
        rec = dict(color=color.cpu().numpy(), radii=radii.cpu().numpy(), num_rendered=np.int64(R),
                   n_contrib=ref.fetch("n_contrib"), point_list=ref.fetch("point_list"),
                   ranges=ref.fetch("ranges").reshape(-1, 2), checksum=np.array(input_checksum(kw)))
        for k, v in g.items():
            rec[k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "R", R, "visible", int((radii > 0).sum()), "color mean", float(color.mean()))


if __name__ == "__main__":
    main()
