"""CPU tests: the oracle (oracle/gs_oracle.c) against the reference's golden vectors and against analytic cases.
The golden vectors were produced by the UNMODIFIED reference CUDA library on a B200 (tests/golden/make_golden.py)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import scenes
from conftest import GOLDEN_NAMES, ROOT
from make_golden import loss_weights

PIX_TOL = 1e-4  # BASELINE.json north_star: rendered pixels within 1e-4 max abs (fp32)
GRAD_RTOL = 1e-3  # gradients within rtol 1e-3 (reference backward is atomics-order dependent)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_forward_matches_reference_golden(name, golden, oracle32):
    _, kw, g = golden[name]
    f = oracle32.forward(**kw)
    assert f["num_rendered"] == int(g["num_rendered"])
    assert np.array_equal(f["radii"], g["radii"])
    assert np.abs(f["color"] - g["color"]).max() <= PIX_TOL
    # same per-tile lists in the same order (stable (tile|depth) sort, ties by Gaussian index)
    assert np.array_equal(f["point_list"], g["point_list"])
    ne = g["ranges"][:, 0] != g["ranges"][:, 1]
    assert np.array_equal(f["ranges"][ne], g["ranges"][ne])
    assert np.array_equal(f["n_contrib"].ravel(), g["n_contrib"])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_backward_matches_reference_golden(name, golden, oracle32):
    _, kw, g = golden[name]
    f = oracle32.forward(**kw)
    gr = oracle32.backward(f, loss_weights(g["color"].shape), **{k: v for k, v in kw.items() if k != "opacities"})
    for k, v in gr.items():
        ref = g[k].reshape(v.shape)
        if ref.size == 0:
            continue
        assert np.abs(v - ref).max() <= GRAD_RTOL * (np.abs(ref).max() + 1e-12), k


def test_orbit_matches_reference_fixture():
    """scenes.orbit_c2w(12) == H_c2w of the reference's validate/temp_state_dict.pt (generate_cam, r=3, 12 views)."""
    fix = np.load(os.path.join(ROOT, "tests", "golden", "orbit12_H_c2w.npy"))
    assert np.abs(scenes.orbit_c2w(12) - fix).max() < 1e-5
    v = scenes.make_view(fix[0], 512, 512)
    assert v.tanfovx == pytest.approx(1.0)  # tan(45 deg): the full-angle quirk (simple_raw_render.py:101)
    # intrinsics of the fixture: f = 256 / tan(22.5 deg) = 618.04 -> projection uses the half angle
    assert scenes.projection_matrix(0.01, 100, math.pi / 4, math.pi / 4)[0, 0] == pytest.approx(618.0387 / 256, rel=1e-5)


def _single(oracle, mean, scale, opacity, rgb, W=64, H=64, bg=(0.0, 0.0, 0.0), extra=None):
    """One (or a few) isotropic Gaussians in front of an identity camera looking down +z."""
    view = np.eye(4, dtype=np.float32)
    fov = math.pi / 2
    P = scenes.projection_matrix(0.01, 100.0, fov, fov)
    proj = np.ascontiguousarray((view @ P.T).astype(np.float32))
    means = np.asarray(mean, np.float32).reshape(-1, 3)
    n = means.shape[0]
    kw = dict(means3D=means, opacities=np.asarray(opacity, np.float32).reshape(n, 1), W=W, H=H, viewmatrix=view,
              projmatrix=proj, campos=np.zeros(3, np.float32), bg=np.asarray(bg, np.float32), tanfovx=1.0, tanfovy=1.0,
              sh_degree=0, colors_precomp=np.asarray(rgb, np.float32).reshape(n, 3),
              scales=np.asarray(scale, np.float32).reshape(n, 3), rotations=np.tile([1.0, 0, 0, 0], (n, 1)).astype(np.float32))
    if extra:
        kw.update(extra)
    return oracle.forward(**kw), kw


def test_single_gaussian_closed_form(oracle64):
    """alpha(d) = min(.99, o exp(-d^2 / (2 sigma_px^2))), sigma_px^2 = (f s / z)^2 + 0.3 (forward.cu:111-112)."""
    s, z, o, W = 0.05, 2.0, 0.8, 64
    f, _ = _single(oracle64, [0, 0, z], [s, s, s], [o], [1.0, 0.5, 0.25], W=W, H=W)
    focal = W / 2.0  # tanfov = 1
    var = (focal * s / z) ** 2 + 0.3
    cx = ((0.0 + 1.0) * W - 1.0) * 0.5  # ndc2Pix of ndc 0
    assert f["means2D"][0] == pytest.approx([cx, cx])
    # eigenvalues use sqrt(max(0.1, mid^2 - det)) (forward.cu:232-233): an isotropic splat gets lambda = var + sqrt(.1)
    assert f["radii"][0] == math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))
    for (px, py) in [(31, 31), (32, 32), (30, 33), (28, 31), (35, 35)]:
        d2 = (cx - px) ** 2 + (cx - py) ** 2
        a = min(0.99, o * math.exp(-0.5 * d2 / var))
        exp_c = a * np.array([1.0, 0.5, 0.25]) if a >= 1 / 255 else np.zeros(3)
        assert f["color"][:, py, px] == pytest.approx(exp_c, abs=1e-6)
        assert f["final_T"][py, px] == pytest.approx(1 - a if a >= 1 / 255 else 1.0, abs=1e-6)


def test_sh_degree0_colour_and_clamp(oracle32):
    """colour = 0.28209479 * sh0 + 0.5, clamped at 0 with the clamp recorded (forward.cu:63-70)."""
    sh = np.zeros((2, 1, 3), np.float32)
    sh[0, 0] = [1.0, -1.0, 0.2]
    sh[1, 0] = [-3.0, 0.0, 0.0]  # 0.5 - 0.846 < 0 -> clamped
    f, _ = _single(oracle32, [[0, 0, 2], [0.3, 0, 2]], [[.05] * 3] * 2, [1, 1], [[0] * 3] * 2,
                   extra=dict(colors_precomp=None, shs=sh, sh_degree=0))
    assert f["rgb"][0] == pytest.approx(0.28209479177387814 * sh[0, 0] + 0.5, abs=1e-6)
    assert f["rgb"][1, 0] == 0.0 and f["clamped"][1].tolist() == [1, 0, 0] and f["clamped"][0].tolist() == [0, 0, 0]


def test_depth_order_and_saturation(oracle64):
    """Two opaque overlapping Gaussians: the nearer one is composited first; with opacity 1 alpha clamps at .99 and
    the pixel stops before the second full hit because (1-.99)^2 < 1e-4 (forward.cu:343-347)."""
    f, _ = _single(oracle64, [[0, 0, 3.0], [0, 0, 2.0]], [[1.0] * 3] * 2, [1.0, 1.0], [[1, 0, 0], [0, 1, 0]])
    c = f["color"][:, 32, 32]
    assert f["point_list"][f["ranges"][f["ranges"][:, 1] > 0][0, 0]] in (0, 1)
    # nearer (index 1, green) first with alpha .99; the red one would leave T = 1e-4 * ... -> not blended
    assert c == pytest.approx([0.0, 0.99, 0.0], abs=1e-6)
    assert f["n_contrib"][32, 32] == 1 and f["final_T"][32, 32] == pytest.approx(0.01, abs=1e-7)


def test_near_plane_cull_and_mark_visible(oracle32):
    f, kw = _single(oracle32, [[0, 0, 0.2], [0, 0, 0.2001], [0, 0, -1.0]], [[.01] * 3] * 3, [1, 1, 1], [[1, 1, 1]] * 3)
    assert f["radii"][0] == 0 and f["radii"][2] == 0 and f["radii"][1] > 0  # p_view.z <= 0.2 is culled
    vis = oracle32.mark_visible(kw["means3D"], kw["viewmatrix"], kw["projmatrix"])
    assert vis.tolist() == [False, True, False]


def test_empty_cloud_and_background(oracle32):
    f, _ = _single(oracle32, np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 1)), np.zeros((0, 3)), bg=(0.1, 0.2, 0.3))
    assert f["num_rendered"] == 0
    assert np.allclose(f["color"], np.array([0.1, 0.2, 0.3], np.float32)[:, None, None])


def test_higher_msb_and_sort_stability(oracle32):
    import ctypes as C
    lib = oracle32.lib
    assert [int(lib.gso_higher_msb(C.c_uint32(n))) for n in (4096, 8160, 16384, 1, 255, 256)] == [13, 13, 15, 1, 8, 9]
    rng = np.random.default_rng(0)
    keys = (rng.integers(0, 50, 20000).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 7, 20000).astype(np.uint64)
    vals = np.arange(20000, dtype=np.uint32)
    ko, vo = np.zeros_like(keys), np.zeros_like(vals)
    lib.gso_sort_pairs(C.c_uint32(20000), C.c_int(32 + 6), keys.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p),
                       ko.ctypes.data_as(C.c_void_p), vo.ctypes.data_as(C.c_void_p))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(vo, vals[order]) and np.array_equal(ko, keys[order])


def test_fp32_matches_fp64_oracle(oracle32, oracle64, golden):
    _, kw, _ = golden["human_m13"]
    a, b = oracle32.forward(**kw), oracle64.forward(**kw)
    assert np.abs(a["color"] - b["color"]).max() < 2e-5


def test_backward_against_finite_differences(oracle64):
    """Analytic gradients of the fp64 oracle vs central differences of its own forward.  The scene is built so that
    the loss is smooth: every splat covers the whole 16x16 frame with alpha > 1/255, nothing saturates (.99 clamp,
    T < 1e-4 stop) and no view-space clamp is active, so no discontinuity of App. A sits inside the stencil."""
    rng = np.random.default_rng(4)
    n = 8
    cl = scenes.tiny_cloud(n, seed=4, sh_degree=1, spread=0.05)
    cl["scales"] = torch.from_numpy(rng.uniform(4.0, 8.0, (n, 3)).astype(np.float32))
    cl["opacities"] = torch.from_numpy(rng.uniform(0.05, 0.3, (n, 1)).astype(np.float32))
    v = scenes.make_view(scenes.orbit_c2w(12)[2], 16, 16)
    base = dict(W=16, H=16, viewmatrix=v.viewmatrix, projmatrix=v.projmatrix, campos=v.campos,
                bg=np.array([0.3, 0.1, 0.2], np.float32), tanfovx=v.tanfovx, tanfovy=v.tanfovy, sh_degree=1)
    params = {k: cl[k].numpy().astype(np.float64) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    w = loss_weights((3, 16, 16)).astype(np.float64)

    def loss(p):
        f = oracle64.forward(**base, **{k: x.astype(np.float32) for k, x in p.items()})
        return float((f["color"] * w).sum()), f

    l0, f0 = loss(params)
    assert f0["n_contrib"].min() == n and f0["final_T"].min() > 1e-3  # smooth regime reached
    g = oracle64.backward(f0, w, **base, **{k: x.astype(np.float32) for k, x in params.items() if k != "opacities"})
    names = {"means3D": "dL_dmeans3D", "opacities": "dL_dopacity", "shs": "dL_dsh", "scales": "dL_dscales",
             "rotations": "dL_drotations"}
    for k, gk in names.items():
        x = params[k]
        for _ in range(8):
            idx = tuple(rng.integers(0, s) for s in x.shape)
            eps = 2.0 ** -7 * max(1e-1, abs(x[idx]))  # inputs are float32: the step must survive the cast
            xp, xm = {**params, k: x.copy()}, {**params, k: x.copy()}
            xp[k][idx] = np.float32(x[idx] + eps)
            xm[k][idx] = np.float32(x[idx] - eps)
            num = (loss(xp)[0] - loss(xm)[0]) / (float(xp[k][idx]) - float(xm[k][idx]))
            ana = g[gk].reshape(x.shape)[idx]
            assert ana == pytest.approx(num, rel=2e-2, abs=2e-3 * (np.abs(g[gk]).max() + 1e-9)), (k, idx, ana, num)


@pytest.mark.parametrize("fov", [45, 40])
def test_camera_oracle_matches_reference_camera_setup(fov):
    """oracle/camera.py against the outputs of the reference's own get_rasterize_param_from_camera /
    inv_homogeneous_tensors (tests/golden/camera_params.npz, generated by tests/golden/make_camera_golden.py)."""
    from oracle import camera
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_params.npz"))
    r = camera.raster_params(g["c2w"], float(fov), float(fov), 512, 512, 2)
    assert np.array_equal(r["viewmatrix"][:, :, :3][:, :3], g[f"view_{fov}"][:, :3, :3])      # transposed rotation: exact
    np.testing.assert_allclose(r["viewmatrix"], g[f"view_{fov}"], rtol=0, atol=2e-6)          # -R^T t: summation order
    np.testing.assert_allclose(r["projmatrix"], g[f"proj_{fov}"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(r["campos"], g[f"campos_{fov}"])
    assert [r["tanfovx"], r["tanfovy"], r["image_height"], r["image_width"]] == list(g[f"scalars_{fov}"])
    # the bench's own cameras (scenes.make_view) are the same construction
    v = scenes.make_view(g["c2w"][3], 512, 512, fov_deg=float(fov), super_sample=2)
    np.testing.assert_allclose(v.viewmatrix, g[f"view_{fov}"][3], atol=2e-6)
    np.testing.assert_allclose(v.projmatrix, g[f"proj_{fov}"][3], rtol=2e-6, atol=2e-6)
    assert (v.image_height, v.image_width, v.tanfovx) == (1024, 1024, g[f"scalars_{fov}"][0])


def _head_case(name):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_head_golden as mh
    cfg = dict(mh.CONFIGS[name])
    cfg.pop("C")
    return mh.inputs(name), cfg


@pytest.mark.parametrize("name", ["shipped", "all_heads", "bare", "raw_normal"])
def test_head_oracle_matches_reference_head_decode(name):
    """oracle/head.py against the outputs of the reference's own head-decode statements (models/model_v2.py:286-375 +
    the caller's glue), tests/golden/head_decode.npz generated by tests/golden/make_head_golden.py."""
    from oracle import head
    (feat, rgb, prim), cfg = _head_case(name)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "head_decode.npz"))
    r = head.decode_head(feat, rgb, prim, scale_factor=448, xyz_offset=512, **cfg)
    for k in ("means3D", "rotations", "scales", "opacities", "shs"):
        assert np.array_equal(r[k].numpy(), g[f"{name}.{k}"]), k
    if cfg["est_normal"]:
        assert np.array_equal(r["normals"].numpy(), g[f"{name}.normals"])
    # the GPU flavour of the two scalar divisions stays within one ulp of the true division
    r2 = head.decode_head(feat, rgb, prim, scale_factor=448, xyz_offset=512, cuda_scalar_division=True, **cfg)
    for k in ("means3D", "shs"):
        np.testing.assert_allclose(r2[k].numpy(), r[k].numpy(), rtol=2.4e-7, atol=3e-7)


def test_preprocess_backward_derivation(oracle64):
    """The per-Gaussian backward kernel implements gradients DERIVED in matrix form (csrc/preprocess_backward.cu header);
    tests/derivation_preprocess_backward.py states the same formulas in numpy.  They must reproduce the oracle's
    restatement of the reference (backward.cu:144-396) to double-precision round-off, for all SH degrees, clamped
    colour channels, clamped view-space positions and culled points."""
    import ctypes as C

    from derivation_preprocess_backward import matrix_form_backward
    from oracle.oracle import _ptr
    rng = np.random.default_rng(0)
    for D, M, tanx, tany in ((3, 16, 1.0, 0.8), (1, 13, 0.35, 0.3), (0, 1, 1.0, 1.0), (2, 9, 0.6, 0.9)):
        P, W, H = 300, 640, 480
        means = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
        means[:, 2] += 3.0
        scales = np.exp(rng.uniform(-4, -2, (P, 3))).astype(np.float32)
        rots = rng.standard_normal((P, 4)).astype(np.float32)
        mod = 1.3
        shs = (0.4 * rng.standard_normal((P, M, 3))).astype(np.float32)
        clamped = (rng.uniform(size=(P, 3)) < 0.2).astype(np.uint8)
        radii = (rng.uniform(size=P) < 0.9).astype(np.int32)
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        Rw = Q * np.sign(np.linalg.det(Q))
        tr = np.array([0.1, -0.2, 0.4])
        w2c = np.eye(4)
        w2c[:3, :3], w2c[:3, 3] = Rw, tr
        view = np.ascontiguousarray(w2c.T.astype(np.float32)).reshape(16)
        Pm = np.zeros((4, 4))
        Pm[0, 0], Pm[1, 1], Pm[2, 2], Pm[2, 3], Pm[3, 2] = 1 / tanx, 1 / tany, 1.0001, -0.01, 1
        proj = np.ascontiguousarray((Pm @ w2c).T.astype(np.float32)).reshape(16)
        campos = (-Rw.T @ tr).astype(np.float32)
        cov = rng.standard_normal((P, 3, 3))
        cov = 1e-3 * np.einsum("pij,pkj->pik", cov, cov)                     # any symmetric PSD world covariance
        cov = np.ascontiguousarray(np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2],
                                             cov[:, 2, 2]], 1))
        g2, gc, gcol = rng.standard_normal((P, 3)), rng.standard_normal((P, 4)), rng.standard_normal((P, 3))
        out = dict(m3=np.zeros((P, 3)), c3=np.zeros((P, 6)), sh=np.zeros((P, M, 3)), sc=np.zeros((P, 3)),
                   rt=np.zeros((P, 4)))
        oracle64.lib.gso_preprocess_backward(
            C.c_int(P), C.c_int(D), C.c_int(M), _ptr(means), _ptr(radii), _ptr(shs), _ptr(clamped), _ptr(scales),
            _ptr(rots), C.c_float(mod), _ptr(cov), _ptr(view), _ptr(proj), _ptr(campos), C.c_int(W), C.c_int(H),
            C.c_float(tanx), C.c_float(tany), _ptr(g2), _ptr(gc), _ptr(gcol), _ptr(out["m3"]), _ptr(out["c3"]),
            _ptr(out["sh"]), _ptr(out["sc"]), _ptr(out["rt"]))
        mine = matrix_form_backward(means=means, radii=radii, shs=shs, clamped=clamped, scales=scales, rots=rots, mod=mod,
                                    cov=cov, view=view, proj=proj, campos=campos, W=W, H=H, tanx=tanx, tany=tany, D=D,
                                    g2=g2, gc=gc, gcol=gcol)
        assert (np.abs(np.concatenate([means @ Rw.T[:, :1] + tr[0]]) / (means @ Rw.T[:, 2:] + tr[2])) > 1.3 * tanx).any() \
            or tanx >= 0.6                                                   # the narrow case exercises the clamp
        for k in out:
            err = np.abs(out[k] - mine[k]).max() / (np.abs(out[k]).max() + 1e-30)
            assert err < 2e-5, (D, k, err)  # (un-normalised quaternions: large, cancelling rotation entries)
