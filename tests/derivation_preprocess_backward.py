"""numpy mirror (float64) of the matrix-form gradients implemented in
gaussian-pcloud-render_b200/csrc/preprocess_backward.cu -- TEST INFRASTRUCTURE: the formulas are stated once more,
independently of CUDA, so that the derivation itself can be checked on the CPU against the C oracle's restatement of the
reference (oracle/gs_oracle.c::gso_preprocess_backward, following backward.cu:144-396).  See the header of the .cu file
for the derivation."""
import numpy as np

SH0, SH1 = 0.28209479177387814, 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]
def basis_and_grad(deg, x, y, z):
    """B_i(d) and its gradient w.r.t. d = (x,y,z) for the real SH basis in the reference's sign convention."""
    B = [SH0]; G = [(0., 0., 0.)]
    if deg > 0:
        B += [-SH1 * y, SH1 * z, -SH1 * x]
        G += [(0, -SH1, 0), (0, 0, SH1), (-SH1, 0, 0)]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        B += [C2[0] * x * y, C2[1] * y * z, C2[2] * (2 * zz - xx - yy), C2[3] * x * z, C2[4] * (xx - yy)]
        G += [(C2[0] * y, C2[0] * x, 0), (0, C2[1] * z, C2[1] * y), (-2 * C2[2] * x, -2 * C2[2] * y, 4 * C2[2] * z),
              (C2[3] * z, 0, C2[3] * x), (2 * C2[4] * x, -2 * C2[4] * y, 0)]
    if deg > 2:
        B += [C3[0] * y * (3 * xx - yy), C3[1] * x * y * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
        G += [(C3[0] * 6 * x * y, C3[0] * (3 * xx - 3 * yy), 0),
              (C3[1] * y * z, C3[1] * x * z, C3[1] * x * y),
              (C3[2] * -2 * x * y, C3[2] * (4 * zz - xx - 3 * yy), C3[2] * 8 * y * z),
              (C3[3] * -6 * x * z, C3[3] * -6 * y * z, C3[3] * (6 * zz - 3 * xx - 3 * yy)),
              (C3[4] * (4 * zz - 3 * xx - yy), C3[4] * -2 * x * y, C3[4] * 8 * x * z),
              (C3[5] * 2 * x * z, C3[5] * -2 * y * z, C3[5] * (xx - yy)),
              (C3[6] * (3 * xx - 3 * yy), C3[6] * -6 * x * y, 0)]
    return B, G



def matrix_form_backward(*, means, radii, shs, clamped, scales, rots, mod, cov, view, proj, campos, W, H, tanx, tany, D,
                         g2, gc, gcol):
    """All inputs as in gso_preprocess_backward; returns dict(m3, c3, sh, sc, rt) = dL/d(mean3D, cov3D, sh, scale, rot)."""
    P, M = means.shape[0], shs.shape[1]
    mine = dict(m3=np.zeros((P, 3)), c3=np.zeros((P, 6)), sh=np.zeros((P, M, 3)), sc=np.zeros((P, 3)), rt=np.zeros((P, 4)))
    fx, fy = W / (2 * tanx), H / (2 * tany)
    V4 = view.astype(np.float64)
    Rv = np.array([[V4[c], V4[4 + c], V4[8 + c]] for c in range(3)])  # rows r_c of the view rotation
    tv = np.array([V4[12], V4[13], V4[14]])
    PJ = proj.astype(np.float64)
    for i in range(P):
        if radii[i] <= 0: continue
        m = means[i].astype(np.float64)
        t = Rv @ m + tv
        limx, limy = 1.3 * tanx, 1.3 * tany
        u, v = t[0] / t[2], t[1] / t[2]
        inx, iny = (-limx <= u <= limx), (-limy <= v <= limy)
        tx, ty, tz = min(limx, max(-limx, u)) * t[2], min(limy, max(-limy, v)) * t[2], t[2]
        iz = 1 / tz
        j0 = np.array([fx * iz, 0, -fx * tx * iz * iz]); j1 = np.array([0, fy * iz, -fy * ty * iz * iz])
        m0 = j0[0] * Rv[0] + j0[2] * Rv[2]; m1 = j1[1] * Rv[1] + j1[2] * Rv[2]      # rows of M = J R
        c6 = cov[i]
        V = np.array([[c6[0], c6[1], c6[2]], [c6[1], c6[3], c6[4]], [c6[2], c6[4], c6[5]]])
        u0, u1 = V @ m0, V @ m1
        a, b, c = m0 @ u0 + 0.3, m0 @ u1, m1 @ u1 + 0.3
        det = a * c - b * b
        w = 1 / (det * det + 1e-7)
        gx, gy, gz = gc[i, 0], gc[i, 1], gc[i, 3]
        # D = -K G K  (K = inverse of [[a,b],[b,c]], G the symmetric gradient w.r.t. K), scaled by det^2 * w
        k0 = np.array([c, -b]); k1 = np.array([-b, a])                 # det * rows of K
        Gm = np.array([[gx, gy], [gy, gz]])
        D11, D12, D22 = -w * (k0 @ Gm @ k0), -w * (k0 @ Gm @ k1), -w * (k1 @ Gm @ k1)
        if w == 0: D11 = D12 = D22 = 0
        # dL/dV = M^T D M, off-diagonals doubled (V's off-diagonal entries appear twice)
        e0 = D11 * m0 + D12 * m1; e1 = D12 * m0 + D22 * m1           # rows of D M
        E = np.outer(m0, e0) + np.outer(m1, e1)
        mine["c3"][i] = [E[0, 0], 2 * E[0, 1], 2 * E[0, 2], E[1, 1], 2 * E[1, 2], E[2, 2]]
        # dL/dM = 2 D (M V): rows
        q0 = 2 * (D11 * u0 + D12 * u1); q1 = 2 * (D12 * u0 + D22 * u1)
        dJ00, dJ02 = q0 @ Rv[0], q0 @ Rv[2]
        dJ11, dJ12 = q1 @ Rv[1], q1 @ Rv[2]
        dtx = (-fx * iz * iz * dJ02) if inx else 0.0
        dty = (-fy * iz * iz * dJ12) if iny else 0.0
        dtz = -iz * iz * (fx * dJ00 + fy * dJ11) + 2 * iz ** 3 * (fx * tx * dJ02 + fy * ty * dJ12)
        dm = Rv.T @ np.array([dtx, dty, dtz])
        # projected mean: p = (h.x, h.y) / (h.w + eps)
        h = PJ.reshape(4, 4).T @ np.array([m[0], m[1], m[2], 1.0])
        iw = 1 / (h[3] + 1e-7)
        cx, cy, cw = PJ[[0, 4, 8]], PJ[[1, 5, 9]], PJ[[3, 7, 11]]       # d h.x/dm, d h.y/dm, d h.w/dm
        dm += g2[i, 0] * (cx * iw - cw * h[0] * iw * iw) + g2[i, 1] * (cy * iw - cw * h[1] * iw * iw)
        # SH
        vdir = m - campos.astype(np.float64)
        n = np.linalg.norm(vdir); d = vdir / n
        dRGB = gcol[i] * (1 - clamped[i])
        B, G = basis_and_grad(D, *d)
        gd = np.zeros(3)
        for k in range((D + 1) ** 2):
            mine["sh"][i, k] = B[k] * dRGB
            gd += np.array(G[k]) * (shs[i, k].astype(np.float64) @ dRGB)
        dm += (gd - d * (d @ gd)) / n
        mine["m3"][i] = dm
        # scale / rotation: Sigma = R S^2 R^T, E3 = dL/dSigma (symmetric, off-diagonals halved)
        dc = mine["c3"][i]
        E3 = np.array([[dc[0], .5 * dc[1], .5 * dc[2]], [.5 * dc[1], dc[3], .5 * dc[4]], [.5 * dc[2], .5 * dc[4], dc[5]]])
        r, x, y, z = rots[i].astype(np.float64)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                      [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                      [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
        s = np.float64(np.float32(mod)) * scales[i].astype(np.float64)
        ER = E3 @ R                                   # columns E a_k
        mine["sc"][i] = 2 * s * np.einsum('jk,jk->k', R, ER)          # 2 s_k a_k^T E a_k   (no factor `mod`: reference quirk)
        Aq = 2 * ER * (s * s)[None, :]                # dL/dR, column k = 2 s_k^2 E a_k
        mine["rt"][i] = [2 * (z * (Aq[1, 0] - Aq[0, 1]) + y * (Aq[0, 2] - Aq[2, 0]) + x * (Aq[2, 1] - Aq[1, 2])),
                         2 * (y * (Aq[0, 1] + Aq[1, 0]) + z * (Aq[0, 2] + Aq[2, 0]) + r * (Aq[2, 1] - Aq[1, 2])) - 4 * x * (Aq[1, 1] + Aq[2, 2]),
                         2 * (x * (Aq[0, 1] + Aq[1, 0]) + r * (Aq[0, 2] - Aq[2, 0]) + z * (Aq[1, 2] + Aq[2, 1])) - 4 * y * (Aq[0, 0] + Aq[2, 2]),
                         2 * (r * (Aq[1, 0] - Aq[0, 1]) + x * (Aq[0, 2] + Aq[2, 0]) + y * (Aq[1, 2] + Aq[2, 1])) - 4 * z * (Aq[0, 0] + Aq[1, 1])]
    return mine
