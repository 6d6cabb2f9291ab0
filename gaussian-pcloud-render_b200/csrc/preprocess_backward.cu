// preprocess_backward.cu -- per-Gaussian backward (K8 + K9 fused; replaces computeCov2DCUDA and the backward
// preprocessCUDA, dgr/cuda_rasterizer/backward.cu:144-274 and :346-396, with the SH backward :20-139 and the
// scale/rotation backward :278-341).  One thread per Gaussian, one launch instead of two.
//
// The gradients are derived here from the forward model in matrix form (not transcribed term by term), and checked
// against the C oracle's restatement of the reference in tests/test_oracle.py::test_preprocess_backward_derivation
// (a numpy mirror of exactly these formulas) and against the reference library on the GPU (<= 1e-3, tests/test_gpu.py):
//
//   view space     t = R m + t0            R = rows r0, r1, r2 of the view rotation
//   projection     Sigma' = M V M^T + 0.3 I,   M = J R (2 x 3),   J = [fx/tz 0 -fx tx/tz^2; 0 fy/tz -fy ty/tz^2]
//   conic          K = Sigma'^-1;  blend backward delivers G = dL/dK as the symmetric matrix [gx gy; gy gz]
//   => dL/dSigma' = D = -K G K                (the reference's dL_da, dL_dc are D11, D22, its dL_db is 2 D12)
//      dL/dV      = M^T D M                   (stored as the upper triangle with doubled off-diagonals: an off-diagonal
//                                              entry of the symmetric V appears twice)
//      dL/dM      = 2 D M V,  dL/dJ_ic = (dL/dM)_i . r_c,  dL/dt through the four non-zero entries of J
//                                              (no gradient through tx, ty where the reference clamps them)
//      dL/dm      = R^T dL/dt + the projected-mean term + the view-direction term of the SH colour
//   SH colour      c = sum_k B_k(d) sh_k + 0.5,  d = (m - cam) / |m - cam|
//   => dL/dsh_k = B_k(d) dL/dc,   dL/dd = sum_k grad B_k(d) (sh_k . dL/dc),   dL/dm += (g - d (d . g)) / |m - cam|
//   world cov.     V = A S^2 A^T,  A = rotation of the (un-normalised) quaternion with columns a_k,  S = mod * scale
//   => dL/dS_k = 2 S_k a_k^T E a_k,   dL/dA = 2 E A S^2   (E = dL/dV as a symmetric matrix, off-diagonals halved),
//      dL/dq by the chain rule through the nine entries of A.  Two quirks of the reference are kept on purpose
//      (SURVEY App. A item 17): dL/dscale is dL/dS without the factor `mod`, and the quaternion is not normalised.
// Streaming, HBM-bound: ~80 B + SH in, ~64 B + 12*(D+1)^2 out per visible Gaussian.
#include "gs_common.cuh"
#include "gs_math.cuh"

namespace {

struct BwdArgs {
    int P, D, M;
    float fx, fy, tanx, tany, mod;
    const float* means; const int32_t* radii; const float* shs; const uint8_t* clamp; const float* scales;
    const float* rots; const float* cov3D; const float* view; const float* proj; const float* campos;
    const float* dL_dmean2D; const float* dL_dconic; const float* dL_dcolor;
    float* dL_dmean3D; float* dL_dcov3D; float* dL_dsh; float* dL_dscale; float* dL_drot;
};

__device__ __forceinline__ float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 lin2(float s, float3 a, float t, float3 b) {  // s a + t b
    return v3(s * a.x + t * b.x, s * a.y + t * b.y, s * a.z + t * b.z);
}
__device__ __forceinline__ float3 symmul(const float* c6, float3 v) {  // V v for V = upper triangle c6
    return v3(c6[0] * v.x + c6[1] * v.y + c6[2] * v.z, c6[1] * v.x + c6[3] * v.y + c6[4] * v.z,
              c6[2] * v.x + c6[4] * v.y + c6[5] * v.z);
}

// One SH coefficient: dL/dsh_k = B dRGB, and its share of dL/dd.
__device__ __forceinline__ void sh_term(int k, float B, float3 gradB, const float* __restrict__ sh, float3 dRGB,
                                        float* __restrict__ dL_dsh, float3& g) {
    dL_dsh[3 * k] = B * dRGB.x;
    dL_dsh[3 * k + 1] = B * dRGB.y;
    dL_dsh[3 * k + 2] = B * dRGB.z;
    const float s = sh[3 * k] * dRGB.x + sh[3 * k + 1] * dRGB.y + sh[3 * k + 2] * dRGB.z;
    g.x += s * gradB.x; g.y += s * gradB.y; g.z += s * gradB.z;
}

// SH backward of one Gaussian: writes dL_dsh rows 0..(deg+1)^2-1 and returns dL/dmean through the view direction.
// B_k and grad B_k are the real SH basis polynomials in the reference's sign convention (forward.cu:20-71).
__device__ __forceinline__ float3 sh_backward(int deg, const float* __restrict__ sh, float3 mean, float3 cam,
                                              unsigned clamp_bits, float3 dcol, float* __restrict__ dL_dsh) {
    const float3 v = v3(mean.x - cam.x, mean.y - cam.y, mean.z - cam.z);
    const float inv_len = 1.0f / sqrtf(dot3(v, v));
    const float x = v.x * inv_len, y = v.y * inv_len, z = v.z * inv_len;
    // a clamped channel (colour < 0 in the forward pass) passes no gradient
    const float3 dRGB = v3((clamp_bits & 1u) ? 0.f : dcol.x, (clamp_bits & 2u) ? 0.f : dcol.y,
                           (clamp_bits & 4u) ? 0.f : dcol.z);
    float3 g = v3(0.f, 0.f, 0.f);
    sh_term(0, GS_SH_C0, v3(0.f, 0.f, 0.f), sh, dRGB, dL_dsh, g);
    if (deg > 0) {
        const float c1 = GS_SH_C1;
        sh_term(1, -c1 * y, v3(0.f, -c1, 0.f), sh, dRGB, dL_dsh, g);
        sh_term(2, c1 * z, v3(0.f, 0.f, c1), sh, dRGB, dL_dsh, g);
        sh_term(3, -c1 * x, v3(-c1, 0.f, 0.f), sh, dRGB, dL_dsh, g);
    }
    if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z;
        const float a0 = GS_SH_C2[0], a1 = GS_SH_C2[1], a2 = GS_SH_C2[2], a3 = GS_SH_C2[3], a4 = GS_SH_C2[4];
        sh_term(4, a0 * x * y, v3(a0 * y, a0 * x, 0.f), sh, dRGB, dL_dsh, g);
        sh_term(5, a1 * y * z, v3(0.f, a1 * z, a1 * y), sh, dRGB, dL_dsh, g);
        sh_term(6, a2 * (2.f * zz - xx - yy), v3(-2.f * a2 * x, -2.f * a2 * y, 4.f * a2 * z), sh, dRGB, dL_dsh, g);
        sh_term(7, a3 * x * z, v3(a3 * z, 0.f, a3 * x), sh, dRGB, dL_dsh, g);
        sh_term(8, a4 * (xx - yy), v3(2.f * a4 * x, -2.f * a4 * y, 0.f), sh, dRGB, dL_dsh, g);
        if (deg > 2) {
            const float b0 = GS_SH_C3[0], b1 = GS_SH_C3[1], b2 = GS_SH_C3[2], b3 = GS_SH_C3[3], b4 = GS_SH_C3[4],
                        b5 = GS_SH_C3[5], b6 = GS_SH_C3[6];
            sh_term(9, b0 * y * (3.f * xx - yy), v3(6.f * b0 * x * y, 3.f * b0 * (xx - yy), 0.f), sh, dRGB, dL_dsh, g);
            sh_term(10, b1 * x * y * z, v3(b1 * y * z, b1 * x * z, b1 * x * y), sh, dRGB, dL_dsh, g);
            sh_term(11, b2 * y * (4.f * zz - xx - yy),
                    v3(-2.f * b2 * x * y, b2 * (4.f * zz - xx - 3.f * yy), 8.f * b2 * y * z), sh, dRGB, dL_dsh, g);
            sh_term(12, b3 * z * (2.f * zz - 3.f * xx - 3.f * yy),
                    v3(-6.f * b3 * x * z, -6.f * b3 * y * z, 3.f * b3 * (2.f * zz - xx - yy)), sh, dRGB, dL_dsh, g);
            sh_term(13, b4 * x * (4.f * zz - xx - yy),
                    v3(b4 * (4.f * zz - 3.f * xx - yy), -2.f * b4 * x * y, 8.f * b4 * x * z), sh, dRGB, dL_dsh, g);
            sh_term(14, b5 * z * (xx - yy), v3(2.f * b5 * x * z, -2.f * b5 * y * z, b5 * (xx - yy)), sh, dRGB, dL_dsh, g);
            sh_term(15, b6 * x * (xx - 3.f * yy), v3(3.f * b6 * (xx - yy), -6.f * b6 * x * y, 0.f), sh, dRGB, dL_dsh, g);
        }
    }
    // d = v / |v|:  dL/dv = (g - d (d . g)) / |v|
    const float dg = x * g.x + y * g.y + z * g.z;
    return v3((g.x - x * dg) * inv_len, (g.y - y * dg) * inv_len, (g.z - z * dg) * inv_len);
}

__global__ void __launch_bounds__(256) preprocess_backward_kernel(const BwdArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P || !(a.radii[i] > 0)) return;
    const float3 mean = v3(a.means[3 * i], a.means[3 * i + 1], a.means[3 * i + 2]);
    // world covariance: the caller's, or recomputed from scale / rotation with the forward pass's own function (same
    // bits; 24 B per Gaussian that the forward pass no longer stores and this kernel no longer reads back)
    float c6[6];
    if (a.cov3D == nullptr) {
        cov3d_from_scale_rot(v3(a.scales[3 * (size_t)i], a.scales[3 * (size_t)i + 1], a.scales[3 * (size_t)i + 2]), a.mod,
                             make_float4(a.rots[4 * (size_t)i], a.rots[4 * (size_t)i + 1], a.rots[4 * (size_t)i + 2],
                                         a.rots[4 * (size_t)i + 3]), c6);
    } else if ((reinterpret_cast<uintptr_t>(a.cov3D) & 7u) == 0) {  // 24-byte rows of an 8-byte aligned array: three 8-byte loads
        const float2* cp = reinterpret_cast<const float2*>(a.cov3D + 6 * (size_t)i);
        const float2 p0 = cp[0], p1 = cp[1], p2 = cp[2];
        c6[0] = p0.x; c6[1] = p0.y; c6[2] = p1.x; c6[3] = p1.y; c6[4] = p2.x; c6[5] = p2.y;
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = a.cov3D[6 * (size_t)i + k];
    }
    const float* vm = a.view;
    const float3 r0 = v3(vm[0], vm[4], vm[8]), r1 = v3(vm[1], vm[5], vm[9]), r2 = v3(vm[2], vm[6], vm[10]);

    // ---- forward quantities: clamped view-space position, M = J R, U = M V, Sigma' ----
    const float tz = dot3(r2, mean) + vm[14];
    const float limx = 1.3f * a.tanx, limy = 1.3f * a.tany;
    const float ux = (dot3(r0, mean) + vm[12]) / tz, uy = (dot3(r1, mean) + vm[13]) / tz;
    const bool in_x = !(ux < -limx || ux > limx), in_y = !(uy < -limy || uy > limy);
    const float tx = fminf(limx, fmaxf(-limx, ux)) * tz, ty = fminf(limy, fmaxf(-limy, uy)) * tz;
    const float iz = 1.f / tz, iz2 = iz * iz;
    const float j00 = a.fx * iz, j02 = -(a.fx * tx) * iz2, j11 = a.fy * iz, j12 = -(a.fy * ty) * iz2;
    const float3 m0 = lin2(j00, r0, j02, r2), m1 = lin2(j11, r1, j12, r2);
    const float3 u0 = symmul(c6, m0), u1 = symmul(c6, m1);
    const float ca = dot3(m0, u0) + 0.3f, cb = dot3(m0, u1), cc = dot3(m1, u1) + 0.3f;

    // ---- D = -K G K with the reference's regularised 1 / det^2 ----
    const float det = ca * cc - cb * cb;
    const float w = 1.0f / (det * det + 0.0000001f);
    // [P][2][2] rows: gx, gy, (unused), gz (caller memory of any 4-byte alignment)
    const float gx = a.dL_dconic[4 * (size_t)i], gy = a.dL_dconic[4 * (size_t)i + 1], gz = a.dL_dconic[4 * (size_t)i + 3];
    float D11 = 0.f, D12 = 0.f, D22 = 0.f;
    if (w != 0.f) {  // det * K = [cc -cb; -cb ca]
        const float p0 = cc * gx - cb * gy, p1 = cc * gy - cb * gz;   // (det K) G, first row
        const float q0 = ca * gy - cb * gx, q1 = ca * gz - cb * gy;   // second row
        D11 = -w * (p0 * cc - p1 * cb);
        D12 = -w * (p1 * ca - p0 * cb);
        D22 = -w * (q1 * ca - q0 * cb);
    }

    // ---- dL/dV = M^T D M ----
    const float3 e0 = lin2(D11, m0, D12, m1), e1 = lin2(D12, m0, D22, m1);  // rows of D M
    float dcov[6];
    dcov[0] = m0.x * e0.x + m1.x * e1.x;
    dcov[3] = m0.y * e0.y + m1.y * e1.y;
    dcov[5] = m0.z * e0.z + m1.z * e1.z;
    dcov[1] = 2.f * (m0.x * e0.y + m1.x * e1.y);
    dcov[2] = 2.f * (m0.x * e0.z + m1.x * e1.z);
    dcov[4] = 2.f * (m0.y * e0.z + m1.y * e1.z);
    if ((reinterpret_cast<uintptr_t>(a.dL_dcov3D) & 7u) == 0) {
        float2* dp = reinterpret_cast<float2*>(a.dL_dcov3D + 6 * (size_t)i);
        dp[0] = make_float2(dcov[0], dcov[1]); dp[1] = make_float2(dcov[2], dcov[3]); dp[2] = make_float2(dcov[4], dcov[5]);
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)i + k] = dcov[k];
    }

    // ---- dL/dM = 2 D (M V) -> dL/dJ -> dL/dt -> dL/dmean ----
    const float3 q0 = lin2(2.f * D11, u0, 2.f * D12, u1), q1 = lin2(2.f * D12, u0, 2.f * D22, u1);
    const float dJ00 = dot3(q0, r0), dJ02 = dot3(q0, r2), dJ11 = dot3(q1, r1), dJ12 = dot3(q1, r2);
    const float dtx = in_x ? -a.fx * iz2 * dJ02 : 0.f;
    const float dty = in_y ? -a.fy * iz2 * dJ12 : 0.f;
    const float dtz = -iz2 * (a.fx * dJ00 + a.fy * dJ11) + 2.f * iz2 * iz * (a.fx * tx * dJ02 + a.fy * ty * dJ12);
    float3 dmean = v3(r0.x * dtx + r1.x * dty + r2.x * dtz, r0.y * dtx + r1.y * dty + r2.y * dtz,
                      r0.z * dtx + r1.z * dty + r2.z * dtz);

    // ---- projected 2D mean p = h.xy / (h.w + eps), h = P [m 1] ----
    {
        const float* pj = a.proj;
        const float hx = pj[0] * mean.x + pj[4] * mean.y + pj[8] * mean.z + pj[12];
        const float hy = pj[1] * mean.x + pj[5] * mean.y + pj[9] * mean.z + pj[13];
        const float hw = pj[3] * mean.x + pj[7] * mean.y + pj[11] * mean.z + pj[15];
        const float iw = 1.0f / (hw + 0.0000001f);
        const float g2x = a.dL_dmean2D[3 * (size_t)i], g2y = a.dL_dmean2D[3 * (size_t)i + 1];
        const float sx = g2x * iw, sy = g2y * iw;                // weights of d h.x/dm, d h.y/dm
        const float sw = -(sx * hx + sy * hy) * iw;              // weight of d h.w/dm
        dmean.x += sx * pj[0] + sy * pj[1] + sw * pj[3];
        dmean.y += sx * pj[4] + sy * pj[5] + sw * pj[7];
        dmean.z += sx * pj[8] + sy * pj[9] + sw * pj[11];
    }

    // ---- colour -> SH ----
    if (a.shs) {
        const float3 dcol = v3(a.dL_dcolor[3 * (size_t)i], a.dL_dcolor[3 * (size_t)i + 1], a.dL_dcolor[3 * (size_t)i + 2]);
        const float3 g = sh_backward(a.D, a.shs + (size_t)i * a.M * 3, mean, v3(a.campos[0], a.campos[1], a.campos[2]),
                                     a.clamp[i], dcol, a.dL_dsh + (size_t)i * a.M * 3);
        dmean.x += g.x; dmean.y += g.y; dmean.z += g.z;
    }
    a.dL_dmean3D[3 * (size_t)i] = dmean.x;
    a.dL_dmean3D[3 * (size_t)i + 1] = dmean.y;
    a.dL_dmean3D[3 * (size_t)i + 2] = dmean.z;

    // ---- covariance -> scale / rotation:  V = A S^2 A^T ----
    if (a.scales) {
        const float r = a.rots[4 * (size_t)i], x = a.rots[4 * (size_t)i + 1], y = a.rots[4 * (size_t)i + 2],
                    z = a.rots[4 * (size_t)i + 3];
        const float s[3] = {a.mod * a.scales[3 * i], a.mod * a.scales[3 * i + 1], a.mod * a.scales[3 * i + 2]};
        // columns a_k of the rotation
        const float3 ak[3] = {v3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + r * z), 2.f * (x * z - r * y)),
                              v3(2.f * (x * y - r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + r * x)),
                              v3(2.f * (x * z + r * y), 2.f * (y * z - r * x), 1.f - 2.f * (x * x + y * y))};
        const float e6[6] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], dcov[3], 0.5f * dcov[4], dcov[5]};
        float ds[3];
        float A[3][3];  // dL/dA, A[row][col]
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float3 ea = symmul(e6, ak[k]);          // E a_k
            ds[k] = 2.f * s[k] * dot3(ak[k], ea);
            const float f = 2.f * s[k] * s[k];
            A[0][k] = f * ea.x; A[1][k] = f * ea.y; A[2][k] = f * ea.z;
        }
        a.dL_dscale[3 * (size_t)i] = ds[0];
        a.dL_dscale[3 * (size_t)i + 1] = ds[1];
        a.dL_dscale[3 * (size_t)i + 2] = ds[2];
        float4 dq;
        dq.x = 2.f * (z * (A[1][0] - A[0][1]) + y * (A[0][2] - A[2][0]) + x * (A[2][1] - A[1][2]));
        dq.y = 2.f * (y * (A[0][1] + A[1][0]) + z * (A[0][2] + A[2][0]) + r * (A[2][1] - A[1][2])) - 4.f * x * (A[1][1] + A[2][2]);
        dq.z = 2.f * (x * (A[0][1] + A[1][0]) + r * (A[0][2] - A[2][0]) + z * (A[1][2] + A[2][1])) - 4.f * y * (A[0][0] + A[2][2]);
        dq.w = 2.f * (r * (A[1][0] - A[0][1]) + x * (A[0][2] + A[2][0]) + y * (A[1][2] + A[2][1])) - 4.f * z * (A[0][0] + A[1][1]);
        a.dL_drot[4 * (size_t)i] = dq.x; a.dL_drot[4 * (size_t)i + 1] = dq.y;
        a.dL_drot[4 * (size_t)i + 2] = dq.z; a.dL_drot[4 * (size_t)i + 3] = dq.w;
    }
}

}  // namespace

cudaError_t gs_launch_preprocess_backward(const GsFrame& f, const GsGeom& g, const int32_t* radii,
                                          const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                                          float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                                          float* dL_drot) {
    const GsScene& s = f.s;
    BwdArgs a;
    a.P = s.P; a.D = s.sh_degree; a.M = s.sh_stride;
    a.fx = f.focal_x; a.fy = f.focal_y; a.tanx = s.tan_fovx; a.tany = s.tan_fovy; a.mod = s.scale_modifier;
    a.means = s.means3D; a.radii = radii; a.shs = s.shs; a.clamp = g.clamp; a.scales = s.scales; a.rots = s.rotations;
    a.cov3D = s.cov3D_precomp;  // nullptr: recomputed from scale / rotation
    a.view = s.viewmatrix; a.proj = s.projmatrix; a.campos = s.campos;
    a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dcolor = dL_dcolor;
    a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscale = dL_dscale; a.dL_drot = dL_drot;
    preprocess_backward_kernel<<<(unsigned)gs_div_up(s.P, 256), 256, 0, f.stream>>>(a);
    gs_note_launch();
    return cudaGetLastError();
}
