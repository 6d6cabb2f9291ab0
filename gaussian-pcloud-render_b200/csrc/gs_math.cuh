// gs_math.cuh -- per-Gaussian math shared by the forward and backward preprocess kernels.
//
// The expression trees below restate the semantics listed in SURVEY.md App. A (items 1-9) in the same operand
// order the reference evaluates them (dgr/cuda_rasterizer/forward.cu:20-155, auxiliary.h:41-97, and GLM's
// column-major mat3 product order, third_party/glm/glm/detail/type_mat3x3.inl:486-519), so that with the default
// nvcc flags (no fast-math, -fmad=true) the discontinuous decisions downstream -- ceil(3 sigma), the tile
// rectangle, alpha < 1/255, T < 1e-4 -- fall on the same side as in the reference.
#pragma once
#include <cuda_runtime.h>

struct M3 {  // column-major: c[col][row]
    float c[3][3];
};

__device__ __forceinline__ M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1,
                                      float c2) {
    M3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}

__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B) {
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.c[j][i] = A.c[0][i] * B.c[j][0] + A.c[1][i] * B.c[j][1] + A.c[2][i] * B.c[j][2];
    return r;
}

__device__ __forceinline__ M3 m3_t(const M3& A) {
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) r.c[j][i] = A.c[i][j];
    return r;
}

__device__ __forceinline__ float3 xform43(const float* __restrict__ m, float3 p) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform44(const float* __restrict__ m, float3 p) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

// World covariance from scale + (un-normalised) quaternion; upper triangle in cov6.
__device__ __forceinline__ void cov3d_from_scale_rot(float3 scale, float mod, float4 q, float* cov6) {
    M3 S = m3_cols(mod * scale.x, 0.f, 0.f, 0.f, mod * scale.y, 0.f, 0.f, 0.f, mod * scale.z);
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 M = m3_mul(S, R);
    M3 Sg = m3_mul(m3_t(M), M);
    cov6[0] = Sg.c[0][0]; cov6[1] = Sg.c[0][1]; cov6[2] = Sg.c[0][2];
    cov6[3] = Sg.c[1][1]; cov6[4] = Sg.c[1][2]; cov6[5] = Sg.c[2][2];
}

struct Cov2D {
    M3 T, Vrk, W;
    float tx, ty, tz, txtz, tytz, limx, limy;
    float a, b, c;  // dilated 2D covariance (a b; b c)
};

// EWA projection of the 3D covariance to screen space, with the reference's clamp of the view-space direction
// and its +0.3 dilation of the diagonal.
__device__ __forceinline__ void cov2d_eval(float3 mean, float fx, float fy, float tanx, float tany, const float* cov6,
                                           const float* __restrict__ view, Cov2D& k) {
    float3 t = xform43(view, mean);
    k.limx = 1.3f * tanx;
    k.limy = 1.3f * tany;
    k.txtz = t.x / t.z;
    k.tytz = t.y / t.z;
    t.x = fminf(k.limx, fmaxf(-k.limx, k.txtz)) * t.z;
    t.y = fminf(k.limy, fmaxf(-k.limy, k.tytz)) * t.z;
    k.tx = t.x; k.ty = t.y; k.tz = t.z;
    M3 J = m3_cols(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z), 0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z), 0.f, 0.f, 0.f);
    k.W = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    k.T = m3_mul(k.W, J);
    k.Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
    M3 cov = m3_mul(m3_mul(m3_t(k.T), m3_t(k.Vrk)), k.T);
    k.a = cov.c[0][0] + 0.3f;
    k.b = cov.c[0][1];
    k.c = cov.c[1][1] + 0.3f;
}

__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }  // in double

// Tile rectangle of a splat of integer radius r around pixel p, clipped to the grid (and to the tile-row shard).
__device__ __forceinline__ void tile_rect(float px, float py, int r, int gx, int gy, int& x0, int& y0, int& x1,
                                          int& y1) {
    x0 = min(gx, max(0, (int)((px - r) / GS_TILE)));
    y0 = min(gy, max(0, (int)((py - r) / GS_TILE)));
    x1 = min(gx, max(0, (int)((px + r + GS_TILE - 1) / GS_TILE)));
    y1 = min(gy, max(0, (int)((py + r + GS_TILE - 1) / GS_TILE)));
}

#define GS_SH_C0 0.28209479177387814f
#define GS_SH_C1 0.4886025119029199f
__device__ const float GS_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                      -1.0925484305920792f, 0.5462742152960396f};
__device__ const float GS_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                      0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                      -0.5900435899266435f};
