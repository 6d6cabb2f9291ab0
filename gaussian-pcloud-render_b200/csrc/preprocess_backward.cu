// preprocess_backward.cu -- per-Gaussian backward (K8 + K9 fused; replaces computeCov2DCUDA and the backward
// preprocessCUDA, dgr/cuda_rasterizer/backward.cu:144-274 and :346-396, with the SH backward :20-139 and the
// scale/rotation backward :278-341).  One thread per Gaussian, one launch instead of two: both reference kernels
// read the same mean / covariance / view data, and the fused kernel keeps dL_dmean3D in registers between the
// covariance part (assignment, :273) and the projection + SH parts (+=, :387 and :138).
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

// SH backward: writes dL_dsh rows 0..(D+1)^2-1 of this Gaussian and returns the gradient w.r.t. the mean that
// flows through the view direction.
__device__ __forceinline__ float3 sh_backward(int deg, const float* __restrict__ sh, float3 mean, float3 cam,
                                              unsigned clamp_bits, float3 dL_dcol, float* __restrict__ dL_dsh) {
    const float3 dir_orig = make_float3(mean.x - cam.x, mean.y - cam.y, mean.z - cam.z);
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    float dRGB[3] = {dL_dcol.x, dL_dcol.y, dL_dcol.z};
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dRGB[ch] *= ((clamp_bits >> ch) & 1u) ? 0 : 1;
    float ddx[3] = {0.f, 0.f, 0.f}, ddy[3] = {0.f, 0.f, 0.f}, ddz[3] = {0.f, 0.f, 0.f};
#define SHV(i, ch) sh[(i) * 3 + (ch)]
#define DSH(i, wgt)                                                          \
    {                                                                        \
        const float w_ = (wgt);                                              \
        _Pragma("unroll") for (int ch = 0; ch < 3; ch++) dL_dsh[(i) * 3 + ch] = w_ * dRGB[ch]; \
    }
    DSH(0, GS_SH_C0);
    if (deg > 0) {
        DSH(1, -GS_SH_C1 * y);
        DSH(2, GS_SH_C1 * z);
        DSH(3, -GS_SH_C1 * x);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            ddx[ch] = -GS_SH_C1 * SHV(3, ch);
            ddy[ch] = -GS_SH_C1 * SHV(1, ch);
            ddz[ch] = GS_SH_C1 * SHV(2, ch);
        }
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, GS_SH_C2[0] * xy);
            DSH(5, GS_SH_C2[1] * yz);
            DSH(6, GS_SH_C2[2] * (2.f * zz - xx - yy));
            DSH(7, GS_SH_C2[3] * xz);
            DSH(8, GS_SH_C2[4] * (xx - yy));
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                ddx[ch] += GS_SH_C2[0] * y * SHV(4, ch) + GS_SH_C2[2] * 2.f * -x * SHV(6, ch) +
                           GS_SH_C2[3] * z * SHV(7, ch) + GS_SH_C2[4] * 2.f * x * SHV(8, ch);
                ddy[ch] += GS_SH_C2[0] * x * SHV(4, ch) + GS_SH_C2[1] * z * SHV(5, ch) +
                           GS_SH_C2[2] * 2.f * -y * SHV(6, ch) + GS_SH_C2[4] * 2.f * -y * SHV(8, ch);
                ddz[ch] += GS_SH_C2[1] * y * SHV(5, ch) + GS_SH_C2[2] * 2.f * 2.f * z * SHV(6, ch) +
                           GS_SH_C2[3] * x * SHV(7, ch);
            }
            if (deg > 2) {
                DSH(9, GS_SH_C3[0] * y * (3.f * xx - yy));
                DSH(10, GS_SH_C3[1] * xy * z);
                DSH(11, GS_SH_C3[2] * y * (4.f * zz - xx - yy));
                DSH(12, GS_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                DSH(13, GS_SH_C3[4] * x * (4.f * zz - xx - yy));
                DSH(14, GS_SH_C3[5] * z * (xx - yy));
                DSH(15, GS_SH_C3[6] * x * (xx - 3.f * yy));
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    ddx[ch] += (GS_SH_C3[0] * SHV(9, ch) * 3.f * 2.f * xy + GS_SH_C3[1] * SHV(10, ch) * yz +
                                GS_SH_C3[2] * SHV(11, ch) * -2.f * xy + GS_SH_C3[3] * SHV(12, ch) * -3.f * 2.f * xz +
                                GS_SH_C3[4] * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                GS_SH_C3[5] * SHV(14, ch) * 2.f * xz + GS_SH_C3[6] * SHV(15, ch) * 3.f * (xx - yy));
                    ddy[ch] += (GS_SH_C3[0] * SHV(9, ch) * 3.f * (xx - yy) + GS_SH_C3[1] * SHV(10, ch) * xz +
                                GS_SH_C3[2] * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                GS_SH_C3[3] * SHV(12, ch) * -3.f * 2.f * yz + GS_SH_C3[4] * SHV(13, ch) * -2.f * xy +
                                GS_SH_C3[5] * SHV(14, ch) * -2.f * yz + GS_SH_C3[6] * SHV(15, ch) * -3.f * 2.f * xy);
                    ddz[ch] += (GS_SH_C3[1] * SHV(10, ch) * xy + GS_SH_C3[2] * SHV(11, ch) * 4.f * 2.f * yz +
                                GS_SH_C3[3] * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                GS_SH_C3[4] * SHV(13, ch) * 4.f * 2.f * xz + GS_SH_C3[5] * SHV(14, ch) * (xx - yy));
                }
            }
        }
    }
#undef SHV
#undef DSH
    const float3 dL_ddir = make_float3(ddx[0] * dRGB[0] + ddx[1] * dRGB[1] + ddx[2] * dRGB[2],
                                       ddy[0] * dRGB[0] + ddy[1] * dRGB[1] + ddy[2] * dRGB[2],
                                       ddz[0] * dRGB[0] + ddz[1] * dRGB[1] + ddz[2] * dRGB[2]);
    // gradient through v / |v| (auxiliary.h:108-119)
    const float3 v = dir_orig, dv = dL_ddir;
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// dL/dcov3D -> dL/dscale, dL/drot (no quaternion-normalisation backward)
__device__ __forceinline__ void cov3d_backward(float3 scale, float mod, float4 rot, const float* dc, float* dL_dscale,
                                               float* dL_drot) {
    const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    const float3 s = make_float3(mod * scale.x, mod * scale.y, mod * scale.z);
    M3 S = m3_cols(s.x, 0.f, 0.f, 0.f, s.y, 0.f, 0.f, 0.f, s.z);
    M3 M = m3_mul(S, R);
    M3 dSig = m3_cols(dc[0], 0.5f * dc[1], 0.5f * dc[2], 0.5f * dc[1], dc[3], 0.5f * dc[4], 0.5f * dc[2],
                      0.5f * dc[4], dc[5]);
    M3 M2;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) M2.c[j][i] = 2.0f * M.c[j][i];
    M3 dM = m3_mul(M2, dSig);
    M3 Rt = m3_t(R);
    M3 dMt = m3_t(dM);
    const float sv[3] = {s.x, s.y, s.z};
#pragma unroll
    for (int k = 0; k < 3; k++)
        dL_dscale[k] = Rt.c[k][0] * dMt.c[k][0] + Rt.c[k][1] * dMt.c[k][1] + Rt.c[k][2] * dMt.c[k][2];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) dMt.c[k][i] *= sv[k];
#define D(a, b) dMt.c[a][b]
    dL_drot[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
    dL_drot[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) -
                 4 * x * (D(2, 2) + D(1, 1));
    dL_drot[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) -
                 4 * y * (D(2, 2) + D(0, 0));
    dL_drot[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) -
                 4 * z * (D(1, 1) + D(0, 0));
#undef D
}

__global__ void __launch_bounds__(256) preprocess_backward_kernel(const BwdArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P || !(a.radii[i] > 0)) return;
    const float3 mean = make_float3(a.means[3 * i], a.means[3 * i + 1], a.means[3 * i + 2]);
    const float* cov3D = a.cov3D + 6 * (size_t)i;
    float c6[6];
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = cov3D[k];

    // ---- gradient of the conic w.r.t. the 2D covariance, then w.r.t. cov3D and the view-space mean ----
    Cov2D k;
    cov2d_eval(mean, a.fx, a.fy, a.tanx, a.tany, c6, a.view, k);
    const float3 dL_dconic = make_float3(a.dL_dconic[4 * i], a.dL_dconic[4 * i + 1], a.dL_dconic[4 * i + 3]);
    const float x_grad_mul = k.txtz < -k.limx || k.txtz > k.limx ? 0 : 1;
    const float y_grad_mul = k.tytz < -k.limy || k.tytz > k.limy ? 0 : 1;
    const float ca = k.a, cb = k.b, cc = k.c;
    const float denom = ca * cc - cb * cb;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcov[6];
#define TT(col, row) k.T.c[col][row]
#define VV(col, row) k.Vrk.c[col][row]
#define WW(col, row) k.W.c[col][row]
    if (denom2inv != 0) {
        dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * cb * cc * dL_dconic.y + (denom - ca * cc) * dL_dconic.z);
        dL_dc = denom2inv * (-ca * ca * dL_dconic.z + 2 * ca * cb * dL_dconic.y + (denom - ca * cc) * dL_dconic.x);
        dL_db = denom2inv * 2 * (cb * cc * dL_dconic.x - (denom + 2 * cb * cb) * dL_dconic.y + ca * cb * dL_dconic.z);
        dcov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
        dcov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
        dcov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
        dcov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db +
                  2 * TT(1, 0) * TT(1, 1) * dL_dc;
        dcov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db +
                  2 * TT(1, 0) * TT(1, 2) * dL_dc;
        dcov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db +
                  2 * TT(1, 1) * TT(1, 2) * dL_dc;
    } else {
#pragma unroll
        for (int q = 0; q < 6; q++) dcov[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) a.dL_dcov3D[6 * (size_t)i + q] = dcov[q];

    const float dL_dT00 = 2 * (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_da +
                          (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_db;
    const float dL_dT01 = 2 * (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_da +
                          (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_db;
    const float dL_dT02 = 2 * (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_da +
                          (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_db;
    const float dL_dT10 = 2 * (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_dc +
                          (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_db;
    const float dL_dT11 = 2 * (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_dc +
                          (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_db;
    const float dL_dT12 = 2 * (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_dc +
                          (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_db;
    const float dL_dJ00 = WW(0, 0) * dL_dT00 + WW(0, 1) * dL_dT01 + WW(0, 2) * dL_dT02;
    const float dL_dJ02 = WW(2, 0) * dL_dT00 + WW(2, 1) * dL_dT01 + WW(2, 2) * dL_dT02;
    const float dL_dJ11 = WW(1, 0) * dL_dT10 + WW(1, 1) * dL_dT11 + WW(1, 2) * dL_dT12;
    const float dL_dJ12 = WW(2, 0) * dL_dT10 + WW(2, 1) * dL_dT11 + WW(2, 2) * dL_dT12;
#undef TT
#undef VV
#undef WW
    const float tz = 1.f / k.tz;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;
    const float dL_dtx = x_grad_mul * -a.fx * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -a.fy * tz2 * dL_dJ12;
    const float dL_dtz = -a.fx * tz2 * dL_dJ00 - a.fy * tz2 * dL_dJ11 + (2 * a.fx * k.tx) * tz3 * dL_dJ02 +
                         (2 * a.fy * k.ty) * tz3 * dL_dJ12;
    const float* vm = a.view;
    float3 dmean = make_float3(vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz,
                               vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
                               vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz);

    // ---- gradient of the projected 2D mean ----
    const float* proj = a.proj;
    const float4 m_hom = xform44(proj, mean);
    const float m_w = 1.0f / (m_hom.w + 0.0000001f);
    const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
    const float g2x = a.dL_dmean2D[3 * (size_t)i], g2y = a.dL_dmean2D[3 * (size_t)i + 1];
    float3 dm2;
    dm2.x = (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
    dm2.y = (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
    dm2.z = (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
    dmean.x += dm2.x; dmean.y += dm2.y; dmean.z += dm2.z;

    // ---- colour -> SH ----
    if (a.shs) {
        const float3 dcol = make_float3(a.dL_dcolor[3 * (size_t)i], a.dL_dcolor[3 * (size_t)i + 1],
                                        a.dL_dcolor[3 * (size_t)i + 2]);
        const float3 dsh_mean = sh_backward(a.D, a.shs + (size_t)i * a.M * 3, mean,
                                            make_float3(a.campos[0], a.campos[1], a.campos[2]), a.clamp[i], dcol,
                                            a.dL_dsh + (size_t)i * a.M * 3);
        dmean.x += dsh_mean.x; dmean.y += dsh_mean.y; dmean.z += dsh_mean.z;
    }
    a.dL_dmean3D[3 * (size_t)i] = dmean.x;
    a.dL_dmean3D[3 * (size_t)i + 1] = dmean.y;
    a.dL_dmean3D[3 * (size_t)i + 2] = dmean.z;

    // ---- covariance -> scale / rotation ----
    if (a.scales) {
        float ds[3], dq[4];
        cov3d_backward(make_float3(a.scales[3 * i], a.scales[3 * i + 1], a.scales[3 * i + 2]), a.mod,
                       reinterpret_cast<const float4*>(a.rots)[i], dcov, ds, dq);
        a.dL_dscale[3 * (size_t)i] = ds[0];
        a.dL_dscale[3 * (size_t)i + 1] = ds[1];
        a.dL_dscale[3 * (size_t)i + 2] = ds[2];
        reinterpret_cast<float4*>(a.dL_drot)[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
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
    a.cov3D = s.cov3D_precomp ? s.cov3D_precomp : g.cov3D;
    a.view = s.viewmatrix; a.proj = s.projmatrix; a.campos = s.campos;
    a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dcolor = dL_dcolor;
    a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscale = dL_dscale; a.dL_drot = dL_drot;
    preprocess_backward_kernel<<<(unsigned)gs_div_up(s.P, 256), 256, 0, f.stream>>>(a);
    gs_note_launch();
    return cudaGetLastError();
}
