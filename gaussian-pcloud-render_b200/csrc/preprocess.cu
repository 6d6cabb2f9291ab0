// preprocess.cu -- per-Gaussian forward stage (K1) and markVisible (K10).
//
// Replaces dgr/cuda_rasterizer/forward.cu:158-259 (preprocessCUDA) + rasterizer_impl.cu:277 (the InclusiveSum that
// only served to size the instance buffers) + rasterizer_impl.cu:54-66 (checkFrustum).
//
// One thread per Gaussian.  Output is ONE packed 48-B record per visible Gaussian (xy, conic, opacity, alpha
// cut-off threshold, depth, rgb) that the blend kernels gather with three 16-B loads, plus the depth-sort key,
// the tile rectangle and the tile count.  The total instance count (num_rendered) and the visible count are
// reduced per block and added to the frame header with one atomic each, so no device-wide scan is needed.
// HBM-bound: algorithmic bytes per point = 44 + 12*(D+1)^2 in, 8 out, + 67 per visible point.
#include "gs_common.cuh"
#include "gs_math.cuh"

namespace {

struct PreArgs {
    int P, D, M, W, H, gx, gy, row0, row1;
    float tanx, tany, fx, fy, mod;
    int prefiltered;
    const float* means; const float* scales; const float* rots; const float* opac; const float* shs;
    const float* cov3D_pre; const float* colors_pre; const float* view; const float* proj; const float* campos;
    int32_t* radii;
    GsRec* rec; uint32_t* key; ushort4* rect; uint32_t* ntile; uint8_t* clamp;
    int* rdiff;
    GsHeader* hdr;
    const uint32_t* cand;  // shard cull: thread t handles Gaussian cand[t], t < hdr->num_cand (nullptr: Gaussian t, t < P)
};

// SH -> RGB for one Gaussian; coefficient stride is M (may exceed (D+1)^2), SURVEY App. A item 9.
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* __restrict__ sh, float3 mean, float3 cam,
                                            unsigned& clamp_bits) {
    float3 dir = make_float3(mean.x - cam.x, mean.y - cam.y, mean.z - cam.z);
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    const float x = dir.x / len, y = dir.y / len, z = dir.z / len;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
#define SHC(i) sh[(i) * 3 + ch]
        // rounded product: on the degree-0 path the compiler would otherwise fuse it with the `+ 0.5f` below, which
        // the reference's build does not do (measured against the live reference library: 1 ulp in the colours)
        float r = __fmul_rn(GS_SH_C0, SHC(0));
        if (deg > 0) {
            r = r - GS_SH_C1 * y * SHC(1) + GS_SH_C1 * z * SHC(2) - GS_SH_C1 * x * SHC(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + GS_SH_C2[0] * xy * SHC(4) + GS_SH_C2[1] * yz * SHC(5) +
                    GS_SH_C2[2] * (2.0f * zz - xx - yy) * SHC(6) + GS_SH_C2[3] * xz * SHC(7) +
                    GS_SH_C2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    r = r + GS_SH_C3[0] * y * (3.0f * xx - yy) * SHC(9) + GS_SH_C3[1] * xy * z * SHC(10) +
                        GS_SH_C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
                        GS_SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
                        GS_SH_C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + GS_SH_C3[5] * z * (xx - yy) * SHC(14) +
                        GS_SH_C3[6] * x * (xx - 3.0f * yy) * SHC(15);
                }
            }
        }
#undef SHC
        r += 0.5f;
        if (r < 0.f) clamp_bits |= 1u << ch;
        res[ch] = (r < 0.0f) ? 0.0f : r;
    }
    return make_float3(res[0], res[1], res[2]);
}

// The same for degree <= 1 with the (at most) 12 coefficients already in registers (loaded up front by the caller, so
// that a Gaussian costs one round trip to memory instead of one per input array).
__device__ __forceinline__ float3 sh_to_rgb_low(int deg, const float (&sh)[12], float3 mean, float3 cam,
                                                unsigned& clamp_bits) {
    float3 dir = make_float3(mean.x - cam.x, mean.y - cam.y, mean.z - cam.z);
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    const float x = dir.x / len, y = dir.y / len, z = dir.z / len;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float r = __fmul_rn(GS_SH_C0, sh[ch]);
        if (deg > 0) r = r - GS_SH_C1 * y * sh[3 + ch] + GS_SH_C1 * z * sh[6 + ch] - GS_SH_C1 * x * sh[9 + ch];
        r += 0.5f;
        if (r < 0.f) clamp_bits |= 1u << ch;
        res[ch] = (r < 0.0f) ? 0.0f : r;
    }
    return make_float3(res[0], res[1], res[2]);
}

// ---- TMA bulk staging (sm_90+/sm_100a): one elected thread issues 1-D bulk copies global -> shared that complete
// on an mbarrier; the inputs of a block are five contiguous slabs of the caller's AoS arrays.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// STAGED: the block's inputs (256 means / scales / quaternions / opacities / SH rows) are brought into shared memory
// with TMA bulk copies and read from there: the 12-B / 16-B / (12 M)-B strided per-thread reads hit shared-memory
// banks (conflict-free for the benchmark's M = 13) instead of fetching partial sectors through L1.
template <bool STAGED>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;  // position in the depth-sort input
    const int n = a.cand ? (int)a.hdr->num_cand : a.P;
    if (a.cand && blockIdx.x * blockDim.x >= n) return;  // (whole block past the candidates)
    const int i = (a.cand && pos < n) ? (int)a.cand[pos] : pos;  // the Gaussian
    extern __shared__ __align__(128) unsigned char s_stage[];
    const float* mp = a.means + 3 * (size_t)i;
    const float* sp = a.scales ? a.scales + 3 * (size_t)i : nullptr;
    const float* rp = a.rots ? a.rots + 4 * (size_t)i : nullptr;
    const float* op_ptr = a.opac + i;
    const float* shp = a.shs ? a.shs + (size_t)i * a.M * 3 : nullptr;
    if (STAGED && (blockIdx.x + 1) * 256 <= a.P) {  // full blocks only; the last partial block reads global memory
        __shared__ __align__(8) unsigned long long s_bar;
        float* s_means = reinterpret_cast<float*>(s_stage);
        float* s_scales = s_means + 3 * 256;
        float* s_rots = s_scales + 3 * 256;
        float* s_opac = s_rots + 4 * 256;
        float* s_sh = s_opac + 256;
        const unsigned sh_bytes = 256u * (unsigned)a.M * 12u;
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const size_t b0 = (size_t)blockIdx.x * 256;
            mbar_expect_tx(&s_bar, 3072u + 3072u + 4096u + 1024u + sh_bytes);
            bulk_g2s(s_means, a.means + 3 * b0, 3072u, &s_bar);
            bulk_g2s(s_scales, a.scales + 3 * b0, 3072u, &s_bar);
            bulk_g2s(s_rots, a.rots + 4 * b0, 4096u, &s_bar);
            bulk_g2s(s_opac, a.opac + b0, 1024u, &s_bar);
            bulk_g2s(s_sh, a.shs + b0 * a.M * 3, sh_bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
        mp = s_means + 3 * threadIdx.x;
        sp = s_scales + 3 * threadIdx.x;
        rp = s_rots + 4 * threadIdx.x;
        op_ptr = s_opac + threadIdx.x;
        shp = s_sh + (size_t)threadIdx.x * a.M * 3;
    }
    unsigned my_tiles = 0, my_vis = 0, my_rows = 0;
    bool bad = false;
    __shared__ int s_rd[GS_MAX_GRID + 1];  // this block's share of the row difference array
    for (int y = threadIdx.x; y <= a.gy; y += 256) s_rd[y] = 0;
    __syncthreads();
    if (pos < n) {
        int radius_out = 0;
        uint32_t key = 0xFFFFFFFFu;  // culled Gaussians sort to the end of the depth order
        ushort4 rect = make_ushort4(0, 0, 0, 0);
        // every input of this Gaussian is requested before the first early-out below can hold a load back: one round
        // trip to memory per thread instead of four (position -> scale / rotation -> SH -> opacity)
        const float3 mean = make_float3(mp[0], mp[1], mp[2]);
        float3 sc_in = make_float3(0.f, 0.f, 0.f);
        float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.cov3D_pre == nullptr) {
            sc_in = make_float3(sp[0], sp[1], sp[2]);
            q_in = *reinterpret_cast<const float4*>(rp);
        }
        const float op = *op_ptr;
        const bool sh_low = !STAGED && a.colors_pre == nullptr && a.D <= 1;
        float sh_in[12];
#pragma unroll
        for (int k = 0; k < 12; k++) sh_in[k] = (sh_low && (k < 3 || a.D == 1)) ? shp[k] : 0.f;
        do {
            const float3 p_view = xform43(a.view, mean);
            if (p_view.z <= 0.2f) {  // near plane only (auxiliary.h:154)
                bad = a.prefiltered != 0;
                break;
            }
            const float4 p_hom = xform44(a.proj, mean);
            const float p_w = 1.0f / (p_hom.w + 0.0000001f);
            const float projx = p_hom.x * p_w, projy = p_hom.y * p_w;

            float c6[6];
            if (a.cov3D_pre != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) c6[k] = a.cov3D_pre[6 * (size_t)i + k];
            } else {
                cov3d_from_scale_rot(sc_in, a.mod, q_in, c6);  // (not stored: the backward pass recomputes it)
            }
            Cov2D k2;
            cov2d_eval(mean, a.fx, a.fy, a.tanx, a.tany, c6, a.view, k2);
            const float3 cov = make_float3(k2.a, k2.b, k2.c);
            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det == 0.0f) break;
            const float det_inv = 1.f / det;
            const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
            const float mid = 0.5f * (cov.x + cov.z);
            const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
            const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
            const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
            const float px = ndc_to_pix(projx, a.W), py = ndc_to_pix(projy, a.H);
            int x0, y0, x1, y1;
            tile_rect(px, py, (int)my_radius, a.gx, a.gy, x0, y0, x1, y1);
            if ((x1 - x0) * (y1 - y0) == 0) break;

            float3 rgb = make_float3(0.f, 0.f, 0.f);
            unsigned clamp_bits = 0;
            if (a.colors_pre == nullptr) {
                const float3 cam = make_float3(a.campos[0], a.campos[1], a.campos[2]);
                rgb = sh_low ? sh_to_rgb_low(a.D, sh_in, mean, cam, clamp_bits) : sh_to_rgb(a.D, shp, mean, cam, clamp_bits);
                a.clamp[i] = (uint8_t)clamp_bits;
            } else {
                rgb = make_float3(a.colors_pre[3 * i], a.colors_pre[3 * i + 1], a.colors_pre[3 * i + 2]);
            }
            // Conservative cut-off on `power`: below it, op*exp(power) < (1/255)(1 - 1e-3), so the blend kernels
            // may skip the exponential with no change to the result (alpha < 1/255 is skipped anyway).
            const float thr = -logf(255.0f * op) - 1.0e-3f;
            // op <= 0: alpha <= 0, always skipped (thr 0 skips every power < 0); NaN opacity: never skip.
            // -B/A and -B/C for the blend kernels' box-maximum cull.  The box maximum is only valid for a positive
            // definite conic (concave exponent): anything else (det < 0 from fp32 cancellation on huge thin splats, a
            // non-PSD cov3D_precomp) gets the cut-off -inf, i.e. is never culled and always evaluated per pixel.
            const bool pd = det > 0.f && conic.x > 0.f && conic.z > 0.f;
            const float thr_rec = !pd ? -INFINITY : ((op > 0.f) ? thr : ((op <= 0.f) ? 0.0f : -INFINITY));
            const float nBA = pd ? -conic.y / conic.x : __int_as_float(0x7fc00000);
            const float nBC = pd ? -conic.y / conic.z : __int_as_float(0x7fc00000);
            GsRec r;
            r.a = make_float4(px, py, conic.x, conic.y);
            r.b = make_float4(conic.z, op, thr_rec, nBC);
            r.c = make_float4(rgb.x, rgb.y, rgb.z, nBA);
            a.rec[i] = r;
            radius_out = (int)my_radius;
            // shard clip: this rank only bins tile rows [row0,row1); radii/records stay those of the full frame
            y0 = max(y0, a.row0);
            y1 = min(y1, a.row1);
            if (y1 > y0) {
                rect = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
                my_tiles = (unsigned)((y1 - y0) * (x1 - x0));
                my_rows = (unsigned)(y1 - y0);
                // row difference array, integrated by the row pass (binning.cu) into items per tile row
                atomicAdd(&s_rd[y0], 1);
                atomicAdd(&s_rd[y1], -1);
            }
            key = __float_as_uint(p_view.z);
            my_vis = 1;
        } while (false);
        if (a.radii) a.radii[i] = radius_out;
        a.key[pos] = key;
        a.rect[i] = rect;
        a.ntile[i] = my_tiles;
    }
    // block reduction of the instance / visible counts -> two atomics per block
    __shared__ unsigned s_t[8], s_v[8], s_r[8];
    unsigned t = my_tiles, v = my_vis, r = my_rows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(GS_FULL, t, o);
        v += __shfl_xor_sync(GS_FULL, v, o);
        r += __shfl_xor_sync(GS_FULL, r, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_t[w] = t; s_v[w] = v; s_r[w] = r; }
    if (__syncthreads_or(bad)) {
        if (threadIdx.x == 0) a.hdr->code = GS_ERR_PREFILTERED;
    }
    if (threadIdx.x == 0) {
        unsigned long long tt = 0;
        unsigned vv = 0, rr = 0;
        for (int k = 0; k < 8; k++) { tt += s_t[k]; vv += s_v[k]; rr += s_r[k]; }
        if (tt) atomicAdd(&a.hdr->num_rendered, tt);
        if (vv) atomicAdd(&a.hdr->num_visible, vv);
        if (rr) atomicAdd(&a.hdr->num_row_items, rr);
    }
    for (int y = threadIdx.x; y <= a.gy; y += 256) {  // the __syncthreads_or above ordered the shared atomics
        const int d = s_rd[y];
        if (d) atomicAdd(&a.rdiff[y], d);
    }
}


// ---- Shard cull (GsScene.shard_cull): which Gaussians can reach tile rows [row0, row1)? ---------------------------------
// A Gaussian is kept unless its splat provably misses the shard: same near-plane test and the same pixel y as the
// per-Gaussian stage (identical expressions), and an upper bound R on its screen radius ceil(3 sqrt(lambda_max)):
//   lambda_max(Sigma') <= trace(Sigma') = a + c,   a <= |j0|^2 |W|_2^2 lambda_max(V) + 0.3 <= |j0|^2 g trace(V) + 0.3,
// c likewise with j1 (j0, j1 the rows of the projection Jacobian, W the 3x3 part of the view matrix with
// |W|_2^2 <= g = the largest absolute row sum of W^T W -- Gershgorin; exactly 1 for a rigid view --, V the world
// covariance, trace(V) = sum_k S_k^2 |a_k|^2 for V = A S^2 A^T), plus a relative allowance for rounding and the
// reference's max(0.1, .) under the square root.  Reads 40 B per Gaussian instead of the 92 - 236 B of the full stage;
// everything that depends on the camera alone is computed once per thread (CullCam).  Measured and not kept: the exact
// radius of the full stage (cov3d_from_scale_rot + cov2d_eval) with two pixels of margin -- fewer candidates at C2
// (cull + per-Gaussian stage of a shard 0.082 -> 0.063 ms) but 110 more instructions per Gaussian, slower where the
// splats are small against a shard (C4: 0.245 -> 0.257 ms).
// The three kernels (flags + counts per 4096-block, scan of the block counts, ordered scatter) give the ascending
// candidate list: the depth sort's stable tie order is the Gaussian index, so the compaction must keep it.
struct CullCam {
    float v[12];   // view matrix, rows 0-2 of the column-major 4x4: v[3 c + r] = view[4 c + r]
    float p1[4], p3[4];  // rows 1 (y) and 3 (w) of the projection matrix
    float g;       // bound on |W|_2^2
};
__device__ __forceinline__ CullCam make_cull_cam(const PreArgs& a) {
    CullCam c;
#pragma unroll
    for (int col = 0; col < 4; col++) {
#pragma unroll
        for (int r = 0; r < 3; r++) c.v[3 * col + r] = a.view[4 * col + r];
        c.p1[col] = a.proj[4 * col + 1];
        c.p3[col] = a.proj[4 * col + 3];
    }
    float g = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {  // row i of W^T W: dot products of the columns of W
        float rs = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++)
            rs += fabsf(c.v[3 * i] * c.v[3 * j] + c.v[3 * i + 1] * c.v[3 * j + 1] + c.v[3 * i + 2] * c.v[3 * j + 2]);
        g = fmaxf(g, rs);
    }
    c.g = g * 1.0001f;
    return c;
}

__device__ __forceinline__ bool shard_candidate(const PreArgs& a, const CullCam& cam, int i) {
    const float3 mean = make_float3(a.means[3 * (size_t)i], a.means[3 * (size_t)i + 1], a.means[3 * (size_t)i + 2]);
    // (all inputs are requested before the near-plane early-out: one round trip to memory per Gaussian)
    float tr_in[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.cov3D_pre != nullptr) {
        tr_in[0] = a.cov3D_pre[6 * (size_t)i]; tr_in[1] = a.cov3D_pre[6 * (size_t)i + 3]; tr_in[2] = a.cov3D_pre[6 * (size_t)i + 5];
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) tr_in[k] = a.rots[4 * (size_t)i + k];
#pragma unroll
        for (int k = 0; k < 3; k++) tr_in[4 + k] = a.scales[3 * (size_t)i + k];
    }
    // view-space position and pixel y: the expressions of xform43 / xform44 / ndc_to_pix (same operand order)
    const float3 p_view = make_float3(cam.v[0] * mean.x + cam.v[3] * mean.y + cam.v[6] * mean.z + cam.v[9],
                                      cam.v[1] * mean.x + cam.v[4] * mean.y + cam.v[7] * mean.z + cam.v[10],
                                      cam.v[2] * mean.x + cam.v[5] * mean.y + cam.v[8] * mean.z + cam.v[11]);
    if (p_view.z <= 0.2f) return false;
    const float hy = cam.p1[0] * mean.x + cam.p1[1] * mean.y + cam.p1[2] * mean.z + cam.p1[3];
    const float hw = cam.p3[0] * mean.x + cam.p3[1] * mean.y + cam.p3[2] * mean.z + cam.p3[3];
    const float p_w = 1.0f / (hw + 0.0000001f);
    const float py = ndc_to_pix(hy * p_w, a.H);
    float tr;  // trace of the world covariance
    if (a.cov3D_pre != nullptr) {
        tr = tr_in[0] + tr_in[1] + tr_in[2];
    } else {
        const float r = tr_in[0], x = tr_in[1], y = tr_in[2], z = tr_in[3];
        const float s0 = a.mod * tr_in[4], s1 = a.mod * tr_in[5], s2 = a.mod * tr_in[6];
        // squared column norms of the (un-normalised) quaternion's rotation matrix
        const float a00 = 1.f - 2.f * (y * y + z * z), a10 = 2.f * (x * y + r * z), a20 = 2.f * (x * z - r * y);
        const float a01 = 2.f * (x * y - r * z), a11 = 1.f - 2.f * (x * x + z * z), a21 = 2.f * (y * z + r * x);
        const float a02 = 2.f * (x * z + r * y), a12 = 2.f * (y * z - r * x), a22 = 1.f - 2.f * (x * x + y * y);
        tr = s0 * s0 * (a00 * a00 + a10 * a10 + a20 * a20) + s1 * s1 * (a01 * a01 + a11 * a11 + a21 * a21) +
             s2 * s2 * (a02 * a02 + a12 * a12 + a22 * a22);
    }
    const float iz = 1.0f / p_view.z;
    const float ux = fminf(1.3f * a.tanx, fabsf(p_view.x * iz)), uy = fminf(1.3f * a.tany, fabsf(p_view.y * iz));
    const float j0 = (a.fx * iz) * (a.fx * iz) * (1.f + ux * ux), j1 = (a.fy * iz) * (a.fy * iz) * (1.f + uy * uy);
    const float lam = (j0 + j1) * cam.g * tr * 1.01f + 1.0f;  // >= a + c + 0.32 with room for rounding
    if (!(lam < 1.0e12f)) return true;                        // overflow / NaN: let the full stage decide
    const float R = ceilf(3.f * sqrtf(lam)) + 1.f;
    const int y0 = min(a.gy, max(0, (int)((py - R) / GS_TILE)));
    const int y1 = min(a.gy, max(0, (int)((py + R + GS_TILE - 1) / GS_TILE)));
    return max(y0, a.row0) < min(y1, a.row1);
}

__global__ void __launch_bounds__(256) shard_flag_kernel(const PreArgs a, uint32_t* __restrict__ cmask,
                                                         uint32_t* __restrict__ ccount) {
    __shared__ unsigned s_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * GS_CULL_CHUNK;
    unsigned mine = 0;
    const CullCam cam = make_cull_cam(a);
#pragma unroll 4
    for (int r = 0; r < GS_CULL_CHUNK / 256; r++) {
        const size_t i = base + (size_t)r * 256 + threadIdx.x;
        const bool keep = i < (size_t)a.P && shard_candidate(a, cam, (int)i);
        const unsigned m = __ballot_sync(GS_FULL, keep);
        if (lane == 0) {
            cmask[(base >> 5) + (size_t)r * 8 + warp] = m;
            mine += __popc(m);
        }
    }
    if (lane == 0) s_cnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int k = 0; k < 8; k++) tot += s_cnt[k];
        ccount[blockIdx.x] = tot;
    }
}

// One CTA: exclusive scan of the block counts (in place), total -> hdr->num_cand.
__global__ void __launch_bounds__(1024) shard_scan_kernel(uint32_t* __restrict__ ccount, int nblocks,
                                                          GsHeader* __restrict__ hdr) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const unsigned v = b < nblocks ? ccount[b] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(GS_FULL, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned pre = s_run;
        for (int w = 0; w < warp; w++) pre += s_w[w];
        if (b < nblocks) ccount[b] = pre + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = pre + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) hdr->num_cand = s_run;
}

__global__ void __launch_bounds__(256) shard_scatter_kernel(const uint32_t* __restrict__ cmask,
                                                            const uint32_t* __restrict__ ccount, int P,
                                                            uint32_t* __restrict__ cand) {
    __shared__ unsigned s_pre[GS_CULL_CHUNK / 32];  // exclusive prefix of the 128 mask words of this block
    __shared__ unsigned s_w[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t wbase = (size_t)blockIdx.x * (GS_CULL_CHUNK / 32);
    unsigned word = 0, incl = 0;
    if (threadIdx.x < GS_CULL_CHUNK / 32) {
        word = cmask[wbase + threadIdx.x];
        incl = __popc(word);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(GS_FULL, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_w[warp] = incl;
    }
    __syncthreads();
    if (threadIdx.x < GS_CULL_CHUNK / 32) {
        unsigned pre = ccount[blockIdx.x];
        for (int w = 0; w < warp; w++) pre += s_w[w];
        s_pre[threadIdx.x] = pre + incl - __popc(word);
    }
    __syncthreads();
    for (int k = warp; k < GS_CULL_CHUNK / 32; k += 8) {  // warp k-th word: lane l <-> Gaussian 32 (wbase + k) + l
        const unsigned m = cmask[wbase + k];
        if ((m >> lane) & 1u) {
            const size_t i = ((wbase + k) << 5) + lane;
            if (i < (size_t)P) cand[s_pre[k] + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
    }
}

// Colour-only pass over an already preprocessed frame (gs_forward_recolor): rewrites rec.c.xyz and the clamp bits.
// A Gaussian culled by preprocess has no tile instance, so its record is never read and needs no update.
__global__ void __launch_bounds__(256) recolor_kernel(int P, int D, int M, const float* __restrict__ means,
                                                      const float* __restrict__ shs, const float* __restrict__ colors_pre,
                                                      const float* __restrict__ campos, const uint32_t* __restrict__ ntile,
                                                      GsRec* __restrict__ rec, uint8_t* __restrict__ clamp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || ntile[i] == 0u) return;
    float3 rgb;
    if (colors_pre != nullptr) {
        rgb = make_float3(colors_pre[3 * i], colors_pre[3 * i + 1], colors_pre[3 * i + 2]);
    } else {
        unsigned clamp_bits = 0;
        const float3 mean = make_float3(means[3 * i], means[3 * i + 1], means[3 * i + 2]);
        rgb = sh_to_rgb(D, shs + (size_t)i * M * 3, mean, make_float3(campos[0], campos[1], campos[2]), clamp_bits);
        clamp[i] = (uint8_t)clamp_bits;
    }
    float* c = reinterpret_cast<float*>(&rec[i].c);
    c[0] = rgb.x; c[1] = rgb.y; c[2] = rgb.z;
}

// Extra colour passes (GsScene.extra_colors): pack up to three [P][3] arrays into 16-B aligned entries that the
// blend kernel can gather with cp.async next to the 48-B record.
__global__ void __launch_bounds__(256) pack_extra_kernel(int P, int K, const float* __restrict__ c0,
                                                         const float* __restrict__ c1, const float* __restrict__ c2,
                                                         float4* __restrict__ xrec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float* src[3] = {c0, c1, c2};
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (k < K) xrec[3 * (size_t)i + k] = make_float4(src[k][3 * i], src[k][3 * i + 1], src[k][3 * i + 2], 0.f);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 p = xform43(view, make_float3(means[3 * i], means[3 * i + 1], means[3 * i + 2]));
    present[i] = p.z > 0.2f;
}

// One thread per view: rigid inverse (R^T, -R^T t), transposes, and the product with the sparse projection matrix.
__global__ void make_views_kernel(const float* __restrict__ c2w, int N, float p00, float p11, float p22, float p23,
                                  float* __restrict__ views) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* h = c2w + 16 * (size_t)n;
    float* o = views + GS_VIEW_STRIDE * (size_t)n;
    // w2c = [R^T | -(R^T t)]; viewmatrix = w2c^T, i.e. vt[i][k] = w2c[k][i]
    float w2c[4][4];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) w2c[i][j] = h[4 * j + i];
        w2c[i][3] = -1.0f * (h[4 * 0 + i] * h[3] + h[4 * 1 + i] * h[7] + h[4 * 2 + i] * h[11]);
    }
    w2c[3][0] = w2c[3][1] = w2c[3][2] = 0.f;
    w2c[3][3] = 1.f;
    for (int i = 0; i < 4; i++) {
        const float v0 = w2c[0][i], v1 = w2c[1][i], v2 = w2c[2][i], v3 = w2c[3][i];  // row i of viewmatrix
        o[4 * i + 0] = v0; o[4 * i + 1] = v1; o[4 * i + 2] = v2; o[4 * i + 3] = v3;
        // row i of viewmatrix * P^T; P^T has (0,0)=p00 (1,1)=p11 (2,2)=p22 (3,2)=p23 (2,3)=1
        o[16 + 4 * i + 0] = v0 * p00;
        o[16 + 4 * i + 1] = v1 * p11;
        o[16 + 4 * i + 2] = v2 * p22 + v3 * p23;
        o[16 + 4 * i + 3] = v2;
    }
    o[32] = h[3]; o[33] = h[7]; o[34] = h[11];
    for (int k = 35; k < GS_VIEW_STRIDE; k++) o[k] = 0.f;
}

// Head decode: 128 points per CTA; the feature rows are staged in shared memory with coalesced loads, then one thread
// decodes one point.  Every operation is a single IEEE fp32 op in the reference's order, so results are bit-identical
// to the torch expressions of model_v2.py:287-375 (clamps written as compares so that NaNs propagate like torch.clamp).
#define HEAD_PTS 128
__global__ void __launch_bounds__(HEAD_PTS) decode_head_kernel(const float* __restrict__ feat,
                                                               const float* __restrict__ rgb,
                                                               const float* __restrict__ prim, int P, GsHeadLayout L,
                                                               float* __restrict__ means3D, float* __restrict__ rot,
                                                               float* __restrict__ scales, float* __restrict__ opac,
                                                               float* __restrict__ shs, float* __restrict__ normals) {
    extern __shared__ float s_feat[];
    const int C = L.C;
    const size_t p0 = (size_t)blockIdx.x * HEAD_PTS;
    const int np = (int)min((size_t)HEAD_PTS, (size_t)P - p0);
    for (int i = threadIdx.x; i < np * C; i += HEAD_PTS) s_feat[i] = feat[p0 * C + i];
    __syncthreads();
    if ((int)threadIdx.x >= np) return;
    const float* f = s_feat + threadIdx.x * C;
    const size_t p = p0 + threadIdx.x;
    // torch divides a CUDA tensor by a Python scalar as a multiplication with the reciprocal (ATen
    // BinaryDivTrueKernel.cu); measured on torch 2.11 the reciprocal is taken in double and then narrowed
    // (tools/div_probe.py).  The reference runs these expressions on the GPU, so that is what is reproduced.
    const float inv_factor = (float)(1.0 / (double)L.xyz_factor), inv_c0 = (float)(1.0 / 0.28209479177387814);
    int used = 0;
    float q0 = 1.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    if (L.use_rotation) { q0 = f[0] + 1.f; q1 = f[1] + 0.f; q2 = f[2] + 0.f; q3 = f[3] + 0.f; used += 4; }
    rot[4 * p + 0] = q0; rot[4 * p + 1] = q1; rot[4 * p + 2] = q2; rot[4 * p + 3] = q3;
    for (int c = 0; c < 3; c++) {
        float sc = 1.f;
        if (L.use_scale) { sc = f[used + c] + 1.f; sc = sc < 0.f ? 0.f : sc; }
        scales[3 * p + c] = sc * L.radius;
    }
    if (L.use_scale) used += 3;
    float o = 1.f;
    if (L.use_opacity) {
        const float v = f[used];
        if (L.enable_opacity) o = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
        used += 1;
    }
    opac[p] = o;
    for (int c = 0; c < 3; c++) {
        float x = prim[3 * p + c];
        if (L.use_offset) x = x + f[used + c];
        means3D[3 * p + c] = __fmul_rn(x - L.xyz_offset, inv_factor);
    }
    if (L.use_offset) used += 3;
    const int M = 1 + L.sh_ac_coeffs;
    for (int c = 0; c < 3; c++) {
        const float dc = __fmul_rn(rgb[3 * p + c] - 0.5f, inv_c0);  // RGB2SH, models/sh_utils.py:114-115
        shs[(size_t)3 * M * p + c] = L.use_dc_offset ? __fadd_rn(f[used + c], dc) : dc;
    }
    if (L.use_dc_offset) used += 3;
    if (L.est_normal) {
        if (normals) {
            float n0 = f[used], n1 = f[used + 1], n2 = f[used + 2];
            if (L.normalize_normal) {
                const float d = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-12f);
                n0 /= d; n1 /= d; n2 /= d;
            }
            normals[3 * p] = n0; normals[3 * p + 1] = n1; normals[3 * p + 2] = n2;
        }
        used += 3;
    }
    for (int k = 0; k < 3 * L.sh_ac_coeffs; k++) shs[(size_t)3 * M * p + 3 + k] = f[used + k];
}

}  // namespace

static PreArgs make_pre_args(const GsFrame& f, const GsGeom& g, const int* rdiff, int32_t* radii) {
    const GsScene& s = f.s;
    PreArgs a;
    a.P = s.P; a.D = s.sh_degree; a.M = s.sh_stride; a.W = s.width; a.H = s.height; a.gx = f.gx; a.gy = f.gy;
    a.row0 = f.row0; a.row1 = f.row1;
    a.tanx = s.tan_fovx; a.tany = s.tan_fovy; a.fx = f.focal_x; a.fy = f.focal_y; a.mod = s.scale_modifier;
    a.prefiltered = s.prefiltered;
    a.means = s.means3D; a.scales = s.scales; a.rots = s.rotations; a.opac = s.opacities; a.shs = s.shs;
    a.cov3D_pre = s.cov3D_precomp; a.colors_pre = s.colors_precomp; a.view = s.viewmatrix; a.proj = s.projmatrix;
    a.campos = s.campos;
    a.radii = radii;
    a.rec = g.rec; a.key = g.key[0]; a.rect = g.rect; a.ntile = g.ntile;
    a.clamp = g.clamp; a.rdiff = const_cast<int*>(rdiff); a.hdr = g.hdr;
    a.cand = f.cull ? g.cand : nullptr;
    return a;
}

cudaError_t gs_launch_shard_cull(const GsFrame& f, const GsGeom& g) {
    if (!f.cull) return cudaSuccess;
    const PreArgs a = make_pre_args(f, g, nullptr, nullptr);
    const unsigned nblocks = (unsigned)g.cull_chunks;
    // ntile of a Gaussian outside the shard must read 0 (gs_forward_recolor, gs_fetch): the full stage will not visit it
    cudaError_t e = cudaMemsetAsync(g.ntile, 0, sizeof(uint32_t) * (size_t)f.s.P, f.stream);
    if (e != cudaSuccess) return e;
    shard_flag_kernel<<<nblocks, 256, 0, f.stream>>>(a, g.cmask, g.ccount);
    gs_note_launch();
    shard_scan_kernel<<<1, 1024, 0, f.stream>>>(g.ccount, (int)nblocks, g.hdr);
    gs_note_launch();
    shard_scatter_kernel<<<nblocks, 256, 0, f.stream>>>(g.cmask, g.ccount, f.s.P, g.cand);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_preprocess(const GsFrame& f, const GsGeom& g, const GsImage& im, int32_t* radii) {
    const GsScene& s = f.s;
    const PreArgs a = make_pre_args(f, g, im.rdiff, radii);
    // TMA-staged variant: scale/rotation + SH inputs, every slab 16-B aligned, and the staging buffer fits
    const size_t stage_bytes = (size_t)(3 + 3 + 4 + 1) * 256 * 4 + (size_t)256 * s.sh_stride * 12;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    // ... and at least half of every SH row is actually read (bulk copies fetch whole rows; the pcrender shape reads 4
    // of its 13 coefficients, there the direct path moves fewer bytes).  Measured on B200: neutral at C2 and C4 --
    // the kernel is bound by its scattered 48-B record stores and per-thread latency, not by the input reads.
    // (Not with shard cull: the candidates of a block are not contiguous.)
    const bool sh_dense = 2 * (s.sh_degree + 1) * (s.sh_degree + 1) > s.sh_stride;
    const bool staged = !f.cull && sh_dense && s.P >= 256 && s.scales && s.rotations && s.shs && !s.cov3D_precomp &&
                        !s.colors_precomp && stage_bytes <= 100 * 1024 && al16(s.means3D) && al16(s.scales) &&
                        al16(s.rotations) && al16(s.opacities) && al16(s.shs);
    if (staged) {
        static GsPerDevice per_dev;
        const int* dv = nullptr;
        cudaError_t e = per_dev.get(&dv, [](int, int*) {
            return cudaFuncSetAttribute(preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        });
        if (e != cudaSuccess) return e;
    }
    if (staged) preprocess_kernel<true><<<(unsigned)gs_div_up(s.P, 256), 256, stage_bytes, f.stream>>>(a);
    else preprocess_kernel<false><<<(unsigned)gs_div_up(s.P, 256), 256, 0, f.stream>>>(a);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_pack_extra(const GsFrame& f, const GsGeom& g) {
    const GsScene& s = f.s;
    if (s.num_extra <= 0) return cudaSuccess;
    pack_extra_kernel<<<(unsigned)gs_div_up(s.P, 256), 256, 0, f.stream>>>(s.P, s.num_extra, s.extra_colors[0],
                                                                          s.extra_colors[1], s.extra_colors[2], g.xrec);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_recolor(const GsFrame& f, const GsGeom& g) {
    const GsScene& s = f.s;
    recolor_kernel<<<(unsigned)gs_div_up(s.P, 256), 256, 0, f.stream>>>(s.P, s.sh_degree, s.sh_stride, s.means3D, s.shs,
                                                                       s.colors_precomp, s.campos, g.ntile, g.rec, g.clamp);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                                   cudaStream_t stream) {
    mark_visible_kernel<<<(unsigned)gs_div_up(P, 256), 256, 0, stream>>>(P, means3D, view, present);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_make_views(const float* c2w, int N, const float* p4, float* views, cudaStream_t stream) {
    if (N <= 0) return cudaSuccess;
    make_views_kernel<<<(unsigned)gs_div_up(N, 64), 64, 0, stream>>>(c2w, N, p4[0], p4[1], p4[2], p4[3], views);
    gs_note_launch();
    return cudaGetLastError();
}

cudaError_t gs_launch_decode_head(const float* feat, const float* rgb, const float* prim, int P, const GsHeadLayout& L,
                                  float* means3D, float* rot, float* scales, float* opac, float* shs, float* normals,
                                  cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    decode_head_kernel<<<(unsigned)gs_div_up(P, HEAD_PTS), HEAD_PTS, (size_t)HEAD_PTS * L.C * sizeof(float), stream>>>(
        feat, rgb, prim, P, L, means3D, rot, scales, opac, shs, normals);
    gs_note_launch();
    return cudaGetLastError();
}
