// blend_backward.cu -- backward of the per-tile alpha compositing (K7; replaces renderCUDA backward,
// dgr/cuda_rasterizer/backward.cu:399-557).
//
// Same work decomposition as the forward kernel (blend_forward.cu): persistent grid, the unit of work is ONE WARP
// and an 8x4 pixel block, units pulled longest-list-first from an atomic queue, records gathered into a per-warp
// cp.async ring, every lane culls ITS Gaussian against the block with the exact box maximum of the exponent.
// The block's list is walked back to front.  Structural differences against the
// reference, none of which changes a term of any gradient:
//   * the walk starts at the deepest last contributor of the BLOCK's 32 pixels instead of at the end of the tile's
//     list: entries behind every pixel's last contributor are skipped by the reference one by one
//     (`contributor >= last_contributor`), here they are never fetched (typically ~75 % of the list);
//   * instances whose alpha is provably < 1/255 on the whole block never enter the per-pixel loop;
//   * the 9 per-(pixel,Gaussian) partial derivatives are summed across the warp with a 14-shuffle transposing
//     reduction before they reach memory, so a Gaussian receives 9 atomics per WARP that touches it instead of
//     9 per PIXEL (backward.cu:523-554): 10-32x fewer L2 atomics.  Summation order differs from the reference's
//     (which is itself non-deterministic); gradients agree to fp32 round-off (measured 1e-6 relative at C3).
// Measured and rejected: two pixels per lane (8x8 blocks, partials pre-added in the lane, one reduction per 64
// pixels).  106 registers (2 CTAs/SM instead of 3) and two divergent per-pixel regions per hit: backward 2.19 ms
// against 1.77 ms for this kernel at C3 (tools/bench_backward.py, median of 40 views).
#include "gs_common.cuh"

namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Same conservative bound as blend_forward.cu (exact box maximum of the concave exponent + rounding allowance).
__device__ __forceinline__ float box_max_power(float A, float B, float C, float nBA, float nBC, float xlo, float xhi,
                                               float ylo, float yhi) {
    const bool in_x = (xlo <= 0.f) && (xhi >= 0.f);
    const bool in_y = (ylo <= 0.f) && (yhi >= 0.f);
    if (in_x && in_y) return 0.f;
    const float xe = (xlo > 0.f) ? xlo : xhi;
    const float ye = (ylo > 0.f) ? ylo : yhi;
    float best = -3.0e38f;
    if (!in_x) {
        const float y = fminf(yhi, fmaxf(ylo, nBC * xe));
        const float p1 = 0.5f * A * xe * xe, p2 = 0.5f * C * y * y, p3 = B * xe * y;
        best = -(p1 + p2) - p3 + (4.0e-6f * (p1 + p2 + fabsf(p3)) + 0.01f);
    }
    if (!in_y) {
        const float x = fminf(xhi, fmaxf(xlo, nBA * ye));
        const float p1 = 0.5f * A * x * x, p2 = 0.5f * C * ye * ye, p3 = B * x * ye;
        const float v = -(p1 + p2) - p3 + (4.0e-6f * (p1 + p2 + fabsf(p3)) + 0.01f);
        best = fmaxf(best, v);
    }
    return best;
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#define BB_WARPS 8
#define BB_STAGES 3

struct BbStage {
    float4 a[32];  // x, y, conic.x, conic.y
    float4 b[32];  // conic.z, opacity, thr, -B/C
    float4 c[32];  // r, g, b, -B/A
    uint32_t id[32];
};

__global__ void __launch_bounds__(BB_WARPS * 32, 3) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, const GsHeader* __restrict__ hdr,
    unsigned int* __restrict__ queue, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, int ds, float* __restrict__ dL_dmean2D,
    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
    __shared__ BbStage s_ring[BB_WARPS][BB_STAGES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
    BbStage* __restrict__ ring = s_ring[warp];
    const uint32_t num_units = hdr->nonempty_tiles * 8u;  // order[] lists the non-empty tiles first (plan kernel)

    while (true) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(queue, 1u);
        unit = __shfl_sync(GS_FULL, unit, 0);
        if (unit >= num_units) break;
        const uint32_t tile = order[unit >> 3];
        const int sub = unit & 7;
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 4;
        if (bx0 >= W || by0 >= H) continue;
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = px < W && py < H;
        const float pfx = (float)px, pfy = (float)py;
        const size_t pid = (size_t)W * py + px;
        const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;

        // deepest last contributor of the block = where its back-to-front walk starts
        uint32_t need = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) need = max(need, __shfl_xor_sync(GS_FULL, need, o));
        if (need == 0) continue;
        // cull box = bounding box of the pixels that have any contributor
        const unsigned live = __ballot_sync(GS_FULL, last_contributor > 0u);
        const unsigned cols = (live | (live >> 8) | (live >> 16) | (live >> 24)) & 0xffu;
        const unsigned rows = ((live & 0xffu) ? 1u : 0u) | ((live & 0xff00u) ? 2u : 0u) |
                              ((live & 0xff0000u) ? 4u : 0u) | ((live & 0xff000000u) ? 8u : 0u);
        const float fx0 = (float)(bx0 + __ffs(cols) - 1), fx1 = (float)(bx0 + 31 - __clz(cols));
        const float fy0 = (float)(by0 + __ffs(rows) - 1), fy1 = (float)(by0 + 31 - __clz(rows));

        const uint32_t* __restrict__ lst = list + ranges[tile].x;
        const float T_final = inside ? final_T[pid] : 0.f;
        float T = T_final;
        float dpx = 0.f, dpy = 0.f, dpz = 0.f;
        if (inside) {
            if (ds) {  // gradient of the 2x2 box mean (GsScene.downsample): a quarter of the half-resolution pixel's
                const size_t plane2 = (size_t)(H >> 1) * (W >> 1);
                const size_t q = (size_t)(W >> 1) * (py >> 1) + (px >> 1);
                dpx = 0.25f * dL_dpix[q];
                dpy = 0.25f * dL_dpix[plane2 + q];
                dpz = 0.25f * dL_dpix[2 * plane2 + q];
            } else {
                dpx = dL_dpix[pid];
                dpy = dL_dpix[plane + pid];
                dpz = dL_dpix[2 * plane + pid];
            }
        }
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;

        // batch k, slot l holds list position need-1 - (32k + l): slot 0 of batch 0 is the farthest needed entry
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (p * 32 + lane < need) {
                const uint32_t id = lst[need - 1 - (p * 32 + lane)];
                const GsRec* r = rec + id;
                cp_async16(&ring[p].a[lane], &r->a);
                cp_async16(&ring[p].b[lane], &r->b);
                cp_async16(&ring[p].c[lane], &r->c);
                ring[p].id[lane] = id;
            }
            cp_async_commit();
        }
        uint32_t id_next = (64 + lane < need) ? lst[need - 1 - (64 + lane)] : 0u;

        int stage = 0;
        for (uint32_t base = 0; base < need; base += 32) {
            cp_async_wait<1>();
            __syncwarp();
            {
                int nst = stage + 2; if (nst >= BB_STAGES) nst -= BB_STAGES;
                if (base + 64 + lane < need) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&ring[nst].a[lane], &r->a);
                    cp_async16(&ring[nst].b[lane], &r->b);
                    cp_async16(&ring[nst].c[lane], &r->c);
                    ring[nst].id[lane] = id_next;
                }
                cp_async_commit();
                if (base + 96 + lane < need) id_next = lst[need - 1 - (base + 96 + lane)];
            }
            const BbStage& st = ring[stage];
            stage = (stage + 1 == BB_STAGES) ? 0 : stage + 1;

            bool hit = false;
            if (base + lane < need) {
                const float4 a = st.a[lane], b = st.b[lane];
                const float nBA = st.c[lane].w;
                const float bound = box_max_power(a.z, a.w, b.x, nBA, b.w, a.x - fx1, a.x - fx0, a.y - fy1, a.y - fy0);
                hit = !(bound < b.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const uint32_t pos = need - 1u - (base + (uint32_t)j);  // == reference's `contributor`
                bool active = pos < last_contributor;  // false for pixels outside the image (last_contributor = 0)
                const float4 a = st.a[j];
                const float4 bq = st.b[j];
                const float dx = a.x - pfx, dy = a.y - pfy;
                const float power = -0.5f * (a.z * dx * dx + bq.x * dy * dy) - a.w * dx * dy;
                active = active && !(power > 0.0f) && !(power < bq.z);
                float G = 0.f, alpha = 0.f;
                if (active) {
                    G = expf(power);
                    alpha = fminf(0.99f, bq.y * G);
                    active = !(alpha < 1.0f / 255.0f);
                }
                if (!__any_sync(GS_FULL, active)) continue;

                float v[8], v8 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = 0.f;
                if (active) {
                    const float4 c = st.c[j];
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = c.x; dL_dalpha += (c.x - acc0) * dpx;
                    acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = c.y; dL_dalpha += (c.y - acc1) * dpy;
                    acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = c.z; dL_dalpha += (c.z - acc2) * dpz;
                    v[0] = dchannel_dcolor * dpx;
                    v[1] = dchannel_dcolor * dpy;
                    v[2] = dchannel_dcolor * dpz;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot_dpixel = 0;
                    bg_dot_dpixel += bg0 * dpx;
                    bg_dot_dpixel += bg1 * dpy;
                    bg_dot_dpixel += bg2 * dpz;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = bq.y * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * a.z - gdy * a.w;
                    const float dG_ddely = -gdy * bq.x - gdx * a.w;
                    v[3] = dL_dG * dG_ddelx * ddelx_dx;
                    v[4] = dL_dG * dG_ddely * ddely_dy;
                    v[5] = -0.5f * gdx * dx * dL_dG;
                    v[6] = -0.5f * gdx * dy * dL_dG;
                    v[7] = -0.5f * gdy * dy * dL_dG;
                    v8 = G * dL_dalpha;
                }
                // transposing warp reduction: 8 values -> lane 4k holds the warp total of value k (4+2+1+1+1 shuffles)
                float w4[4], w2[2], w1;
                {
                    const bool hi = (lane & 16) != 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float send = hi ? v[k] : v[k + 4];
                        const float keep = hi ? v[k + 4] : v[k];
                        w4[k] = keep + __shfl_xor_sync(GS_FULL, send, 16);
                    }
                }
                {
                    const bool hi = (lane & 8) != 0;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const float send = hi ? w4[k] : w4[k + 2];
                        const float keep = hi ? w4[k + 2] : w4[k];
                        w2[k] = keep + __shfl_xor_sync(GS_FULL, send, 8);
                    }
                }
                {
                    const bool hi = (lane & 4) != 0;
                    const float send = hi ? w2[0] : w2[1];
                    const float keep = hi ? w2[1] : w2[0];
                    w1 = keep + __shfl_xor_sync(GS_FULL, send, 4);
                }
                w1 += __shfl_xor_sync(GS_FULL, w1, 2);
                w1 += __shfl_xor_sync(GS_FULL, w1, 1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(GS_FULL, v8, o);

                const uint32_t id = st.id[j];
                if ((lane & 3) == 0) {
                    const int k = lane >> 2;  // value index
                    float* dst;
                    if (k < 3) dst = dL_dcolor + 3 * (size_t)id + k;
                    else if (k < 5) dst = dL_dmean2D + 3 * (size_t)id + (k - 3);
                    else dst = dL_dconic + 4 * (size_t)id + (k == 7 ? 3 : k - 5);
                    if (w1 != 0.f) atomicAdd(dst, w1);
                } else if (lane == 1) {
                    if (v8 != 0.f) atomicAdd(dL_dopacity + id, v8);
                }
            }
        }
        cp_async_wait<0>();
    }
}

GsPerDevice g_bwd_dev;  // value[0] = resident CTAs of the kernel on this device

}  // namespace

cudaError_t gs_launch_blend_backward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                     const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                     float* dL_dcolor) {
    const uint32_t num_tiles = (uint32_t)f.gx * (uint32_t)(f.row1 - f.row0);
    if (num_tiles == 0) return cudaSuccess;
    const int* dv = nullptr;
    cudaError_t e = g_bwd_dev.get(&dv, [](int dev, int* v) {
        int sms = 0, per_sm = 0;
        cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_backward_kernel, BB_WARPS * 32, 0);
        if (e != cudaSuccess) return e;
        v[0] = sms * (per_sm > 0 ? per_sm : 1);
        return cudaSuccess;
    });
    if (e != cudaSuccess) return e;
    const int g_bwd_grid = dv[0];
    unsigned int* queue = &g.hdr->tickets[10];  // the forward pass left the header in place; only the queue restarts
    e = cudaMemsetAsync(queue, 0, sizeof(unsigned int), f.stream);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)min((uint32_t)g_bwd_grid, num_tiles);
    blend_backward_kernel<<<grid, BB_WARPS * 32, 0, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height,
                                                               f.gx, g.hdr, queue, f.s.background, im.final_T,
                                                               im.n_contrib, dL_dpix, f.s.downsample == 2 ? 1 : 0, dL_dmean2D,
                                                               dL_dconic, dL_dopacity,
                                                               dL_dcolor);
    gs_note_launch();
    return cudaGetLastError();
}
