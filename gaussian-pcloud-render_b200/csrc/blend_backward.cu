// blend_backward.cu -- backward of the per-tile alpha compositing (K7; replaces renderCUDA backward,
// dgr/cuda_rasterizer/backward.cu:399-557).
//
// Same tile/pixel mapping and staging as the forward kernel, walking the tile's list back to front
// (SURVEY App. A item 15).  Two structural changes against the reference, neither of which changes a term of
// any gradient:
//   * the walk starts at need_t = max over the tile's pixels of n_contrib instead of at the end of the list:
//     entries behind every pixel's last contributor are skipped by the reference one by one
//     (`contributor >= last_contributor`), here they are never fetched (typically ~75 % of the list);
//   * the 9 per-(pixel,Gaussian) partial derivatives are summed across the warp with a 14-shuffle transposing
//     reduction before they reach memory, so a Gaussian receives 9 atomics per WARP that touches it instead of
//     9 per PIXEL (backward.cu:523-554): 10-32x fewer L2 atomics.  Summation order differs from the reference's
//     (which is itself non-deterministic); gradients agree to fp32 round-off.
#include "gs_common.cuh"

namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__global__ void __launch_bounds__(GS_TILE_PIX) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ list, const GsRec* __restrict__ rec, int W, int H,
    int gx, int row0, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float* __restrict__ dL_dmean2D,
    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
    __shared__ float4 sA[2][GS_TILE_PIX];
    __shared__ float4 sB[2][GS_TILE_PIX];
    __shared__ float4 sC[2][GS_TILE_PIX];
    __shared__ uint32_t sId[2][GS_TILE_PIX];
    __shared__ uint32_t s_max[GS_TILE_PIX / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + row0;
    const int px = tile_x * GS_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * GS_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    const size_t pid = (size_t)W * py + px;
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[tile_y * gx + tile_x];
    const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;

    // need = deepest last contributor of the tile
    uint32_t m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(GS_FULL, m, o));
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    uint32_t need = 0;
#pragma unroll
    for (int k = 0; k < GS_TILE_PIX / 32; k++) need = max(need, s_max[k]);
    if (need == 0) return;
    const int n = (int)need;
    const int rounds = (n + GS_TILE_PIX - 1) / GS_TILE_PIX;

    const float T_final = inside ? final_T[pid] : 0.f;
    float T = T_final;
    float dpx = 0.f, dpy = 0.f, dpz = 0.f;
    if (inside) {
        dpx = dL_dpix[pid];
        dpy = dL_dpix[plane + pid];
        dpz = dL_dpix[2 * plane + pid];
    }
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;

    // batch i, slot t holds list position p = n-1 - (i*256 + t): slot 0 is the farthest entry
    uint32_t next_id = 0;
    if (tid < n) {
        const uint32_t id = list[range.x + (n - 1 - tid)];
        const GsRec* r = rec + id;
        cp_async16(&sA[0][tid], &r->a);
        cp_async16(&sB[0][tid], &r->b);
        cp_async16(&sC[0][tid], &r->c);
        sId[0][tid] = id;
    }
    cp_async_commit();
    if (GS_TILE_PIX + tid < n) next_id = list[range.x + (n - 1 - (GS_TILE_PIX + tid))];

    int stage = 0;
    for (int i = 0; i < rounds; i++) {
        cp_async_wait_all();
        __syncthreads();
        if (i + 1 < rounds) {
            const int o = (i + 1) * GS_TILE_PIX + tid;
            if (o < n) {
                const GsRec* r = rec + next_id;
                cp_async16(&sA[stage ^ 1][tid], &r->a);
                cp_async16(&sB[stage ^ 1][tid], &r->b);
                cp_async16(&sC[stage ^ 1][tid], &r->c);
                sId[stage ^ 1][tid] = next_id;
            }
            cp_async_commit();
            if (o + GS_TILE_PIX < n) next_id = list[range.x + (n - 1 - (o + GS_TILE_PIX))];
        }
        const int nj = min(GS_TILE_PIX, n - i * GS_TILE_PIX);
        for (int j = 0; j < nj; j++) {
            const uint32_t pos = (uint32_t)(n - 1 - (i * GS_TILE_PIX + j));  // == reference's `contributor`
            bool active = pos < last_contributor;  // false for pixels outside the image (last_contributor = 0)
            const float4 a = sA[stage][j];
            const float4 bq = sB[stage][j];
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
                const float4 c = sC[stage][j];
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

            const uint32_t id = sId[stage][j];
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
        stage ^= 1;
    }
    cp_async_wait_all();
}

}  // namespace

cudaError_t gs_launch_blend_backward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                     const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                     float* dL_dcolor) {
    dim3 grid((unsigned)f.gx, (unsigned)(f.row1 - f.row0), 1);
    if (grid.y == 0 || grid.x == 0) return cudaSuccess;
    blend_backward_kernel<<<grid, GS_TILE_PIX, 0, f.stream>>>(im.ranges, b.list, g.rec, f.s.width, f.s.height, f.gx,
                                                             f.row0, f.s.background, im.final_T, im.n_contrib, dL_dpix,
                                                             dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    gs_note_launch();
    return cudaGetLastError();
}
