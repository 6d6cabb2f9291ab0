// blend_forward.cu -- per-tile front-to-back alpha compositing (K6; replaces renderCUDA,
// dgr/cuda_rasterizer/forward.cu:264-377).
//
// Work decomposition (B200: 148 SMs, one persistent grid):
//   * the unit of work is ONE WARP blending an 8x4 pixel block (1/8 of a 16x16 tile) against the tile's
//     depth-ordered instance list; warps never synchronise with each other (no __syncthreads in the loop: the
//     reference's batch barrier was the top stall reason and its CTA-per-tile grid left the average SM idle 60 % of
//     the frame behind a few silhouette tiles);
//   * units are handed out through a global atomic queue in descending order of list length (longest-processing-
//     time-first), 8 consecutive units = the 8 blocks of one tile, so warps of one CTA walk the same list and share
//     it through L1;
//   * each lane gathers one packed 48-B record per step (3 x 16-B loads, table is L2 resident), software-pipelined
//     one batch ahead (indices two batches ahead);
//   * before blending, every lane tests ITS Gaussian against the warp's 8x4 block with an exact box-maximum of the
//     (concave) exponent; instances whose alpha is provably < 1/255 on the whole block are dropped by a ballot and
//     never enter the per-pixel loop.  The bound is conservative (it only removes evaluations the reference
//     skips through `alpha < 1/255`), so the image is unchanged.
// Blending semantics are exactly SURVEY App. A item 14; expression order of power / alpha / colour accumulation is
// the reference's, so results are bit-identical to the reference kernels compiled for the same GPU.
//
// Bound: FP32 issue + MUFU, not HBM.  Algorithmic HBM bytes: 40*sum(need_t) + 20*N + 8*Tn.
#include "gs_common.cuh"

namespace {

#define BF_WARPS 8

// Upper bound of power(d) = -0.5 (A dx^2 + C dy^2) - B dx dy over the box [xlo,xhi] x [ylo,yhi] of d = mean - pixel,
// plus a rounding allowance.  Exact box maximum of a concave quadratic: 0 if the box contains the origin, otherwise
// the best of the 1-D maxima on the (at most two) box edges that face the origin.  nBA = -B/A, nBC = -B/C
// (NaN when the conic is not positive definite -> the comparison in the caller fails -> instance is kept).
__device__ __forceinline__ float box_max_power(float A, float B, float C, float nBA, float nBC, float xlo, float xhi,
                                               float ylo, float yhi) {
    const bool in_x = (xlo <= 0.f) && (xhi >= 0.f);
    const bool in_y = (ylo <= 0.f) && (yhi >= 0.f);
    if (in_x && in_y) return 0.f;
    const float xe = (xlo > 0.f) ? xlo : xhi;  // box x nearest to 0 (only used when !in_x)
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

__global__ void __launch_bounds__(BF_WARPS * 32, 4) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_units, unsigned int* __restrict__ queue,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    float* __restrict__ out_color) {
    __shared__ float4 sA[BF_WARPS][32];  // x, y, conic.x, conic.y
    __shared__ float4 sB[BF_WARPS][32];  // conic.z, opacity, thr, -B/C
    __shared__ float4 sC[BF_WARPS][32];  // r, g, b, -B/A

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    float4* __restrict__ wA = sA[warp];
    float4* __restrict__ wB = sB[warp];
    float4* __restrict__ wC = sC[warp];

    while (true) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(queue, 1u);
        unit = __shfl_sync(GS_FULL, unit, 0);
        if (unit >= num_units) break;
        const uint32_t tile = order[unit >> 3];
        const int sub = unit & 7;
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 4;
        if (bx0 >= W || by0 >= H) continue;  // block entirely outside the image
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        const bool inside = px < W && py < H;
        const float pfx = (float)px, pfy = (float)py;
        const float fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 3);

        const uint2 range = ranges[tile];
        const uint32_t total = range.y - range.x;

        bool done = !inside;
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
        uint32_t last_contributor = 0;

        // software pipeline: records one batch ahead, list indices two batches ahead
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na, nc = na;
        uint32_t id2 = 0;
        if (lane < total) {
            const uint32_t id = list[range.x + lane];
            const float4* r = reinterpret_cast<const float4*>(rec + id);
            na = __ldg(r); nb = __ldg(r + 1); nc = __ldg(r + 2);
        }
        if (32 + lane < total) id2 = list[range.x + 32 + lane];

        for (uint32_t base = 0; base < total; base += 32) {
            const float4 a = na, b = nb, c = nc;
            const bool have = base + lane < total;
            if (base + 32 + lane < total) {
                const float4* r = reinterpret_cast<const float4*>(rec + id2);
                na = __ldg(r); nb = __ldg(r + 1); nc = __ldg(r + 2);
            }
            if (base + 64 + lane < total) id2 = list[range.x + base + 64 + lane];

            // conservative cull of this lane's Gaussian against the warp's 8x4 pixel block
            bool hit = false;
            if (have) {
                const float bound = box_max_power(a.z, a.w, b.x, c.w, b.w, a.x - fx1, a.x - fx0, a.y - fy1, a.y - fy0);
                hit = !(bound < b.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
            if (mask == 0) continue;
            __syncwarp();  // previous batch's readers are done with the staging rows
            wA[lane] = a; wB[lane] = b; wC[lane] = c;
            __syncwarp();
            if (!done) {
                do {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4 ga = wA[j];
                    const float4 gb = wB[j];
                    const float dx = ga.x - pfx, dy = ga.y - pfy;
                    const float power = -0.5f * (ga.z * dx * dx + gb.x * dy * dy) - ga.w * dx * dy;
                    if (power > 0.0f) continue;
                    if (power < gb.z) continue;  // provably alpha < 1/255
                    const float alpha = fminf(0.99f, gb.y * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) {
                        done = true;
                        break;
                    }
                    const float4 gc = wC[j];
                    C0 += gc.x * alpha * T;
                    C1 += gc.y * alpha * T;
                    C2 += gc.z * alpha * T;
                    T = test_T;
                    last_contributor = base + (uint32_t)j + 1u;
                } while (mask);
            }
            if (__ballot_sync(GS_FULL, !done) == 0) break;
        }

        if (inside) {
            const size_t pid = (size_t)W * py + px;
            final_T[pid] = T;
            n_contrib[pid] = last_contributor;
            out_color[pid] = C0 + T * bg0;
            out_color[plane + pid] = C1 + T * bg1;
            out_color[2 * plane + pid] = C2 + T * bg2;
        }
    }
}

int g_blend_grid = 0;

}  // namespace

cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color) {
    const uint32_t num_units = (uint32_t)f.gx * (uint32_t)(f.row1 - f.row0) * 8u;
    if (num_units == 0) return cudaSuccess;
    if (g_blend_grid == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_kernel, BF_WARPS * 32, 0);
        if (e != cudaSuccess) return e;
        g_blend_grid = sms * (per_sm > 0 ? per_sm : 1);
    }
    const unsigned grid = (unsigned)min((uint32_t)g_blend_grid, (num_units + BF_WARPS - 1) / BF_WARPS);
    blend_forward_kernel<<<grid, BF_WARPS * 32, 0, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height,
                                                              f.gx, num_units, &g.hdr->tickets[6], f.s.background,
                                                              im.final_T, im.n_contrib, out_color);
    gs_note_launch();
    return cudaGetLastError();
}
