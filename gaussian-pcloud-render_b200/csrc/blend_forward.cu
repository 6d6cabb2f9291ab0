// blend_forward.cu -- per-tile front-to-back alpha compositing (K6; replaces renderCUDA,
// dgr/cuda_rasterizer/forward.cu:264-377).
//
// One CTA of 256 threads per 16x16 tile, one thread per pixel; warp w covers an 8x4 pixel block so that a
// Gaussian's footprint is warp-coherent.  The tile's depth-ordered instance list is consumed in batches of 256:
// every thread gathers ONE packed 48-B record (3 x cp.async 16 B, L2-resident table) of the NEXT batch into a
// double-buffered shared-memory ring while the current batch is blended, so the gather latency is hidden behind
// the arithmetic.  Colours are staged too (the reference reads them from global memory inside the inner loop,
// forward.cu:358).  Blending semantics are exactly SURVEY App. A item 14: skip power > 0, alpha = min(.99,
// o*exp(power)), skip alpha < 1/255, stop before T*(1-alpha) < 1e-4; the extra `power < thr` test only skips
// evaluations whose alpha is provably < 1/255 (thr is computed per Gaussian in preprocess), which removes the
// exponential from the common far-field case without changing any result.
//
// Bound: FP32 issue + MUFU, not HBM.  Algorithmic HBM bytes: 40*sum(need_t) + 20*N + 8*Tn.
#include "gs_common.cuh"

namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__global__ void __launch_bounds__(GS_TILE_PIX) blend_forward_kernel(const uint2* __restrict__ ranges,
                                                                   const uint32_t* __restrict__ list,
                                                                   const GsRec* __restrict__ rec, int W, int H, int gx,
                                                                   int row0, const float* __restrict__ bg,
                                                                   float* __restrict__ final_T,
                                                                   uint32_t* __restrict__ n_contrib,
                                                                   float* __restrict__ out_color) {
    __shared__ float4 sA[2][GS_TILE_PIX];  // x, y, conic.x, conic.y
    __shared__ float4 sB[2][GS_TILE_PIX];  // conic.z, opacity, thr, depth
    __shared__ float4 sC[2][GS_TILE_PIX];  // r, g, b, -

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + row0;
    const int px = tile_x * GS_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * GS_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;

    const uint2 range = ranges[tile_y * gx + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + GS_TILE_PIX - 1) / GS_TILE_PIX;

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last_contributor = 0;

    // prologue: gather batch 0, prefetch the indices of batch 1
    uint32_t next_id = 0;
    if (tid < total) {
        const uint32_t id = list[range.x + tid];
        const GsRec* r = rec + id;
        cp_async16(&sA[0][tid], &r->a);
        cp_async16(&sB[0][tid], &r->b);
        cp_async16(&sC[0][tid], &r->c);
    }
    cp_async_commit();
    if (GS_TILE_PIX + tid < total) next_id = list[range.x + GS_TILE_PIX + tid];

    int stage = 0;
    int toDo = total;
    for (int b = 0; b < rounds; b++, toDo -= GS_TILE_PIX) {
        cp_async_wait_all();
        // one barrier: (a) batch b has landed for every thread, (b) everybody left batch b-1, (c) early-out vote
        if (__syncthreads_count(done) == GS_TILE_PIX) break;
        if (b + 1 < rounds) {
            const int o = (b + 1) * GS_TILE_PIX + tid;
            if (o < total) {
                const GsRec* r = rec + next_id;
                cp_async16(&sA[stage ^ 1][tid], &r->a);
                cp_async16(&sB[stage ^ 1][tid], &r->b);
                cp_async16(&sC[stage ^ 1][tid], &r->c);
            }
            cp_async_commit();
            if (o + GS_TILE_PIX < total) next_id = list[range.x + o + GS_TILE_PIX];
        }
        const int nj = min(GS_TILE_PIX, toDo);
        const float4* __restrict__ A = sA[stage];
        const float4* __restrict__ B = sB[stage];
        const float4* __restrict__ Cc = sC[stage];
        for (int j = 0; !done && j < nj; j++) {
            const float4 a = A[j];
            const float4 bq = B[j];
            const float dx = a.x - pfx, dy = a.y - pfy;
            const float power = -0.5f * (a.z * dx * dx + bq.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            if (power < bq.z) continue;  // provably alpha < 1/255
            const float alpha = fminf(0.99f, bq.y * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            const float4 c = Cc[j];
            C0 += c.x * alpha * T;
            C1 += c.y * alpha * T;
            C2 += c.z * alpha * T;
            T = test_T;
            last_contributor = (uint32_t)(b * GS_TILE_PIX + j + 1);
        }
        stage ^= 1;
    }
    cp_async_wait_all();  // nothing may be in flight when the CTA retires

    if (inside) {
        const size_t pid = (size_t)W * py + px;
        const size_t plane = (size_t)H * W;
        final_T[pid] = T;
        n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * bg[0];
        out_color[plane + pid] = C1 + T * bg[1];
        out_color[2 * plane + pid] = C2 + T * bg[2];
    }
}

}  // namespace

cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color) {
    dim3 grid((unsigned)f.gx, (unsigned)(f.row1 - f.row0), 1);
    if (grid.y == 0 || grid.x == 0) return cudaSuccess;
    blend_forward_kernel<<<grid, GS_TILE_PIX, 0, f.stream>>>(im.ranges, b.list, g.rec, f.s.width, f.s.height, f.gx,
                                                            f.row0, f.s.background, im.final_T, im.n_contrib,
                                                            out_color);
    gs_note_launch();
    return cudaGetLastError();
}
