// blend_forward.cu -- per-tile front-to-back alpha compositing (K6; replaces renderCUDA,
// dgr/cuda_rasterizer/forward.cu:264-377).
//
// Work decomposition (B200: 148 SMs, one persistent grid of 3 CTAs x 8 warps per SM):
//   * the unit of work is ONE WARP blending an 8x8 pixel block (1/4 of a 16x16 tile) against the tile's
//     depth-ordered instance list, two pixels per lane; warps never synchronise with each other (the reference's
//     per-batch __syncthreads was its top stall reason and its CTA-per-tile grid left the average SM idle 60 % of
//     the frame behind a few silhouette tiles);
//   * units are handed out through a global atomic queue in descending order of list length (longest-processing-
//     time-first); empty tiles (85 % of a THuman frame) are not blended at all: one warp fills a whole tile with
//     the background;
//   * each lane gathers one packed 48-B record per step straight into a per-warp shared-memory ring with cp.async
//     (3 x 16 B, no register staging; the table is L2 resident), two batches ahead, list indices three ahead;
//   * before blending, every lane tests ITS Gaussian against the warp's pixel block with an exact box-maximum of
//     the (concave) exponent; instances whose alpha is provably < 1/255 on the whole block are dropped by a ballot
//     and never enter the per-pixel loop (the reference skips those per pixel through `alpha < 1/255`, so the image
//     is unchanged).  Every 16 batches the box shrinks to the pixels that are still live;
//   * lane (lx, ly) owns pixels (lx, ly) and (lx, ly + 4).  They share dx, so the exponent of both costs 3 scalar +
//     6 packed FP32x2 instructions (FADD2 / FMUL2 / FFMA2 of sm_100), the colour accumulation 6 packed ones; every
//     packed operation is, per half, the scalar operation the reference executes (same operands, same order, same
//     fused multiply-adds, libdevice expf), so results are bit-identical to the reference kernels compiled for the
//     same GPU (tests compare n_contrib exactly and pixels to 1e-6).
//   * tile-row sharding with peer stores: the epilogue can write every pixel into the images of all ranks (NVLink
//     peer mappings), so the frame is assembled on every GPU by this kernel and only a barrier follows.
// A few silhouette blocks walk 10-20 K-entry lists without saturating and set the single-frame kernel time (700 us
// for one block while 95 % of all blocks are done after 250 us); the tail is hidden by keeping several frames in
// flight (renderer.FramePipeline).  Measured on B200 at C2 and rejected: handing half of a long block's live pixels
// to idle warps through a split queue (slower: the critical path is the list walk, which pixel splitting
// replicates), one instance per iteration with per-lane early outs (2.6x slower: the divergent loop stops
// reconverging per instance), a warp vote that skips the exponentials nobody needs (+6 %), a packed FP32x2
// transcription of expf (bit-identical, not faster: FP32x2 saves issue slots, not FMA-pipe cycles), 4 or 5 CTAs
// per SM instead of 3 (slower tail, no throughput gain).
// Round 2, three bit-identical restructurings of the tail, all measured and NOT kept (profiles/r02a_*): (1) "lane per
// Gaussian" -- once <= 24 pixels of a block are live, queue the surviving instances in a shared-memory FIFO, let lane j
// evaluate alpha of instance j for every live pixel (16 instances x 2 half-warps, four pixels in flight per lane) and
// let lane r walk the 16 alphas of ITS pixel in list order with the reference's exact T / C recurrence: 2.4x fewer
// cycles per instance (80 against 190), but the longest blocks keep > 32 live pixels for the first half of their walk
// (silhouette blocks lie mostly OUTSIDE the body), and FIFO + alpha matrix take the kernel from 37 to 74 KB of shared
// memory per CTA, which leaves no room for the binning kernels of the next frames: single frame 0.60 -> 0.60 ms,
// six frames in flight -17 %; (2) evaluating the alphas of 4 or 8 queued instances together before blending them in
// order (independent expf chains): +45 % instructions, the same 0.135 us per instance; (3) producer / consumer warp
// pairs (one warp walks and culls, its partner blends; FIFO, named barrier per pair): the consumer is the bottleneck,
// 0.63-0.69 ms.  What the per-unit timelines say: the long blocks are the FIRST units started, share their scheduler
// with five other warps for the 250 us in which the queue of fresh units drains -- whatever their instruction-level
// parallelism, they get a sixth of the issue slots -- and half of all unit-time is still outstanding at that point,
// spread over ~1000 half-finished walks.  Shortening the frame below ~0.45 ms therefore needs the remaining walks
// re-distributed over the idle warps (alpha digests handed from helper warps to the owning warp through L2), not a
// faster walk; not built.
//
// Bound: FP32 issue, not HBM.  Algorithmic HBM bytes: 40*sum(need_t) + 20*N + 8*Tn.
#include "gs_common.cuh"

namespace {

#define BF_WARPS 8

#ifdef GS_TIMELINE  // developer build only (make timeline): per-unit start/end time, SM, batches, hits
__device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}
#endif

// Upper bound of power(d) = -0.5 (A dx^2 + C dy^2) - B dx dy over the box [xlo,xhi] x [ylo,yhi] of d = mean - pixel,
// plus a rounding allowance.  Exact box maximum of a concave quadratic: 0 if the box contains the origin, otherwise
// the best of the 1-D maxima on the (at most two) box edges that face the origin.  nBA = -B/A, nBC = -B/C
// (only meaningful for a positive definite conic; preprocess stores the cut-off -inf otherwise, so the caller's
// `bound < cut-off` is false and the instance is kept whatever this function returns).
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#define BF_STAGES 3  // per-warp ring of 32-record batches: one being blended, two in flight

template <int K>
struct BfStage {
    float4 a[32];  // x, y, conic.x, conic.y
    float4 b[32];  // conic.z, opacity, thr, -B/C
    float4 c[32];  // r, g, b, -B/A
    float4 e[K > 0 ? K : 1][32];  // colours of the extra passes (K > 0)
};

// Images of the extra colour passes blended in the same list walk (GsScene.extra_colors / extra_out).
struct BfExtra {
    const float4* xrec;  // [P][3]
    float* out[3];
};

#define BF_CHECK 16  // batches between two looks at which pixels of the block are still live

// Output images of the epilogue: this rank's out_color, or (tile-row sharding with peer stores) the images of all
// ranks, addressed through NVLink peer mappings -- the blend kernel assembles the frame on every GPU itself.
struct BfTargets {
    int n;
    int ds;  // 1: the epilogue stores the 2x2 box mean of the frame, images are [3][H/2][W/2] (GsScene.downsample)
    float* img[8];
};
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void fma2_acc(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// 2x2 box mean of a value held by the four lanes (lx, lx^1) x (ly, ly^1) of a block: bit for bit what
// F.interpolate(mode="bilinear", align_corners=False) computes when halving an image (all four weights are 0.5 * 0.5;
// simple_raw_render.py:281-284).
__device__ __forceinline__ float box4(float v) {
    const float h = v + __shfl_xor_sync(GS_FULL, v, 1);
    return 0.25f * (h + __shfl_xor_sync(GS_FULL, h, 8));
}

#ifndef BF_THR_TEST
#define BF_THR_TEST 0
#endif
#ifndef BF_DONE_PER_HIT
#define BF_DONE_PER_HIT 0
#endif
#ifndef BF_SLOT_NUM  // fraction of the CTA slots the blend grid may occupy
#define BF_SLOT_NUM 3
#define BF_SLOT_DEN 5
#endif
#ifndef BF_PX2_OCC
#define BF_PX2_OCC 3
#endif
template <int K>
__global__ void __launch_bounds__(BF_WARPS * 32, (K == 0) ? BF_PX2_OCC : 2) blend_forward_px2_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_tiles, GsHeader* __restrict__ hdr,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    const BfTargets tg, const BfExtra ex, int quota) {
    extern __shared__ __align__(16) unsigned char s_ring_raw[];
    BfStage<K>(*s_ring)[BF_STAGES] = reinterpret_cast<BfStage<K>(*)[BF_STAGES]>(s_ring_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    BfStage<K>* __restrict__ ring = s_ring[warp];
    unsigned* const q_fresh = &hdr->tickets[6];
    const uint32_t nonempty = hdr->nonempty_tiles;
    const uint32_t blend_units = nonempty * 4u;
    const uint32_t num_units = blend_units + (num_tiles - nonempty);

    // Every warp serves at most `quota` units and the grid is sized for ~60 % of the CTA slots of the GPU: with several
    // frames in flight the short, latency-bound binning kernels of the next frames then find room next to this
    // kernel instead of waiting for its persistent CTAs to drain (+15 % frames/s at C2, same single-frame time).
    for (int served = 0; served < quota; served++) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(q_fresh, 1u);
        unit = __shfl_sync(GS_FULL, unit, 0);
        if (unit >= num_units) break;
        if (unit >= blend_units) {  // ---- empty tile: colour = background, T = 1, no contributor
            const uint32_t tile = order[nonempty + (unit - blend_units)];
            const int x0 = (int)(tile % gx) * GS_TILE, y0 = (int)(tile / gx) * GS_TILE;
            if (tg.ds) {  // half-resolution colour (8x8 per tile), full-resolution T / contributor count
                const int W2 = W >> 1, H2 = H >> 1;
                const size_t plane2 = (size_t)W2 * H2;
                const int qx = (x0 >> 1) + (lane & 7);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int qy = (y0 >> 1) + r * 4 + (lane >> 3);
                    if (qx < W2 && qy < H2) {
                        const size_t pid = (size_t)W2 * qy + qx;
                        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                            float* oc = tg.img[k];
                            oc[pid] = bg0; oc[plane2 + pid] = bg1; oc[2 * plane2 + pid] = bg2;
                        }
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            float* oc = ex.out[k];
                            oc[pid] = bg0; oc[plane2 + pid] = bg1; oc[2 * plane2 + pid] = bg2;
                        }
                    }
                }
                for (int r = 0; r < 8; r++) {
                    const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
                    if (px < W && py < H) {
                        const size_t pid = (size_t)W * py + px;
                        final_T[pid] = 1.f; n_contrib[pid] = 0u;
                    }
                }
            } else
            if ((W & 3) == 0 && x0 + GS_TILE <= W) {
                const int px = x0 + (lane & 3) * 4;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int py = y0 + r * 8 + (lane >> 2);
                    if (py < H) {
                        const size_t pid = (size_t)W * py + px;
                        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                            float* oc = tg.img[k];
                            *reinterpret_cast<float4*>(oc + pid) = make_float4(bg0, bg0, bg0, bg0);
                            *reinterpret_cast<float4*>(oc + plane + pid) = make_float4(bg1, bg1, bg1, bg1);
                            *reinterpret_cast<float4*>(oc + 2 * plane + pid) = make_float4(bg2, bg2, bg2, bg2);
                        }
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            float* oc = ex.out[k];
                            *reinterpret_cast<float4*>(oc + pid) = make_float4(bg0, bg0, bg0, bg0);
                            *reinterpret_cast<float4*>(oc + plane + pid) = make_float4(bg1, bg1, bg1, bg1);
                            *reinterpret_cast<float4*>(oc + 2 * plane + pid) = make_float4(bg2, bg2, bg2, bg2);
                        }
                        *reinterpret_cast<float4*>(final_T + pid) = make_float4(1.f, 1.f, 1.f, 1.f);
                        *reinterpret_cast<uint4*>(n_contrib + pid) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                for (int r = 0; r < 8; r++) {
                    const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
                    if (px < W && py < H) {
                        const size_t pid = (size_t)W * py + px;
                        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                            float* oc = tg.img[k];
                            oc[pid] = bg0; oc[plane + pid] = bg1; oc[2 * plane + pid] = bg2;
                        }
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            float* oc = ex.out[k];
                            oc[pid] = bg0; oc[plane + pid] = bg1; oc[2 * plane + pid] = bg2;
                        }
                        final_T[pid] = 1.f; n_contrib[pid] = 0u;
                    }
                }
            }
            continue;
        }
        const uint32_t tile = order[unit >> 2];
        const int sub = unit & 3;
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 8;
        if (bx0 >= W || by0 >= H) continue;  // block entirely outside the image
        const int px = bx0 + lx, pyA = by0 + ly, pyB = by0 + ly + 4;
        const bool insA = px < W && pyA < H, insB = px < W && pyB < H;
        const float pfx = (float)px;
        const f2 pfy2 = pk((float)pyA, (float)pyB);
        float fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 7);

        const uint2 range = ranges[tile];
        const uint32_t total = range.y - range.x;
        const uint32_t* __restrict__ lst = list + range.x;

#ifdef GS_TIMELINE
        const unsigned long long tl_t0 = gtime();
        unsigned tl_batches = 0, tl_hits = 0;
        long long tl_wait = 0, tl_loop = 0;
#endif
        bool doneA = !insA, doneB = !insB;
        f2 T2 = bc(1.0f);
        float c0A = 0.f, c0B = 0.f, c1A = 0.f, c1B = 0.f, c2A = 0.f, c2B = 0.f;
        f2 E[K > 0 ? K : 1][3];
#pragma unroll
        for (int k = 0; k < K; k++) { E[k][0] = bc(0.f); E[k][1] = bc(0.f); E[k][2] = bc(0.f); }
        uint32_t lastA = 0, lastB = 0;

        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (p * 32 + lane < total) {
                const uint32_t id = lst[p * 32 + lane];
                const GsRec* r = rec + id;
                cp_async16(&ring[p].a[lane], &r->a);
                cp_async16(&ring[p].b[lane], &r->b);
                cp_async16(&ring[p].c[lane], &r->c);
#pragma unroll
                for (int k = 0; k < K; k++) cp_async16(&ring[p].e[k][lane], ex.xrec + 3 * (size_t)id + k);
            }
            cp_async_commit();
        }
        uint32_t id_next = (64 + lane < total) ? lst[64 + lane] : 0u;

        int stage = 0;
        uint32_t next_check = BF_CHECK * 32;
        for (uint32_t base = 0; base < total; base += 32) {
            if (base == next_check) {  // shrink the cull box to the pixels that are still live
                next_check += BF_CHECK * 32;
                const unsigned aliveA = __ballot_sync(GS_FULL, !doneA), aliveB = __ballot_sync(GS_FULL, !doneB);
                const unsigned both = aliveA | aliveB;
                const unsigned cols = (both | (both >> 8) | (both >> 16) | (both >> 24)) & 0xffu;
                unsigned rows = 0;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    rows |= ((aliveA >> (8 * r)) & 0xffu) ? (1u << r) : 0u;
                    rows |= ((aliveB >> (8 * r)) & 0xffu) ? (16u << r) : 0u;
                }
                fx0 = (float)(bx0 + __ffs(cols) - 1); fx1 = (float)(bx0 + 31 - __clz(cols));
                fy0 = (float)(by0 + __ffs(rows) - 1); fy1 = (float)(by0 + 31 - __clz(rows));
            }
#ifdef GS_TIMELINE
            tl_batches++;
            const long long tl_c0 = clock64();
#endif
            cp_async_wait<1>();
            __syncwarp();
#ifdef GS_TIMELINE
            tl_wait += clock64() - tl_c0;
#endif
            {
                int nst = stage + 2; if (nst >= BF_STAGES) nst -= BF_STAGES;
                if (base + 64 + lane < total) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&ring[nst].a[lane], &r->a);
                    cp_async16(&ring[nst].b[lane], &r->b);
                    cp_async16(&ring[nst].c[lane], &r->c);
#pragma unroll
                    for (int k = 0; k < K; k++) cp_async16(&ring[nst].e[k][lane], ex.xrec + 3 * (size_t)id_next + k);
                }
                cp_async_commit();
                if (base + 96 + lane < total) id_next = lst[base + 96 + lane];
            }
            const BfStage<K>& st = ring[stage];
            stage = (stage + 1 == BF_STAGES) ? 0 : stage + 1;

            bool hit = false;
            if (base + lane < total) {
                const float4 a = st.a[lane], b = st.b[lane];
                const float nBA = st.c[lane].w;
                const float bound = box_max_power(a.z, a.w, b.x, nBA, b.w, a.x - fx1, a.x - fx0, a.y - fy1, a.y - fy0);
                hit = !(bound < b.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
#ifdef GS_TIMELINE
            tl_hits += __popc(mask);
            const long long tl_c1 = clock64();
#endif
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 ga = st.a[j], gb = st.b[j];
                // power of both pixels, operation for operation as forward.cu:336-338 compiles:
                //   power = fma(fma(dx, A*dx, dy*(C*dy)), -0.5, -(dy*(B*dx)))
                const float dx = ga.x - pfx;
                const f2 dy2 = sub2(bc(ga.y), pfy2);
                f2 t1 = mul2(bc(gb.x), dy2);
                t1 = mul2(dy2, t1);
                const float t2 = ga.z * dx, t3n = (-ga.w) * dx;
                const f2 t3 = mul2(dy2, bc(t3n));
                const f2 sm = fma2(bc(dx), bc(t2), t1);
                const f2 p2 = fma2(sm, bc(-0.5f), t3);
                const float pA = lo(p2), pB = hi(p2);
                // (a packed FP32x2 transcription of libdevice's expf was bit-identical but not faster: FP32x2
                // instructions save issue slots, not FMA-pipe cycles)
                float alphaA = fminf(0.99f, gb.y * expf(pA)), alphaB = fminf(0.99f, gb.y * expf(pB));
#if BF_THR_TEST
                bool okA = !doneA && !(pA > 0.0f) && !(pA < gb.z) && !(alphaA < 1.0f / 255.0f);
                bool okB = !doneB && !(pB > 0.0f) && !(pB < gb.z) && !(alphaB < 1.0f / 255.0f);
#else           // power below the record's cut-off already implies alpha < 1/255 (the cut-off has a safety margin)
                bool okA = !doneA && !(pA > 0.0f) && !(alphaA < 1.0f / 255.0f);
                bool okB = !doneB && !(pB > 0.0f) && !(alphaB < 1.0f / 255.0f);
#endif
                if (!__any_sync(GS_FULL, okA || okB)) continue;
                alphaA = okA ? alphaA : 0.0f;  // alpha 0 leaves T and the colour exactly unchanged
                alphaB = okB ? alphaB : 0.0f;
                const f2 tt2 = mul2(T2, sub2(bc(1.0f), pk(alphaA, alphaB)));
                const bool stopA = okA && lo(tt2) < 0.0001f, stopB = okB && hi(tt2) < 0.0001f;
                doneA = doneA || stopA;
                doneB = doneB || stopB;
                alphaA = stopA ? 0.0f : alphaA;
                alphaB = stopB ? 0.0f : alphaB;
                const f2 a2 = pk(alphaA, alphaB);
                const float4 gc = st.c[j];
                {   // scalar FFMAs: a loop-carried FFMA2 accumulator costs two extra register moves per iteration
                    const float TA = lo(T2), TB = hi(T2);
                    const f2 w0 = mul2(bc(gc.x), a2), w1 = mul2(bc(gc.y), a2), w2 = mul2(bc(gc.z), a2);
                    c0A = __fmaf_rn(lo(w0), TA, c0A); c0B = __fmaf_rn(hi(w0), TB, c0B);
                    c1A = __fmaf_rn(lo(w1), TA, c1A); c1B = __fmaf_rn(hi(w1), TB, c1B);
                    c2A = __fmaf_rn(lo(w2), TA, c2A); c2B = __fmaf_rn(hi(w2), TB, c2B);
                }
#pragma unroll
                for (int k = 0; k < K; k++) {  // the extra passes: same alpha, same T, other colours
                    const float4 ge = st.e[k][j];
                    fma2_acc(E[k][0], mul2(bc(ge.x), a2), T2);
                    fma2_acc(E[k][1], mul2(bc(ge.y), a2), T2);
                    fma2_acc(E[k][2], mul2(bc(ge.z), a2), T2);
                }
                T2 = pk(stopA ? lo(T2) : lo(tt2), stopB ? hi(T2) : hi(tt2));
                if (okA && !stopA) lastA = base + (uint32_t)j + 1u;
                if (okB && !stopB) lastB = base + (uint32_t)j + 1u;
#if BF_DONE_PER_HIT
                if (__all_sync(GS_FULL, doneA && doneB)) break;
#endif          // otherwise: looked at once per batch below; hits after the last live pixel fail the vote above
            }
#ifdef GS_TIMELINE
            tl_loop += clock64() - tl_c1;
#endif
            if (__all_sync(GS_FULL, doneA && doneB)) break;
        }
        cp_async_wait<0>();
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0) {
            unsigned long long* tl = g_timeline + 6ull * unit;
            tl[0] = tl_t0; tl[1] = gtime();
            tl[2] = ((unsigned long long)smid() << 32) | total;
            tl[3] = ((unsigned long long)tl_batches << 32) | tl_hits;
            tl[4] = (unsigned long long)tl_wait; tl[5] = (unsigned long long)tl_loop;
        }
#endif
        const f2 C0 = pk(c0A, c0B), C1 = pk(c1A, c1B), C2 = pk(c2A, c2B);

        const float TA = lo(T2), TB = hi(T2);
        const size_t pidA = (size_t)W * pyA + px, pidB = (size_t)W * pyB + px;
        if (insA) { final_T[pidA] = TA; n_contrib[pidA] = lastA; }
        if (insB) { final_T[pidB] = TB; n_contrib[pidB] = lastB; }
        float oA[3 * (K + 1)], oB[3 * (K + 1)];
        oA[0] = lo(C0) + TA * bg0; oA[1] = lo(C1) + TA * bg1; oA[2] = lo(C2) + TA * bg2;
        oB[0] = hi(C0) + TB * bg0; oB[1] = hi(C1) + TB * bg1; oB[2] = hi(C2) + TB * bg2;
#pragma unroll
        for (int k = 0; k < K; k++) {
            oA[3 * k + 3] = lo(E[k][0]) + TA * bg0; oA[3 * k + 4] = lo(E[k][1]) + TA * bg1;
            oA[3 * k + 5] = lo(E[k][2]) + TA * bg2;
            oB[3 * k + 3] = hi(E[k][0]) + TB * bg0; oB[3 * k + 4] = hi(E[k][1]) + TB * bg1;
            oB[3 * k + 5] = hi(E[k][2]) + TB * bg2;
        }
        size_t oplane = plane, opA = pidA, opB = pidB;
        bool wA = insA, wB = insB;
        if (tg.ds) {  // W, H even and blocks start on even pixels: a 2x2 group is inside or outside as a whole
#pragma unroll
            for (int c = 0; c < 3 * (K + 1); c++) { oA[c] = box4(oA[c]); oB[c] = box4(oB[c]); }
            const int W2 = W >> 1;
            oplane = (size_t)W2 * (H >> 1);
            opA = (size_t)W2 * (pyA >> 1) + (px >> 1);
            opB = (size_t)W2 * (pyB >> 1) + (px >> 1);
            const bool writer = (lane & 9) == 0;  // even lx, even ly
            wA = insA && writer; wB = insB && writer;
        }
        if (wA) {
            _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                float* oc = tg.img[k];
                oc[opA] = oA[0]; oc[oplane + opA] = oA[1]; oc[2 * oplane + opA] = oA[2];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                float* oc = ex.out[k];
                oc[opA] = oA[3 * k + 3]; oc[oplane + opA] = oA[3 * k + 4]; oc[2 * oplane + opA] = oA[3 * k + 5];
            }
        }
        if (wB) {
            _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                float* oc = tg.img[k];
                oc[opB] = oB[0]; oc[oplane + opB] = oB[1]; oc[2 * oplane + opB] = oB[2];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                float* oc = ex.out[k];
                oc[opB] = oB[3 * k + 3]; oc[oplane + opB] = oB[3 * k + 4]; oc[2 * oplane + opB] = oB[3 * k + 5];
            }
        }
    }
}

GsPerDevice g_blend_dev[4];  // per extra-pass count K: value[0] = resident CTAs of the kernel on this device

template <int K>
cudaError_t launch_blend(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im, float* out_color,
                         uint32_t num_tiles) {
    const size_t smem = sizeof(BfStage<K>) * BF_STAGES * BF_WARPS;
    const int* dv = nullptr;
    {
        cudaError_t e = g_blend_dev[K].get(&dv, [smem](int dev, int* v) {
            int sms = 0, per_sm = 0;
            cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(blend_forward_px2_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_px2_kernel<K>, BF_WARPS * 32, smem);
            if (e != cudaSuccess) return e;
            v[0] = sms * (per_sm > 0 ? per_sm : 1);
            return cudaSuccess;
        });
        if (e != cudaSuccess) return e;
    }
    const int resident = dv[0];
    BfTargets tg;
    tg.n = f.s.num_peers > 0 ? f.s.num_peers : 1;
    tg.ds = f.s.downsample == 2 ? 1 : 0;
    for (int k = 0; k < 8; k++) tg.img[k] = f.s.num_peers > 0 ? f.s.peer_out_color[k] : out_color;
    BfExtra ex;
    ex.xrec = g.xrec;
    for (int k = 0; k < 3; k++) ex.out[k] = f.s.extra_out[k];
    // grid x 8 warps x quota covers the upper bound of units (4 per tile); quota >= 16, grid <= 60 % of the slots
    const uint32_t units_max = num_tiles * 4u;
    const uint32_t slots = (uint32_t)((resident * BF_SLOT_NUM + BF_SLOT_DEN - 1) / BF_SLOT_DEN);
    uint32_t quota = (units_max + BF_WARPS * slots - 1) / (BF_WARPS * slots);
    if (quota < 16u) quota = 16u;
    const unsigned grid = (units_max + BF_WARPS * quota - 1) / (BF_WARPS * quota);
    blend_forward_px2_kernel<K><<<grid, BF_WARPS * 32, smem, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width,
                                                                        f.s.height, f.gx, num_tiles, g.hdr,
                                                                        f.s.background, im.final_T, im.n_contrib, tg, ex,
                                                                        (int)quota);
    gs_note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color) {
    const uint32_t num_tiles = (uint32_t)f.gx * (uint32_t)(f.row1 - f.row0);
    if (num_tiles == 0) return cudaSuccess;
    switch (f.s.num_extra) {
        case 1: return launch_blend<1>(f, g, b, im, out_color, num_tiles);
        case 2: return launch_blend<2>(f, g, b, im, out_color, num_tiles);
        case 3: return launch_blend<3>(f, g, b, im, out_color, num_tiles);
        default: return launch_blend<0>(f, g, b, im, out_color, num_tiles);
    }
}

#ifdef GS_TIMELINE
extern "C" int gs_debug_timeline(void* dev_buf) {
    unsigned long long* p = (unsigned long long*)dev_buf;
    return (int)cudaMemcpyToSymbol(g_timeline, &p, sizeof(p));
}
#endif
