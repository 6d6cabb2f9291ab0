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
//   * each lane gathers one packed 48-B record per step straight into a per-warp shared-memory ring with cp.async
//     (3 x 16 B, no register staging; the table is L2 resident), two batches ahead, list indices three ahead;
//   * a few blocks on silhouettes walk very long lists without ever saturating; with one warp per block they alone
//     set the kernel time (measured: 700 us for one block, 250 us for 95 % of all blocks).  Every 16 batches a
//     warp shrinks its cull box to the pixels that are still live and, once the queue of fresh units is empty (so
//     idle warps exist), hands half of its live pixels -- with their T / colour / contributor state -- to an idle
//     warp through a split queue.  Each pixel still sees the same operations in the same order: bit-identical;
//   * empty tiles (85 % of a THuman frame) are not blended at all: one warp fills a whole tile with the background;
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#define BF_STAGES 3  // per-warp ring of 32-record batches: one being blended, two in flight

struct BfStage {
    float4 a[32];  // x, y, conic.x, conic.y
    float4 b[32];  // conic.z, opacity, thr, -B/C
    float4 c[32];  // r, g, b, -B/A
};

// One (pixel, Gaussian) evaluation up to alpha, exactly as forward.cu:330-349, in two steps.
// power > 0: reference skips; power < thr: provably alpha < 1/255; alpha < 1/255: reference skips.
__device__ __forceinline__ bool eval_power(const float4 ga, const float4 gb, float pfx, float pfy, float& power) {
    const float dx = ga.x - pfx, dy = ga.y - pfy;
    power = -0.5f * (ga.z * dx * dx + gb.x * dy * dy) - ga.w * dx * dy;
    return !(power > 0.0f) && !(power < gb.z);
}
__device__ __forceinline__ bool eval_alpha(const float4 gb, float power, float& alpha) {
    alpha = fminf(0.99f, gb.y * expf(power));
    return !(alpha < 1.0f / 255.0f);
}

__device__ __forceinline__ unsigned ldv(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

#define BF_CHECK 16  // batches between two looks at the pixel block (shrink the cull box / hand half of it away)
// BF_SPLIT 1: once no fresh unit is left, running warps hand half of their live pixels to idle warps through the split
// queue.  Bit-identical, but measured SLOWER on B200 (0.75 vs 0.66 ms at C2): the critical path of a long silhouette
// block is its list walk (~1 us per 32 instances for a lone warp), which pixel splitting replicates instead of
// shortening.  Kept for experiments; the tail is hidden by overlapping frames on streams instead (renderer.py).
#ifndef BF_SPLIT
#define BF_SPLIT 0
#endif
// Measured and rejected on B200 (C2): one instance per iteration with per-lane early outs (2.6x slower: the divergent
// loop no longer reconverges per instance), and a warp vote that skips the exponentials of a pair nobody needs
// (+6 % kernel time: the vote costs more than the rare skip saves).

__global__ void __launch_bounds__(BF_WARPS * 32, 5) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_tiles, GsHeader* __restrict__ hdr,
    uint4* __restrict__ q_task, float* __restrict__ q_state, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color) {
    __shared__ BfStage s_ring[BF_WARPS][BF_STAGES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;  // this lane's pixel inside the warp's 8x4 block
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    BfStage* __restrict__ ring = s_ring[warp];
    unsigned* const q_fresh = &hdr->tickets[6];    // next fresh unit
    unsigned* const q_head = &hdr->tickets[7];     // next split task to claim
    unsigned* const q_tail = &hdr->tickets[8];     // split tasks allocated
    unsigned* const q_pending = &hdr->tickets[9];  // tasks (fresh + split) not finished yet
    // order[] lists the shard's tiles longest list first, empty tiles last: units [0, 8*nonempty) are 8x4 pixel
    // blocks of non-empty tiles, the remaining units are whole empty tiles that only receive the background.
    const uint32_t nonempty = hdr->nonempty_tiles;
    const uint32_t blend_units = nonempty * 8u;
    const uint32_t num_units = blend_units + (num_tiles - nonempty);
    bool fresh_left = true;

    while (true) {
        // ---- next task: a fresh unit while there are any, then split-off tasks until every task has finished
        uint32_t unit = 0, base0 = 0, slot = 0;
        int rx0 = 0, rx1 = 7, ry0 = 0, ry1 = 3;  // owned pixel rectangle inside the 8x4 block (inclusive)
        bool child = false;
        if (fresh_left) {
            if (lane == 0) unit = atomicAdd(q_fresh, 1u);
            unit = __shfl_sync(GS_FULL, unit, 0);
            fresh_left = unit < num_units;
        }
        if (!fresh_left && !BF_SPLIT) break;
        if (!fresh_left) {
            unsigned got = 0;
            if (lane == 0) {
                slot = atomicAdd(q_head, 1u);
                if (slot < GS_BF_QCAP) {
                    const unsigned* ready = reinterpret_cast<const unsigned*>(q_task + slot) + 3;
                    while (true) {
                        if (ldv(ready) != 0u) { got = 1; break; }
                        if (ldv(q_pending) == 0u) break;
                        __nanosleep(200);
                    }
                }
            }
            got = __shfl_sync(GS_FULL, got, 0);
            if (!got) break;
            slot = __shfl_sync(GS_FULL, slot, 0);
            __threadfence();
            const unsigned* tq = reinterpret_cast<const unsigned*>(q_task + slot);
            unit = ldv(tq); base0 = ldv(tq + 1);
            const unsigned tz = ldv(tq + 2);
            rx0 = tz & 15; rx1 = (tz >> 4) & 15; ry0 = (tz >> 8) & 15; ry1 = (tz >> 12) & 15;
            child = true;
        }
        if (unit >= blend_units) {  // ---- empty tile: colour = background, T = 1, no contributor
            const uint32_t tile = order[nonempty + (unit - blend_units)];
            const int x0 = (int)(tile % gx) * GS_TILE, y0 = (int)(tile / gx) * GS_TILE;
            if ((W & 3) == 0 && x0 + GS_TILE <= W) {
                const int px = x0 + (lane & 3) * 4;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int py = y0 + r * 8 + (lane >> 2);
                    if (py < H) {
                        const size_t pid = (size_t)W * py + px;
                        *reinterpret_cast<float4*>(out_color + pid) = make_float4(bg0, bg0, bg0, bg0);
                        *reinterpret_cast<float4*>(out_color + plane + pid) = make_float4(bg1, bg1, bg1, bg1);
                        *reinterpret_cast<float4*>(out_color + 2 * plane + pid) = make_float4(bg2, bg2, bg2, bg2);
                        *reinterpret_cast<float4*>(final_T + pid) = make_float4(1.f, 1.f, 1.f, 1.f);
                        *reinterpret_cast<uint4*>(n_contrib + pid) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                for (int r = 0; r < 8; r++) {
                    const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
                    if (px < W && py < H) {
                        const size_t pid = (size_t)W * py + px;
                        out_color[pid] = bg0; out_color[plane + pid] = bg1; out_color[2 * plane + pid] = bg2;
                        final_T[pid] = 1.f; n_contrib[pid] = 0u;
                    }
                }
            }
            if (lane == 0) atomicSub(q_pending, 1u);
            continue;
        }
#ifdef GS_TIMELINE
        const unsigned long long t_start = gtime();
        unsigned tl_batches = 0, tl_hits = 0;
        long long tl_wait = 0, tl_loop = 0;
#endif
        const uint32_t tile = order[unit >> 3];
        const int sub = unit & 7;
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 4;
        if (bx0 >= W || by0 >= H) {  // block entirely outside the image
            if (lane == 0) atomicSub(q_pending, 1u);
            continue;
        }
        const int px = bx0 + lx, py = by0 + ly;
        const bool inside = px < W && py < H;
        const float pfx = (float)px, pfy = (float)py;
        bool owned = inside && lx >= rx0 && lx <= rx1 && ly >= ry0 && ly <= ry1;
        float fx0 = (float)(bx0 + rx0), fx1 = (float)(bx0 + rx1), fy0 = (float)(by0 + ry0), fy1 = (float)(by0 + ry1);

        const uint2 range = ranges[tile];
        const uint32_t total = range.y - range.x;
        const uint32_t* __restrict__ lst = list + range.x;

        bool done = !owned;
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
        uint32_t last_contributor = 0;
        if (child) {  // continue where the task that split this one off stopped
            const float* st = q_state + (size_t)slot * (5 * 32);
            T = __ldcg(st + lane); C0 = __ldcg(st + 32 + lane); C1 = __ldcg(st + 64 + lane); C2 = __ldcg(st + 96 + lane);
            const uint32_t lc = __float_as_uint(__ldcg(st + 128 + lane));
            last_contributor = lc & 0x7fffffffu;
            done = done || (lc >> 31) != 0u;
        }

        // prologue: batches 0 and 1 in flight, indices of batch 2 in a register
        __syncwarp();  // the previous task's readers are done with the ring
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (base0 + p * 32 + lane < total) {
                const GsRec* r = rec + lst[base0 + p * 32 + lane];
                cp_async16(&ring[p].a[lane], &r->a);
                cp_async16(&ring[p].b[lane], &r->b);
                cp_async16(&ring[p].c[lane], &r->c);
            }
            cp_async_commit();
        }
        uint32_t id_next = (base0 + 64 + lane < total) ? lst[base0 + 64 + lane] : 0u;

        int stage = 0;
        uint32_t next_check = base0 + BF_CHECK * 32;
        for (uint32_t base = base0; base < total; base += 32) {
            if (base == next_check) {
                next_check += BF_CHECK * 32;
                // Pixels that have finished no longer need any Gaussian: shrink the cull box to the live ones.
                const unsigned alive = __ballot_sync(GS_FULL, !done);
                const unsigned cols = (alive | (alive >> 8) | (alive >> 16) | (alive >> 24)) & 0xffu;
                rx0 = __ffs(cols) - 1; rx1 = 31 - __clz(cols);
                const unsigned rows = ((alive & 0xffu) ? 1u : 0u) | ((alive & 0xff00u) ? 2u : 0u) |
                                      ((alive & 0xff0000u) ? 4u : 0u) | ((alive & 0xff000000u) ? 8u : 0u);
                ry0 = __ffs(rows) - 1; ry1 = 31 - __clz(rows);
                // Once no fresh unit is left, idle warps exist: hand half of the live pixels (with their state) to
                // one of them.  Every pixel still sees exactly the same operation sequence.
                const int w = rx1 - rx0 + 1, h = ry1 - ry0 + 1;
                if (BF_SPLIT && w * h >= 2 && total - base >= 2 * BF_CHECK * 32) {
                    unsigned s2 = GS_BF_QCAP;
                    if (lane == 0 && ldv(q_fresh) >= num_units && ldv(q_tail) < GS_BF_QCAP) {
                        s2 = atomicAdd(q_tail, 1u);
                        if (s2 < GS_BF_QCAP) atomicAdd(q_pending, 1u);
                    }
                    s2 = __shfl_sync(GS_FULL, s2, 0);
                    if (s2 < GS_BF_QCAP) {
                        float* st = q_state + (size_t)s2 * (5 * 32);
                        __stcg(st + lane, T); __stcg(st + 32 + lane, C0); __stcg(st + 64 + lane, C1);
                        __stcg(st + 96 + lane, C2);
                        __stcg(st + 128 + lane, __uint_as_float(last_contributor | (done ? 0x80000000u : 0u)));
                        int bx_0 = rx0, bx_1 = rx1, by_0 = ry0, by_1 = ry1;  // the half that is given away
                        if (w >= 2 * h || (w >= 2 && h < 2)) { bx_0 = rx0 + w / 2; rx1 = bx_0 - 1; }
                        else { by_0 = ry0 + h / 2; ry1 = by_0 - 1; }
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) {
                            uint4* e = q_task + s2;
                            e->x = unit; e->y = base;
                            e->z = (unsigned)bx_0 | ((unsigned)bx_1 << 4) | ((unsigned)by_0 << 8) | ((unsigned)by_1 << 12);
                            __threadfence();
                            *reinterpret_cast<volatile unsigned*>(&e->w) = 1u;
                        }
                        if (lx >= bx_0 && lx <= bx_1 && ly >= by_0 && ly <= by_1) { owned = false; done = true; }
                    }
                }
                fx0 = (float)(bx0 + rx0); fx1 = (float)(bx0 + rx1); fy0 = (float)(by0 + ry0); fy1 = (float)(by0 + ry1);
            }
#ifdef GS_TIMELINE
            const long long tw0 = clock64();
#endif
            cp_async_wait<1>();  // this lane's copies of the current batch have landed
            __syncwarp();        // ... and everybody else's; all lanes are done reading the stage refilled below
#ifdef GS_TIMELINE
            const long long tw1 = clock64();
            tl_wait += tw1 - tw0;
#endif
            {
                int nst = stage + 2; if (nst >= BF_STAGES) nst -= BF_STAGES;
                if (base + 64 + lane < total) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&ring[nst].a[lane], &r->a);
                    cp_async16(&ring[nst].b[lane], &r->b);
                    cp_async16(&ring[nst].c[lane], &r->c);
                }
                cp_async_commit();
                if (base + 96 + lane < total) id_next = lst[base + 96 + lane];
            }
            const BfStage& st = ring[stage];
            stage = (stage + 1 == BF_STAGES) ? 0 : stage + 1;

            // conservative cull of this lane's Gaussian against the live pixels' bounding box
            bool hit = false;
            if (base + lane < total) {
                const float4 a = st.a[lane], b = st.b[lane];
                const float nBA = st.c[lane].w;
                const float bound = box_max_power(a.z, a.w, b.x, nBA, b.w, a.x - fx1, a.x - fx0, a.y - fy1, a.y - fy0);
                hit = !(bound < b.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
#ifdef GS_TIMELINE
            tl_batches++;
            tl_hits += __popc(mask);
            const long long tl0 = clock64();
#endif
            // surviving instances in list order, two per iteration: the two alpha evaluations are independent
            // (instruction-level parallelism when few warps are resident); T and the colour are then updated
            // strictly in order, so every pixel sees the reference's operation sequence.
            while (mask) {
                const int j0 = __ffs(mask) - 1;
                mask &= mask - 1;
                const bool two = mask != 0;
                const int j1 = two ? __ffs(mask) - 1 : j0;
                mask &= mask - 1;  // no-op when mask == 0
                float alpha0, alpha1, power0, power1;
                const float4 gb0 = st.b[j0], gb1 = st.b[j1];
                bool ok0 = eval_power(st.a[j0], gb0, pfx, pfy, power0) && !done;
                bool ok1 = eval_power(st.a[j1], gb1, pfx, pfy, power1) && two && !done;
                ok0 = eval_alpha(gb0, power0, alpha0) && ok0;
                ok1 = eval_alpha(gb1, power1, alpha1) && ok1;
                if (ok0) {
                    const float test_T = T * (1 - alpha0);
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const float4 gc = st.c[j0];
                        C0 += gc.x * alpha0 * T;
                        C1 += gc.y * alpha0 * T;
                        C2 += gc.z * alpha0 * T;
                        T = test_T;
                        last_contributor = base + (uint32_t)j0 + 1u;
                    }
                }
                ok1 = ok1 && !done;
                if (ok1) {
                    const float test_T = T * (1 - alpha1);
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const float4 gc = st.c[j1];
                        C0 += gc.x * alpha1 * T;
                        C1 += gc.y * alpha1 * T;
                        C2 += gc.z * alpha1 * T;
                        T = test_T;
                        last_contributor = base + (uint32_t)j1 + 1u;
                    }
                }
                if (__all_sync(GS_FULL, done)) break;
            }
#ifdef GS_TIMELINE
            tl_loop += clock64() - tl0;
#endif
            if (__all_sync(GS_FULL, done)) break;
        }
        cp_async_wait<0>();  // nothing of this task may land in the ring after the next one starts filling it

        if (owned) {
            const size_t pid = (size_t)W * py + px;
            final_T[pid] = T;
            n_contrib[pid] = last_contributor;
            out_color[pid] = C0 + T * bg0;
            out_color[plane + pid] = C1 + T * bg1;
            out_color[2 * plane + pid] = C2 + T * bg2;
        }
        if (lane == 0) atomicSub(q_pending, 1u);
#ifdef GS_TIMELINE
        if (lane == 0 && g_timeline && !child) {
            unsigned long long* e = g_timeline + 6ull * unit;
            e[0] = t_start; e[1] = gtime();
            e[2] = ((unsigned long long)smid() << 32) | total;
            e[3] = ((unsigned long long)tl_batches << 32) | tl_hits;
            e[4] = (unsigned long long)tl_wait; e[5] = (unsigned long long)tl_loop;
        }
#endif
    }
}

// ---------------------------------------------------------------------------------------------------
// Two pixels per lane (BF_PX2): the warp blends an 8x8 block, lane (lx, ly) owns pixels (lx, ly) and (lx, ly + 4).
// The two pixels share dx, so the exponent of both costs 3 scalar + 6 packed FP32x2 instructions (FADD2 / FMUL2 /
// FFMA2 of sm_100: one issue slot, two IEEE-rn results), the colour accumulation 6 packed ones; every packed
// operation is, per half, the scalar operation the reference executes (same operands, same order, same fused
// multiply-adds), so results stay bit-identical.  Per-batch work (gather, cull) is shared by 64 pixels.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void fma2_acc(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

#ifndef BF_PX2_OCC
#define BF_PX2_OCC 3
#endif
__global__ void __launch_bounds__(BF_WARPS * 32, BF_PX2_OCC) blend_forward_px2_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_tiles, GsHeader* __restrict__ hdr,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    float* __restrict__ out_color) {
    __shared__ BfStage s_ring[BF_WARPS][BF_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    BfStage* __restrict__ ring = s_ring[warp];
    unsigned* const q_fresh = &hdr->tickets[6];
    const uint32_t nonempty = hdr->nonempty_tiles;
    const uint32_t blend_units = nonempty * 4u;
    const uint32_t num_units = blend_units + (num_tiles - nonempty);

    while (true) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(q_fresh, 1u);
        unit = __shfl_sync(GS_FULL, unit, 0);
        if (unit >= num_units) break;
        if (unit >= blend_units) {  // ---- empty tile: colour = background, T = 1, no contributor
            const uint32_t tile = order[nonempty + (unit - blend_units)];
            const int x0 = (int)(tile % gx) * GS_TILE, y0 = (int)(tile / gx) * GS_TILE;
            if ((W & 3) == 0 && x0 + GS_TILE <= W) {
                const int px = x0 + (lane & 3) * 4;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int py = y0 + r * 8 + (lane >> 2);
                    if (py < H) {
                        const size_t pid = (size_t)W * py + px;
                        *reinterpret_cast<float4*>(out_color + pid) = make_float4(bg0, bg0, bg0, bg0);
                        *reinterpret_cast<float4*>(out_color + plane + pid) = make_float4(bg1, bg1, bg1, bg1);
                        *reinterpret_cast<float4*>(out_color + 2 * plane + pid) = make_float4(bg2, bg2, bg2, bg2);
                        *reinterpret_cast<float4*>(final_T + pid) = make_float4(1.f, 1.f, 1.f, 1.f);
                        *reinterpret_cast<uint4*>(n_contrib + pid) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                for (int r = 0; r < 8; r++) {
                    const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
                    if (px < W && py < H) {
                        const size_t pid = (size_t)W * py + px;
                        out_color[pid] = bg0; out_color[plane + pid] = bg1; out_color[2 * plane + pid] = bg2;
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

        bool doneA = !insA, doneB = !insB;
        f2 T2 = bc(1.0f), C0 = bc(0.f), C1 = bc(0.f), C2 = bc(0.f);
        uint32_t lastA = 0, lastB = 0;

        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (p * 32 + lane < total) {
                const GsRec* r = rec + lst[p * 32 + lane];
                cp_async16(&ring[p].a[lane], &r->a);
                cp_async16(&ring[p].b[lane], &r->b);
                cp_async16(&ring[p].c[lane], &r->c);
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
            cp_async_wait<1>();
            __syncwarp();
            {
                int nst = stage + 2; if (nst >= BF_STAGES) nst -= BF_STAGES;
                if (base + 64 + lane < total) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&ring[nst].a[lane], &r->a);
                    cp_async16(&ring[nst].b[lane], &r->b);
                    cp_async16(&ring[nst].c[lane], &r->c);
                }
                cp_async_commit();
                if (base + 96 + lane < total) id_next = lst[base + 96 + lane];
            }
            const BfStage& st = ring[stage];
            stage = (stage + 1 == BF_STAGES) ? 0 : stage + 1;

            bool hit = false;
            if (base + lane < total) {
                const float4 a = st.a[lane], b = st.b[lane];
                const float nBA = st.c[lane].w;
                const float bound = box_max_power(a.z, a.w, b.x, nBA, b.w, a.x - fx1, a.x - fx0, a.y - fy1, a.y - fy0);
                hit = !(bound < b.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
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
                bool okA = !doneA && !(pA > 0.0f) && !(pA < gb.z) && !(alphaA < 1.0f / 255.0f);
                bool okB = !doneB && !(pB > 0.0f) && !(pB < gb.z) && !(alphaB < 1.0f / 255.0f);
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
                fma2_acc(C0, mul2(bc(gc.x), a2), T2);
                fma2_acc(C1, mul2(bc(gc.y), a2), T2);
                fma2_acc(C2, mul2(bc(gc.z), a2), T2);
                T2 = pk(stopA ? lo(T2) : lo(tt2), stopB ? hi(T2) : hi(tt2));
                if (okA && !stopA) lastA = base + (uint32_t)j + 1u;
                if (okB && !stopB) lastB = base + (uint32_t)j + 1u;
                if (__all_sync(GS_FULL, doneA && doneB)) break;
            }
            if (__all_sync(GS_FULL, doneA && doneB)) break;
        }
        cp_async_wait<0>();

        if (insA) {
            const size_t pid = (size_t)W * pyA + px;
            const float T = lo(T2);
            final_T[pid] = T;
            n_contrib[pid] = lastA;
            out_color[pid] = lo(C0) + T * bg0;
            out_color[plane + pid] = lo(C1) + T * bg1;
            out_color[2 * plane + pid] = lo(C2) + T * bg2;
        }
        if (insB) {
            const size_t pid = (size_t)W * pyB + px;
            const float T = hi(T2);
            final_T[pid] = T;
            n_contrib[pid] = lastB;
            out_color[pid] = hi(C0) + T * bg0;
            out_color[plane + pid] = hi(C1) + T * bg1;
            out_color[2 * plane + pid] = hi(C2) + T * bg2;
        }
    }
}

// BF_PX2 1 (default): the two-pixels-per-lane kernel.  Measured on B200 (C2): +8.5 % frames/s with frames in flight
// (1717 vs 1583) for +3 % single-frame blend time (coarser units lengthen the tail), bit-identical output.
#ifndef BF_PX2
#define BF_PX2 1
#endif

int g_blend_grid = 0;

}  // namespace

#ifdef GS_TIMELINE
extern "C" int gs_debug_timeline(void* dev_buf) {
    unsigned long long* p = (unsigned long long*)dev_buf;
    return (int)cudaMemcpyToSymbol(g_timeline, &p, sizeof(p));
}
#endif

cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color) {
    const uint32_t num_tiles = (uint32_t)f.gx * (uint32_t)(f.row1 - f.row0);
    if (num_tiles == 0) return cudaSuccess;
    if (g_blend_grid == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (BF_PX2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_px2_kernel, BF_WARPS * 32, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_kernel, BF_WARPS * 32, 0);
        if (e != cudaSuccess) return e;
        g_blend_grid = sms * (per_sm > 0 ? per_sm : 1);
    }
    const unsigned grid = (unsigned)min((uint32_t)g_blend_grid, num_tiles);
    if (BF_PX2)
        blend_forward_px2_kernel<<<grid, BF_WARPS * 32, 0, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width,
                                                                      f.s.height, f.gx, num_tiles, g.hdr, f.s.background,
                                                                      im.final_T, im.n_contrib, out_color);
    else
    blend_forward_kernel<<<grid, BF_WARPS * 32, 0, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height,
                                                              f.gx, num_tiles, g.hdr, im.bf_task, im.bf_state,
                                                              f.s.background, im.final_T, im.n_contrib, out_color);
    gs_note_launch();
    return cudaGetLastError();
}
