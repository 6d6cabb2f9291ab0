// blend_forward.cu -- per-tile front-to-back alpha compositing (K6; replaces renderCUDA,
// dgr/cuda_rasterizer/forward.cu:264-377).
//
// The kernel that runs is blend_forward_grouped_kernel<K> (second half of this file; K = extra colour passes of the
// same walk); blend_forward_px2_kernel is the round-1 loop (one instance per iteration) that it grew out of -- kept for
// the opt-in latency experiments built on it (TEAM = 1: exact CTA teams, GsScene.team_after; TEAM = 2: associative split
// walks, GsScene.blend_split) and as the A/B partner of the grouped loop (developer switch GSPLAT_B200_BLEND_PLAIN=1).
// Work decomposition, common to both (B200: 148 SMs):
//   * the unit of work is ONE WARP blending an 8x8 pixel block (1/4 of a 16x16 tile) against the tile's
//     depth-ordered instance list, two pixels per lane; warps never synchronise with each other (the reference's
//     per-batch __syncthreads was its top stall reason and its CTA-per-tile grid left the average SM idle 60 % of
//     the frame behind a few silhouette tiles);
//   * units are handed out through a global atomic queue in descending order of list length (longest-processing-
//     time-first); empty tiles (85 % of a THuman frame) are not blended at all: one warp fills a whole tile with
//     the background;
//   * each lane gathers one packed 48-B record per step straight into a per-warp shared-memory ring with cp.async
//     (3 x 16 B, no register staging; the table is L2 resident), two batches ahead, list indices further ahead;
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
// spread over ~1000 half-finished walks.  (4) That re-distribution was then built (TEAM variant below, opt-in through
// GsScene.team_after): long walks are parked and finished by CTA teams -- six warps cull + evaluate alphas four
// instances at a time, two warps apply them in list order with the reference's exact T / C recurrence, bit-identical
// in every test.  Measured at C2 (profiles/r02b_blend_timeline_teams_c2.txt): a team finishes a block at 0.095 us per
// surviving instance against 0.13 (lone warp) / 0.23 (warp under contention), i.e. the serial recurrence warp with its
// in-order issue (~19 dependent-ish instructions per instance) is the new limit, and eight warps per block cost
// 3x the warp-time of one; frame 0.594 ms at the best hand-over threshold (96 batches) against 0.602 without teams.
// The distribution of work is the obstacle, not a few outliers: 400 of the 4888 blocks carry > 1200 surviving
// instances (>= 150 us of dependent chain each), the queue of fresh blocks drains at ~220 us, and everything started
// in its second half finishes late.  Off by default.
//
// Bound: FP32 issue, not HBM.  Algorithmic HBM bytes: 40*sum(need_t) + 20*N + 8*Tn.
#include <cstdlib>

#include "gs_common.cuh"

namespace {

#define BF_WARPS 8
#ifndef BF_TEAM_AFTER
#define BF_TEAM_AFTER 0   // library default of GsScene.team_after: 0 = teams off (measured: no gain at C2, see below)
#endif
#ifndef BF_TEAM_CTAS
#define BF_TEAM_CTAS 1    // CTAs per SM that serve parked blocks from the start of the kernel (latency mode)
#endif

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

// The same bound without branches (both edge maxima are always evaluated and selected afterwards): for a warp that has
// its scheduler to itself a predicate-to-branch round trip costs more than the arithmetic it skips.
__device__ __forceinline__ float box_max_power_sel(float A, float B, float C, float nBA, float nBC, float xlo, float xhi,
                                                   float ylo, float yhi) {
    const bool in_x = (xlo <= 0.f) && (xhi >= 0.f);
    const bool in_y = (ylo <= 0.f) && (yhi >= 0.f);
    const float xe = (xlo > 0.f) ? xlo : xhi;
    const float ye = (ylo > 0.f) ? ylo : yhi;
    const float y = fminf(yhi, fmaxf(ylo, nBC * xe));
    const float a1 = 0.5f * A * xe * xe, a2 = 0.5f * C * y * y, a3 = B * xe * y;
    const float v1 = -(a1 + a2) - a3 + (4.0e-6f * (a1 + a2 + fabsf(a3)) + 0.01f);
    const float x = fminf(xhi, fmaxf(xlo, nBA * ye));
    const float b1 = 0.5f * A * x * x, b2 = 0.5f * C * ye * ye, b3 = B * x * ye;
    const float v2 = -(b1 + b2) - b3 + (4.0e-6f * (b1 + b2 + fabsf(b3)) + 0.01f);
    const float e1 = in_x ? -3.0e38f : v1;
    const float best = in_y ? e1 : fmaxf(e1, v2);
    return (in_x && in_y) ? 0.f : best;
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

#ifndef BF_CHECK
#define BF_CHECK 16  // batches between two looks at which pixels of the block are still live
#endif

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


// ---------------------------------------------------------------------------------------------------------------------
// Teams (latency mode, template parameter TEAM): a pixel block whose list walk is still running after `team_after`
// batches is parked by its warp (state of the 64 pixels to global memory) and finished by a CTA working as a team.
// What one warp cannot shorten is the per-instance chain -- exponent, expf, tests, blend: ~58 dependent instructions
// for the 64 pixels of a block -- so eight warps split it along the only cut that keeps the result bit for bit: the
// alpha of an instance at a pixel does not depend on the compositing state, the transmittance recurrence does.
// Warps 1..7 ("producers") walk the list, batch b to producer b mod 7: gather the records (cp.async, two batches in
// flight per warp), cull against the block, evaluate alpha of every surviving instance for all 64 pixels with exactly
// the operations of the warp-per-block path, two instances at a time, and append them to a shared-memory FIFO whose
// entries are numbered in LIST order (the hit counts of the batches are chained from producer to producer).  Warp 0
// (the "consumer") owns T, colour, done flags and contributor index of the 64 pixels and applies the FIFO batch by
// batch with the reference's recurrence -- test_T = T (1 - alpha), stop test, C = fma(c alpha, T, C): ~25
// instructions per instance instead of ~58, overlapped with the seven producers.  Flow control: a producer waits
// until the FIFO has room for its whole batch (consumer's head counter) and until the consumer is less than two
// rounds behind; the consumer waits for the batch flag of the producer whose turn it is.  When every pixel is done the
// consumer raises `stop`.  The cull box shrinks to the live pixels as in the warp-per-block path (published by the
// consumer; a producer that still sees the older, larger box only evaluates alphas the consumer then ignores).
// Roles are given out by ARRIVAL order of the CTAs (first n_fresh arrivals work the fresh queue, later ones serve
// parked blocks from the start), so a team CTA only ever waits for CTAs that are already running; a CTA that runs
// out of fresh blocks becomes a team too.  Everybody leaves when all fresh CTAs are done and the parked queue is empty.
#define TM_PROD 6    // producer warps (warps 2..7); warps 0 and 1 are the consumers of pixel rows 0-3 / 4-7 of the block
#define TM_RING 128  // FIFO entries (power of two)
#define TM_STAGES 2
#define TM_FAR 0x40000000u  // a consumer whose pixels are all done reports "everything consumed"
struct TmStage {
    float4 a[32], b[32], c[32];
};
struct TmShared {
    TmStage ring[TM_PROD][TM_STAGES];
    float fa[2][TM_RING][32];    // alpha of the entry's instance at pixel (lane & 7, (lane >> 3) + 4 c) (0 = skip)
    float4 fc[TM_RING];          // r, g, b, list position + 1
    volatile unsigned tok_seq[8], tok_cum[8];  // producer q: tok_cum[q] = FIFO entry number of batch tok_seq[q]'s first hit
    volatile unsigned bflag[8];                // producer q has published all its batches < bflag[q]
    volatile unsigned bcnt[8][2];              // hits of producer q's batch, by round parity
    volatile unsigned head[2], cbatch[2];      // per consumer: entries consumed, batches whose count it has read
    volatile unsigned stop;
    volatile float box[2][4];                  // per consumer: bounding box of its live pixels (x0 > x1: none)
    unsigned slot;
};

#ifndef TM_SLEEP_NS
#define TM_SLEEP_NS 0
#endif
__device__ __forceinline__ void tm_sleep() {
#if TM_SLEEP_NS > 0
    __nanosleep(TM_SLEEP_NS);
#endif
}
__device__ __forceinline__ unsigned ldv(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

// alpha of instance j of a staged batch at the two pixels of this lane (0 = the reference skips it there)
__device__ __forceinline__ float2 tm_alpha(const TmStage& st, int j, float pfx, f2 pfy2) {
    const float4 ga = st.a[j], gb = st.b[j];
    const float dx = ga.x - pfx;
    const f2 dy2 = sub2(bc(ga.y), pfy2);
    f2 t1 = mul2(bc(gb.x), dy2);
    t1 = mul2(dy2, t1);
    const float t2 = ga.z * dx, t3n = (-ga.w) * dx;
    const f2 t3 = mul2(dy2, bc(t3n));
    const f2 sm = fma2(bc(dx), bc(t2), t1);
    const f2 p2 = fma2(sm, bc(-0.5f), t3);
    const float pA = lo(p2), pB = hi(p2);
    const float alphaA = fminf(0.99f, gb.y * expf(pA)), alphaB = fminf(0.99f, gb.y * expf(pB));
    const bool okA = !(pA > 0.0f) && !(alphaA < 1.0f / 255.0f);
    const bool okB = !(pB > 0.0f) && !(alphaB < 1.0f / 255.0f);
    return make_float2(okA ? alphaA : 0.0f, okB ? alphaB : 0.0f);
}

struct TmPix {  // consumer state of one lane's pixel
    float T, c0, c1, c2;
    uint32_t last;
    bool done;
};
// The warp-per-block path's blend step with ok = !done && (the alpha passed its tests), arranged so that the
// loop-carried chain is done -> select -> multiply -> compare: the factor (1 - alpha) does not depend on the state
// (a skipped pixel multiplies by exactly 1), the colour terms hang off the chain.  Same operations on the same
// operands as the warp-per-block path: T (1 - alpha), the stop test, fma(c alpha, T, C).
__device__ __forceinline__ void tm_apply(TmPix& s, float al, float4 gc) {
    const float om = __fsub_rn(1.0f, al);  // 1 - alpha (alpha = 0: the instance skips this pixel)
    const float T = s.T;
    const float tt = __fmul_rn(T, s.done ? 1.0f : om);
    const bool stop = tt < 0.0001f;  // T itself never is: only a hit can stop
    s.done = s.done || stop;
    s.T = stop ? T : tt;
    const float a = s.done ? 0.0f : al;  // done before or stopped here
    s.c0 = __fmaf_rn(__fmul_rn(gc.x, a), T, s.c0);
    s.c1 = __fmaf_rn(__fmul_rn(gc.y, a), T, s.c1);
    s.c2 = __fmaf_rn(__fmul_rn(gc.z, a), T, s.c2);
    if (a != 0.0f) s.last = __float_as_uint(gc.w);
}

// bounding box of the live pixels of one consumer (rows ly0 .. ly0 + 3 of the block); x0 > x1 if none is live
__device__ __forceinline__ unsigned tm_live_box(bool done, int bx0, int by0, float box[4]) {
    const unsigned alive = __ballot_sync(GS_FULL, !done);
    const unsigned cols = (alive | (alive >> 8) | (alive >> 16) | (alive >> 24)) & 0xffu;
    unsigned rows = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) rows |= ((alive >> (8 * r)) & 0xffu) ? (1u << r) : 0u;
    if (alive) {
        box[0] = (float)(bx0 + __ffs(cols) - 1); box[1] = (float)(bx0 + 31 - __clz(cols));
        box[2] = (float)(by0 + __ffs(rows) - 1); box[3] = (float)(by0 + 31 - __clz(rows));
    } else {
        box[0] = 1.0f; box[1] = 0.0f; box[2] = 1.0f; box[3] = 0.0f;
    }
    return alive;
}

// One parked block, all eight warps of the CTA.  Called between two __syncthreads of the caller.
__device__ __forceinline__ void team_block(TmShared& S, unsigned slot, const uint2* __restrict__ ranges,
                                           const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
                                           const GsRec* __restrict__ rec, int W, int H, int gx,
                                           const float* __restrict__ bg, float* __restrict__ final_T,
                                           uint32_t* __restrict__ n_contrib, const BfTargets& tg,
                                           const uint2* __restrict__ park_units, const float* __restrict__ park_state) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const uint2 pu = __ldcg(park_units + slot);
    const uint32_t unit = pu.x, base0 = pu.y;
    const uint32_t tile = order[unit >> 2];
    const int sub = unit & 3;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 8;
    const uint2 range = ranges[tile];
    const uint32_t total = range.y - range.x;
    const uint32_t* __restrict__ lst = list + range.x;
    const uint32_t first_b = base0 >> 5, nb = (total + 31u) >> 5;
    const float* stp = park_state + (size_t)slot * GS_PARK_WORDS + lane;
    if (warp < 2) {  // initial cull boxes from the saved done flags
        const unsigned flags = __float_as_uint(__ldcg(stp + 320));
        float box[4];
        tm_live_box((flags >> warp) & 1u, bx0, by0 + 4 * warp, box);
        if (lane < 4) S.box[warp][lane] = box[lane];
    }
    __syncthreads();
#ifdef GS_TIMELINE
    const unsigned long long tl_t0 = gtime();
    unsigned tl_hits = 0;
#endif
    if (warp < 2) {
        // ---------------------------------------------------------------- consumer c: pixel rows 4c .. 4c + 3
        const int c = warp;
        const int px = bx0 + lx, py = by0 + ly + 4 * c;
        TmPix s;
        s.T = __ldcg(stp + 32 * c);
        s.c0 = __ldcg(stp + 64 + 32 * c); s.c1 = __ldcg(stp + 128 + 32 * c); s.c2 = __ldcg(stp + 192 + 32 * c);
        s.last = __float_as_uint(__ldcg(stp + 256 + 32 * c));
        s.done = (__float_as_uint(__ldcg(stp + 320)) >> c) & 1u;
        const float (*fa)[32] = S.fa[c];
        // The FIFO as a stream: `pub` entries are published (batch flags are looked at without waiting as long as
        // entries are left), entries are applied four at a time -- eight loads in flight, four short dependent steps.
        unsigned e = 0, pub = 0, groups = 0;
        uint32_t bnext = first_b;
        int q = 0, par = 0;
        bool live = __any_sync(GS_FULL, !s.done);
#ifdef GS_TIMELINE
        long long tl_cwait = 0;
#endif
        while (live) {
#ifdef GS_TIMELINE
            const long long tl_c0 = clock64();
#endif
            while (bnext < nb) {
                if (S.bflag[q] < bnext + 1u) {
                    if (e < pub) break;
                    continue;  // nothing left to apply: wait for the producer whose turn it is
                }
                pub += S.bcnt[q][par];
                bnext++;
                if (lane == 0) S.cbatch[c] = bnext;  // (a run of batches without hits must not stall the producers)
                if (++q == TM_PROD) { q = 0; par ^= 1; }
                if (pub - e >= 4u) break;
            }
#ifdef GS_TIMELINE
            tl_cwait += clock64() - tl_c0;
#endif
            if (e == pub) break;  // bnext == nb: the list is finished
            __syncwarp();
            const unsigned n = min(pub - e, 4u);
            float al[4];
            float4 g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned r = (e + (unsigned)u) & (TM_RING - 1);
                al[u] = fa[r][lane];
                g[u] = S.fc[r];
                if ((unsigned)u >= n) {  // not published yet: applies as a no-op
                    al[u] = 0.0f;
                    g[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) tm_apply(s, al[u], g[u]);
            e += n;
#ifdef GS_TIMELINE
            tl_hits += n;
#endif
            __syncwarp();
            if (lane == 0) S.head[c] = e;
            groups++;
            if ((groups & 3u) == 0u) {
                float box[4];
                live = tm_live_box(s.done, bx0, by0 + 4 * c, box) != 0u;
                if ((!live || (groups & 31u) == 0u) && lane < 4) S.box[c][lane] = box[lane];  // shrink the cull box
            }
        }
        if (lane == 0) {  // this half of the block is finished: nothing more to wait for on its behalf
            S.head[c] = TM_FAR; S.cbatch[c] = TM_FAR;
            __threadfence_block();
            if (S.head[c ^ 1] == TM_FAR) S.stop = 1u;  // both halves: producers still at work quit
        }
        // ---- epilogue (the stores of the warp-per-block path)
        const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
        const bool ins = px < W && py < H;
        const float T = s.T;
        const size_t pid = (size_t)W * py + px;
        if (ins) { final_T[pid] = T; n_contrib[pid] = s.last; }
        float o[3];
        o[0] = s.c0 + T * bg0; o[1] = s.c1 + T * bg1; o[2] = s.c2 + T * bg2;
        size_t oplane = (size_t)H * W, op = pid;
        bool wr = ins;
        if (tg.ds) {
#pragma unroll
            for (int k = 0; k < 3; k++) o[k] = box4(o[k]);
            const int W2 = W >> 1;
            oplane = (size_t)W2 * (H >> 1);
            op = (size_t)W2 * (py >> 1) + (px >> 1);
            wr = ins && (lane & 9) == 0;
        }
        if (wr) {
            _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
                float* oc = tg.img[k];
                oc[op] = o[0]; oc[oplane + op] = o[1]; oc[2 * oplane + op] = o[2];
            }
        }
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0 && c == 0) {
            unsigned long long* tl = g_timeline + 12ull * unit;
            tl[6] = tl_t0; tl[7] = gtime(); tl[8] = ((unsigned long long)(nb - first_b) << 32) | tl_hits;
            tl[9] = (unsigned long long)tl_cwait;
        }
#endif
    } else {
        // ---------------------------------------------------------------- producer q: batches first_b + q, + 6, ...
        const int q = warp - 2;
        TmStage* __restrict__ ring = S.ring[q];
        const int px = bx0 + lx, pyA = by0 + ly, pyB = by0 + ly + 4;
        const float pfx = (float)px;
        const f2 pfy2 = pk((float)pyA, (float)pyB);
        uint32_t b = first_b + (uint32_t)q;
#pragma unroll
        for (int p = 0; p < TM_STAGES; p++) {
            const uint32_t en = (b + (uint32_t)p * TM_PROD) * 32u + lane;
            if (en < total) {
                const GsRec* r = rec + lst[en];
                cp_async16(&ring[p].a[lane], &r->a);
                cp_async16(&ring[p].b[lane], &r->b);
                cp_async16(&ring[p].c[lane], &r->c);
            }
            cp_async_commit();
        }
        int stage = 0, par = 0;
        bool quit = false;
#ifdef GS_TIMELINE
        long long tl_rec = 0, tl_tok = 0, tl_room = 0, tl_eval = 0;
        const long long tl_p0 = clock64();
#endif
        for (; b < nb; b += TM_PROD, par ^= 1) {
            if (S.stop != 0u) break;
#ifdef GS_TIMELINE
            long long tl_a = clock64();
#endif
            cp_async_wait<TM_STAGES - 1>();
            __syncwarp();
#ifdef GS_TIMELINE
            tl_rec += clock64() - tl_a;
#endif
            const TmStage& st = ring[stage];
            const uint32_t base = b * 32u;
            bool hit = false;
            if (base + lane < total) {
                const float4 a = st.a[lane], bb = st.b[lane];
                const float nBA = st.c[lane].w;
                // union of the two consumers' live boxes (an empty one has x0 > x1 and is ignored)
                const float ax0 = S.box[0][0], ax1 = S.box[0][1], ay0 = S.box[0][2], ay1 = S.box[0][3];
                const float bx0f = S.box[1][0], bx1f = S.box[1][1], by0f = S.box[1][2], by1f = S.box[1][3];
                const bool ea = ax0 > ax1, eb = bx0f > bx1f;
                const float x0 = ea ? bx0f : (eb ? ax0 : fminf(ax0, bx0f)), x1 = ea ? bx1f : (eb ? ax1 : fmaxf(ax1, bx1f));
                const float y0 = ea ? by0f : (eb ? ay0 : fminf(ay0, by0f)), y1 = ea ? by1f : (eb ? ay1 : fmaxf(ay1, by1f));
                const float bound = box_max_power(a.z, a.w, bb.x, nBA, bb.w, a.x - x1, a.x - x0, a.y - y1, a.y - y0);
                hit = !(ea && eb) && !(bound < bb.z);
            }
            unsigned mask = __ballot_sync(GS_FULL, hit);
            const unsigned cnt = __popc(mask);
            // FIFO entry number of this batch's first hit: chained from the producer of the previous batch
            unsigned cum = 0u;
#ifdef GS_TIMELINE
            tl_a = clock64();
#endif
            if (b != first_b) {
                while (S.tok_seq[q] != b) {
                    if (S.stop != 0u) { quit = true; break; }
                    tm_sleep();
                }
                cum = S.tok_cum[q];
            }
            if (quit) break;
            if (lane == 0) {
                const int nq = (q + 1 == TM_PROD) ? 0 : q + 1;
                S.tok_cum[nq] = cum + cnt;
                __threadfence_block();
                S.tok_seq[nq] = b + 1u;
            }
#ifdef GS_TIMELINE
            tl_tok += clock64() - tl_a;
            tl_a = clock64();
#endif
            // room for the whole batch, and both consumers have read this producer's record of two rounds ago
            while (true) {
                const unsigned head = min(S.head[0], S.head[1]), cb = min(S.cbatch[0], S.cbatch[1]);
                if (cum + cnt - head <= (unsigned)TM_RING && !(b >= first_b + 2u * TM_PROD && cb + 2u * TM_PROD <= b)) break;
                if (S.stop != 0u) { quit = true; break; }
#ifdef TM_ROOM_SLEEP_NS
                __nanosleep(TM_ROOM_SLEEP_NS);
#else
                tm_sleep();
#endif
            }
            if (quit) break;
#ifdef GS_TIMELINE
            tl_room += clock64() - tl_a;
            tl_a = clock64();
#endif
            unsigned e = cum;
            while (mask) {  // four instances at a time: four independent exponent / expf chains
                int j[4];
                unsigned n = 0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    j[u] = mask ? __ffs(mask) - 1 : j[0];
                    n += mask ? 1u : 0u;
                    mask &= mask - 1;
                }
                float2 av[4];
#pragma unroll
                for (int u = 0; u < 4; u++) av[u] = tm_alpha(st, j[u], pfx, pfy2);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if ((unsigned)u < n) {
                        const unsigned r = (e + (unsigned)u) & (TM_RING - 1);
                        S.fa[0][r][lane] = av[u].x;
                        S.fa[1][r][lane] = av[u].y;
                        if (lane == u) {
                            const float4 gc = st.c[j[u]];
                            S.fc[r] = make_float4(gc.x, gc.y, gc.z, __uint_as_float(base + (uint32_t)j[u] + 1u));
                        }
                    }
                }
                e += n;
            }
            __syncwarp();  // every lane has written its alphas and read this stage's records
#ifdef GS_TIMELINE
            tl_eval += clock64() - tl_a;
#endif
            if (lane == 0) {
                S.bcnt[q][par] = cnt;
                __threadfence_block();
                S.bflag[q] = b + 1u;
            }
            {
                const uint32_t en = (b + (uint32_t)TM_STAGES * TM_PROD) * 32u + lane;
                if (en < total) {
                    const GsRec* r = rec + lst[en];
                    cp_async16(&ring[stage].a[lane], &r->a);
                    cp_async16(&ring[stage].b[lane], &r->b);
                    cp_async16(&ring[stage].c[lane], &r->c);
                }
                cp_async_commit();
            }
            stage = (stage + 1 == TM_STAGES) ? 0 : stage + 1;
        }
        cp_async_wait<0>();
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0 && q == 0) {
            unsigned long long* tl = g_timeline + 12ull * unit;
            tl[10] = ((unsigned long long)tl_rec << 32) | (unsigned long long)(tl_tok & 0xffffffff);
            tl[11] = ((unsigned long long)tl_room << 32) | (unsigned long long)(tl_eval & 0xffffffff);
            tl[5] = (unsigned long long)(clock64() - tl_p0);  // producer 0's total (overwrites the warp part's loop time)
        }
#endif
    }
}

// Serves parked blocks until every fresh CTA is done and the parked queue is empty.  All eight warps of the CTA.
__device__ __forceinline__ void team_serve(TmShared& S, GsHeader* __restrict__ hdr, unsigned n_fresh,
                                           const unsigned* __restrict__ park_ready, const uint2* __restrict__ ranges,
                                           const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
                                           const GsRec* __restrict__ rec, int W, int H, int gx,
                                           const float* __restrict__ bg, float* __restrict__ final_T,
                                           uint32_t* __restrict__ n_contrib, const BfTargets& tg,
                                           const uint2* __restrict__ park_units, const float* __restrict__ park_state) {
    while (true) {
        __syncthreads();  // the previous block is finished by everybody
        if (threadIdx.x == 0) {
            unsigned slot = 0xFFFFFFFFu;
            while (true) {
                const unsigned taken = ldv(&hdr->tickets[8]);
                unsigned avail = min(ldv(&hdr->tickets[7]), (unsigned)GS_PARK_CAP);
                if (taken < avail) {
                    if (atomicCAS(&hdr->tickets[8], taken, taken + 1u) == taken) { slot = taken; break; }
                    continue;
                }
                if (ldv(&hdr->tickets[9]) >= n_fresh) {  // nobody parks any more: is the count we compared with final?
                    avail = min(ldv(&hdr->tickets[7]), (unsigned)GS_PARK_CAP);
                    if (ldv(&hdr->tickets[8]) >= avail) break;
                    continue;
                }
                __nanosleep(400);
            }
            if (slot != 0xFFFFFFFFu) {
                while (ldv(park_ready + slot) == 0u) __nanosleep(100);  // the parking warp is still writing the state
                __threadfence();
            }
            S.slot = slot;
            S.head[0] = S.head[1] = 0u; S.cbatch[0] = S.cbatch[1] = 0u; S.stop = 0u;
        }
        if (threadIdx.x < 8) { S.tok_seq[threadIdx.x] = 0u; S.bflag[threadIdx.x] = 0u; }
        __syncthreads();
        const unsigned slot = S.slot;
        if (slot == 0xFFFFFFFFu) break;
        team_block(S, slot, ranges, order, list, rec, W, H, gx, bg, final_T, n_contrib, tg, park_units, park_state);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Split walks (latency mode 2, GsScene.blend_split; NOT bit-identical, pixels within ~1e-6 of the exact walk): what no
// exact restructuring shortens is the serial walk of ONE long list (a silhouette block: 19 K entries, 3000 surviving
// instances, 0.45 ms in one warp while the rest of the GPU is idle).  Front-to-back compositing is associative --
// (C, T) of a list = (C_a + T_a C_b, T_a T_b) for any cut -- so a parked block's remaining list is cut into rounds of
// eight segments, one per warp of a CTA: every warp walks its segment with the ordinary loop from the neutral state
// (T = 1, C = 0), the eight partial results are merged in list order, and only a pixel whose early stop
// (T (1 - alpha) < 1e-4, forward.cu:353-358) falls inside the round is walked again from the merged state with the
// exact recurrence (by warp 0, cull box = just those pixels), so the stop position, n_contrib and final_T follow the
// reference's rule on the merged transmittance.  Only the association of the sums differs from the reference.
struct BfPix2 {  // compositing state of the two pixels of a lane
    f2 T2;
    float c0A, c0B, c1A, c1B, c2A, c2B;
    uint32_t lastA, lastB;
    bool doneA, doneB;
};

// The warp-per-block loop of the kernel below over list entries [beg, end) (beg a multiple of 32), state in / out.
__device__ __forceinline__ void bf_walk(BfStage<0>* __restrict__ ring, const uint32_t* __restrict__ lst,
                                        const GsRec* __restrict__ rec, uint32_t beg, uint32_t end, int bx0, int by0,
                                        float pfx, f2 pfy2, BfPix2& s, int lane) {
    float fx0 = 0.f, fx1 = 0.f, fy0 = 0.f, fy1 = 0.f;
    auto live_box = [&]() -> bool {
        const unsigned aliveA = __ballot_sync(GS_FULL, !s.doneA), aliveB = __ballot_sync(GS_FULL, !s.doneB);
        const unsigned both = aliveA | aliveB;
        if (both == 0u) return false;
        const unsigned cols = (both | (both >> 8) | (both >> 16) | (both >> 24)) & 0xffu;
        unsigned rows = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            rows |= ((aliveA >> (8 * r)) & 0xffu) ? (1u << r) : 0u;
            rows |= ((aliveB >> (8 * r)) & 0xffu) ? (16u << r) : 0u;
        }
        fx0 = (float)(bx0 + __ffs(cols) - 1); fx1 = (float)(bx0 + 31 - __clz(cols));
        fy0 = (float)(by0 + __ffs(rows) - 1); fy1 = (float)(by0 + 31 - __clz(rows));
        return true;
    };
    if (beg >= end || !live_box()) return;
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 2; p++) {
        if (beg + p * 32 + lane < end) {
            const GsRec* r = rec + lst[beg + p * 32 + lane];
            cp_async16(&ring[p].a[lane], &r->a);
            cp_async16(&ring[p].b[lane], &r->b);
            cp_async16(&ring[p].c[lane], &r->c);
        }
        cp_async_commit();
    }
    uint32_t id_next = (beg + 64 + lane < end) ? lst[beg + 64 + lane] : 0u;
    int stage = 0;
    uint32_t next_check = beg + BF_CHECK * 32;
    for (uint32_t base = beg; base < end; base += 32) {
        if (base == next_check) {
            next_check += BF_CHECK * 32;
            if (!live_box()) break;
        }
        cp_async_wait<1>();
        __syncwarp();
        {
            int nst = stage + 2; if (nst >= BF_STAGES) nst -= BF_STAGES;
            if (base + 64 + lane < end) {
                const GsRec* r = rec + id_next;
                cp_async16(&ring[nst].a[lane], &r->a);
                cp_async16(&ring[nst].b[lane], &r->b);
                cp_async16(&ring[nst].c[lane], &r->c);
            }
            cp_async_commit();
            if (base + 96 + lane < end) id_next = lst[base + 96 + lane];
        }
        const BfStage<0>& st = ring[stage];
        stage = (stage + 1 == BF_STAGES) ? 0 : stage + 1;
        bool hit = false;
        if (base + lane < end) {
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
            const float dx = ga.x - pfx;
            const f2 dy2 = sub2(bc(ga.y), pfy2);
            f2 t1 = mul2(bc(gb.x), dy2);
            t1 = mul2(dy2, t1);
            const float t2 = ga.z * dx, t3n = (-ga.w) * dx;
            const f2 t3 = mul2(dy2, bc(t3n));
            const f2 sm = fma2(bc(dx), bc(t2), t1);
            const f2 p2 = fma2(sm, bc(-0.5f), t3);
            const float pA = lo(p2), pB = hi(p2);
            float alphaA = fminf(0.99f, gb.y * expf(pA)), alphaB = fminf(0.99f, gb.y * expf(pB));
            const bool okA = !s.doneA && !(pA > 0.0f) && !(alphaA < 1.0f / 255.0f);
            const bool okB = !s.doneB && !(pB > 0.0f) && !(alphaB < 1.0f / 255.0f);
            if (!__any_sync(GS_FULL, okA || okB)) continue;
            alphaA = okA ? alphaA : 0.0f;
            alphaB = okB ? alphaB : 0.0f;
            const f2 tt2 = mul2(s.T2, sub2(bc(1.0f), pk(alphaA, alphaB)));
            const bool stopA = okA && lo(tt2) < 0.0001f, stopB = okB && hi(tt2) < 0.0001f;
            s.doneA = s.doneA || stopA;
            s.doneB = s.doneB || stopB;
            alphaA = stopA ? 0.0f : alphaA;
            alphaB = stopB ? 0.0f : alphaB;
            const f2 a2 = pk(alphaA, alphaB);
            const float4 gc = st.c[j];
            const float TA = lo(s.T2), TB = hi(s.T2);
            const f2 w0 = mul2(bc(gc.x), a2), w1 = mul2(bc(gc.y), a2), w2 = mul2(bc(gc.z), a2);
            s.c0A = __fmaf_rn(lo(w0), TA, s.c0A); s.c0B = __fmaf_rn(hi(w0), TB, s.c0B);
            s.c1A = __fmaf_rn(lo(w1), TA, s.c1A); s.c1B = __fmaf_rn(hi(w1), TB, s.c1B);
            s.c2A = __fmaf_rn(lo(w2), TA, s.c2A); s.c2B = __fmaf_rn(hi(w2), TB, s.c2B);
            s.T2 = pk(stopA ? TA : lo(tt2), stopB ? TB : hi(tt2));
            if (okA && !stopA) s.lastA = base + (uint32_t)j + 1u;
            if (okB && !stopB) s.lastB = base + (uint32_t)j + 1u;
        }
        if (__all_sync(GS_FULL, s.doneA && s.doneB)) break;
    }
    cp_async_wait<0>();
    __syncwarp();
}

#ifndef SP_GMIN
#define SP_GMIN 4    // batches per segment: at least / at most
#define SP_GMAX 32
#endif
struct SpShared {  // behind the record rings of the eight warps
    float T[BF_WARPS][64], c0[BF_WARPS][64], c1[BF_WARPS][64], c2[BF_WARPS][64];  // pixel A of lane l at l, pixel B at 32 + l
    unsigned last[BF_WARPS][64];
    unsigned stopped[BF_WARPS][32];  // bit 0 / 1: pixel A / B stopped inside the segment
    float pT[64], p0[64], p1[64], p2[64];  // state after the exact second walk (published by warp 0)
    unsigned plast[64], pdone[32];
    unsigned slot;
};

// One parked block, all eight warps of the CTA.  Called between two __syncthreads of the caller.
__device__ __forceinline__ void split_block(BfStage<0>* __restrict__ ring, SpShared& S, unsigned slot,
                                            const uint2* __restrict__ ranges, const uint32_t* __restrict__ order,
                                            const uint32_t* __restrict__ list, const GsRec* __restrict__ rec, int W, int H,
                                            int gx, const float* __restrict__ bg, float* __restrict__ final_T,
                                            uint32_t* __restrict__ n_contrib, const BfTargets& tg,
                                            const uint2* __restrict__ park_units, const float* __restrict__ park_state) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const uint2 pu = __ldcg(park_units + slot);
    const uint32_t unit = pu.x;
    const uint32_t tile = order[unit >> 2];
    const int sub = unit & 3;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int bx0 = tile_x * GS_TILE + (sub & 1) * 8, by0 = tile_y * GS_TILE + (sub >> 1) * 8;
    const uint2 range = ranges[tile];
    const uint32_t total = range.y - range.x;
    const uint32_t* __restrict__ lst = list + range.x;
    const int px = bx0 + lx, pyA = by0 + ly, pyB = by0 + ly + 4;
    const float pfx = (float)px;
    const f2 pfy2 = pk((float)pyA, (float)pyB);
    // every warp keeps the merged state of the 64 pixels (identical in all eight: the merge is deterministic)
    BfPix2 s;
    {
        const float* stp = park_state + (size_t)slot * GS_PARK_WORDS + lane;
        s.T2 = pk(__ldcg(stp), __ldcg(stp + 32));
        s.c0A = __ldcg(stp + 64); s.c0B = __ldcg(stp + 96); s.c1A = __ldcg(stp + 128); s.c1B = __ldcg(stp + 160);
        s.c2A = __ldcg(stp + 192); s.c2B = __ldcg(stp + 224);
        s.lastA = __float_as_uint(__ldcg(stp + 256)); s.lastB = __float_as_uint(__ldcg(stp + 288));
        const unsigned fl = __float_as_uint(__ldcg(stp + 320));
        s.doneA = fl & 1u; s.doneB = (fl >> 1) & 1u;
    }
#ifdef GS_TIMELINE
    const unsigned long long tl_t0 = gtime();
    unsigned tl_rounds = 0, tl_second = 0;
#endif
    uint32_t pos = pu.y;
    while (pos < total && __any_sync(GS_FULL, !s.doneA || !s.doneB)) {
        const uint32_t nb_rem = (total - pos + 31u) >> 5;
        uint32_t G = (nb_rem + BF_WARPS - 1) / BF_WARPS;
        G = G < SP_GMIN ? SP_GMIN : (G > SP_GMAX ? SP_GMAX : G);
        const uint32_t seg = G * 32u;
        {   // ---- this warp's segment from the neutral state
            BfPix2 l;
            l.T2 = bc(1.0f);
            l.c0A = l.c0B = l.c1A = l.c1B = l.c2A = l.c2B = 0.0f;
            l.lastA = l.lastB = 0u;
            l.doneA = s.doneA; l.doneB = s.doneB;
            const uint32_t sb = pos + (uint32_t)warp * seg;
            if (sb < total) bf_walk(ring, lst, rec, sb, min(total, sb + seg), bx0, by0, pfx, pfy2, l, lane);
            S.T[warp][lane] = lo(l.T2); S.T[warp][32 + lane] = hi(l.T2);
            S.c0[warp][lane] = l.c0A; S.c0[warp][32 + lane] = l.c0B;
            S.c1[warp][lane] = l.c1A; S.c1[warp][32 + lane] = l.c1B;
            S.c2[warp][lane] = l.c2A; S.c2[warp][32 + lane] = l.c2B;
            S.last[warp][lane] = l.lastA; S.last[warp][32 + lane] = l.lastB;
            S.stopped[warp][lane] = ((l.doneA && !s.doneA) ? 1u : 0u) | ((l.doneB && !s.doneB) ? 2u : 0u);
        }
        __syncthreads();
        // ---- merge in list order; a pixel whose stop may lie in segment k keeps its state of the start of k
        float TA = lo(s.T2), TB = hi(s.T2);
        int kA = BF_WARPS, kB = BF_WARPS;
#pragma unroll
        for (int k = 0; k < BF_WARPS; k++) {
            const unsigned sf = S.stopped[k][lane];
            if (!s.doneA && kA == BF_WARPS) {
                const float tt = __fmul_rn(TA, S.T[k][lane]);
                if ((sf & 1u) || tt < 0.0001f) kA = k;
                else {
                    s.c0A = __fmaf_rn(TA, S.c0[k][lane], s.c0A); s.c1A = __fmaf_rn(TA, S.c1[k][lane], s.c1A);
                    s.c2A = __fmaf_rn(TA, S.c2[k][lane], s.c2A);
                    const unsigned la = S.last[k][lane];
                    if (la) s.lastA = la;
                    TA = tt;
                }
            }
            if (!s.doneB && kB == BF_WARPS) {
                const float tt = __fmul_rn(TB, S.T[k][32 + lane]);
                if ((sf & 2u) || tt < 0.0001f) kB = k;
                else {
                    s.c0B = __fmaf_rn(TB, S.c0[k][32 + lane], s.c0B); s.c1B = __fmaf_rn(TB, S.c1[k][32 + lane], s.c1B);
                    s.c2B = __fmaf_rn(TB, S.c2[k][32 + lane], s.c2B);
                    const unsigned lb = S.last[k][32 + lane];
                    if (lb) s.lastB = lb;
                    TB = tt;
                }
            }
        }
        s.T2 = pk(TA, TB);
        const int kmin = (int)__reduce_min_sync(GS_FULL, (unsigned)min(kA, kB));
#ifdef GS_TIMELINE
        tl_rounds++;
#endif
        if (kmin < BF_WARPS) {  // (the same in all eight warps)
#ifdef GS_TIMELINE
            tl_second++;
#endif
            if (warp == 0) {  // exact second walk of the flagged pixels, each from the start of its segment
                BfPix2 r = s;
                r.doneA = true; r.doneB = true;
                for (int k = kmin; k < BF_WARPS; k++) {
                    if (kA == k) r.doneA = false;
                    if (kB == k) r.doneB = false;
                    const uint32_t sb = pos + (uint32_t)k * seg;
                    if (sb >= total) break;
                    bf_walk(ring, lst, rec, sb, min(total, sb + seg), bx0, by0, pfx, pfy2, r, lane);
                }
                if (kA < BF_WARPS) { S.pT[lane] = lo(r.T2); S.p0[lane] = r.c0A; S.p1[lane] = r.c1A; S.p2[lane] = r.c2A; S.plast[lane] = r.lastA; }
                if (kB < BF_WARPS) { S.pT[32 + lane] = hi(r.T2); S.p0[32 + lane] = r.c0B; S.p1[32 + lane] = r.c1B; S.p2[32 + lane] = r.c2B; S.plast[32 + lane] = r.lastB; }
                S.pdone[lane] = (r.doneA ? 1u : 0u) | (r.doneB ? 2u : 0u);
            }
            __syncthreads();
            const unsigned pd = S.pdone[lane];
            if (kA < BF_WARPS) {
                TA = S.pT[lane]; s.c0A = S.p0[lane]; s.c1A = S.p1[lane]; s.c2A = S.p2[lane]; s.lastA = S.plast[lane];
                s.doneA = pd & 1u;
            }
            if (kB < BF_WARPS) {
                TB = S.pT[32 + lane]; s.c0B = S.p0[32 + lane]; s.c1B = S.p1[32 + lane]; s.c2B = S.p2[32 + lane];
                s.lastB = S.plast[32 + lane];
                s.doneB = (pd >> 1) & 1u;
            }
            s.T2 = pk(TA, TB);
        }
        __syncthreads();  // everybody has read this round's partial results
        pos += BF_WARPS * seg;
    }
    if (warp == 0) {  // ---- epilogue (the stores of the warp-per-block path)
        const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
        const bool insA = px < W && pyA < H, insB = px < W && pyB < H;
        const float TA = lo(s.T2), TB = hi(s.T2);
        const size_t pidA = (size_t)W * pyA + px, pidB = (size_t)W * pyB + px;
        if (insA) { final_T[pidA] = TA; n_contrib[pidA] = s.lastA; }
        if (insB) { final_T[pidB] = TB; n_contrib[pidB] = s.lastB; }
        float oA[3], oB[3];
        oA[0] = s.c0A + TA * bg0; oA[1] = s.c1A + TA * bg1; oA[2] = s.c2A + TA * bg2;
        oB[0] = s.c0B + TB * bg0; oB[1] = s.c1B + TB * bg1; oB[2] = s.c2B + TB * bg2;
        size_t oplane = (size_t)H * W, opA = pidA, opB = pidB;
        bool wA = insA, wB = insB;
        if (tg.ds) {
#pragma unroll
            for (int c = 0; c < 3; c++) { oA[c] = box4(oA[c]); oB[c] = box4(oB[c]); }
            const int W2 = W >> 1;
            oplane = (size_t)W2 * (H >> 1);
            opA = (size_t)W2 * (pyA >> 1) + (px >> 1);
            opB = (size_t)W2 * (pyB >> 1) + (px >> 1);
            const bool writer = (lane & 9) == 0;
            wA = insA && writer; wB = insB && writer;
        }
        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
            float* oc = tg.img[k];
            if (wA) { oc[opA] = oA[0]; oc[oplane + opA] = oA[1]; oc[2 * oplane + opA] = oA[2]; }
            if (wB) { oc[opB] = oB[0]; oc[oplane + opB] = oB[1]; oc[2 * oplane + opB] = oB[2]; }
        }
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0) {
            unsigned long long* tl = g_timeline + 12ull * unit;
            tl[6] = tl_t0; tl[7] = gtime(); tl[8] = ((unsigned long long)tl_rounds << 32) | tl_second;
        }
#endif
    }
}

// Serves parked blocks (split walks) until every fresh CTA is done and the parked queue is empty.
__device__ __forceinline__ void split_serve(BfStage<0>* __restrict__ ring, SpShared& S, GsHeader* __restrict__ hdr,
                                            unsigned n_fresh, const unsigned* __restrict__ park_ready,
                                            const uint2* __restrict__ ranges, const uint32_t* __restrict__ order,
                                            const uint32_t* __restrict__ list, const GsRec* __restrict__ rec, int W, int H,
                                            int gx, const float* __restrict__ bg, float* __restrict__ final_T,
                                            uint32_t* __restrict__ n_contrib, const BfTargets& tg,
                                            const uint2* __restrict__ park_units, const float* __restrict__ park_state) {
    while (true) {
        __syncthreads();  // the previous block is finished by everybody
        if (threadIdx.x == 0) {
            unsigned slot = 0xFFFFFFFFu;
            while (true) {
                const unsigned taken = ldv(&hdr->tickets[8]);
                unsigned avail = min(ldv(&hdr->tickets[7]), (unsigned)GS_PARK_CAP);
                if (taken < avail) {
                    if (atomicCAS(&hdr->tickets[8], taken, taken + 1u) == taken) { slot = taken; break; }
                    continue;
                }
                if (ldv(&hdr->tickets[9]) >= n_fresh) {  // nobody parks any more: is the count we compared with final?
                    avail = min(ldv(&hdr->tickets[7]), (unsigned)GS_PARK_CAP);
                    if (ldv(&hdr->tickets[8]) >= avail) break;
                    continue;
                }
                __nanosleep(200);
            }
            if (slot != 0xFFFFFFFFu) {
                while (ldv(park_ready + slot) == 0u) __nanosleep(100);  // the parking warp is still writing the state
                __threadfence();
            }
            S.slot = slot;
        }
        __syncthreads();
        const unsigned slot = S.slot;
        if (slot == 0xFFFFFFFFu) break;
        split_block(ring, S, slot, ranges, order, list, rec, W, H, gx, bg, final_T, n_contrib, tg, park_units, park_state);
    }
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
template <int K, int TEAM>  // TEAM: 0 = off, 1 = exact teams, 2 = split walks (associative merge)
__global__ void __launch_bounds__(BF_WARPS * 32, (K == 0) ? BF_PX2_OCC : 2) blend_forward_px2_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_tiles, GsHeader* __restrict__ hdr,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    const BfTargets tg, const BfExtra ex, int quota, uint32_t park_at, uint2* __restrict__ park_units,
    float* __restrict__ park_state, unsigned* __restrict__ park_ready, unsigned n_fresh) {
    extern __shared__ __align__(16) unsigned char s_ring_raw[];
    static_assert(!TEAM || K == 0, "teams finish plain frames only");
    __shared__ unsigned s_arrival;
    if (TEAM) {  // roles by arrival order: a team CTA only ever waits for CTAs that are already running
        if (threadIdx.x == 0) s_arrival = atomicAdd(&hdr->tickets[11], 1u);
        __syncthreads();
    }
    const bool fresh = !TEAM || s_arrival < n_fresh;
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
    for (int served = 0; fresh && served < quota; served++) {
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
        uint32_t park_slot = GS_PARK_CAP, park_base = 0;
        for (uint32_t base = 0; base < total; base += 32) {
            if (TEAM && base == park_at) {  // a long walk: hand the block over to a team (if the queue has room)
                unsigned slot = 0;
                if (lane == 0) slot = atomicAdd(&hdr->tickets[7], 1u);
                slot = __shfl_sync(GS_FULL, slot, 0);
                if (slot < GS_PARK_CAP) { park_slot = slot; park_base = base; break; }
            }
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
        if (TEAM && park_slot < GS_PARK_CAP) {  // state of the 64 pixels, word-major so that the stores coalesce
            float* st = park_state + (size_t)park_slot * GS_PARK_WORDS + lane;
            st[0] = lo(T2); st[32] = hi(T2);
            st[64] = c0A; st[96] = c0B; st[128] = c1A; st[160] = c1B; st[192] = c2A; st[224] = c2B;
            st[256] = __uint_as_float(lastA); st[288] = __uint_as_float(lastB);
            st[320] = __uint_as_float((doneA ? 1u : 0u) | (doneB ? 2u : 0u));
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                park_units[park_slot] = make_uint2(unit, park_base);
                __threadfence();
                *reinterpret_cast<volatile unsigned*>(park_ready + park_slot) = 1u;
            }
#ifdef GS_TIMELINE
            if (g_timeline && lane == 0) {
                unsigned long long* tl = g_timeline + 12ull * unit;
                tl[0] = tl_t0; tl[1] = gtime();
                tl[2] = ((unsigned long long)smid() << 32) | total;
                tl[3] = ((unsigned long long)tl_batches << 32) | tl_hits;
                tl[4] = (unsigned long long)tl_wait; tl[5] = (unsigned long long)tl_loop;
            }
#endif
            continue;
        }
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0) {
            unsigned long long* tl = g_timeline + 12ull * unit;
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
    if (TEAM) {
        __syncthreads();  // every warp of this CTA is out of fresh blocks: the record rings are free
        if (fresh && threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&hdr->tickets[9], 1u);  // this CTA parks nothing any more
        }
        if (TEAM == 2) {
            SpShared& S = *reinterpret_cast<SpShared*>(s_ring_raw + sizeof(BfStage<0>) * BF_STAGES * BF_WARPS);
            split_serve(reinterpret_cast<BfStage<0>(*)[BF_STAGES]>(s_ring_raw)[warp], S, hdr, n_fresh, park_ready, ranges,
                        order, list, rec, W, H, gx, bg, final_T, n_contrib, tg, park_units, park_state);
        } else {
            TmShared& S = *reinterpret_cast<TmShared*>(s_ring_raw);
            team_serve(S, hdr, n_fresh, park_ready, ranges, order, list, rec, W, H, gx, bg, final_T, n_contrib, tg,
                       park_units, park_state);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Grouped hit loop (the K = 0 kernel of plain frames).  The hit loop above is one dependent chain per instance --
// ffs -> record -> exponent -> expf -> tests -> vote -> blend, ~30 dependent steps, ~280 cycles for a warp that has its
// scheduler to itself -- although only FOUR of those steps depend on the compositing state: alpha of an instance at a
// pixel is a function of the record alone.  Here the instances that survive the block cull are taken BI_G at a time:
// phase A evaluates the BI_G alphas (independent chains the scheduler interleaves, no votes, no branches), phase B
// applies them in list order with the reference's recurrence arranged so that the loop-carried chain is
// multiply -> compare -> select: test_T = T (1 - alpha) with alpha = 0 for a pixel that is done or fails the
// reference's skip tests (the factor is then exactly 1), stop test, C = fma(c alpha, T, C).  Same operations on the same
// operands in the same order as the loop above, so every result stays bit-identical.  The survivors of a batch are
// compacted into a small per-warp FIFO (every surviving lane stores ITS record at its rank among the survivors: no
// serial find-first-set chain, and a group reads BI_G consecutive FIFO entries at compile-time offsets); survivors that
// do not fill a group simply stay in the FIFO for the next batch, so a group is always full except at the end of a list
// (padded with null records: opacity 0 -> alpha 0).
// Measured at C2 (profiles/r02e_*): blend 0.604 -> 0.447 ms for one frame at a time (the longest walk -- 19 K entries,
// 3121 survivors -- 0.71 -> 0.44 ms), 2340 -> 2500 frames/s with six frames in flight.  The single-warp issue model of the
// loop (tools/sass_lonewarp.py: stall fields + scoreboards of the SASS) says 108 cycles per instance for a warp alone
// on its scheduler against 245 for the loop above; groups of 6 / 8 gain < 10 % more.  Tried on top and not kept:
// two batches per step (each lane culls two records, one wait / vote / loop overhead per 64 entries; 73 KB of shared
// memory and 110 registers: blend 0.422 ms for one frame but 2390 frames/s in flight), warps 4-7 (or 0-3) take the longest
// lists and the others the shortest (no sign of a warp-id priority in the scheduler), the first round handed out
// warp-major, 8-32 SMs reserved for the longest lists with one warp per scheduler (list length does not predict walk
// length: the longest lists are body tiles that saturate after ~40 batches), a long walk asking the warp that shares
// its scheduler to leave (the queue then drains too slowly), 185-370 CTAs instead of one per SM.
#ifndef BI_G
#define BI_G 4
#endif
#define BI_FIFO 48  // FIFO entries: a multiple of BI_G (a group never wraps) >= 32 + BI_G - 1 (one batch + left-overs)
static_assert(BI_FIFO % BI_G == 0 && BI_FIFO >= 32 + 2 * BI_G - 1, "FIFO size");
template <int K>  // K = extra colour passes blended in the same list walk (GsScene.extra_colors)
struct BiRing {  // one per warp
    float4 a[BF_STAGES][32];  // x, y, conic.x, conic.y                      } the record ring (cp.async), as BfStage
    float4 b[BF_STAGES][32];  // conic.z, opacity, thr, -B/C                 }
    float4 c[BF_STAGES][32];  // r, g, b, -B/A                               }
    float4 e[K > 0 ? K : 1][BF_STAGES][32];  // colours of the extra passes  }
    float4 fa[BI_FIFO];       // FIFO of survivors: x, y, conic.x, conic.y
    float4 fc[BI_FIFO];       //                    r, g, b, list position + 1
    float4 fe[K > 0 ? K : 1][BI_FIFO];  //          colours of the extra passes
    float2 fb[BI_FIFO];       //                    conic.z, opacity
};

// Predicated shared-memory stores (no branch, and nothing is written when the predicate is false -- a dummy "scratch"
// target for the lanes without a survivor would be a write-write race in the eyes of compute-sanitizer).
__device__ __forceinline__ void sts_if(bool p, float4* dst, float4 v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n}"
                 ::"r"((unsigned)p), "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_if(bool p, float2* dst, float2 v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.shared.v2.f32 [%1], {%2, %3};\n}"
                 ::"r"((unsigned)p), "r"(a), "f"(v.x), "f"(v.y) : "memory");
}

template <int K>
__device__ __forceinline__ void bf_fill_empty_tile(uint32_t tile, int gx, int W, int H, int lane, float bg0, float bg1,
                                                   float bg2, float* __restrict__ final_T,
                                                   uint32_t* __restrict__ n_contrib, const BfTargets& tg,
                                                   const BfExtra& ex) {
    const size_t plane = (size_t)H * W;
    const int x0 = (int)(tile % gx) * GS_TILE, y0 = (int)(tile / gx) * GS_TILE;
    // colour images: the frame's targets (this rank's image or all peers') and the images of the K extra passes
    auto for_each_image = [&](auto&& store) {
        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) store(tg.img[k]);
        _Pragma("unroll") for (int k = 0; k < K; k++) store(ex.out[k]);
    };
    if (tg.ds) {  // half-resolution colour (8x8 per tile), full-resolution T / contributor count
        const int W2 = W >> 1, H2 = H >> 1;
        const size_t plane2 = (size_t)W2 * H2;
        const int qx = (x0 >> 1) + (lane & 7);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int qy = (y0 >> 1) + r * 4 + (lane >> 3);
            if (qx < W2 && qy < H2) {
                const size_t pid = (size_t)W2 * qy + qx;
                for_each_image([&](float* oc) { oc[pid] = bg0; oc[plane2 + pid] = bg1; oc[2 * plane2 + pid] = bg2; });
            }
        }
        for (int r = 0; r < 8; r++) {
            const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
            if (px < W && py < H) {
                const size_t pid = (size_t)W * py + px;
                final_T[pid] = 1.f; n_contrib[pid] = 0u;
            }
        }
    } else if ((W & 3) == 0 && x0 + GS_TILE <= W) {
        const int px = x0 + (lane & 3) * 4;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int py = y0 + r * 8 + (lane >> 2);
            if (py < H) {
                const size_t pid = (size_t)W * py + px;
                for_each_image([&](float* oc) {
                    *reinterpret_cast<float4*>(oc + pid) = make_float4(bg0, bg0, bg0, bg0);
                    *reinterpret_cast<float4*>(oc + plane + pid) = make_float4(bg1, bg1, bg1, bg1);
                    *reinterpret_cast<float4*>(oc + 2 * plane + pid) = make_float4(bg2, bg2, bg2, bg2);
                });
                *reinterpret_cast<float4*>(final_T + pid) = make_float4(1.f, 1.f, 1.f, 1.f);
                *reinterpret_cast<uint4*>(n_contrib + pid) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else {
        for (int r = 0; r < 8; r++) {
            const int px = x0 + (lane & 15), py = y0 + r * 2 + (lane >> 4);
            if (px < W && py < H) {
                const size_t pid = (size_t)W * py + px;
                for_each_image([&](float* oc) { oc[pid] = bg0; oc[plane + pid] = bg1; oc[2 * plane + pid] = bg2; });
                final_T[pid] = 1.f; n_contrib[pid] = 0u;
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(BF_WARPS * 32, (K == 0) ? BF_PX2_OCC : 2) blend_forward_grouped_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, uint32_t num_tiles, GsHeader* __restrict__ hdr,
    const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, const BfTargets tg,
    const BfExtra ex, int quota) {
    extern __shared__ __align__(16) unsigned char s_ring_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    BiRing<K>& R = reinterpret_cast<BiRing<K>*>(s_ring_raw)[warp];
    unsigned* const q_fresh = &hdr->tickets[6];
    const uint32_t nonempty = hdr->nonempty_tiles;
    const uint32_t blend_units = nonempty * 4u;
    const uint32_t num_units = blend_units + (num_tiles - nonempty);

    for (int served = 0; served < quota; served++) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(q_fresh, 1u);
        unit = __shfl_sync(GS_FULL, unit, 0);
        if (unit >= num_units) break;
        if (unit >= blend_units) {  // ---- empty tile: colour = background, T = 1, no contributor
            bf_fill_empty_tile<K>(order[nonempty + (unit - blend_units)], gx, W, H, lane, bg0, bg1, bg2, final_T, n_contrib, tg, ex);
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
        long long tl_wait = 0, tl_loop = 0, tl_issue = 0, tl_cull = 0;
#endif
        bool doneA = !insA, doneB = !insB;
        f2 T2 = bc(1.0f);
        float c0A = 0.f, c0B = 0.f, c1A = 0.f, c1B = 0.f, c2A = 0.f, c2B = 0.f;
        f2 E[K > 0 ? K : 1][3];  // colour accumulators of the extra passes
#pragma unroll
        for (int k = 0; k < K; k++) { E[k][0] = bc(0.f); E[k][1] = bc(0.f); E[k][2] = bc(0.f); }
        uint32_t lastA = 0, lastB = 0;

        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (p * 32 + lane < total) {
                const uint32_t id = lst[p * 32 + lane];
                const GsRec* r = rec + id;
                cp_async16(&R.a[p][lane], &r->a);
                cp_async16(&R.b[p][lane], &r->b);
                cp_async16(&R.c[p][lane], &r->c);
#pragma unroll
                for (int k = 0; k < K; k++) cp_async16(&R.e[k][p][lane], ex.xrec + 3 * (size_t)id + k);
            }
            cp_async_commit();
        }
        // list indices run two batches ahead of the record prefetch (a dependent load chain: index -> record)
        uint32_t id_next = (64 + lane < total) ? lst[64 + lane] : 0u;
        uint32_t id_next2 = (96 + lane < total) ? lst[96 + lane] : 0u;

        int stage = 0;
        uint32_t next_check = BF_CHECK * 32;
        unsigned head = 0, avail = 0;  // FIFO: first unconsumed entry (a multiple of BI_G, < BI_FIFO), entries waiting
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
            const long long tl_c2 = clock64();
#endif
            {
                int nst = stage + 2; if (nst >= BF_STAGES) nst -= BF_STAGES;
                if (base + 64 + lane < total) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&R.a[nst][lane], &r->a);
                    cp_async16(&R.b[nst][lane], &r->b);
                    cp_async16(&R.c[nst][lane], &r->c);
#pragma unroll
                    for (int k = 0; k < K; k++) cp_async16(&R.e[k][nst][lane], ex.xrec + 3 * (size_t)id_next + k);
                }
                cp_async_commit();
                id_next = id_next2;
                if (base + 128 + lane < total) id_next2 = lst[base + 128 + lane];
            }
#ifdef GS_TIMELINE
            const long long tl_c3 = clock64();
            tl_issue += tl_c3 - tl_c2;
#endif
            const int sb = stage;
            stage = (stage + 1 == BF_STAGES) ? 0 : stage + 1;

            // cull, without branches: lanes behind the end of the list read stale ring slots and are masked out
            const float4 ra = R.a[sb][lane], rb = R.b[sb][lane], rc = R.c[sb][lane];
            const float bound = box_max_power_sel(ra.z, ra.w, rb.x, rc.w, rb.w, ra.x - fx1, ra.x - fx0, ra.y - fy1, ra.y - fy0);
            const bool hit = (base + lane < total) && !(bound < rb.z);
            const unsigned mask = __ballot_sync(GS_FULL, hit);
            const bool final_batch = base + 32 >= total;
#ifdef GS_TIMELINE
            tl_hits += __popc(mask);
            const long long tl_c1 = clock64();
            tl_cull += tl_c1 - tl_c3;
#endif
            {   // compaction: the survivor of lane l goes to FIFO entry head + avail + (survivors in lower lanes)
                unsigned pos = head + avail + (unsigned)__popc(mask & ((1u << lane) - 1u));
                if (pos >= (unsigned)BI_FIFO) pos -= (unsigned)BI_FIFO;
                sts_if(hit, &R.fa[pos], ra);
                sts_if(hit, &R.fb[pos], make_float2(rb.x, rb.y));
                sts_if(hit, &R.fc[pos], make_float4(rc.x, rc.y, rc.z, __uint_as_float(base + (uint32_t)lane + 1u)));
#pragma unroll
                for (int k = 0; k < K; k++) sts_if(hit, &R.fe[k][pos], R.e[k][sb][lane]);
                avail += (unsigned)__popc(mask);
                // the last batch pads the last group with null records (opacity 0 -> alpha 0)
                const unsigned pad = final_batch ? ((unsigned)BI_G - avail % BI_G) % BI_G : 0u;
                unsigned pp = head + avail + (unsigned)lane;
                if (pp >= (unsigned)BI_FIFO) pp -= (unsigned)BI_FIFO;
                const bool padder = (unsigned)lane < pad;
                sts_if(padder, &R.fa[pp], make_float4(0.f, 0.f, 0.f, 0.f));
                sts_if(padder, &R.fb[pp], make_float2(0.f, 0.f));
                sts_if(padder, &R.fc[pp], make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
                for (int k = 0; k < K; k++) sts_if(padder, &R.fe[k][pp], make_float4(0.f, 0.f, 0.f, 0.f));
                avail += pad;
                __syncwarp();
            }
            while (avail >= (unsigned)BI_G) {
                const float4* __restrict__ Fa = R.fa + head;
                const float2* __restrict__ Fb = R.fb + head;
                const float4* __restrict__ Fc = R.fc + head;
                const unsigned head0 = head;
                head = (head + BI_G == (unsigned)BI_FIFO) ? 0u : head + BI_G;
                avail -= BI_G;
                // ---- phase A: alpha of each instance at the two pixels of this lane (0 = the reference skips it);
                // operation for operation the loop above (forward.cu:336-338 as compiled, libdevice expf)
                f2 al[BI_G];
#pragma unroll
                for (int u = 0; u < BI_G; u++) {
                    const float4 ga = Fa[u];
                    const float2 gb = Fb[u];
                    const float dx = ga.x - pfx;
                    const f2 dy2 = sub2(bc(ga.y), pfy2);
                    f2 t1 = mul2(bc(gb.x), dy2);
                    t1 = mul2(dy2, t1);
                    const float t2 = ga.z * dx, t3n = (-ga.w) * dx;
                    const f2 t3 = mul2(dy2, bc(t3n));
                    const f2 sm = fma2(bc(dx), bc(t2), t1);
                    const f2 p2 = fma2(sm, bc(-0.5f), t3);
                    const float pA = lo(p2), pB = hi(p2);
                    const float alphaA = fminf(0.99f, gb.y * expf(pA)), alphaB = fminf(0.99f, gb.y * expf(pB));
                    const bool okA = !(pA > 0.0f) && !(alphaA < 1.0f / 255.0f);
                    const bool okB = !(pB > 0.0f) && !(alphaB < 1.0f / 255.0f);
                    al[u] = pk(okA ? alphaA : 0.0f, okB ? alphaB : 0.0f);
                }
                // ---- phase B: the recurrence, in list order
#pragma unroll
                for (int u = 0; u < BI_G; u++) {
                    const float aA = lo(al[u]), aB = hi(al[u]);
                    const f2 om = sub2(bc(1.0f), al[u]);
                    const f2 tt2 = mul2(T2, pk(doneA ? 1.0f : lo(om), doneB ? 1.0f : hi(om)));
                    // T never is below the stop threshold itself, so a pixel that is done or skipped cannot stop here
                    const bool dA = doneA || lo(tt2) < 0.0001f, dB = doneB || hi(tt2) < 0.0001f;
                    const f2 eff = pk(dA ? 0.0f : aA, dB ? 0.0f : aB);
                    const float4 gc = Fc[u];
                    const float TA = lo(T2), TB = hi(T2);
                    const f2 w0 = mul2(bc(gc.x), eff), w1 = mul2(bc(gc.y), eff), w2 = mul2(bc(gc.z), eff);
                    c0A = __fmaf_rn(lo(w0), TA, c0A); c0B = __fmaf_rn(hi(w0), TB, c0B);
                    c1A = __fmaf_rn(lo(w1), TA, c1A); c1B = __fmaf_rn(hi(w1), TB, c1B);
                    c2A = __fmaf_rn(lo(w2), TA, c2A); c2B = __fmaf_rn(hi(w2), TB, c2B);
#pragma unroll
                    for (int k = 0; k < K; k++) {  // the extra passes: same alpha, same T, other colours
                        const float4 ge = R.fe[k][head0 + u];
                        fma2_acc(E[k][0], mul2(bc(ge.x), eff), T2);
                        fma2_acc(E[k][1], mul2(bc(ge.y), eff), T2);
                        fma2_acc(E[k][2], mul2(bc(ge.z), eff), T2);
                    }
                    T2 = pk(dA ? TA : lo(tt2), dB ? TB : hi(tt2));
                    if (!dA && aA != 0.0f) lastA = __float_as_uint(gc.w);
                    if (!dB && aB != 0.0f) lastB = __float_as_uint(gc.w);
                    doneA = dA; doneB = dB;
                }
            }
            __syncwarp();  // the groups' reads are over before the next batch appends
#ifdef GS_TIMELINE
            tl_loop += clock64() - tl_c1;
#endif
            if (__all_sync(GS_FULL, doneA && doneB)) break;  // (pending survivors would change nothing)
        }
        cp_async_wait<0>();
#ifdef GS_TIMELINE
        if (g_timeline && lane == 0) {
            unsigned long long* tl = g_timeline + 12ull * unit;
            tl[0] = tl_t0; tl[1] = gtime();
            tl[2] = ((unsigned long long)smid() << 32) | total;
            tl[3] = ((unsigned long long)tl_batches << 32) | tl_hits;
            tl[4] = (unsigned long long)tl_wait; tl[5] = (unsigned long long)tl_loop;
            tl[9] = (unsigned long long)tl_issue; tl[10] = (unsigned long long)tl_cull;
        }
#endif
        // ---- epilogue
        const float TA = lo(T2), TB = hi(T2);
        const size_t pidA = (size_t)W * pyA + px, pidB = (size_t)W * pyB + px;
        if (insA) { final_T[pidA] = TA; n_contrib[pidA] = lastA; }
        if (insB) { final_T[pidB] = TB; n_contrib[pidB] = lastB; }
        float oA[3 * (K + 1)], oB[3 * (K + 1)];
        oA[0] = c0A + TA * bg0; oA[1] = c1A + TA * bg1; oA[2] = c2A + TA * bg2;
        oB[0] = c0B + TB * bg0; oB[1] = c1B + TB * bg1; oB[2] = c2B + TB * bg2;
#pragma unroll
        for (int k = 0; k < K; k++) {
            oA[3 * k + 3] = lo(E[k][0]) + TA * bg0; oA[3 * k + 4] = lo(E[k][1]) + TA * bg1;
            oA[3 * k + 5] = lo(E[k][2]) + TA * bg2;
            oB[3 * k + 3] = hi(E[k][0]) + TB * bg0; oB[3 * k + 4] = hi(E[k][1]) + TB * bg1;
            oB[3 * k + 5] = hi(E[k][2]) + TB * bg2;
        }
        size_t oplane = (size_t)H * W, opA = pidA, opB = pidB;
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
        _Pragma("unroll") for (int k = 0; k < 8; k++) if (k < tg.n) {
            float* oc = tg.img[k];
            if (wA) { oc[opA] = oA[0]; oc[oplane + opA] = oA[1]; oc[2 * oplane + opA] = oA[2]; }
            if (wB) { oc[opB] = oB[0]; oc[oplane + opB] = oB[1]; oc[2 * oplane + opB] = oB[2]; }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            float* oc = ex.out[k];
            if (wA) { oc[opA] = oA[3 * k + 3]; oc[oplane + opA] = oA[3 * k + 4]; oc[2 * oplane + opA] = oA[3 * k + 5]; }
            if (wB) { oc[opB] = oB[3 * k + 3]; oc[oplane + opB] = oB[3 * k + 4]; oc[2 * oplane + opB] = oB[3 * k + 5]; }
        }
    }
}

// per extra-pass count K (index 4 / 5: the K = 0 kernel with teams / split walks): value[0] = resident CTAs of the kernel on this
// device, value[1] = default hand-over threshold
GsPerDevice g_blend_dev[6];

template <int K, int TEAM>
cudaError_t launch_blend(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im, float* out_color,
                         uint32_t num_tiles, int team_after) {
    const size_t ring_bytes = sizeof(BfStage<K>) * BF_STAGES * BF_WARPS;
    const size_t smem = TEAM == 2 ? ring_bytes + sizeof(SpShared)  // split walks keep the rings
                                  : (TEAM == 1 && sizeof(TmShared) > ring_bytes ? sizeof(TmShared) : ring_bytes);  // teams alias them
    const int* dv = nullptr;
    {
        cudaError_t e = g_blend_dev[TEAM ? 3 + TEAM : K].get(&dv, [smem](int dev, int* v) {
            int sms = 0, per_sm = 0;
            cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(blend_forward_px2_kernel<K, TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_px2_kernel<K, TEAM>, BF_WARPS * 32, smem);
            if (e != cudaSuccess) return e;
            v[0] = sms * (per_sm > 0 ? per_sm : 1);
            v[1] = sms;
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
    const uint32_t units_max = num_tiles * 4u;
    unsigned grid, n_fresh;
    uint32_t quota;
    if (TEAM) {
        // latency mode: the whole GPU is this frame's; BF_TEAM_CTAS CTAs per SM serve parked blocks from the start, the
        // others work the fresh queue (and become teams when it is empty).  The grid is never larger than what is
        // resident at once, and roles go by arrival order, so a team never waits for a CTA that cannot start.
        const unsigned teams = (unsigned)dv[1] * BF_TEAM_CTAS < (unsigned)resident ? (unsigned)dv[1] * BF_TEAM_CTAS
                                                                                     : (unsigned)resident / 2u;
        n_fresh = (unsigned)resident - teams;
        quota = (units_max + BF_WARPS * n_fresh - 1) / (BF_WARPS * n_fresh);
        if (quota < 4u) quota = 4u;
        n_fresh = (units_max + BF_WARPS * quota - 1) / (BF_WARPS * quota);
        grid = n_fresh + teams;
    } else {
        // grid x 8 warps x quota covers the upper bound of units (4 per tile); quota >= 16, grid <= 60 % of the slots
        const uint32_t slots = (uint32_t)((resident * BF_SLOT_NUM + BF_SLOT_DEN - 1) / BF_SLOT_DEN);
        quota = (units_max + BF_WARPS * slots - 1) / (BF_WARPS * slots);
        if (quota < 16u) quota = 16u;
        grid = (units_max + BF_WARPS * quota - 1) / (BF_WARPS * quota);
        n_fresh = grid;
    }
    const uint32_t park_at = TEAM ? (uint32_t)team_after * 32u : 0xFFFFFFFFu;
    blend_forward_px2_kernel<K, TEAM><<<grid, BF_WARPS * 32, smem, f.stream>>>(
        im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height, f.gx, num_tiles, g.hdr, f.s.background, im.final_T,
        im.n_contrib, tg, ex, (int)quota, park_at, im.park_units, im.park_state, im.park_ready, n_fresh);
    gs_note_launch();
    return cudaGetLastError();
}

GsPerDevice g_grouped_dev[4];  // per extra-pass count K: value[1] = SMs of the device

template <int K>
cudaError_t launch_blend_grouped(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                 float* out_color, uint32_t num_tiles) {
    const size_t smem = sizeof(BiRing<K>) * BF_WARPS;
    const int* dv = nullptr;
    {
        cudaError_t e = g_grouped_dev[K].get(&dv, [smem](int dev, int* v) {
            int sms = 0, per_sm = 0;
            cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(blend_forward_grouped_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_forward_grouped_kernel<K>, BF_WARPS * 32, smem);
            if (e != cudaSuccess) return e;
            v[0] = sms * (per_sm > 0 ? per_sm : 1);
            v[1] = sms;
            return cudaSuccess;
        });
        if (e != cudaSuccess) return e;
    }
    BfTargets tg;
    tg.n = f.s.num_peers > 0 ? f.s.num_peers : 1;
    tg.ds = f.s.downsample == 2 ? 1 : 0;
    for (int k = 0; k < 8; k++) tg.img[k] = f.s.num_peers > 0 ? f.s.peer_out_color[k] : out_color;
    BfExtra ex;
    ex.xrec = g.xrec;
    for (int k = 0; k < 3; k++) ex.out[k] = f.s.extra_out[k];
    const uint32_t units_max = num_tiles * 4u;
    // One CTA per SM (two warps per scheduler) and a per-warp unit quota that covers the upper bound of units: measured
    // best for a single frame (long walks share their scheduler with one other warp instead of three) AND with
    // frames in flight (the binning kernels of the next frames find room next to it); profiles/r02e_blend_grid_c2.txt
    // (the quota is twice the even share, so that warps held up by long walks do not leave work behind in a dense frame)
    const uint32_t slots = (uint32_t)dv[1];
    const uint32_t share = (units_max + BF_WARPS * slots - 1) / (BF_WARPS * slots);
    const uint32_t quota = 2u * (share < 1u ? 1u : share);
    const unsigned grid = units_max < BF_WARPS * slots ? (units_max + BF_WARPS - 1) / BF_WARPS : slots;
    blend_forward_grouped_kernel<K><<<grid, BF_WARPS * 32, smem, f.stream>>>(
        im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height, f.gx, num_tiles, g.hdr, f.s.background, im.final_T,
        im.n_contrib, tg, ex, (int)quota);
    gs_note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color) {
    const uint32_t num_tiles = (uint32_t)f.gx * (uint32_t)(f.row1 - f.row0);
    if (num_tiles == 0) return cudaSuccess;
    // Teams (latency mode): GsScene.team_after > 0 = hand-over threshold in batches, 0 = library default
    // (BF_TEAM_AFTER, or the environment variable GSPLAT_B200_TEAM_AFTER), < 0 = off (throughput mode)
    int after = f.s.team_after;
    if (after == 0) {
        static const int env_after = [] {
            const char* e = getenv("GSPLAT_B200_TEAM_AFTER");
            return e ? atoi(e) : BF_TEAM_AFTER;
        }();
        after = env_after;
    }
    static const int plain_k = [] { const char* e = getenv("GSPLAT_B200_BLEND_PLAIN"); return e ? atoi(e) : 0; }();
    switch (f.s.num_extra) {
        case 1: return plain_k ? launch_blend<1, 0>(f, g, b, im, out_color, num_tiles, 0) : launch_blend_grouped<1>(f, g, b, im, out_color, num_tiles);
        case 2: return plain_k ? launch_blend<2, 0>(f, g, b, im, out_color, num_tiles, 0) : launch_blend_grouped<2>(f, g, b, im, out_color, num_tiles);
        case 3: return plain_k ? launch_blend<3, 0>(f, g, b, im, out_color, num_tiles, 0) : launch_blend_grouped<3>(f, g, b, im, out_color, num_tiles);
        default:
            if (f.s.blend_split > 0) return launch_blend<0, 2>(f, g, b, im, out_color, num_tiles, f.s.blend_split);
            if (after > 0) return launch_blend<0, 1>(f, g, b, im, out_color, num_tiles, after);
            {   // developer switch: GSPLAT_B200_BLEND_PLAIN=1 selects the one-instance-per-iteration loop (A/B runs)
                static const int plain = [] { const char* e = getenv("GSPLAT_B200_BLEND_PLAIN"); return e ? atoi(e) : 0; }();
                if (plain) return launch_blend<0, 0>(f, g, b, im, out_color, num_tiles, 0);
            }
            return launch_blend_grouped<0>(f, g, b, im, out_color, num_tiles);
    }
}

#ifdef GS_TIMELINE
extern "C" int gs_debug_timeline(void* dev_buf) {
    unsigned long long* p = (unsigned long long*)dev_buf;
    return (int)cudaMemcpyToSymbol(g_timeline, &p, sizeof(p));
}
#endif
