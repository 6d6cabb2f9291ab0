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
//   * the 9 per-(pixel,Gaussian) partial derivatives are summed across the warp with a transposing
//     shuffle reduction before they reach memory, so a Gaussian receives 9 atomics per WARP that touches it instead of
//     9 per PIXEL (backward.cu:523-554): 10-32x fewer L2 atomics.  Summation order differs from the reference's
//     (which is itself non-deterministic); gradients agree to fp32 round-off (measured 1e-6 relative at C3).
//   * the survivors of the block cull are taken two at a time (BbRing below): 16-value transposing reduction, one
//     atomic instruction per pair; C3 backward 1.79 -> 1.44 ms (profiles/r02f_backward_c3_*.json).
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

// The same bound without branches (blend_forward.cu: box_max_power_sel).
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

// Predicated shared-memory store (no branch; nothing is written when the predicate is false).
__device__ __forceinline__ void sts_if(bool p, float4* dst, float4 v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n}"
                 ::"r"((unsigned)p), "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#define BB_WARPS 8
#define BB_STAGES 3
#define BB_FIFO 36  // FIFO entries: even (a pair never wraps), >= 32 + 1 left-over + 2 padding

// Per-warp shared memory: the record ring (cp.async, as in the forward kernel) and the FIFO of the instances that
// survive the block cull.  As in blend_forward_grouped_kernel the survivors of a batch are compacted into the FIFO (every
// surviving lane stores ITS record at its rank) and taken two at a time: everything that decides whether a pixel takes
// part -- list position against the pixel's last contributor, exponent, alpha -- depends on the record alone, so the
// two alpha chains are evaluated together, the per-pixel recurrences (T / (1 - alpha), the running colour behind the
// instance) follow in walk order, and the 2 x 8 partial derivatives are summed across the warp by ONE 16-value
// transposing reduction (8 + 4 + 2 + 1 + 1 shuffles) plus 5 for the two opacity terms: 10.5 shuffles and one atomic
// instruction per instance instead of 19 and two.
struct BbRing {
    float4 a[BB_STAGES][32];   // x, y, conic.x, conic.y
    float4 b[BB_STAGES][32];   // conic.z, opacity, thr, -B/C
    float4 c[BB_STAGES][32];   // r, g, b, -B/A
    uint32_t id[BB_STAGES][32];
    float4 fa[BB_FIFO];        // FIFO: x, y, conic.x, conic.y
    float4 fb[BB_FIFO];        //       conic.z, opacity, thr, list position (the reference's `contributor`)
    float4 fc[BB_FIFO];        //       r, g, b, Gaussian index
};

#ifndef BB_OCC
#define BB_OCC 3
#endif
__global__ void __launch_bounds__(BB_WARPS * 32, BB_OCC) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const uint32_t* __restrict__ list,
    const GsRec* __restrict__ rec, int W, int H, int gx, const GsHeader* __restrict__ hdr,
    unsigned int* __restrict__ queue, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, int ds, float* __restrict__ dL_dmean2D,
    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
    extern __shared__ __align__(16) unsigned char s_raw[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    const size_t plane = (size_t)H * W;
    const float ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
    BbRing& R = reinterpret_cast<BbRing*>(s_raw)[warp];
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
        const float bg_dot_dpixel = bg0 * dpx + bg1 * dpy + bg2 * dpz;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;

        // batch k, slot l holds list position need-1 - (32k + l): slot 0 of batch 0 is the farthest needed entry
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (p * 32 + lane < need) {
                const uint32_t id = lst[need - 1 - (p * 32 + lane)];
                const GsRec* r = rec + id;
                cp_async16(&R.a[p][lane], &r->a);
                cp_async16(&R.b[p][lane], &r->b);
                cp_async16(&R.c[p][lane], &r->c);
                R.id[p][lane] = id;
            }
            cp_async_commit();
        }
        uint32_t id_next = (64 + lane < need) ? lst[need - 1 - (64 + lane)] : 0u;
        uint32_t id_next2 = (96 + lane < need) ? lst[need - 1 - (96 + lane)] : 0u;

        int stage = 0;
        unsigned head = 0, avail = 0;  // FIFO: first unconsumed entry (even, < BB_FIFO), entries waiting
        for (uint32_t base = 0; base < need; base += 32) {
            cp_async_wait<1>();
            __syncwarp();
            {
                int nst = stage + 2; if (nst >= BB_STAGES) nst -= BB_STAGES;
                if (base + 64 + lane < need) {
                    const GsRec* r = rec + id_next;
                    cp_async16(&R.a[nst][lane], &r->a);
                    cp_async16(&R.b[nst][lane], &r->b);
                    cp_async16(&R.c[nst][lane], &r->c);
                    R.id[nst][lane] = id_next;
                }
                cp_async_commit();
                id_next = id_next2;
                if (base + 128 + lane < need) id_next2 = lst[need - 1 - (base + 128 + lane)];
            }
            const int sb = stage;
            stage = (stage + 1 == BB_STAGES) ? 0 : stage + 1;

            // cull, without branches: lanes past the start of the walk read stale ring slots and are masked out
            const float4 ra = R.a[sb][lane], rb = R.b[sb][lane], rc = R.c[sb][lane];
            const uint32_t rid = R.id[sb][lane];
            const float bound = box_max_power_sel(ra.z, ra.w, rb.x, rc.w, rb.w, ra.x - fx1, ra.x - fx0, ra.y - fy1, ra.y - fy0);
            const bool hit = (base + lane < need) && !(bound < rb.z);
            const unsigned mask = __ballot_sync(GS_FULL, hit);
            const bool final_batch = base + 32 >= need;
            {   // compaction: the survivor of lane l goes to FIFO entry head + avail + (survivors in lower lanes)
                unsigned pos = head + avail + (unsigned)__popc(mask & ((1u << lane) - 1u));
                if (pos >= (unsigned)BB_FIFO) pos -= (unsigned)BB_FIFO;
                sts_if(hit, &R.fa[pos], ra);
                sts_if(hit, &R.fb[pos], make_float4(rb.x, rb.y, rb.z, __uint_as_float(need - 1u - (base + (uint32_t)lane))));
                sts_if(hit, &R.fc[pos], make_float4(rc.x, rc.y, rc.z, __uint_as_float(rid)));
                avail += (unsigned)__popc(mask);
                // the last batch pads the last pair with a null record (list position 2^32 - 1: no pixel takes part)
                const unsigned pad = final_batch ? (avail & 1u) : 0u;
                unsigned pp = head + avail;
                if (pp >= (unsigned)BB_FIFO) pp -= (unsigned)BB_FIFO;
                const bool padder = pad != 0u && lane == 0;
                sts_if(padder, &R.fa[pp], make_float4(0.f, 0.f, 0.f, 0.f));
                sts_if(padder, &R.fb[pp], make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu)));
                sts_if(padder, &R.fc[pp], make_float4(0.f, 0.f, 0.f, 0.f));
                avail += pad;
                __syncwarp();
            }
            while (avail >= 2u) {
                const float4* __restrict__ Fa = R.fa + head;
                const float4* __restrict__ Fb = R.fb + head;
                const float4* __restrict__ Fc = R.fc + head;
                head = (head + 2u == (unsigned)BB_FIFO) ? 0u : head + 2u;
                avail -= 2u;
                // ---- phase A: which pixels take part, G and alpha of both instances (record data only)
                float dxs[2], dys[2], Gs[2], als[2], ops[2];
                bool act[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const float4 a = Fa[u], bq = Fb[u];
                    const float dx = a.x - pfx, dy = a.y - pfy;
                    const float power = -0.5f * (a.z * dx * dx + bq.x * dy * dy) - a.w * dx * dy;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, bq.y * G);
                    act[u] = __float_as_uint(bq.w) < last_contributor && !(power > 0.0f) && !(power < bq.z) &&
                             !(alpha < 1.0f / 255.0f);
                    dxs[u] = dx; dys[u] = dy; Gs[u] = G; als[u] = alpha; ops[u] = bq.y;
                }
                if (!__any_sync(GS_FULL, act[0] || act[1])) continue;
                // ---- phase B: the per-pixel recurrences and the nine partial derivatives, in walk order
                float v[2][8], v8[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[u][k] = 0.f;
                    v8[u] = 0.f;
                    if (act[u]) {
                        const float4 a = Fa[u], bq = Fb[u], c = Fc[u];
                        const float alpha = als[u], G = Gs[u], dx = dxs[u], dy = dys[u];
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.0f;
                        acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = c.x; dL_dalpha += (c.x - acc0) * dpx;
                        acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = c.y; dL_dalpha += (c.y - acc1) * dpy;
                        acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = c.z; dL_dalpha += (c.z - acc2) * dpz;
                        v[u][0] = dchannel_dcolor * dpx;
                        v[u][1] = dchannel_dcolor * dpy;
                        v[u][2] = dchannel_dcolor * dpz;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                        const float dL_dG = ops[u] * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * a.z - gdy * a.w;
                        const float dG_ddely = -gdy * bq.x - gdx * a.w;
                        v[u][3] = dL_dG * dG_ddelx * ddelx_dx;
                        v[u][4] = dL_dG * dG_ddely * ddely_dy;
                        v[u][5] = -0.5f * gdx * dx * dL_dG;
                        v[u][6] = -0.5f * gdx * dy * dL_dG;
                        v[u][7] = -0.5f * gdy * dy * dL_dG;
                        v8[u] = G * dL_dalpha;
                    }
                }
                // ---- phase C: transposing warp reduction of the 16 values: afterwards every lane holds the warp total
                // of value (instance lane >> 4, component (lane >> 1) & 7)
                float w8[8], w4[4], w2[2], w1;
                {
                    const bool hi = (lane & 16) != 0;  // upper half keeps instance 1, lower half instance 0
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float send = hi ? v[0][k] : v[1][k];
                        const float keep = hi ? v[1][k] : v[0][k];
                        w8[k] = keep + __shfl_xor_sync(GS_FULL, send, 16);
                    }
                }
                {
                    const bool hi = (lane & 8) != 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float send = hi ? w8[k] : w8[k + 4];
                        const float keep = hi ? w8[k + 4] : w8[k];
                        w4[k] = keep + __shfl_xor_sync(GS_FULL, send, 8);
                    }
                }
                {
                    const bool hi = (lane & 4) != 0;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const float send = hi ? w4[k] : w4[k + 2];
                        const float keep = hi ? w4[k + 2] : w4[k];
                        w2[k] = keep + __shfl_xor_sync(GS_FULL, send, 4);
                    }
                }
                {
                    const bool hi = (lane & 2) != 0;
                    const float send = hi ? w2[0] : w2[1];
                    const float keep = hi ? w2[1] : w2[0];
                    w1 = keep + __shfl_xor_sync(GS_FULL, send, 2);
                }
                w1 += __shfl_xor_sync(GS_FULL, w1, 1);
                float o8;  // opacity term: upper half keeps instance 1
                {
                    const bool hi = (lane & 16) != 0;
                    const float send = hi ? v8[0] : v8[1];
                    const float keep = hi ? v8[1] : v8[0];
                    o8 = keep + __shfl_xor_sync(GS_FULL, send, 16);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) o8 += __shfl_xor_sync(GS_FULL, o8, o);
                // ---- phase D: one atomic instruction for the 16 totals, one for the two opacity terms
                const uint32_t id = __float_as_uint((lane & 16) ? Fc[1].w : Fc[0].w);
                if ((lane & 1) == 0) {
                    const int k = (lane >> 1) & 7;  // component: bit 2 <- lane bit 3, bit 1 <- lane bit 2, bit 0 <- lane bit 1
                    float* dst;
                    if (k < 3) dst = dL_dcolor + 3 * (size_t)id + k;
                    else if (k < 5) dst = dL_dmean2D + 3 * (size_t)id + (k - 3);
                    else dst = dL_dconic + 4 * (size_t)id + (k == 7 ? 3 : k - 5);
                    if (w1 != 0.f) atomicAdd(dst, w1);
                } else if ((lane & 15) == 1) {
                    if (o8 != 0.f) atomicAdd(dL_dopacity + id, o8);
                }
            }
            __syncwarp();  // the pairs' reads are over before the next batch appends
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
        e = cudaFuncSetAttribute(blend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(BbRing) * BB_WARPS));
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blend_backward_kernel, BB_WARPS * 32, sizeof(BbRing) * BB_WARPS);
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
    blend_backward_kernel<<<grid, BB_WARPS * 32, sizeof(BbRing) * BB_WARPS, f.stream>>>(im.ranges, im.order, b.list, g.rec, f.s.width, f.s.height,
                                                               f.gx, g.hdr, queue, f.s.background, im.final_T,
                                                               im.n_contrib, dL_dpix, f.s.downsample == 2 ? 1 : 0, dL_dmean2D,
                                                               dL_dconic, dL_dopacity,
                                                               dL_dcolor);
    gs_note_launch();
    return cudaGetLastError();
}
