// binning.cu -- depth ordering + per-tile instance lists (replaces K2-K5 of the reference:
// cub::DeviceScan::InclusiveSum, duplicateWithKeys, cub::DeviceRadixSort::SortPairs on 64-bit (tile|depth) keys and
// identifyTileRanges; dgr/cuda_rasterizer/rasterizer_impl.cu:70-138,277-318).
//
// The reference sorts R = sum(tiles touched) 64-bit keys + 32-bit values: ~152 B of HBM traffic per instance
// (SURVEY 8a row a10).  Here the same final order -- per tile, ascending depth, ties by ascending Gaussian index
//  -- is produced by
//   1. a stable LSD radix sort of the P Gaussians by their 32-bit depth key       (P  x 4 passes, 8 B/elem/pass)
//   2. tile pass 1: walk the Gaussians in depth order, expand each tile rectangle on the fly and scatter the
//      instances by the LOW 8 bits of the tile id, writing one packed u32 (tile_hi | gaussian) per instance
//   3. tile pass 2: stable scatter by the HIGH tile bits, writing the final u32 Gaussian-index list; the tile
//      ranges fall out of the pass-2 prefix table (no key comparison, no memset of ranges)
// i.e. ~16 B per instance instead of ~172 B, with no 64-bit keys ever materialised.
//
// Every radix pass is count -> scan -> scatter.  The unit of work is a WARP: a warp owns a contiguous slice of the
// input, ranks its elements with __match_any_sync + a private 256-entry shared-memory counter row (stable by
// construction: rounds are processed in order, lanes in order), and owns one column of the digit-major histogram
// table.  The scan over the table is a single-pass decoupled look-back scan with ticketed block ids.
#include "gs_common.cuh"

namespace {

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GS_FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(GS_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Stable rank of this lane's element among the elements of the same digit seen so far by this warp.
// wcount = this warp's 256 counters in shared memory.  All 32 lanes call it.
__device__ __forceinline__ uint32_t warp_rank(uint32_t digit, bool valid, uint32_t* wcount, int lane) {
    const unsigned act = __ballot_sync(GS_FULL, valid);
    uint32_t off = 0;
    if (valid) {
        const unsigned m = __match_any_sync(act, digit);
        const unsigned rank = __popc(m & ((1u << lane) - 1u));
        const uint32_t base = wcount[digit];
        __syncwarp(act);
        if (rank == 0) wcount[digit] = base + __popc(m);
        off = base + rank;
    }
    __syncwarp();
    return off;
}

// ---------------------------------------------------------------------------------------------------
// Single-pass exclusive scan (in place) of data[0..n), total written to data[n].
// state[] must not contain this pass's flags on entry (zeroed once per frame; flags are epoch-tagged so several
// scans can reuse one state array within a frame).
__global__ void __launch_bounds__(GS_SCAN_THREADS) scan_kernel(uint32_t* __restrict__ data, uint32_t n,
                                                               unsigned long long* __restrict__ state,
                                                               unsigned int* __restrict__ ticket, uint32_t epoch) {
    __shared__ uint32_t s_bid, s_warp[GS_SCAN_THREADS / 32], s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const unsigned long long FLAG_AGG = (unsigned long long)(2 * epoch + 1) << 32;
    const unsigned long long FLAG_INC = (unsigned long long)(2 * epoch + 2) << 32;

    const size_t base = (size_t)bid * GS_SCAN_TILE + (size_t)tid * GS_SCAN_ITEMS;
    uint32_t v[GS_SCAN_ITEMS];
    if (base + GS_SCAN_ITEMS <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(data + base);
        const uint4 b = *reinterpret_cast<const uint4*>(data + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < GS_SCAN_ITEMS; k++) v[k] = (base + k < n) ? data[base + k] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < GS_SCAN_ITEMS; k++) tsum += v[k];
    const uint32_t incl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wpre = 0, btotal = 0;
#pragma unroll
    for (int k = 0; k < GS_SCAN_THREADS / 32; k++) {
        const uint32_t w = s_warp[k];
        if (k < warp) wpre += w;
        btotal += w;
    }
    if (warp == 0) {
        if (lane == 0) st_volatile_u64(state + bid, (bid == 0 ? FLAG_INC : FLAG_AGG) | btotal);
        uint32_t excl = 0;
        if (bid > 0) {
            int look = (int)bid - 1;
            while (true) {
                const int j = look - lane;
                unsigned long long st = FLAG_INC;  // positions before block 0 act as "inclusive 0"
                if (j >= 0) {
                    do { st = ld_volatile_u64(state + j); } while ((st & ~0xffffffffull) != FLAG_AGG &&
                                                                 (st & ~0xffffffffull) != FLAG_INC);
                }
                const bool inc = (st & ~0xffffffffull) == FLAG_INC;
                const uint32_t val = (uint32_t)st;
                const unsigned im = __ballot_sync(GS_FULL, inc);
                if (im) {
                    const int first = __ffs(im) - 1;
                    excl += warp_sum(lane <= first ? val : 0u);
                    break;
                }
                excl += warp_sum(val);
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(state + bid, FLAG_INC | (uint32_t)(excl + btotal));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    uint32_t run = s_prefix + wpre + (incl - tsum);
    uint32_t o[GS_SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < GS_SCAN_ITEMS; k++) { o[k] = run; run += v[k]; }
    if (base + GS_SCAN_ITEMS <= n) {
        *reinterpret_cast<uint4*>(data + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(data + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < GS_SCAN_ITEMS; k++)
            if (base + k < n) data[base + k] = o[k];
    }
    // the block that owns the last element also publishes the grand total at data[n]
    if (base <= (size_t)n - 1 && (size_t)n - 1 < base + GS_SCAN_ITEMS) data[n] = run;
}

// ---------------------------------------------------------------------------------------------------
// Depth sort: stable LSD radix sort of (key = depth bits, value = Gaussian index), 8 bits per pass.
#define DS_WARPS 4
#define DS_ROUNDS (GS_DEPTH_UNIT / 32)

template <bool SCATTER>
__global__ void __launch_bounds__(DS_WARPS * 32) depth_pass_kernel(const uint32_t* __restrict__ key_in,
                                                                  const uint32_t* __restrict__ idx_in,
                                                                  uint32_t* __restrict__ key_out,
                                                                  uint32_t* __restrict__ idx_out,
                                                                  uint32_t* __restrict__ hist, uint32_t P,
                                                                  uint32_t units, int shift) {
    __shared__ uint32_t s_cnt[DS_WARPS][GS_RADIX];
    __shared__ uint32_t s_base[SCATTER ? DS_WARPS : 1][GS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * DS_WARPS + warp;
    if (unit >= units) return;  // whole warps only; no block-level barrier below
    uint32_t* cnt = s_cnt[warp];
#pragma unroll
    for (int k = 0; k < GS_RADIX / 32; k++) {
        cnt[k * 32 + lane] = 0;
        if (SCATTER) s_base[warp][k * 32 + lane] = hist[(size_t)(k * 32 + lane) * units + unit];
    }
    __syncwarp();
    const uint32_t base = unit * GS_DEPTH_UNIT;
#pragma unroll 4
    for (int r = 0; r < DS_ROUNDS; r++) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < P;
        const uint32_t k = valid ? key_in[i] : 0u;
        const uint32_t d = (k >> shift) & (GS_RADIX - 1);
        const uint32_t off = warp_rank(d, valid, cnt, lane);
        if (SCATTER && valid) {
            const uint32_t pos = s_base[warp][d] + off;
            key_out[pos] = k;
            idx_out[pos] = idx_in[i];
        }
    }
    if (!SCATTER) {
#pragma unroll
        for (int k = 0; k < GS_RADIX / 32; k++) hist[(size_t)(k * 32 + lane) * units + unit] = cnt[k * 32 + lane];
    }
}

// ---------------------------------------------------------------------------------------------------
// Tile pass 1: a warp owns GS_EMIT_UNIT consecutive Gaussians of the depth order, enumerates their tile
// rectangles row-major (same emission order as rasterizer_impl.cu:98-108) and ranks every instance by the low
// 8 bits of its tile id.
#define EM_WARPS 4
#define EM_PER_LANE (GS_EMIT_UNIT / 32)

template <bool SCATTER>
__global__ void __launch_bounds__(EM_WARPS * 32) tile_pass1_kernel(const uint32_t* __restrict__ sorted_idx,
                                                                  const ushort4* __restrict__ rect,
                                                                  const uint32_t* __restrict__ ntile,
                                                                  uint32_t* __restrict__ hist,
                                                                  uint32_t* __restrict__ stage,
                                                                  GsHeader* __restrict__ hdr, uint32_t P,
                                                                  uint32_t units, int gx, int idx_bits,
                                                                  unsigned long long Rcap) {
    __shared__ uint32_t s_cnt[EM_WARPS][GS_RADIX];
    __shared__ uint32_t s_base[SCATTER ? EM_WARPS : 1][GS_RADIX];
    __shared__ uint32_t s_off[EM_WARPS][GS_EMIT_UNIT + 1];
    __shared__ uint32_t s_gi[EM_WARPS][GS_EMIT_UNIT];
    __shared__ ushort4 s_rect[EM_WARPS][GS_EMIT_UNIT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * EM_WARPS + warp;
    if (unit >= units) return;
    const unsigned long long R = hdr->num_rendered;
    if (R > Rcap) {  // no-sync mode: the preallocated instance buffers are too small -> frame is skipped
        if (unit == 0 && lane == 0) hdr->code = GS_ERR_CAPACITY;
        if (SCATTER) return;
    }
    uint32_t* cnt = s_cnt[warp];
#pragma unroll
    for (int k = 0; k < GS_RADIX / 32; k++) {
        cnt[k * 32 + lane] = 0;
        if (SCATTER) s_base[warp][k * 32 + lane] = hist[(size_t)(k * 32 + lane) * units + unit];
    }
    // stage this warp's Gaussians: index, rectangle, exclusive offsets of their tile counts
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < EM_PER_LANE; k++) {
        const uint32_t i = unit * GS_EMIT_UNIT + k * 32 + lane;
        uint32_t gi = 0, nt = 0;
        ushort4 rc = make_ushort4(0, 0, 0, 0);
        if (i < P) {
            gi = sorted_idx[i];
            nt = ntile[gi];
            rc = rect[gi];
        }
        const uint32_t inc = warp_incl_scan(nt, lane);
        s_off[warp][k * 32 + lane] = carry + inc - nt;
        s_gi[warp][k * 32 + lane] = gi;
        s_rect[warp][k * 32 + lane] = rc;
        carry += __shfl_sync(GS_FULL, inc, 31);
    }
    if (lane == 0) s_off[warp][GS_EMIT_UNIT] = carry;
    __syncwarp();
    const uint32_t n = (R > Rcap) ? 0u : carry;
    const uint32_t* off = s_off[warp];
    for (uint32_t e0 = 0; e0 < n; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool valid = e < n;
        uint32_t d = 0, packed = 0;
        if (valid) {
            // largest g with off[g] <= e  (zero-tile Gaussians are skipped automatically)
            int g = 0;
#pragma unroll
            for (int step = GS_EMIT_UNIT / 2; step > 0; step >>= 1)
                if (off[g + step] <= e) g += step;
            const ushort4 rc = s_rect[warp][g];
            const uint32_t local = e - off[g];
            const uint32_t w = (uint32_t)rc.z - (uint32_t)rc.x;
            const uint32_t row = local / w;
            const uint32_t tile = ((uint32_t)rc.y + row) * (uint32_t)gx + (uint32_t)rc.x + (local - row * w);
            d = tile & (GS_RADIX - 1);
            packed = ((tile >> GS_RADIX_BITS) << idx_bits) | s_gi[warp][g];
        }
        const uint32_t o = warp_rank(d, valid, cnt, lane);
        if (SCATTER && valid) stage[s_base[warp][d] + o] = packed;
    }
    if (!SCATTER) {
#pragma unroll
        for (int k = 0; k < GS_RADIX / 32; k++) hist[(size_t)(k * 32 + lane) * units + unit] = cnt[k * 32 + lane];
    }
}

// ---------------------------------------------------------------------------------------------------
// Tile pass 2: warps own <= GS_TILE2_UNIT consecutive instances that all lie inside ONE low-digit bucket, so the
// pass-2 prefix table, indexed [high digit][unit], is ordered exactly like the final list:
// (high digit, low digit, position) = (tile id, depth order).
#define T2_WARPS 4
#define T2_ROUNDS (GS_TILE2_UNIT / 32)

template <bool SCATTER>
__global__ void __launch_bounds__(T2_WARPS * 32) tile_pass2_kernel(const uint32_t* __restrict__ stage,
                                                                  const uint32_t* __restrict__ hist1,
                                                                  uint32_t units1, uint32_t* __restrict__ hist2,
                                                                  uint32_t units2, uint32_t* __restrict__ list,
                                                                  uint32_t* __restrict__ bucket_unit0,
                                                                  const GsHeader* __restrict__ hdr, int idx_bits,
                                                                  unsigned long long Rcap) {
    __shared__ uint32_t s_cnt[T2_WARPS][GS_RADIX];
    __shared__ uint32_t s_base[SCATTER ? T2_WARPS : 1][GS_RADIX];
    __shared__ uint32_t s_bstart[GS_RADIX + 1];
    __shared__ uint32_t s_unit0[GS_RADIX + 1];
    __shared__ uint32_t s_wsum[T2_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool over = hdr->num_rendered > Rcap;
    // bucket boundaries = scanned pass-1 table at the first unit of each digit; total at [256*units1]
    for (int d = tid; d <= GS_RADIX; d += T2_WARPS * 32) s_bstart[d] = over ? 0u : hist1[(size_t)d * units1];
    __syncthreads();
    // exclusive scan of the per-bucket unit counts (2 buckets per thread)
    uint32_t nb0 = 0, nb1 = 0;
    {
        const int d0 = 2 * tid, d1 = 2 * tid + 1;
        nb0 = (uint32_t)gs_div_up(s_bstart[d0 + 1] - s_bstart[d0], GS_TILE2_UNIT);
        nb1 = (uint32_t)gs_div_up(s_bstart[d1 + 1] - s_bstart[d1], GS_TILE2_UNIT);
    }
    const uint32_t pair = nb0 + nb1;
    const uint32_t inc = warp_incl_scan(pair, lane);
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t wpre = 0;
    for (int k = 0; k < warp; k++) wpre += s_wsum[k];
    const uint32_t ex = wpre + inc - pair;
    s_unit0[2 * tid] = ex;
    s_unit0[2 * tid + 1] = ex + nb0;
    if (tid == T2_WARPS * 32 - 1) s_unit0[GS_RADIX] = ex + pair;
    __syncthreads();
    if (!SCATTER && blockIdx.x == 0)
        for (int d = tid; d <= GS_RADIX; d += T2_WARPS * 32) bucket_unit0[d] = s_unit0[d];

    const uint32_t unit = blockIdx.x * T2_WARPS + warp;
    if (unit >= units2) return;
    uint32_t* cnt = s_cnt[warp];
#pragma unroll
    for (int k = 0; k < GS_RADIX / 32; k++) {
        cnt[k * 32 + lane] = 0;
        if (SCATTER) s_base[warp][k * 32 + lane] = hist2[(size_t)(k * 32 + lane) * units2 + unit];
    }
    __syncwarp();
    uint32_t begin = 0, end = 0;
    if (unit < s_unit0[GS_RADIX]) {
        int b = 0;  // largest bucket with unit0[b] <= unit and at least one unit
#pragma unroll
        for (int step = GS_RADIX / 2; step > 0; step >>= 1)
            if (s_unit0[b + step] <= unit) b += step;
        begin = s_bstart[b] + (unit - s_unit0[b]) * GS_TILE2_UNIT;
        end = min(begin + (uint32_t)GS_TILE2_UNIT, s_bstart[b + 1]);
    }
    const uint32_t mask = (idx_bits >= 32) ? 0xffffffffu : ((1u << idx_bits) - 1u);
    for (uint32_t i0 = begin; i0 < end; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < end;
        const uint32_t e = valid ? stage[i] : 0u;
        const uint32_t d = (idx_bits >= 32) ? 0u : (e >> idx_bits);
        const uint32_t o = warp_rank(d, valid, cnt, lane);
        if (SCATTER && valid) list[s_base[warp][d] + o] = e & mask;
    }
    if (!SCATTER) {
#pragma unroll
        for (int k = 0; k < GS_RADIX / 32; k++) hist2[(size_t)(k * 32 + lane) * units2 + unit] = cnt[k * 32 + lane];
    }
}

cudaError_t launch_scan(uint32_t* data, size_t n, unsigned long long* state, unsigned int* ticket, uint32_t epoch,
                        cudaStream_t stream) {
    scan_kernel<<<(unsigned)gs_div_up(n, GS_SCAN_TILE), GS_SCAN_THREADS, 0, stream>>>(data, (uint32_t)n, state, ticket,
                                                                                     epoch);
    gs_note_launch();
    return cudaGetLastError();
}

}  // namespace

#define GS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t gs_launch_depth_sort(const GsFrame& f, const GsGeom& g, int* sorted_side) {
    const uint32_t P = (uint32_t)f.s.P;
    const uint32_t units = (uint32_t)g.depth_units;
    const unsigned blocks = (unsigned)gs_div_up(units, DS_WARPS);
    int side = 0;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * GS_RADIX_BITS;
        depth_pass_kernel<false><<<blocks, DS_WARPS * 32, 0, f.stream>>>(g.key[side], g.idx[side], nullptr, nullptr,
                                                                        g.dhist, P, units, shift);
        gs_note_launch();
        GS_TRY(cudaGetLastError());
        GS_TRY(launch_scan(g.dhist, (size_t)GS_RADIX * units, g.dstate, &g.hdr->tickets[pass], (uint32_t)pass,
                           f.stream));
        depth_pass_kernel<true><<<blocks, DS_WARPS * 32, 0, f.stream>>>(g.key[side], g.idx[side], g.key[side ^ 1],
                                                                       g.idx[side ^ 1], g.dhist, P, units, shift);
        gs_note_launch();
        GS_TRY(cudaGetLastError());
        side ^= 1;
    }
    *sorted_side = side;
    return cudaSuccess;
}

cudaError_t gs_launch_tile_binning(const GsFrame& f, const GsGeom& g, int sorted_side, const GsBinning& b,
                                   size_t Rcap, const GsImage& im) {
    const uint32_t P = (uint32_t)f.s.P;
    const uint32_t units1 = (uint32_t)b.emit_units, units2 = (uint32_t)b.units2;
    const unsigned blocks1 = (unsigned)gs_div_up(units1, EM_WARPS), blocks2 = (unsigned)gs_div_up(units2, T2_WARPS);
    const uint32_t* sidx = g.idx[sorted_side];
    tile_pass1_kernel<false><<<blocks1, EM_WARPS * 32, 0, f.stream>>>(sidx, g.rect, g.ntile, b.hist1, b.stage, g.hdr,
                                                                     P, units1, f.gx, f.idx_bits, Rcap);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    GS_TRY(launch_scan(b.hist1, (size_t)GS_RADIX * units1, b.state1, &g.hdr->tickets[4], 0u, f.stream));
    tile_pass1_kernel<true><<<blocks1, EM_WARPS * 32, 0, f.stream>>>(sidx, g.rect, g.ntile, b.hist1, b.stage, g.hdr, P,
                                                                    units1, f.gx, f.idx_bits, Rcap);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    tile_pass2_kernel<false><<<blocks2, T2_WARPS * 32, 0, f.stream>>>(b.stage, b.hist1, units1, b.hist2, units2, b.list,
                                                                     b.bucket_unit0, g.hdr, f.idx_bits, Rcap);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    GS_TRY(launch_scan(b.hist2, (size_t)GS_RADIX * units2, b.state2, &g.hdr->tickets[5], 0u, f.stream));
    tile_pass2_kernel<true><<<blocks2, T2_WARPS * 32, 0, f.stream>>>(b.stage, b.hist1, units1, b.hist2, units2, b.list,
                                                                    b.bucket_unit0, g.hdr, f.idx_bits, Rcap);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    return gs_launch_tile_order(f, g, b, Rcap, im);  // ranges + longest-first blend queue (blend_forward.cu)
}
