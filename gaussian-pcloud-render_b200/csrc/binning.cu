// binning.cu -- depth ordering + per-tile instance lists (replaces K2-K5 of the reference:
// cub::DeviceScan::InclusiveSum, duplicateWithKeys, cub::DeviceRadixSort::SortPairs on 64-bit (tile|depth) keys and
// identifyTileRanges; dgr/cuda_rasterizer/rasterizer_impl.cu:70-138,277-318).
//
// The reference sorts R = sum(tiles touched) 64-bit keys + 32-bit values: ~152 B of HBM traffic per instance
// (SURVEY 8a row a10).  Here the same final order -- per tile, ascending depth, ties by ascending Gaussian index
//  -- is produced without ever materialising a key per instance:
//   1. depth sort: stable LSD radix sort of the P Gaussians by their 32-bit depth key, 8 bits per pass, each pass a
//      single "onesweep" kernel (rank in shared memory, chunk chain across CTAs, reorder in shared memory);
//   2. plan: preprocess accumulated a 2-D difference array of the tile rectangles; one CTA integrates it into the
//      exact per-tile instance counts, their exclusive scan (= the tile ranges), the per-row item counts and the
//      longest-list-first tile queue of the blend kernel.  No instance is touched;
//   3. row pass: walk the Gaussians in depth order and stably partition one item per (Gaussian, tile row) by row;
//   4. column pass: walk each row's items (still in depth order) and stably partition one entry per
//      (item, tile column) by column, writing the final u32 Gaussian index straight into the tile's list.
// Passes 3 and 4 exploit that a Gaussian covers a contiguous RANGE of rows / columns and contributes at most one
// element per bin: a warp builds, for 32 items at once, the bitmask of covering items for every bin (two matches,
// group leaders store range starts / ends, prefix-xor over the bins); an element's stable rank is a popcount.
// Each pass is count -> scan -> scatter: a counting kernel writes per-chunk bin counts, a one-warp-per-row scan (rows)
// / extra blocks of the plan kernel (columns) turn them into output positions.
// Measured on B200 and rejected: hardware MATCH.ANY in the sort (no faster than 8 ballots), shared-memory atomicXor
// for the masks (ATOMS serialises), in-kernel decoupled look-back for the row / column passes (all chunks resident at
// once -> O(chunks^2) reads of hot rows, 8-18 us per chunk), 4096-key sort chunks (longer chains), 1024-item column
// chunks with the output staged in shared memory and written as coalesced runs (68 vs 56 us: the extra barriers and
// per-chunk overheads cost more than the scattered 4-byte stores).
// HBM traffic: ~8 B per row item written + read, 4 B per instance written: ~8 B per instance instead of ~172 B.
#include "gs_common.cuh"

namespace {

#ifdef GS_TIMELINE  // developer build: per-CTA phase timestamps (globaltimer, ns) of the binning kernels
__device__ unsigned long long* g_bin_timeline = nullptr;
__device__ __forceinline__ unsigned long long bin_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define BIN_MARK(slot, k) do { if (g_bin_timeline && threadIdx.x == 0) g_bin_timeline[(size_t)(slot) * 8 + (k)] = bin_gtime(); } while (0)
#else
#define BIN_MARK(slot, k) do {} while (0)
#endif

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    return *reinterpret_cast<const volatile unsigned*>(p);
}
__device__ __forceinline__ void st_volatile_u32(unsigned* p, unsigned v) {
    *reinterpret_cast<volatile unsigned*>(p) = v;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(GS_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Lanes of `act` holding the same BITS-bit value as this lane.  MATCH.ANY serialises over the distinct values in
// the warp (measured on B200: ~10x slower when all 32 lanes differ, which is the common case for radix digits and
// tile rows), BITS ballots cost the same for every input.
template <int BITS>
__device__ __forceinline__ unsigned match_bits(unsigned act, unsigned v) {
    unsigned m = act;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (v >> b) & 1u;
        const unsigned bal = __ballot_sync(act, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

// Chained prefix over the chunks of one pass (decoupled look-back, restated for the regime this pipeline runs in:
// a few hundred chunks that are almost all resident at the same time, so a chunk usually finds NO predecessor with a
// finished inclusive prefix and has to add up everything before it).  Per chunk a state word (0 / 1 = aggregates
// published / 2 = inclusive prefixes published) guards two rows of nb counters.  A chunk publishes its aggregates,
// finds the nearest predecessor whose inclusive prefix is ready (one coalesced read of up to NT state words), and
// sums that prefix plus the aggregates in between with ALL its threads: thread = (4 bins, slice of the
// predecessors), 16-byte loads, independent iterations -- a handful of L2 round trips instead of one per
// predecessor.  Called by every thread of the CTA (blockDim.x == NT); s_tot[nb] in, s_excl[nb] out (shared).
template <int NT>
__device__ __forceinline__ void chain_prefix(const GsChain ch, int chunk, int first, int nb,
                                             const uint32_t* s_tot, uint32_t* s_excl, uint32_t* s_part /*[4 NT]*/,
                                             int* s_pstar) {
    const int tid = threadIdx.x, lane = tid & 31;
    const bool head = chunk == first;
    const size_t row = (size_t)chunk * GS_MAX_GRID;
    for (int b = tid; b < nb; b += NT) {
        __stcg(ch.agg + row + b, s_tot[b]);
        if (head) __stcg(ch.inc + row + b, s_tot[b]);
    }
    if (tid == 0) *s_pstar = first - 1;
    __syncthreads();
    if (tid == 0) {  // one fence after the barrier orders every thread's counters before the state word
        __threadfence();
        st_volatile_u32(ch.state + chunk, head ? 2u : 1u);
    }
    if (head) {
        for (int b = tid; b < nb; b += NT) s_excl[b] = 0;
        __syncthreads();
        return;
    }
    // nearest predecessor with an inclusive prefix; every predecessor after it is waited for (aggregate published)
    for (int hi = chunk - 1; hi >= first; hi -= NT) {
        const int pos = hi - tid;
        unsigned st = 0;
        if (pos >= first) {
            do { st = ld_volatile_u32(ch.state + pos); } while (st == 0u);
        }
        const unsigned inc_lanes = __ballot_sync(GS_FULL, st == 2u);
        if (inc_lanes != 0u && lane == __ffs(inc_lanes) - 1) atomicMax(s_pstar, pos);  // lanes descend in position
        __syncthreads();
        const int found = *s_pstar;
        __syncthreads();  // nobody may update s_pstar (next window) before everybody has read it
        if (found >= first) break;
    }
    __threadfence();
    const int pstar = *s_pstar;
    const int nq = (nb + 3) >> 2, S = NT / nq;
    const int q = tid % nq, sl = tid / nq;
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    if (sl < S) {
        if (sl == 0 && pstar >= first) acc = __ldcg(reinterpret_cast<const uint4*>(ch.inc + (size_t)pstar * GS_MAX_GRID) + q);
#pragma unroll 4
        for (int p = chunk - 1 - sl; p > pstar; p -= S) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(ch.agg + (size_t)p * GS_MAX_GRID) + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<uint4*>(s_part)[sl * nq + q] = acc;
    }
    __syncthreads();
    for (int b = tid; b < nb; b += NT) {
        uint32_t e = 0;
        for (int k = 0; k < S; k++) e += s_part[(k * nq + (b >> 2)) * 4 + (b & 3)];
        s_excl[b] = e;
        __stcg(ch.inc + row + b, e + s_tot[b]);
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_volatile_u32(ch.state + chunk, 2u);
    }
}

// Row tables, recomputed by warp 0 of every CTA that needs them (<= 257 entries): prefix sum of the row difference
// array = items per tile row; s_rs = exclusive scan of the items (row_start), s_cf = exclusive scan of the number of
// column-pass chunks per row (chunks never straddle rows).  Entries 0..gy are written.
__device__ __forceinline__ void row_tables(const int* __restrict__ rdiff, int gy, uint32_t* s_rs, uint32_t* s_cf,
                                           int lane) {
    uint32_t carry_items = 0, carry_chunks = 0;
    int carry_diff = 0;
    for (int y0 = 0; y0 <= gy; y0 += 32) {
        const int y = y0 + lane;
        int d = (y <= gy) ? rdiff[y] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(GS_FULL, d, o);
            if (lane >= o) d += t;
        }
        d += carry_diff;  // items of row y
        const uint32_t n = (y < gy) ? (uint32_t)d : 0u;
        const uint32_t nch = (uint32_t)gs_div_up(n, GS_PART_CHUNK);
        const uint32_t in = warp_incl_scan(n, lane), ic = warp_incl_scan(nch, lane);
        if (y <= gy) {
            s_rs[y] = carry_items + in - n;
            s_cf[y] = carry_chunks + ic - nch;
        }
        carry_items += __shfl_sync(GS_FULL, in, 31);
        carry_chunks += __shfl_sync(GS_FULL, ic, 31);
        carry_diff = __shfl_sync(GS_FULL, d, 31);
    }
}
// largest row with s_cf[row] <= chunk (rows without items share a value with their successor; the last one owns it)
__device__ __forceinline__ int chunk_row(const uint32_t* s_cf, int gy, uint32_t chunk) {
    int lo = 0, hi = gy;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_cf[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------
// Depth sort.  Histograms of all four digits in one pass over the keys.
// `pn` (shard cull): the number of keys is hdr->num_cand, known only on the device; nullptr = P.
__global__ void __launch_bounds__(1024) depth_hist_kernel(const uint32_t* __restrict__ key, uint32_t P,
                                                          const unsigned* __restrict__ pn, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s[4][GS_RADIX];
    if (pn) P = *pn;
    const int tid = threadIdx.x, lane = tid & 31;
    (&s[0][0])[tid] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * 1024u;
    const uint32_t iters = (uint32_t)gs_div_up(P, stride);
    for (uint32_t it0 = 0; it0 < iters; it0 += 4) {  // whole warps iterate together; 4 independent loads in flight
        uint32_t k[4];
        bool valid[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t i = (it0 + u) * stride + blockIdx.x * 1024u + tid;
            valid[u] = (it0 + u) < iters && i < P;
            k[u] = valid[u] ? key[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned act = __ballot_sync(GS_FULL, valid[u]);
            if (valid[u]) {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint32_t d = (k[u] >> (8 * p)) & 255u;
                    if (p < 3) {
                        atomicAdd(&s[p][d], 1u);  // mantissa digits: spread out, few conflicts
                    } else {  // top byte (sign + exponent): a handful of distinct values, aggregate within the warp
                        const unsigned m = __match_any_sync(act, d);
                        if (lane == __ffs(m) - 1) atomicAdd(&s[p][d], (uint32_t)__popc(m));
                    }
                }
            }
        }
    }
    __syncthreads();
    const uint32_t v = (&s[0][0])[tid];
    if (v) atomicAdd(&hist[tid], v);
}

// One LSD pass, onesweep style: a CTA owns GS_SORT_CHUNK consecutive keys (warp w the w-th 512, round r of a warp
// the r-th 32), ranks them by digit with a ballot-based match + per-warp counters (stable: rounds in order, lanes in
// order), obtains the number of equal-digit keys in all earlier chunks from the chunk chain, reorders its keys by
// digit in shared memory and writes them out as contiguous runs (a direct scatter of 4-byte elements is limited to
// about one 32-byte sector per clock and SM, measured 18 us per pass; runs are coalesced).
#ifndef SORT_MATCH_HW
#define SORT_MATCH_HW 0
#endif
#if SORT_MATCH_HW
#define SORT_MATCH(act, v) __match_any_sync(act, v)
#else
#define SORT_MATCH(act, v) match_bits<8>(act, v)
#endif
#define SORT_THREADS 512
#define SORT_WARPS (SORT_THREADS / 32)
#define SORT_ROUNDS (GS_SORT_CHUNK / SORT_THREADS)
#define SORT_SMEM ((2 * GS_SORT_CHUNK + SORT_WARPS * GS_RADIX + 4 * GS_RADIX + 4 * SORT_THREADS) * 4 + 64)
#ifndef SORT_MIN_BLOCKS
#define SORT_MIN_BLOCKS 2  // <= 64 registers: leaves room for blend CTAs of other frames next to a sort CTA (+6 % frames/s)
#endif
__global__ void __launch_bounds__(SORT_THREADS, SORT_MIN_BLOCKS) depth_pass_kernel(uint32_t* __restrict__ key0,
                                                                  uint32_t* __restrict__ idx0,
                                                                  uint32_t* __restrict__ key1,
                                                                  uint32_t* __restrict__ idx1,
                                                                  const uint32_t* __restrict__ hist,  // 256 totals
                                                                  const GsChain chain, unsigned* __restrict__ ticket,
                                                                  GsHeader* __restrict__ hdr, uint32_t P, int pass,
                                                                  const uint32_t* __restrict__ cand) {
    if (cand) P = hdr->num_cand;  // shard cull: keys of the candidates only; position t holds Gaussian cand[t]
    // input side of the ping-pong = where the previous pass left its output (preprocess writes the keys to side 0)
    const int shift = pass * GS_RADIX_BITS, first_pass = pass == 0;
    const unsigned side = first_pass ? 0u : hdr->sort_side[pass - 1];
    const uint32_t* __restrict__ key_in = side ? key1 : key0;
    const uint32_t* __restrict__ idx_in = side ? idx1 : idx0;
    uint32_t* __restrict__ key_out = side ? key0 : key1;
    uint32_t* __restrict__ idx_out = side ? idx0 : idx1;
    extern __shared__ __align__(16) uint32_t s_sort[];
    uint32_t* s_key = s_sort;                                // [GS_SORT_CHUNK] keys in digit order
    uint32_t* s_val = s_key + GS_SORT_CHUNK;                 // [GS_SORT_CHUNK]
    uint32_t* s_cnt = s_val + GS_SORT_CHUNK;                 // [SORT_WARPS][256]
    uint32_t* s_tot = s_cnt + SORT_WARPS * GS_RADIX;         // [256] keys of this chunk per digit
    uint32_t* s_excl = s_tot + GS_RADIX;                     // [256] keys of earlier chunks per digit
    uint32_t* s_lstart = s_excl + GS_RADIX;                  // [256] first local position of each digit
    uint32_t* s_gbase = s_lstart + GS_RADIX;                 // [256] global position of the chunk's first key of a digit
    uint32_t* s_part = s_gbase + GS_RADIX;                   // [4 * SORT_THREADS] chain scratch
    __shared__ uint32_t s_wsum[8], s_wsum2[8];
    __shared__ uint32_t s_chunk;
    __shared__ int s_pstar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * GS_RADIX; i += SORT_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const uint32_t chunk = s_chunk;
    if (chunk * GS_SORT_CHUNK >= P && !(chunk == 0)) return;  // past the keys (the grid is sized for the upper bound)
    const uint32_t base = chunk * GS_SORT_CHUNK + warp * (GS_SORT_CHUNK / SORT_WARPS);
    const uint32_t tl_slot = (uint32_t)(shift / 8) * 1024u + chunk;
    BIN_MARK(tl_slot, 0);
    // A digit shared by ALL keys makes the pass the identity permutation (the top byte of the depths of an object
    // that lies within one binade of camera distance, e.g. every frame of C1-C3): nothing moves, the output side is
    // the input side (only the first pass still has to write the index column).  The global histogram is complete
    // before this kernel starts, so every CTA takes the same branch.
    if (hist[(key_in[0] >> shift) & 255u] == P) {
        if (first_pass) {
            const uint32_t beg = chunk * GS_SORT_CHUNK, n = beg < P ? min((uint32_t)GS_SORT_CHUNK, P - beg) : 0u;
            for (uint32_t t = tid; t < n; t += SORT_THREADS) {
                key_out[beg + t] = key_in[beg + t];
                idx_out[beg + t] = cand ? cand[beg + t] : beg + t;
            }
        }
        if (chunk == 0 && tid == 0) hdr->sort_side[pass] = first_pass ? 1u : side;
        return;
    }
    if (chunk == 0 && tid == 0) hdr->sort_side[pass] = side ^ 1u;

    uint32_t k[SORT_ROUNDS], v[SORT_ROUNDS];
    uint16_t rk[SORT_ROUNDS];
    uint32_t* cnt = s_cnt + warp * GS_RADIX;
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = base + r * 32 + lane;
        k[r] = (i < P) ? key_in[i] : 0xFFFFFFFFu;
        v[r] = i >= P ? i : (first_pass ? (cand ? cand[i] : i) : idx_in[i]);
    }
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < P;
        const uint32_t d = (k[r] >> shift) & 255u;
        const unsigned act = __ballot_sync(GS_FULL, valid);
        uint32_t off = 0;
        if (valid) {
            const unsigned m = SORT_MATCH(act, d);
            const unsigned before = __popc(m & ((1u << lane) - 1u));
            const uint32_t c0 = cnt[d];
            __syncwarp(act);
            if (before == 0) cnt[d] = c0 + __popc(m);
            off = c0 + before;
        }
        __syncwarp();
        rk[r] = (uint16_t)off;
    }
    __syncthreads();
    BIN_MARK(tl_slot, 1);
    // thread d < 256: exclusive scan of digit d over the warps, chunk total; exclusive scans over the digits of the
    // chunk totals (local digit starts) and of the global digit totals (global digit starts)
    if (tid < GS_RADIX) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const uint32_t t = s_cnt[w * GS_RADIX + tid];
            s_cnt[w * GS_RADIX + tid] = total;
            total += t;
        }
        s_tot[tid] = total;
        const uint32_t h = hist[tid];
        const uint32_t incl_h = warp_incl_scan(h, lane), incl_t = warp_incl_scan(total, lane);
        if (lane == 31) { s_wsum[warp] = incl_h; s_wsum2[warp] = incl_t; }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // only the first 8 warps take part
        uint32_t wpre_h = 0, wpre_t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++)
            if (w < warp) { wpre_h += s_wsum[w]; wpre_t += s_wsum2[w]; }
        s_gbase[tid] = wpre_h + incl_h - h;
        s_lstart[tid] = wpre_t + incl_t - total;
    }
    __syncthreads();
    BIN_MARK(tl_slot, 2);
    chain_prefix<SORT_THREADS>(chain, (int)chunk, 0, GS_RADIX, s_tot, s_excl, s_part, &s_pstar);
    BIN_MARK(tl_slot, 3);
    // reorder by digit in shared memory
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = base + r * 32 + lane;
        if (i < P) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t lp = s_lstart[d] + cnt[d] + rk[r];
            s_key[lp] = k[r];
            s_val[lp] = v[r];
        }
    }
    __syncthreads();
    BIN_MARK(tl_slot, 4);
    const uint32_t n = chunk * GS_SORT_CHUNK < P ? min((uint32_t)GS_SORT_CHUNK, P - chunk * GS_SORT_CHUNK) : 0u;
#pragma unroll 4
    for (uint32_t t = tid; t < n; t += SORT_THREADS) {
        const uint32_t kk = s_key[t];
        const uint32_t d = (kk >> shift) & 255u;
        const uint32_t pos = s_gbase[d] + s_excl[d] + (t - s_lstart[d]);
        key_out[pos] = kk;
        idx_out[pos] = s_val[t];
    }
    BIN_MARK(tl_slot, 5);
}

// ---------------------------------------------------------------------------------------------------
// Row counts: per chunk of GS_PART_CHUNK depth-sorted Gaussians, the number of row items it emits into each tile row
// (difference array in shared memory, +1 at y0, -1 at y1); row_scan_kernel then turns the table into the output
// positions of the row pass: exclusive prefix over the chunks, per row, on top of the row's first position.
// (Replaces a chained look-back inside the row pass: with a few hundred chunks that are all resident at once every
// chunk had to add up all its predecessors' counters -- O(chunks^2) reads of the same hot rows, 18 us per chunk.)
template <int NB>
__global__ void __launch_bounds__(256) row_count_kernel(const uint32_t* __restrict__ idx0,
                                                        const uint32_t* __restrict__ idx1,
                                                        const GsHeader* __restrict__ hdr,
                                                        const ushort4* __restrict__ rect, uint32_t P, int use_cand,
                                                        int gy, uint32_t* __restrict__ chunk_cnt /*[gy][nchunks]*/) {
    constexpr int G = NB / 32;
    const uint32_t* __restrict__ sorted_idx = hdr->sort_side[3] ? idx1 : idx0;  // output side of the depth sort
    if (use_cand) P = hdr->num_cand;
    __shared__ int s_d[NB + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk = blockIdx.x, nchunks = gridDim.x;
    for (int i = tid; i <= NB; i += 256) s_d[i] = 0;
    __syncthreads();
    const uint32_t ibeg = min(P, chunk * GS_PART_CHUNK), iend = min(P, ibeg + (uint32_t)GS_PART_CHUNK);
    for (uint32_t i = ibeg + tid; i < iend; i += 256) {
        const ushort4 rc = rect[sorted_idx[i]];
        if (rc.w > rc.y) {
            atomicAdd(&s_d[rc.y], 1);
            atomicAdd(&s_d[rc.w], -1);
        }
    }
    __syncthreads();
    if (warp == 0) {
        int v[G], sum = 0;
#pragma unroll
        for (int j = 0; j < G; j++) { sum += s_d[lane * G + j]; v[j] = sum; }
        const int pre = (int)warp_incl_scan((uint32_t)sum, lane) - sum;
#pragma unroll
        for (int j = 0; j < G; j++) {
            const int y = lane * G + j;
            if (y < gy) chunk_cnt[(size_t)y * nchunks + chunk] = (uint32_t)(v[j] + pre);
        }
    }
}

// One warp per tile row: exclusive prefix of the row's chunk counts on top of the row's first output position.
__global__ void __launch_bounds__(32) row_scan_kernel(const int* __restrict__ rdiff, int gy, uint32_t nchunks,
                                                      uint32_t* __restrict__ chunk_cnt /*[gy][nchunks]*/) {
    __shared__ uint32_t s_rs[GS_MAX_GRID + 1], s_cf[GS_MAX_GRID + 1];
    const int lane = threadIdx.x, y = blockIdx.x;
    row_tables(rdiff, gy, s_rs, s_cf, lane);
    __syncwarp();
    uint32_t run = s_rs[y];
    uint32_t* row = chunk_cnt + (size_t)y * nchunks;
    for (uint32_t c0 = 0; c0 < nchunks; c0 += 32) {
        const uint32_t c = c0 + lane;
        const uint32_t v = (c < nchunks) ? row[c] : 0u;
        const uint32_t incl = warp_incl_scan(v, lane);
        if (c < nchunks) row[c] = run + incl - v;
        run += __shfl_sync(GS_FULL, incl, 31);
    }
}

// ---------------------------------------------------------------------------------------------------
// Column histogram: instances per tile = number of row items of the tile's row that cover its column.  A CTA owns
// one column-pass chunk (items of ONE row), builds the column difference array in shared memory (+1 at x0, -1 at
// x1 per item), integrates it and adds the non-zero counts to tcount.
template <int NB>
__global__ void __launch_bounds__(256) column_hist_kernel(const uint2* __restrict__ items, const int* __restrict__ rdiff,
                                                          int gx, int gy, unsigned long long RowCap,
                                                          uint32_t* __restrict__ tcount,
                                                          uint32_t* __restrict__ chunk_cnt /*[chunks][GS_MAX_GRID]*/) {
    constexpr int G = NB / 32;
    __shared__ uint32_t s_rs[GS_MAX_GRID + 1], s_cf[GS_MAX_GRID + 1];
    __shared__ int s_d[NB + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) row_tables(rdiff, gy, s_rs, s_cf, lane);
    __syncthreads();
    if ((unsigned long long)s_rs[gy] > RowCap) return;
    const uint32_t nchunks = s_cf[gy];
    for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int row = chunk_row(s_cf, gy, chunk);
        const uint32_t ibeg = s_rs[row] + (chunk - s_cf[row]) * GS_PART_CHUNK;
        const uint32_t iend = min(s_rs[row + 1], ibeg + (uint32_t)GS_PART_CHUNK);
        for (int i = tid; i <= NB; i += 256) s_d[i] = 0;
        __syncthreads();
        for (uint32_t i = ibeg + tid; i < iend; i += 256) {
            const uint32_t xr = items[i].y;
            atomicAdd(&s_d[xr & 0xffffu], 1);
            atomicAdd(&s_d[xr >> 16], -1);
        }
        __syncthreads();
        if (warp == 0) {
            int v[G], sum = 0;
#pragma unroll
            for (int j = 0; j < G; j++) { sum += s_d[lane * G + j]; v[j] = sum; }
            const int pre = (int)warp_incl_scan((uint32_t)sum, lane) - sum;
#pragma unroll
            for (int j = 0; j < G; j++) {
                const int x = lane * G + j, c = v[j] + pre;
                if (x < gx) {
                    chunk_cnt[(size_t)chunk * GS_MAX_GRID + x] = (uint32_t)c;  // -> exclusive prefix by plan_kernel
                    if (c) atomicAdd(&tcount[row * gx + x], (uint32_t)c);
                }
            }
        }
        __syncthreads();
    }
}

// Plan (one CTA): exclusive scan of the tile counts in tile-id order = the tile ranges; longest-list-first order of
// the shard's tiles for the blend queue; instance-capacity check (no-sync mode).
__global__ void __launch_bounds__(1024) plan_kernel(const uint32_t* __restrict__ tcount, int gx, int gy, int row0,
                                                    int row1, uint32_t* __restrict__ tile_start,
                                                    uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                    GsHeader* __restrict__ hdr, unsigned long long Rcap,
                                                    const int* __restrict__ rdiff, unsigned long long RowCap,
                                                    uint32_t* __restrict__ chunk_cnt /*[chunks][GS_MAX_GRID]*/) {
    __shared__ uint32_t s_wsum[32];
    __shared__ uint32_t s_cnt[33], s_slot[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (blockIdx.x > 0) {
        // blocks 1..gy: exclusive prefix of the column counts over the chunks of tile row blockIdx.x - 1 (the column
        // pass adds the tile's first position); block 0 below does the plan proper
        __shared__ uint32_t s_rs[GS_MAX_GRID + 1], s_cf[GS_MAX_GRID + 1];
        if (warp == 0) row_tables(rdiff, gy, s_rs, s_cf, lane);
        __syncthreads();
        if ((unsigned long long)s_rs[gy] > RowCap) return;
        const int y = (int)blockIdx.x - 1;
        const uint32_t c0 = s_cf[y], c1 = s_cf[y + 1];
        if (tid < gx) {
            uint32_t run = 0;
            uint32_t c = c0;
            for (; c + 8 <= c1; c += 8) {  // eight independent loads in flight, then the dependent running sum
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = __ldcg(chunk_cnt + (size_t)(c + k) * GS_MAX_GRID + tid);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    chunk_cnt[(size_t)(c + k) * GS_MAX_GRID + tid] = run;
                    run += v[k];
                }
            }
            for (; c < c1; c++) {
                const uint32_t v = __ldcg(chunk_cnt + (size_t)c * GS_MAX_GRID + tid);
                chunk_cnt[(size_t)c * GS_MAX_GRID + tid] = run;
                run += v;
            }
        }
        return;
    }
    const int Tn = gx * gy;
    const int per = (Tn + 1023) / 1024;  // each thread owns a contiguous run of tiles
    const int t_begin = min(Tn, tid * per), t_end = min(Tn, t_begin + per);
    uint32_t local = 0;
#pragma unroll 8
    for (int t = t_begin; t < t_end; t++) local += tcount[t];
    const uint32_t incl = warp_incl_scan(local, lane);
    if (lane == 31) s_wsum[warp] = incl;
    if (tid < 33) s_cnt[tid] = 0;
    __syncthreads();
    uint32_t wpre = 0, total = 0;
    for (int w = 0; w < 32; w++) {
        const uint32_t v = s_wsum[w];
        if (w < warp) wpre += v;
        total += v;
    }
    const bool over = (unsigned long long)total > Rcap || hdr->skip != 0;
    if (tid == 0 && over) { hdr->skip = 1u; hdr->code = GS_ERR_CAPACITY; }
    uint32_t run = wpre + incl - local;
    const int s0 = row0 * gx, s1 = row1 * gx;
    for (int t = t_begin; t < t_end; t++) {
        const uint32_t c = tcount[t];
        tile_start[t] = run;
        if (t >= s0 && t < s1) {
            const uint32_t len = over ? 0u : c;
            ranges[t] = over ? make_uint2(0u, 0u) : make_uint2(run, run + c);
            const int cls = len ? 32 - __clz(len) : 0;  // 0 = empty, 1..32
            const unsigned m = match_bits<6>(__activemask(), (unsigned)cls);  // most tiles are empty: aggregate per warp
            if (lane == __ffs(m) - 1) atomicAdd(&s_cnt[32 - cls], (uint32_t)__popc(m));  // slot 0 = longest class
        }
        run += c;
    }
    if (tid == 1023) tile_start[Tn] = total;
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < 33; k++) { s_slot[k] = acc; acc += s_cnt[k]; }
        hdr->nonempty_tiles = acc - s_cnt[32];  // slot 32 = empty tiles, at the end of the order
    }
    __syncthreads();
    for (int t = t_begin; t < t_end; t++) {
        if (t >= s0 && t < s1) {
            const uint32_t len = over ? 0u : tcount[t];
            const int cls = len ? 32 - __clz(len) : 0;
            const unsigned m = match_bits<6>(__activemask(), (unsigned)cls);
            const int leader = __ffs(m) - 1;
            uint32_t slot = 0;
            if (lane == leader) slot = atomicAdd(&s_slot[32 - cls], (uint32_t)__popc(m));
            slot = __shfl_sync(m, slot, leader) + __popc(m & ((1u << lane) - 1u));
            order[slot] = (uint32_t)t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Range partition (row pass: PASS 1, column pass: PASS 2).
// Every input item covers the bin range [lo, hi) and emits one output element into each covered bin; elements of a
// bin keep the input order.  A CTA owns a chunk of GS_PART_CHUNK consecutive items (warp w the w-th 256, round r the
// r-th 32).  NB = number of bins rounded up to 128 or 256.
//
// Ranking without a key per element: for a round of 32 items, T[b] = (lanes whose range starts at b) xor (lanes whose
// range ends at b); the prefix-xor of T over the bins is, for every bin, the bitmask of lanes covering it, so the
// stable rank of lane i's element in bin b is popc(mask[b] & lanes_below_i) + the elements of earlier rounds /
// warps / chunks.  T is built with two ballot-based matches per round (group leaders store; no shared-memory atomics).
#define PART_ROUNDS (GS_PART_CHUNK / 256)
// Hardware MATCH.ANY vs 8-9 ballots: same latency here, but one issue slot instead of ~25; with several frames in
// flight the binning kernels compete with the blend kernels of other frames for issue slots (+1.5 % frames/s).
#ifndef PART_MATCH_HW
#define PART_MATCH_HW 1
#endif
#if PART_MATCH_HW
#define PART_MATCH(act, v) __match_any_sync(act, v)
#else
#define PART_MATCH(act, v) match_bits<(NB == 128) ? 8 : 9>(act, v)
#endif

#ifndef PART_MIN_BLOCKS
#define PART_MIN_BLOCKS 4
#endif
template <int NB, int PASS>
__global__ void __launch_bounds__(256, (NB == 128) ? PART_MIN_BLOCKS : 2) range_partition_kernel(
    const uint32_t* __restrict__ idx0, const uint32_t* __restrict__ idx1,  // PASS 1 input: the depth order (either side)
    const ushort4* __restrict__ rect, uint32_t P, int use_cand,
    const uint2* __restrict__ items_in,                                                      // PASS 2 input
    const int* __restrict__ rdiff, int gx, int gy, const uint32_t* __restrict__ tile_start,
    unsigned long long RowCap, const uint32_t* __restrict__ chunk_base /*[chunks][GS_MAX_GRID]*/,
    unsigned* __restrict__ ticket, GsHeader* __restrict__ hdr, uint2* __restrict__ items_out,
    uint32_t* __restrict__ list_out) {
    constexpr int G = NB / 32;    // bins per lane in the warp-wide scans
    constexpr int MS = NB + 4;    // mask row stride (words), multiple of 4 for the vectorised clear
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t* s_mask = s_dyn;                                         // [8 warps][PART_ROUNDS][MS]
    int* s_cnt = reinterpret_cast<int*>(s_dyn + 8 * PART_ROUNDS * MS);  // [8 warps][NB]: counts -> output positions
    __shared__ uint32_t s_rs[GS_MAX_GRID + 1], s_cf[GS_MAX_GRID + 1];
    __shared__ uint32_t s_chunk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (PASS == 2 && hdr->skip) return;
    const uint32_t* __restrict__ sorted_idx = (PASS == 1 && hdr->sort_side[3]) ? idx1 : idx0;
    const uint32_t Pn = (PASS == 1 && use_cand) ? hdr->num_cand : P;  // row pass input: the depth-sorted candidates
    if (warp == 0) row_tables(rdiff, gy, s_rs, s_cf, lane);
    __syncthreads();
    if ((unsigned long long)s_rs[gy] > RowCap) {  // no-sync mode: row-item buffer too small -> frame is skipped
        if (PASS == 1 && blockIdx.x == 0 && tid == 0) { hdr->skip = 1u; hdr->code = GS_ERR_CAPACITY; }
        return;
    }
    const uint32_t nchunks = (PASS == 1) ? (uint32_t)gs_div_up(P, GS_PART_CHUNK) : s_cf[gy];
    int* cnt = s_cnt + warp * NB;
    uint32_t* msk = s_mask + warp * (PART_ROUNDS * MS);
    const unsigned lt = (1u << lane) - 1u;

    while (true) {
        __syncthreads();  // previous chunk's shared state is dead
        if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
        {   // clear this warp's toggle rows
            uint4* z = reinterpret_cast<uint4*>(msk);
            for (int i = lane; i < PART_ROUNDS * MS / 4; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const uint32_t chunk = s_chunk;
        if (chunk >= nchunks) break;
        const uint32_t tl_slot = 4096u + (uint32_t)(PASS - 1) * 16384u + chunk;
        BIN_MARK(tl_slot, 0);

        // chunk -> item range (column pass: the tile row it belongs to)
        uint32_t ibeg, iend;
        int first = 0, row = 0;
        if (PASS == 1) {
            ibeg = min(Pn, chunk * GS_PART_CHUNK);
            iend = min(Pn, ibeg + (uint32_t)GS_PART_CHUNK);
        } else {
            row = chunk_row(s_cf, gy, chunk);
            first = (int)s_cf[row];
            ibeg = s_rs[row] + (chunk - (uint32_t)first) * GS_PART_CHUNK;
            iend = min(s_rs[row + 1], ibeg + (uint32_t)GS_PART_CHUNK);
        }

        // ---- phase A: load items; toggle rows; masks; per-warp counts
        uint32_t pay[PART_ROUNDS], rng[PART_ROUNDS];  // payload (gaussian), lo | hi << 16
        uint32_t xr[PASS == 1 ? PART_ROUNDS : 1];
#pragma unroll
        for (int r = 0; r < PART_ROUNDS; r++) {
            const uint32_t i = ibeg + warp * (GS_PART_CHUNK / 8) + r * 32 + lane;
            uint32_t lo = 0, hi = 0;
            pay[r] = 0;
            if (i < iend) {
                if (PASS == 1) {
                    const uint32_t gi = sorted_idx[i];
                    const ushort4 rc = rect[gi];
                    pay[r] = gi;
                    xr[r] = (uint32_t)rc.x | ((uint32_t)rc.z << 16);
                    lo = rc.y; hi = rc.w;
                } else {
                    const uint2 it = items_in[i];
                    pay[r] = it.x;
                    lo = it.y & 0xffffu; hi = it.y >> 16;
                }
                if (hi <= lo) { lo = 0; hi = 0; }
            }
            rng[r] = lo | (hi << 16);
        }
#pragma unroll
        for (int r = 0; r < PART_ROUNDS; r++) {  // range starts: one store per distinct start bin
            const uint32_t lo = rng[r] & 0xffffu, hi = rng[r] >> 16;
            const unsigned act = __ballot_sync(GS_FULL, hi > lo);
            if (hi > lo) {
                const unsigned m = PART_MATCH(act, lo);
                if ((m & lt) == 0) msk[r * MS + lo] = m;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < PART_ROUNDS; r++) {  // range ends: xor into the row (one writer per distinct end bin)
            const uint32_t lo = rng[r] & 0xffffu, hi = rng[r] >> 16;
            const unsigned act = __ballot_sync(GS_FULL, hi > lo);
            if (hi > lo) {
                const unsigned m = PART_MATCH(act, hi);
                if ((m & lt) == 0) msk[r * MS + hi] ^= m;
            }
        }
        __syncwarp();
        {
            int csum[G];
#pragma unroll
            for (int j = 0; j < G; j++) csum[j] = 0;
#pragma unroll
            for (int r = 0; r < PART_ROUNDS; r++) {  // prefix-xor over the bins -> covering masks
                uint32_t m[G], acc = 0;
                if (G == 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(&msk[r * MS + lane * 4]);
                    m[0] = q.x; m[1] = q.y; m[2] = q.z; m[3] = q.w;
                } else {
#pragma unroll
                    for (int j = 0; j < G; j++) m[j] = msk[r * MS + lane * G + j];
                }
#pragma unroll
                for (int j = 0; j < G; j++) { acc ^= m[j]; m[j] = acc; }
                uint32_t sc = acc;  // inclusive xor-scan of the lane totals
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(GS_FULL, sc, o);
                    if (lane >= o) sc ^= t;
                }
                const uint32_t pre = sc ^ acc;
#pragma unroll
                for (int j = 0; j < G; j++) {
                    m[j] ^= pre;
                    msk[r * MS + lane * G + j] = m[j];
                    csum[j] += __popc(m[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < G; j++) cnt[lane * G + j] = csum[j];
        }
        __syncthreads();
        BIN_MARK(tl_slot, 1);
        BIN_MARK(tl_slot, 2);
        // ---- phase B: thread b = bin b: exclusive scan over the warps on top of the chunk's first output position
        // (row pass: row_count_kernel's table; column pass: the tile's first position + plan_kernel's table)
        const int nb_used = (PASS == 1) ? gy : gx;
        if (tid < NB) {
            uint32_t run = 0;
            if (tid < nb_used) {
                if (PASS == 1) run = chunk_base[(size_t)tid * nchunks + chunk];  // [row][chunk] (row_scan_kernel)
                else run = chunk_base[(size_t)chunk * GS_MAX_GRID + tid] + tile_start[row * gx + tid];
            }
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t t = (uint32_t)s_cnt[w * NB + tid];
                s_cnt[w * NB + tid] = (int)run;
                run += t;
            }
        }
        __syncthreads();
        BIN_MARK(tl_slot, 3);
        // ---- phase C: scatter, round by round (positions of a bin advance by the round's population)
#pragma unroll
        for (int r = 0; r < PART_ROUNDS; r++) {
            const uint32_t lo = rng[r] & 0xffffu, hi = rng[r] >> 16;
            for (uint32_t b = lo; b < hi; b++) {
                const uint32_t pos = (uint32_t)cnt[b] + __popc(msk[r * MS + b] & lt);
                if (PASS == 1) items_out[pos] = make_uint2(pay[r], xr[PASS == 1 ? r : 0]);
                else list_out[pos] = pay[r];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < G; j++) cnt[lane * G + j] += __popc(msk[r * MS + lane * G + j]);
            __syncwarp();
        }
        __syncthreads();
        BIN_MARK(tl_slot, 4);
    }
}

}  // namespace

#ifdef GS_TIMELINE
extern "C" int gs_debug_bin_timeline(void* dev_buf) {
    unsigned long long* p = (unsigned long long*)dev_buf;
    return (int)cudaMemcpyToSymbol(g_bin_timeline, &p, sizeof(p));
}
#endif

#ifndef HIST_CTAS
#define HIST_CTAS 64
#endif
#define GS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t gs_launch_depth_sort(const GsFrame& f, const GsGeom& g) {
    const uint32_t P = (uint32_t)f.s.P;
    static GsPerDevice per_dev;
    const int* dv = nullptr;
    GS_TRY(per_dev.get(&dv, [](int, int*) {
        return cudaFuncSetAttribute(depth_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM);
    }));
    const unsigned chunks = (unsigned)g.sort_chunks;
    const uint32_t* cand = f.cull ? g.cand : nullptr;
    depth_hist_kernel<<<(unsigned)min((size_t)HIST_CTAS, gs_div_up(P, 4096)), 1024, 0, f.stream>>>(
        g.key[0], P, f.cull ? &g.hdr->num_cand : nullptr, g.dhist);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    for (int pass = 0; pass < 4; pass++) {
        const GsChain ch = {g.dstate + (size_t)pass * chunks, g.dagg + (size_t)pass * chunks * GS_RADIX,
                            g.dinc + (size_t)pass * chunks * GS_RADIX};
        depth_pass_kernel<<<chunks, SORT_THREADS, SORT_SMEM, f.stream>>>(g.key[0], g.idx[0], g.key[1], g.idx[1],
                                                                        g.dhist + pass * GS_RADIX, ch,
                                                                        &g.hdr->tickets[pass], g.hdr, P, pass, cand);
        gs_note_launch();
        GS_TRY(cudaGetLastError());
    }
    return cudaSuccess;  // the sorted order is in key[s] / idx[s], s = hdr->sort_side[3]
}

#define PART_SMEM(NB) ((8 * PART_ROUNDS * ((NB) + 4) + 8 * (NB)) * 4)
#ifndef PART_GRID_MULT
#define PART_GRID_MULT 4
#endif
static GsPerDevice g_part_dev;  // value[0] = grid size of the partition passes on this device

cudaError_t gs_launch_tile_lists(const GsFrame& f, const GsGeom& g, const GsBinning& b, size_t Rcap, size_t RowCap,
                                 const GsImage& im) {
    const uint32_t P = (uint32_t)f.s.P;
    const int* dv = nullptr;
    GS_TRY(g_part_dev.get(&dv, [](int dev, int* v) {
        int sms = 0;
        GS_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GS_TRY(cudaFuncSetAttribute(range_partition_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_SMEM(256)));
        GS_TRY(cudaFuncSetAttribute(range_partition_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_SMEM(256)));
        v[0] = sms * PART_GRID_MULT;
        return cudaSuccess;
    }));
    const int g_part_grid = dv[0];
    const unsigned grid1 = (unsigned)min((size_t)g_part_grid, g.row_chunks);
    const unsigned grid2 = (unsigned)min((size_t)g_part_grid, b.col_chunks);
    const unsigned long long rowcap = RowCap;
#define LAUNCH_PART(NB, PASS, GRID)                                                                               \
    range_partition_kernel<NB, PASS><<<GRID, 256, PART_SMEM(NB), f.stream>>>(                                     \
        g.idx[0], g.idx[1], g.rect, P, f.cull ? 1 : 0, b.items, im.rdiff, f.gx, f.gy, im.tile_start, rowcap,       \
        (PASS == 1) ? g.ragg : b.cagg,                                                                            \
        &g.hdr->tickets[3 + PASS], g.hdr, b.items, b.list)
    // row counts per chunk -> output positions; row pass: Gaussians in depth order -> row items grouped by tile row
    if (f.gy <= 128)
        row_count_kernel<128><<<(unsigned)g.row_chunks, 256, 0, f.stream>>>(g.idx[0], g.idx[1], g.hdr, g.rect, P, f.cull ? 1 : 0, f.gy, g.ragg);
    else
        row_count_kernel<256><<<(unsigned)g.row_chunks, 256, 0, f.stream>>>(g.idx[0], g.idx[1], g.hdr, g.rect, P, f.cull ? 1 : 0, f.gy, g.ragg);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    row_scan_kernel<<<(unsigned)f.gy, 32, 0, f.stream>>>(im.rdiff, f.gy, (uint32_t)g.row_chunks, g.ragg);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    if (f.gy <= 128) LAUNCH_PART(128, 1, grid1); else LAUNCH_PART(256, 1, grid1);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    // column histogram -> per-tile counts, then the plan (ranges, tile_start, blend queue)
    if (f.gx <= 128)
        column_hist_kernel<128><<<grid2, 256, 0, f.stream>>>(b.items, im.rdiff, f.gx, f.gy, rowcap, im.tcount, b.cagg);
    else
        column_hist_kernel<256><<<grid2, 256, 0, f.stream>>>(b.items, im.rdiff, f.gx, f.gy, rowcap, im.tcount, b.cagg);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    plan_kernel<<<1 + f.gy, 1024, 0, f.stream>>>(im.tcount, f.gx, f.gy, f.row0, f.row1, im.tile_start, im.ranges,
                                                im.order, g.hdr, (unsigned long long)Rcap, im.rdiff, rowcap, b.cagg);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
    // column pass: row items -> final per-tile lists
    if (f.gx <= 128) LAUNCH_PART(128, 2, grid2); else LAUNCH_PART(256, 2, grid2);
    gs_note_launch();
    GS_TRY(cudaGetLastError());
#undef LAUNCH_PART
    return cudaSuccess;
}
