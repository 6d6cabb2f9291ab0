// gs_common.cuh -- shared definitions of the B200-native splat rasterizer (sm_100a only).
//
// Data layout in HBM (all carved from the three caller-owned byte buffers, 256-B aligned sub-arrays):
//   geometry buffer (per Gaussian, P entries)
//     GsHeader            status block + counters (256 B)
//     dhist [4][256] u32  digit histograms of the depth keys           } zeroed at the start of every frame
//     dstate [4][C] u32   chunk states of the depth-sort chains          } (C = ceil(P / 8192) chunks)
//     dagg/dinc [4][C][256] u32   per-chunk aggregates / inclusive prefixes of the sort chains (not zeroed)
//     ragg [256][C1] u32  row-pass counts per chunk -> output positions (C1 = ceil(P / 2048) chunks)
//     rec   [P] GsRec     48-B packed record read by the blend kernels (3 x float4)
//     key   [2][P] u32    depth-sort keys (float bits of view-space z; 0xFFFFFFFF = culled), ping/pong
//     idx   [2][P] u32    depth-sort values (Gaussian index), ping/pong
//     rect  [P] ushort4   tile rectangle [x0,x1) x [y0,y1)
//     ntile [P] u32       tiles touched
//     clamp [P] u8        bit c set = SH colour channel c was clamped at 0
//     xrec  [P][3] float4 colours of up to three extra passes blended in the same list walk (GsScene.extra_colors)
//     cand  [P] u32       GsScene.shard_cull: ascending indices of the Gaussians that may reach the shard's tile rows;
//     cmask [P/32] u32, ccount [P/4096 + 1] u32: their bit mask and per-block counts (compaction scratch)
//   binning buffer (per instance, R entries)
//     list  [R] u32       final per-tile, depth-ordered Gaussian index list ("point_list"); FIRST, so that backward
//                         finds it from num_rendered alone
//     cagg [C2][256] u32  column-pass counts per chunk -> output positions
//     items [Rrow] uint2  row items after the row pass: (gaussian, x0 | x1 << 16), grouped by tile row, depth order
//   image buffer
//     rdiff [gy+1] i32    difference array of the row ranges -> row-item counts     } zeroed every
//     tcount [Tn] u32     instances per tile (column histogram of the row items)   } frame
//     final_T [N] f32, n_contrib [N] u32, ranges [Tn] uint2, order [Tn] u32 (longest-list-first tile queue),
//     tile_start [Tn+1] u32
//     park_units [GS_PARK_CAP] uint2 (unit, list position), park_state [GS_PARK_CAP][12][32] f32: pixel blocks whose
//     list walk the blend kernel handed over to the team kernel (blend_forward.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsplat_b200.h"

#define GS_TILE 16
#define GS_TILE_PIX 256
#define GS_ALIGN 256
#define GS_FULL 0xffffffffu

struct __align__(16) GsRec {
    float4 a;  // x, y, conic.x, conic.y
    float4 b;  // conic.z, opacity, alpha-cutoff threshold on `power`, -conic.y/conic.z
    float4 c;  // r, g, b, -conic.y/conic.x
};

struct GsHeader {  // lives at offset 0 of the geometry buffer
    unsigned long long num_rendered;  // sum of tiles touched (atomicAdd from preprocess)
    unsigned int num_visible;
    int code;  // GS_OK / GS_ERR_*
    unsigned int num_row_items;  // sum of tile rows touched (atomicAdd from preprocess)
    unsigned int nonempty_tiles;  // plan kernel: tiles of the shard with at least one instance
    unsigned int skip;           // 1 = instance / row-item capacity exceeded (no-sync mode): the frame is skipped
    unsigned int pad0;
    unsigned int tickets[16];    // 0-3 depth-sort passes, 4 row pass, 5 column pass, 6 blend work queue,
                                 // 7 parked blend units (count), 8 parked units taken by teams, 9 fresh CTAs done,
                                 // 10 blend-backward work queue, 11 CTA arrival order of the blend kernel
    unsigned int sort_side[4];   // depth sort: side (0 / 1) of the key / idx ping-pong that holds the output of pass p
                                 // (an identity pass -- all keys share the digit -- moves nothing and keeps the side)
    unsigned int num_cand;       // GsScene.shard_cull: Gaussians that may reach the shard's tile rows (= entries of cand[])
    unsigned int pad[35];
};
static_assert(sizeof(GsHeader) == 256, "header is one aligned slot");

// ---- buffer carving (host + device agree on this) -------------------------------------------------
struct GsCarver {
    size_t off = 0;
    char* base;
    __host__ __device__ explicit GsCarver(char* b) : base(b) {}
    template <typename T>
    __host__ __device__ T* take(size_t count) {
        off = (off + GS_ALIGN - 1) & ~(size_t)(GS_ALIGN - 1);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};

// chunk sizes of the binning kernels (one CTA of 256 threads per chunk)
#define GS_RADIX_BITS 8
#define GS_RADIX 256
#ifndef GS_SORT_CHUNK
#define GS_SORT_CHUNK 8192   // depth keys per CTA and pass
#endif
#define GS_PART_CHUNK 2048   // items per CTA of the row / column partition passes
#define GS_CULL_CHUNK 4096   // Gaussians per CTA of the shard-cull kernels
#define GS_MAX_GRID 256      // at most 256 x 256 tiles (4096 x 4096 pixels)
#define GS_PARK_CAP 2048     // pixel blocks per frame that can be handed to the blend team kernel
#define GS_PARK_WORDS 384    // saved state of one pixel block: 12 words x 32 lanes

static inline __host__ __device__ size_t gs_div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// One chain of chunks with decoupled look-back (binning.cu: chain_prefix): state[c] = 0 / 1 (aggregate published) /
// 2 (inclusive prefix published); agg / inc hold GS_MAX_GRID (= GS_RADIX) counters per chunk.
struct GsChain {
    unsigned* state;
    uint32_t* agg;
    uint32_t* inc;
};

struct GsGeom {
    GsHeader* hdr;
    uint32_t* dhist;   // [4][256]
    unsigned* dstate;  // [4][sort_chunks]
    uint32_t* dagg; uint32_t* dinc;  // [4][sort_chunks][256]
    uint32_t* ragg;    // [256][row_chunks] row-pass counts per chunk -> output positions (row_count / row_scan)
    GsRec* rec;
    uint32_t* key[2];
    uint32_t* idx[2];
    ushort4* rect;
    uint32_t* ntile;
    uint8_t* clamp;
    float4* xrec;      // [P][3] colours of up to three extra passes (rgb + pad), gathered like rec by the blend
    uint32_t* cand; uint32_t* cmask; uint32_t* ccount;  // shard cull (see above)
    size_t sort_chunks, row_chunks, cull_chunks, zero_bytes, bytes;
    __host__ __device__ GsGeom(char* base, size_t P) {
        GsCarver c(base);
        hdr = c.take<GsHeader>(1);
        sort_chunks = gs_div_up(P, GS_SORT_CHUNK);
        row_chunks = gs_div_up(P, GS_PART_CHUNK);
        dhist = c.take<uint32_t>(4 * GS_RADIX);
        dstate = c.take<unsigned>(4 * sort_chunks);
        zero_bytes = c.off;  // header + histograms + chunk states are zeroed at the start of every frame
        dagg = c.take<uint32_t>(4 * sort_chunks * GS_RADIX);
        dinc = c.take<uint32_t>(4 * sort_chunks * GS_RADIX);
        ragg = c.take<uint32_t>(row_chunks * GS_MAX_GRID);
        rec = c.take<GsRec>(P);
        key[0] = c.take<uint32_t>(P); key[1] = c.take<uint32_t>(P);
        idx[0] = c.take<uint32_t>(P); idx[1] = c.take<uint32_t>(P);
        rect = c.take<ushort4>(P);
        ntile = c.take<uint32_t>(P);
        clamp = c.take<uint8_t>(P);
        xrec = c.take<float4>(3 * P);
        cull_chunks = gs_div_up(P, GS_CULL_CHUNK);
        cand = c.take<uint32_t>(P);
        cmask = c.take<uint32_t>(cull_chunks * (GS_CULL_CHUNK / 32));
        ccount = c.take<uint32_t>(cull_chunks + 1);
        bytes = c.off + GS_ALIGN;
    }
};

struct GsBinning {
    uint32_t* list;    // [Rcap]
    uint32_t* cagg;    // [col_chunks][GS_MAX_GRID] column-pass counts per chunk -> output positions (plan_kernel)
    uint2* items;      // [RowCap]
    size_t col_chunks, zero_off, zero_bytes, bytes;
    // Rcap = instance capacity, RowCap = row-item capacity (<= Rcap always holds for the true counts)
    __host__ __device__ GsBinning(char* base, size_t Rcap, size_t RowCap) {
        GsCarver c(base);
        list = c.take<uint32_t>(Rcap);
        col_chunks = gs_div_up(RowCap, GS_PART_CHUNK) + GS_MAX_GRID;
        zero_off = c.off;
        zero_bytes = 0;  // nothing in this buffer needs clearing
        cagg = c.take<uint32_t>(col_chunks * GS_MAX_GRID);
        items = c.take<uint2>(RowCap);
        bytes = c.off + GS_ALIGN;
    }
};

struct GsImage {
    int* rdiff;        // [gy+1]
    uint32_t* tcount;  // [Tn]
    unsigned* park_ready;   // [GS_PARK_CAP] 1 = the state of the parked block is written (zeroed every frame)
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;
    uint32_t* order;  // tiles of the shard, longest instance list first (blend work queue)
    uint32_t* tile_start;   // [Tn+1]
    uint2* park_units;      // [GS_PARK_CAP]
    float* park_state;      // [GS_PARK_CAP][GS_PARK_WORDS]
    size_t zero_bytes, bytes;
    __host__ __device__ GsImage(char* base, size_t N, size_t gx, size_t gy) {
        GsCarver c(base);
        const size_t Tn = gx * gy;
        rdiff = c.take<int>(gy + 1);
        tcount = c.take<uint32_t>(Tn);
        park_ready = c.take<unsigned>(GS_PARK_CAP);
        zero_bytes = c.off;
        final_T = c.take<float>(N);
        n_contrib = c.take<uint32_t>(N);
        ranges = c.take<uint2>(Tn);
        order = c.take<uint32_t>(Tn);
        tile_start = c.take<uint32_t>(Tn + 1);
        park_units = c.take<uint2>(GS_PARK_CAP);
        park_state = c.take<float>((size_t)GS_PARK_CAP * GS_PARK_WORDS);
        bytes = c.off + GS_ALIGN;
    }
};

// ---- one-time, PER-DEVICE launcher set-up -----------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), occupancy and the SM count are properties of a (kernel, device)
// pair, so every launcher keeps its derived values per device ordinal; the first launch on a device runs `init`
// under a mutex, later launches read the cached values (one acquire load).  Safe for a process that renders on
// several GPUs and for several host threads.
#include <atomic>
#include <mutex>
#define GS_MAX_DEVICES 64
struct GsPerDevice {
    std::mutex mu;
    std::atomic<bool> ready[GS_MAX_DEVICES];
    int value[GS_MAX_DEVICES][4];
    GsPerDevice() { for (auto& r : ready) r.store(false); }
    // init(device, int value[4]) -> cudaError_t; on success *out points at this device's values
    template <typename F>
    cudaError_t get(const int** out, F init) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= GS_MAX_DEVICES) return cudaErrorInvalidDevice;
        if (!ready[dev].load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> lk(mu);
            if (!ready[dev].load(std::memory_order_relaxed)) {
                e = init(dev, value[dev]);
                if (e != cudaSuccess) return e;
                ready[dev].store(true, std::memory_order_release);
            }
        }
        *out = value[dev];
        return cudaSuccess;
    }
};

// launch bookkeeping (gs_launch_count) and error text
void gs_note_launch();
void gs_set_error(const char* what, cudaError_t e);

struct GsFrame {  // host-side derived quantities handed to every launcher
    GsScene s;
    int gx, gy, Tn;        // tile grid
    int row0, row1;        // tile-row shard
    bool cull;             // GsScene.shard_cull on a proper shard: per-Gaussian work runs on the compacted candidates
    float focal_x, focal_y;
    cudaStream_t stream;
};

// stage launchers (each in its own translation unit)
// f.cull: candidates of the tile-row shard -> g.cand / hdr->num_cand (before gs_launch_preprocess)
cudaError_t gs_launch_shard_cull(const GsFrame& f, const GsGeom& g);
cudaError_t gs_launch_preprocess(const GsFrame& f, const GsGeom& g, const GsImage& im, int32_t* radii);
cudaError_t gs_launch_recolor(const GsFrame& f, const GsGeom& g);  // colour-only pass (gs_forward_recolor)
cudaError_t gs_launch_pack_extra(const GsFrame& f, const GsGeom& g);  // GsScene.extra_colors -> g.xrec
cudaError_t gs_launch_depth_sort(const GsFrame& f, const GsGeom& g);  // result in g.key[s] / g.idx[s], s = hdr->sort_side[3]
// row pass -> column histogram -> plan (ranges, tile_start, blend queue) -> column pass
cudaError_t gs_launch_tile_lists(const GsFrame& f, const GsGeom& g, const GsBinning& b, size_t Rcap, size_t RowCap,
                                 const GsImage& im);
cudaError_t gs_launch_blend_forward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                    float* out_color);
cudaError_t gs_launch_blend_backward(const GsFrame& f, const GsGeom& g, const GsBinning& b, const GsImage& im,
                                     const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                     float* dL_dcolor);
cudaError_t gs_launch_preprocess_backward(const GsFrame& f, const GsGeom& g, const int32_t* radii,
                                          const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                                          float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                                          float* dL_drot);
cudaError_t gs_launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                                   cudaStream_t stream);
cudaError_t gs_launch_make_views(const float* c2w, int N, const float* p4, float* views, cudaStream_t stream);
cudaError_t gs_launch_decode_head(const float* feat, const float* rgb, const float* prim, int P, const GsHeadLayout& L,
                                  float* means3D, float* rot, float* scales, float* opac, float* shs, float* normals,
                                  cudaStream_t stream);
