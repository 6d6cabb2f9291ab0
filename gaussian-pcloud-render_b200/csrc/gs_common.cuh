// gs_common.cuh -- shared definitions of the B200-native splat rasterizer (sm_100a only).
//
// Data layout in HBM (all carved from the three caller-owned byte buffers, 256-B aligned sub-arrays):
//   geometry buffer (per Gaussian, P entries)
//     GsHeader            status block + counters (256 B)
//     rec   [P] GsRec     48-B packed record read by the blend kernels (3 x float4)
//     key   [2][P] u32    depth-sort keys (float bits of view-space z; 0xFFFFFFFF = culled), ping/pong
//     idx   [2][P] u32    depth-sort values (Gaussian index), ping/pong
//     rect  [P] ushort4   tile rectangle [x0,x1) x [y0,y1)
//     ntile [P] u32       tiles touched
//     cov3D [P][6] f32    world covariance (only when computed from scale/rotation)
//     clamp [P] u8        bit c set = SH colour channel c was clamped at 0
//     sort scratch        per-warp digit histograms + look-back state of the depth sort
//   binning buffer (per instance, R entries)
//     stage [R] u32       instances after tile pass 1: (tile_hi << idx_bits) | gaussian
//     list  [R] u32       final per-tile, depth-ordered Gaussian index list ("point_list")
//     hist1/hist2 + look-back state of the two tile passes, bucket table
//   image buffer
//     final_T [N] f32, n_contrib [N] u32, ranges [Tn] uint2, order [Tn] u32 (longest-list-first tile queue)
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
    unsigned int tickets[16];  // dynamic block ids of the look-back scans
    unsigned int pad[44];
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

// radix machinery: one WARP is the unit of work of every radix pass.
#define GS_RADIX_BITS 8
#define GS_RADIX 256
#define GS_SCAN_ITEMS 8           // scan kernel: items per thread
#define GS_SCAN_THREADS 256
#define GS_SCAN_TILE (GS_SCAN_ITEMS * GS_SCAN_THREADS)
#define GS_DEPTH_UNIT 512         // keys per warp in a depth-sort pass
#define GS_EMIT_UNIT 128          // depth-sorted Gaussians per warp in tile pass 1
#define GS_TILE2_UNIT 2048        // instances per warp in tile pass 2

static inline __host__ __device__ size_t gs_div_up(size_t a, size_t b) { return (a + b - 1) / b; }

struct GsGeom {
    GsHeader* hdr;
    unsigned long long* dstate;  // look-back state of the depth-sort scans (directly after hdr: one memset)
    GsRec* rec;
    uint32_t* key[2];
    uint32_t* idx[2];
    ushort4* rect;
    uint32_t* ntile;
    float* cov3D;
    uint8_t* clamp;
    uint32_t* dhist;             // [256][depth_units] (+1)
    size_t depth_units, dhist_len, dstate_len, zero_bytes, bytes;
    __host__ __device__ GsGeom(char* base, size_t P) {
        GsCarver c(base);
        hdr = c.take<GsHeader>(1);
        depth_units = gs_div_up(P, GS_DEPTH_UNIT);
        dhist_len = GS_RADIX * depth_units + 1;
        dstate_len = gs_div_up(dhist_len, GS_SCAN_TILE) + 1;
        dstate = c.take<unsigned long long>(dstate_len);
        zero_bytes = c.off;  // header + scan state are zeroed at the start of every frame
        rec = c.take<GsRec>(P);
        key[0] = c.take<uint32_t>(P); key[1] = c.take<uint32_t>(P);
        idx[0] = c.take<uint32_t>(P); idx[1] = c.take<uint32_t>(P);
        rect = c.take<ushort4>(P);
        ntile = c.take<uint32_t>(P);
        cov3D = c.take<float>(6 * P);
        clamp = c.take<uint8_t>(P);
        dhist = c.take<uint32_t>(dhist_len);
        bytes = c.off + GS_ALIGN;
    }
};

struct GsBinning {
    unsigned long long* state1;  // look-back state of the two tile-pass scans (first: one memset)
    unsigned long long* state2;
    uint32_t* stage;
    uint32_t* list;
    uint32_t* hist1;  // [256][emit_units] (+1): per-warp low-digit histogram of tile pass 1
    uint32_t* hist2;  // [256][units2]     (+1): per-warp high-digit histogram of tile pass 2
    uint32_t* bucket_unit0;  // [257] first pass-2 unit of each low-digit bucket
    size_t emit_units, units2, hist1_len, hist2_len, state1_len, state2_len, zero_bytes, bytes;
    __host__ __device__ GsBinning(char* base, size_t Rcap, size_t P) {
        GsCarver c(base);
        emit_units = gs_div_up(P, GS_EMIT_UNIT);
        hist1_len = GS_RADIX * emit_units + 1;
        units2 = gs_div_up(Rcap, GS_TILE2_UNIT) + GS_RADIX;
        hist2_len = GS_RADIX * units2 + 1;
        state1_len = gs_div_up(hist1_len, GS_SCAN_TILE) + 1;
        state2_len = gs_div_up(hist2_len, GS_SCAN_TILE) + 1;
        state1 = c.take<unsigned long long>(state1_len);
        state2 = c.take<unsigned long long>(state2_len);
        zero_bytes = c.off;
        stage = c.take<uint32_t>(Rcap);
        list = c.take<uint32_t>(Rcap);
        hist1 = c.take<uint32_t>(hist1_len);
        hist2 = c.take<uint32_t>(hist2_len);
        bucket_unit0 = c.take<uint32_t>(GS_RADIX + 1);
        bytes = c.off + GS_ALIGN;
    }
};

struct GsImage {
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;
    uint32_t* order;  // tiles of the shard, longest instance list first (blend work queue)
    size_t bytes;
    __host__ __device__ GsImage(char* base, size_t N, size_t Tn) {
        GsCarver c(base);
        final_T = c.take<float>(N);
        n_contrib = c.take<uint32_t>(N);
        ranges = c.take<uint2>(Tn);
        order = c.take<uint32_t>(Tn);
        bytes = c.off + GS_ALIGN;
    }
};

// launch bookkeeping (gs_launch_count) and error text
void gs_note_launch();
void gs_set_error(const char* what, cudaError_t e);

struct GsFrame {  // host-side derived quantities handed to every launcher
    GsScene s;
    int gx, gy, Tn;        // tile grid
    int row0, row1;        // tile-row shard
    int tile_bits, hi_bits, idx_bits;
    float focal_x, focal_y;
    cudaStream_t stream;
};

// stage launchers (each in its own translation unit)
cudaError_t gs_launch_preprocess(const GsFrame& f, const GsGeom& g, int32_t* radii);
cudaError_t gs_launch_depth_sort(const GsFrame& f, const GsGeom& g, int* sorted_side);
cudaError_t gs_launch_tile_binning(const GsFrame& f, const GsGeom& g, int sorted_side, const GsBinning& b,
                                   size_t Rcap, const GsImage& im);
cudaError_t gs_launch_tile_order(const GsFrame& f, const GsGeom& g, const GsBinning& b, size_t Rcap, const GsImage& im);
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
