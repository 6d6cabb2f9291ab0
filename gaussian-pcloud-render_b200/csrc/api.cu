// api.cu -- the C ABI of libgsplat_b200.so (include/gsplat_b200.h) and the host orchestration of one frame.
//
// Replaces CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (dgr/cuda_rasterizer/rasterizer_impl.cu:141-153,198-434).  Differences in structure, not in results:
//   * every launch goes to the caller's stream (the reference uses the legacy default stream);
//   * num_rendered comes from per-block atomics in preprocess, so the one host read-back (kept in gs_forward for
//     drop-in buffer sizing, rasterizer_impl.cu:281) overlaps with the depth sort that is already queued;
//     gs_forward_nosync has no host synchronisation at all;
//   * scan + 64-bit pair sort + range detection are replaced by the depth sort, the plan kernel and the row / column
//     partition passes of binning.cu.
#include <cstdlib>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "gs_common.cuh"

namespace {
std::atomic<long long> g_launches{0};
std::mutex g_err_mu;
std::string g_err;

struct HostCtx {  // per-thread pinned status slot + event for the num_rendered read-back
    GsHeader* pinned = nullptr;
    cudaEvent_t ev = nullptr;
    int device = -1;
    bool ensure() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (pinned && dev == device) return true;
        if (!pinned && cudaHostAlloc((void**)&pinned, sizeof(GsHeader), cudaHostAllocDefault) != cudaSuccess) return false;
        if (ev) cudaEventDestroy(ev);
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
        device = dev;
        return true;
    }
};
thread_local HostCtx t_ctx;

// Optional per-stage CUDA-event timing (bench.py's roofline leg): events are recorded on the caller's stream at the
// stage boundaries of the most recent forward; gs_profile_read synchronises on the last one.
struct Profiler {
    bool on = false;
    bool armed = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t bev[3] = {nullptr, nullptr, nullptr};  // backward: before blend, between the stages, after preprocess
    bool barmed = false;
    void bmark(int k, cudaStream_t st) {
        if (!on) return;
        for (auto& e : bev)
            if (!e && cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(bev[k], st);
        barmed = (k == 2) ? true : barmed;
    }
    bool ensure() {
        for (auto& e : ev)
            if (!e && cudaEventCreate(&e) != cudaSuccess) return false;
        return true;
    }
    void mark(int k, cudaStream_t st) {
        if (on && ensure()) { cudaEventRecord(ev[k], st); armed = (k == 4) ? true : armed; }
    }
};
thread_local Profiler t_prof;

int make_frame(const GsScene* s, void* stream, GsFrame& f, bool forward = true) {
    if (!s || s->P < 0 || s->width <= 0 || s->height <= 0) return GS_ERR_INVALID;
    if (s->P > 0) {
        if (!s->means3D || (forward && !s->opacities) || !s->viewmatrix || !s->projmatrix || !s->campos || !s->background)
            return GS_ERR_INVALID;
        if ((s->shs == nullptr) == (s->colors_precomp == nullptr)) return GS_ERR_INVALID;
        const bool has_sr = s->scales != nullptr && s->rotations != nullptr;
        if (has_sr == (s->cov3D_precomp != nullptr)) return GS_ERR_INVALID;
        if (s->shs && s->sh_stride < (s->sh_degree + 1) * (s->sh_degree + 1)) return GS_ERR_INVALID;
        if (s->sh_degree < 0 || s->sh_degree > 3) return GS_ERR_INVALID;
    }
    if (s->downsample < 0 || s->downsample > 2) return GS_ERR_INVALID;
    if (s->downsample == 2 && ((s->width | s->height) & 1)) return GS_ERR_INVALID;
    if (s->num_peers < 0 || s->num_peers > 8) return GS_ERR_INVALID;
    for (int k = 0; k < s->num_peers; k++)
        if (!s->peer_out_color[k]) return GS_ERR_INVALID;
    if (s->num_extra < 0 || s->num_extra > 3) return GS_ERR_INVALID;
    for (int k = 0; k < s->num_extra; k++)
        if (!s->extra_colors[k] || !s->extra_out[k]) return GS_ERR_INVALID;
    f.s = *s;
    f.gx = (s->width + GS_TILE - 1) / GS_TILE;
    f.gy = (s->height + GS_TILE - 1) / GS_TILE;
    f.Tn = f.gx * f.gy;
    f.row0 = 0;
    f.row1 = f.gy;
    if (s->tile_row_end > s->tile_row_begin) {
        f.row0 = s->tile_row_begin < 0 ? 0 : s->tile_row_begin;
        f.row1 = s->tile_row_end > f.gy ? f.gy : s->tile_row_end;
        if (f.row1 < f.row0) f.row1 = f.row0;
    } else if (s->tile_row_begin != 0 || s->tile_row_end != 0) {
        f.row0 = f.row1 = 0;  // empty shard: every tile rectangle is clipped away, nothing is binned or blended
    }
    f.cull = s->shard_cull != 0 && s->P > 0 && f.row1 > f.row0 && (f.row0 > 0 || f.row1 < f.gy);  // a proper shard only
    if (f.gx > GS_MAX_GRID || f.gy > GS_MAX_GRID) return GS_ERR_UNSUPPORTED;  // at most 4096 x 4096 pixels
    if ((unsigned long long)s->P >= (1ull << 30)) return GS_ERR_UNSUPPORTED;  // instance counters stay well inside 32 bits
    f.focal_y = s->height / (2.0f * s->tan_fovy);  // rasterizer_impl.cu:222-223
    f.focal_x = s->width / (2.0f * s->tan_fovx);
    f.stream = (cudaStream_t)stream;
    return GS_OK;
}

#define GS_CU(x)                                   \
    do {                                           \
        cudaError_t e_ = (x);                      \
        if (e_ != cudaSuccess) {                   \
            gs_set_error(#x, e_);                  \
            return GS_ERR_CUDA;                    \
        }                                          \
    } while (0)

// debug mode = the reference's CHECK_CUDA (auxiliary.h:166-173): synchronise + check after every stage
#define GS_STAGE(x)                                                   \
    do {                                                              \
        GS_CU(x);                                                     \
        if (f.s.debug) GS_CU(cudaStreamSynchronize(f.stream));        \
    } while (0)

}  // namespace

void gs_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void gs_set_error(const char* what, cudaError_t e) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
}

extern "C" {

int64_t gs_launch_count(void) { return g_launches.load(); }
const char* gs_last_error(void) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    static thread_local std::string copy;
    copy = g_err;
    return copy.c_str();
}
int32_t gs_abi_version(void) { return GS_ABI_VERSION; }

size_t gs_geometry_bytes(int32_t P) { return GsGeom(nullptr, (size_t)(P < 0 ? 0 : P)).bytes; }
size_t gs_image_bytes(int32_t width, int32_t height) {
    const size_t gx = (width + GS_TILE - 1) / GS_TILE, gy = (height + GS_TILE - 1) / GS_TILE;
    return GsImage(nullptr, (size_t)width * height, gx, gy).bytes;
}
// Row items never outnumber instances, and a Gaussian has at most gy of them.
static size_t row_capacity(size_t cap, size_t P, size_t gy) { return cap < P * gy ? cap : P * gy; }
size_t gs_binning_bytes(int64_t cap, int32_t P, int32_t, int32_t height) {
    const size_t c = (size_t)(cap < 0 ? 0 : cap), gy = (size_t)((height + GS_TILE - 1) / GS_TILE);
    return GsBinning(nullptr, c, row_capacity(c, (size_t)(P < 0 ? 0 : P), gy)).bytes;
}

int64_t gs_forward(const GsScene* scene, GsBuffer geometry, GsBuffer binning, GsBuffer image, float* out_color,
                   int32_t* radii, void* stream) {
    GsFrame f;
    const int rc = make_frame(scene, stream, f);
    if (rc != GS_OK) return rc;
    if (!geometry.fn || !binning.fn || !image.fn || !out_color) return GS_ERR_INVALID;
    if (f.s.P == 0) return 0;  // rasterize_points.cu:81 -- nothing is launched, outputs stay zero-filled
    if (!t_ctx.ensure()) { gs_set_error("pinned status slot", cudaGetLastError()); return GS_ERR_CUDA; }

    const size_t N = (size_t)f.s.width * f.s.height;
    char* gptr = geometry.fn(geometry.user, GsGeom(nullptr, f.s.P).bytes);
    char* iptr = image.fn(image.user, GsImage(nullptr, N, f.gx, f.gy).bytes);
    if (!gptr || !iptr) return GS_ERR_ALLOC;
    GsGeom g(gptr, f.s.P);
    GsImage im(iptr, N, f.gx, f.gy);

    GS_CU(cudaMemsetAsync(gptr, 0, g.zero_bytes, f.stream));
    GS_CU(cudaMemsetAsync(iptr, 0, im.zero_bytes, f.stream));
    t_prof.mark(0, f.stream);
    GS_STAGE(gs_launch_shard_cull(f, g));
    GS_STAGE(gs_launch_preprocess(f, g, im, radii));
    t_prof.mark(1, f.stream);
    if (f.row1 == f.row0) return 0;  // empty tile-row shard: radii are done, there is nothing to bin or blend
    // read the instance / row-item counts back while the depth sort (which does not depend on them) is already
    // queued behind the copy
    GS_CU(cudaMemcpyAsync(t_ctx.pinned, g.hdr, 32, cudaMemcpyDeviceToHost, f.stream));
    GS_CU(cudaEventRecord(t_ctx.ev, f.stream));
    GS_STAGE(gs_launch_depth_sort(f, g));
    t_prof.mark(2, f.stream);
    GS_CU(cudaEventSynchronize(t_ctx.ev));
    const unsigned long long R = t_ctx.pinned->num_rendered;
    const size_t rows = t_ctx.pinned->num_row_items;
    if (t_ctx.pinned->code == GS_ERR_PREFILTERED) return GS_ERR_PREFILTERED;
    if (R >= (1ull << 30)) return GS_ERR_UNSUPPORTED;

    char* bptr = binning.fn(binning.user, GsBinning(nullptr, (size_t)R, rows).bytes);
    if (!bptr) return GS_ERR_ALLOC;
    GsBinning b(bptr, (size_t)R, rows);
    GS_STAGE(gs_launch_tile_lists(f, g, b, (size_t)R, rows, im));
    t_prof.mark(3, f.stream);
    GS_STAGE(gs_launch_pack_extra(f, g));
    GS_STAGE(gs_launch_blend_forward(f, g, b, im, out_color));
    t_prof.mark(4, f.stream);
    return (int64_t)R;
}

int32_t gs_forward_nosync(const GsScene* scene, char* geometry, char* binning, int64_t cap, char* image,
                          float* out_color, int32_t* radii, void* stream) {
    GsFrame f;
    const int rc = make_frame(scene, stream, f);
    if (rc != GS_OK) return rc;
    if (!geometry || !binning || !image || !out_color || cap < 0) return GS_ERR_INVALID;
    if (f.s.P == 0) return GS_OK;
    const size_t N = (size_t)f.s.width * f.s.height;
    const size_t rowcap = row_capacity((size_t)cap, (size_t)f.s.P, (size_t)f.gy);
    GsGeom g(geometry, f.s.P);
    GsImage im(image, N, f.gx, f.gy);
    GsBinning b(binning, (size_t)cap, rowcap);
    GS_CU(cudaMemsetAsync(geometry, 0, g.zero_bytes, f.stream));
    GS_CU(cudaMemsetAsync(image, 0, im.zero_bytes, f.stream));
    t_prof.mark(0, f.stream);
    GS_STAGE(gs_launch_shard_cull(f, g));
    GS_STAGE(gs_launch_preprocess(f, g, im, radii));
    t_prof.mark(1, f.stream);
    if (f.row1 == f.row0) return GS_OK;  // empty tile-row shard
    // developer switch (tools/frontend_cost.py): stop after the n-th stage to time the front end alone with frames in flight
    const char* stop_env = getenv("GSPLAT_B200_STOP_AFTER");
    const int stop_after = stop_env ? atoi(stop_env) : 0;
    if (stop_after == 1) return GS_OK;
    GS_STAGE(gs_launch_depth_sort(f, g));
    t_prof.mark(2, f.stream);
    if (stop_after == 2) return GS_OK;
    GS_STAGE(gs_launch_tile_lists(f, g, b, (size_t)cap, rowcap, im));
    t_prof.mark(3, f.stream);
    if (stop_after == 3) return GS_OK;
    GS_STAGE(gs_launch_pack_extra(f, g));
    GS_STAGE(gs_launch_blend_forward(f, g, b, im, out_color));
    t_prof.mark(4, f.stream);
    return GS_OK;
}

int32_t gs_forward_recolor(const GsScene* scene, char* geometry, char* binning, char* image, float* out_color,
                           void* stream) {
    GsFrame f;
    const int rc = make_frame(scene, stream, f);
    if (rc != GS_OK) return rc;
    if (!geometry || !binning || !image || !out_color) return GS_ERR_INVALID;
    if (f.s.P == 0) return GS_OK;
    GsGeom g(geometry, f.s.P);
    GsImage im(image, (size_t)f.s.width * f.s.height, f.gx, f.gy);
    GsBinning b(binning, 0, 0);  // only `list` (first array) is used
    // restart the blend work queue; everything else in the header stays as the frame left it
    GS_CU(cudaMemsetAsync(&g.hdr->tickets[6], 0, 6 * sizeof(unsigned int), f.stream));  // blend queue + team counters (6-11)
    GS_CU(cudaMemsetAsync(im.park_ready, 0, GS_PARK_CAP * sizeof(unsigned), f.stream));
    GS_STAGE(gs_launch_recolor(f, g));
    GS_STAGE(gs_launch_pack_extra(f, g));
    GS_STAGE(gs_launch_blend_forward(f, g, b, im, out_color));
    return GS_OK;
}

void gs_profile_enable(int32_t on) { t_prof.on = on != 0; t_prof.armed = false; t_prof.barmed = false; }

int32_t gs_profile_read(float* ms4) {
    if (!ms4 || !t_prof.on || !t_prof.armed) return GS_ERR_INVALID;
    GS_CU(cudaEventSynchronize(t_prof.ev[4]));
    for (int k = 0; k < 4; k++) GS_CU(cudaEventElapsedTime(&ms4[k], t_prof.ev[k], t_prof.ev[k + 1]));
    return GS_OK;
}

int32_t gs_profile_read_backward(float* ms2) {
    if (!ms2 || !t_prof.on || !t_prof.barmed) return GS_ERR_INVALID;
    GS_CU(cudaEventSynchronize(t_prof.bev[2]));
    for (int k = 0; k < 2; k++) GS_CU(cudaEventElapsedTime(&ms2[k], t_prof.bev[k], t_prof.bev[k + 1]));
    return GS_OK;
}

int32_t gs_read_status(const char* geometry, GsStatus* out, void* stream) {
    if (!geometry || !out) return GS_ERR_INVALID;
    static_assert(sizeof(GsStatus) == 16, "GsStatus mirrors the first 16 bytes of GsHeader");
    GS_CU(cudaMemcpyAsync(out, geometry, sizeof(GsStatus), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return GS_OK;
}

int32_t gs_backward_stage(const GsScene* scene, int64_t num_rendered, const int32_t* radii, const char* geometry,
                          const char* binning, const char* image, const float* dL_dpix, float* dL_dmean2D,
                          float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                          float* dL_dsh, float* dL_dscale, float* dL_drot, int32_t stages, void* stream) {
    GsFrame f;
    const int rc = make_frame(scene, stream, f, false);
    if (rc != GS_OK) return rc;
    if ((stages & ~(GS_BWD_BLEND | GS_BWD_PREPROCESS)) || stages == 0) return GS_ERR_INVALID;
    if (f.s.P == 0) return GS_OK;
    const bool blend = stages & GS_BWD_BLEND, pre = stages & GS_BWD_PREPROCESS;
    if (!geometry || !image || !dL_dmean2D || !dL_dconic || !dL_dcolor || num_rendered < 0) return GS_ERR_INVALID;
    if (blend && (!dL_dpix || !dL_dopacity)) return GS_ERR_INVALID;
    if (pre) {
        if (!radii || !dL_dmean3D || !dL_dcov3D) return GS_ERR_INVALID;
        if (f.s.shs && !dL_dsh) return GS_ERR_INVALID;
        if (f.s.scales && (!dL_dscale || !dL_drot)) return GS_ERR_INVALID;
    }
    GsGeom g(const_cast<char*>(geometry), f.s.P);
    GsImage im(const_cast<char*>(image), (size_t)f.s.width * f.s.height, f.gx, f.gy);
    t_prof.bmark(0, f.stream);
    if (blend && num_rendered > 0) {
        if (!binning) return GS_ERR_INVALID;
        GsBinning b(const_cast<char*>(binning), (size_t)num_rendered, 0);  // only `list` (first array) is used
        GS_STAGE(gs_launch_blend_backward(f, g, b, im, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor));
    }
    t_prof.bmark(1, f.stream);
    if (pre)
        GS_STAGE(gs_launch_preprocess_backward(f, g, radii, dL_dmean2D, dL_dconic, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
                                              dL_dscale, dL_drot));
    t_prof.bmark(2, f.stream);
    return GS_OK;
}

int32_t gs_backward(const GsScene* scene, int64_t num_rendered, const int32_t* radii, const char* geometry,
                    const char* binning, const char* image, const float* dL_dpix, float* dL_dmean2D,
                    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                    float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream) {
    return gs_backward_stage(scene, num_rendered, radii, geometry, binning, image, dL_dpix, dL_dmean2D, dL_dconic,
                             dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
                             GS_BWD_BLEND | GS_BWD_PREPROCESS, stream);
}

int32_t gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                        uint8_t* present, void* stream) {
    (void)projmatrix;  // the reference's test only uses the view matrix (auxiliary.h:154)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return GS_ERR_INVALID;
    if (P == 0) return GS_OK;
    GS_CU(gs_launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream));
    return GS_OK;
}

int32_t gs_make_views(const float* c2w, int32_t N, const float* proj4_host, float* views, void* stream) {
    if (N < 0 || !proj4_host || (N > 0 && (!c2w || !views))) return GS_ERR_INVALID;
    GS_CU(gs_launch_make_views(c2w, N, proj4_host, views, (cudaStream_t)stream));
    return GS_OK;
}

int32_t gs_decode_head(const float* features, const float* dc_rgb, const float* primitives, int32_t P,
                       const GsHeadLayout* L, float* means3D, float* rotations, float* scales, float* opacities,
                       float* shs, float* normals, void* stream) {
    if (P < 0 || !L || L->C < 0 || L->C > 96 || L->sh_ac_coeffs < 0) return GS_ERR_INVALID;
    const int need = 4 * !!L->use_rotation + 3 * !!L->use_scale + !!L->use_opacity + 3 * !!L->use_offset +
                     3 * !!L->use_dc_offset + 3 * !!L->est_normal + 3 * L->sh_ac_coeffs;
    if (need > L->C || L->xyz_factor == 0.f) return GS_ERR_INVALID;
    if (P > 0 && ((L->C > 0 && !features) || !dc_rgb || !primitives || !means3D || !rotations || !scales || !opacities ||
                  !shs))
        return GS_ERR_INVALID;
    GS_CU(gs_launch_decode_head(features, dc_rgb, primitives, P, *L, means3D, rotations, scales, opacities, shs, normals,
                                (cudaStream_t)stream));
    return GS_OK;
}

int64_t gs_fetch(const GsScene* scene, const char* geometry, const char* binning, const char* image,
                 int64_t num_rendered, const char* name, void* host_dst, int64_t max_bytes, void* stream) {
    GsFrame f;
    const int rc = make_frame(scene, stream, f, false);
    if (rc != GS_OK) return rc;
    if (!name || !host_dst || !geometry || !image || (num_rendered > 0 && !binning)) return GS_ERR_INVALID;
    const size_t P = f.s.P, N = (size_t)f.s.width * f.s.height, R = (size_t)num_rendered;
    GsGeom g(const_cast<char*>(geometry), P);
    GsImage im(const_cast<char*>(image), N, f.gx, f.gy);
    GsBinning b(const_cast<char*>(binning), R, 0);
    const void* src = nullptr;
    size_t n = 0;
    unsigned side = 0;
    if (!strncmp(name, "sorted_", 7)) {  // which side of the ping-pong holds the depth order (GsHeader::sort_side[3])
        GS_CU(cudaMemcpyAsync(&side, &g.hdr->sort_side[3], sizeof(side), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        GS_CU(cudaStreamSynchronize((cudaStream_t)stream));
        side &= 1u;
    }
    if (!strcmp(name, "records")) { src = g.rec; n = sizeof(GsRec) * P; }
    else if (!strcmp(name, "sorted_idx")) { src = g.idx[side]; n = 4 * P; }
    else if (!strcmp(name, "sorted_key")) { src = g.key[side]; n = 4 * P; }
    else if (!strcmp(name, "clamped")) { src = g.clamp; n = P; }
    else if (!strcmp(name, "tiles_touched")) { src = g.ntile; n = 4 * P; }
    else if (!strcmp(name, "point_list")) { src = b.list; n = 4 * R; }
    else if (!strcmp(name, "ranges")) { src = im.ranges; n = 8 * (size_t)f.Tn; }
    else if (!strcmp(name, "n_contrib")) { src = im.n_contrib; n = 4 * N; }
    else if (!strcmp(name, "final_T")) { src = im.final_T; n = 4 * N; }
    else return GS_ERR_INVALID;
    if ((src == nullptr && n) || (int64_t)n > max_bytes) return GS_ERR_INVALID;
    if (n) {
        GS_CU(cudaMemcpyAsync(host_dst, src, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        GS_CU(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return (int64_t)n;
}

}  // extern "C"
