/*
 * ref_shim.cu -- C-ABI shim around the UNMODIFIED reference CUDA library
 * (CudaRasterizer::Rasterizer, /root/reference/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:24-84).
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles the reference's own forward.cu / backward.cu /
 * rasterizer_impl.cu *where they lie* under /root/reference together with this file into
 * oracle/_ref/libgs_ref.so (git-ignored; it travels to the GPU box with the repo snapshot).  No reference
 * source is copied into the repo.  This file only plays the role of the reference's torch binding
 * (rasterize_points.cu:35-217): it owns the three growable scratch buffers, passes raw device pointers
 * through, and exposes the reference's internal per-stage arrays so tests can compare stage by stage.
 * It is used (a) to generate tests/golden/ fixtures on a B200, (b) as the live GPU parity target in
 * `-m gpu` tests, (c) as the `--impl reference` arm of bench.py.  Product code never loads it.
 */
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>

#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Buf {
    char* p = nullptr;
    size_t cap = 0;
    size_t used = 0;
    char* grow(size_t n) {
        used = n;
        if (n > cap) {
            if (p) cudaFree(p);
            size_t want = n + n / 4 + 256;
            if (cudaMalloc(&p, want) != cudaSuccess) throw std::runtime_error("gsref: cudaMalloc failed");
            cap = want;
        }
        return p;
    }
    ~Buf() { if (p) cudaFree(p); }
};
struct RefState {
    Buf geom, binning, img;
    int P = 0, R = 0, W = 0, H = 0;
};
}  // namespace

extern "C" {

void* gsref_create() { return new RefState(); }
void gsref_destroy(void* s) { delete static_cast<RefState*>(s); }

/* Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:35-115): out_color / radii are zero-filled by the caller.
 * Null pointers select the "compute it" branches exactly as empty tensors do in the reference binding.
 * Returns num_rendered (>= 0) or -1 on a C++ exception. */
int gsref_forward(void* state, int P, int D, int M, const float* background, int W, int H, const float* means3D,
                  const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                  float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, int prefiltered,
                  float* out_color, int* radii, int debug) {
    RefState* st = static_cast<RefState*>(state);
    st->P = P; st->W = W; st->H = H; st->R = 0;
    if (P == 0) return 0;
    try {
        std::function<char*(size_t)> g = [st](size_t n) { return st->geom.grow(n); };
        std::function<char*(size_t)> b = [st](size_t n) { return st->binning.grow(n); };
        std::function<char*(size_t)> i = [st](size_t n) { return st->img.grow(n); };
        st->R = CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, background, W, H, means3D, shs, colors_precomp,
                                                    opacities, scales, scale_modifier, rotations, cov3D_precomp,
                                                    viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
                                                    prefiltered != 0, out_color, radii, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "gsref_forward: %s\n", e.what());
        return -1;
    }
    return st->R;
}

/* Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:117-196); all gradient arrays caller-zeroed. */
int gsref_backward(void* state, int P, int D, int M, const float* background, int W, int H, const float* means3D,
                   const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                   const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                   const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                   const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                   float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug) {
    RefState* st = static_cast<RefState*>(state);
    if (P == 0) return 0;
    try {
        CudaRasterizer::Rasterizer::backward(P, D, M, st->R, background, W, H, means3D, shs, colors_precomp, scales,
                                             scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
                                             tan_fovx, tan_fovy, radii, st->geom.p, st->binning.p, st->img.p, dL_dpix,
                                             dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D,
                                             dL_dsh, dL_dscale, dL_drot, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "gsref_backward: %s\n", e.what());
        return -1;
    }
    return 0;
}

void gsref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    if (P) CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

/* Copy one of the reference's internal arrays (device) into a host buffer.  Returns bytes copied or -1.
 * names: depths means2D cov3D conic_opacity rgb clamped tiles_touched point_offsets
 *        point_list point_list_keys ranges n_contrib accum_alpha */
long long gsref_fetch(void* state, const char* name, void* host_dst, long long max_bytes) {
    RefState* st = static_cast<RefState*>(state);
    if (!st->geom.p) return -1;
    using namespace CudaRasterizer;
    char* gp = st->geom.p;
    GeometryState g = GeometryState::fromChunk(gp, st->P);
    char* ip = st->img.p;
    ImageState im = ImageState::fromChunk(ip, (size_t)st->W * st->H);
    BinningState b{};
    if (st->R > 0) { char* bp = st->binning.p; b = BinningState::fromChunk(bp, st->R); }
    const void* src = nullptr;
    size_t n = 0;
    size_t P = st->P, R = st->R, N = (size_t)st->W * st->H;
    size_t Tn = (size_t)((st->W + 15) / 16) * ((st->H + 15) / 16);
    if (!strcmp(name, "depths")) { src = g.depths; n = 4 * P; }
    else if (!strcmp(name, "means2D")) { src = g.means2D; n = 8 * P; }
    else if (!strcmp(name, "cov3D")) { src = g.cov3D; n = 24 * P; }
    else if (!strcmp(name, "conic_opacity")) { src = g.conic_opacity; n = 16 * P; }
    else if (!strcmp(name, "rgb")) { src = g.rgb; n = 12 * P; }
    else if (!strcmp(name, "clamped")) { src = g.clamped; n = 3 * P; }
    else if (!strcmp(name, "tiles_touched")) { src = g.tiles_touched; n = 4 * P; }
    else if (!strcmp(name, "point_offsets")) { src = g.point_offsets; n = 4 * P; }
    else if (!strcmp(name, "point_list")) { src = b.point_list; n = 4 * R; }
    else if (!strcmp(name, "point_list_keys")) { src = b.point_list_keys; n = 8 * R; }
    else if (!strcmp(name, "ranges")) { src = im.ranges; n = 8 * Tn; }
    else if (!strcmp(name, "n_contrib")) { src = im.n_contrib; n = 4 * N; }
    else if (!strcmp(name, "accum_alpha")) { src = im.accum_alpha; n = 4 * N; }
    else return -1;
    if ((long long)n > max_bytes) return -1;
    if (n && cudaMemcpy(host_dst, src, n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)n;
}

int gsref_num_rendered(void* state) { return static_cast<RefState*>(state)->R; }

}  // extern "C"
