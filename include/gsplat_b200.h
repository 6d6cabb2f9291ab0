/*
 * gsplat_b200.h -- C ABI of the B200-native Gaussian-splat rasterizer (libgsplat_b200.so).
 *
 * Drop-in boundary for the hot path of huzi96/gaussian-pcloud-render.  Each entry point replaces one
 * reference interface (paths relative to the reference repo, dgr/ = diff-gaussian-rasterization/):
 *
 *   gs_forward        <- CudaRasterizer::Rasterizer::forward    dgr/cuda_rasterizer/rasterizer.h:34-57
 *                        (as driven by RasterizeGaussiansCUDA    dgr/rasterize_points.cu:35-115)
 *   gs_backward       <- CudaRasterizer::Rasterizer::backward   dgr/cuda_rasterizer/rasterizer.h:59-84
 *                        (RasterizeGaussiansBackwardCUDA          dgr/rasterize_points.cu:117-196)
 *   gs_mark_visible   <- CudaRasterizer::Rasterizer::markVisible dgr/cuda_rasterizer/rasterizer.h:28-33
 *                        (markVisible                             dgr/rasterize_points.cu:198-217)
 *   gs_resize_fn      <- std::function<char*(size_t)> buffer callbacks, rasterizer.h:35-37 /
 *                        resizeFunctional rasterize_points.cu:27-33
 *
 * Plain pointers and sizes only: no torch / C++ types cross this boundary.  All data pointers are DEVICE
 * pointers on the current CUDA device unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 * Matrices are 16 floats, column-major (i.e. the transposed row-major tensors the reference's callers pass).
 * A NULL optional pointer selects the same branch an empty tensor selects in the reference binding.
 * Errors are returned as negative codes instead of C++ exceptions.
 */
#ifndef GSPLAT_B200_H_
#define GSPLAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 6

enum {
    GS_OK = 0,
    GS_ERR_INVALID = -1,      /* bad argument (null required pointer, negative size, ...) */
    GS_ERR_CUDA = -2,         /* a CUDA call failed; gs_last_error() has the text */
    GS_ERR_ALLOC = -3,        /* a gs_resize_fn returned NULL */
    GS_ERR_CAPACITY = -4,     /* preallocated binning buffer too small for num_rendered (no-sync mode) */
    GS_ERR_UNSUPPORTED = -5,  /* more than 65536 tiles, or P too large for the packed instance format */
    GS_ERR_PREFILTERED = -6   /* `prefiltered` set but a point failed the near-plane test (reference: __trap) */
};

/* Scene + view: the arguments shared by forward and backward (rasterizer.h:38-55 / :60-74). */
typedef struct GsScene {
    int32_t P;              /* number of Gaussians */
    int32_t sh_degree;      /* D: active SH degree 0..3 */
    int32_t sh_stride;      /* M: SH coefficients stored per Gaussian (shs is [P][M][3]); 0 if shs == NULL */
    int32_t width, height;  /* raster size in pixels */
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t prefiltered;
    int32_t debug;          /* !=0: synchronise and check for CUDA errors after every stage (auxiliary.h:166-173) */
    /* Tile-row shard (multi-GPU): only tile rows [tile_row_begin, tile_row_end) are binned and
     * blended; pixels of other rows are left untouched.  0,0 = the whole frame.  begin == end != 0 is an EMPTY shard
     * (work-balanced partitions can produce one): radii are still computed, nothing is binned, blended or stored. */
    int32_t tile_row_begin, tile_row_end;
    const float* background;     /* [3] */
    const float* means3D;        /* [P][3] */
    const float* shs;            /* [P][M][3] or NULL */
    const float* colors_precomp; /* [P][3] or NULL (exactly one of shs / colors_precomp) */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P][3] or NULL */
    const float* rotations;      /* [P][4] or NULL (un-normalised quaternions are used as they are) */
    const float* cov3D_precomp;  /* [P][6] or NULL (exactly one of scales+rotations / cov3D_precomp) */
    const float* viewmatrix;     /* [16] */
    const float* projmatrix;     /* [16] */
    const float* campos;         /* [3] */
    /* Multi-GPU tile-row sharding with peer stores: when num_peers > 0 the blend epilogue writes this
     * shard's pixels into EVERY listed image (peer-mapped device pointers over NVLink, normally including this
     * rank's own image) instead of into out_color, so the frame is assembled on all GPUs by the blend kernel itself
     * and only a barrier follows.  0 = write out_color only. */
    int32_t num_peers;
    /* Super-sample epilogue: 0 or 1 = full-resolution output (reference behaviour).  2 = the blend
     * epilogue stores the 2x2 box mean of the frame, i.e. exactly the image the reference's caller obtains with
     * F.interpolate(size=(H/2, W/2), mode="bilinear", align_corners=False) (simple_raw_render.py:281-284): out_color,
     * peer_out_color[], extra_out[] and gs_backward's dL_dpix are then [3][H/2][W/2]; width and height (the raster
     * size) must be even. */
    int32_t downsample;
    float* peer_out_color[8];    /* each [3][H][W] */
    /* Extra colour passes blended in the SAME list walk as the frame (forward only): up to three more
     * per-Gaussian colour sets, e.g. the position / hit-map / normal passes of the reference's caller
     * (simple_raw_render.py:411-522).  extra_out[k] receives exactly the image a separate forward with
     * colors_precomp = extra_colors[k] would produce. */
    int32_t num_extra;
    /* Blend scheduling hint; never changes a result.  > 0: a pixel block whose list walk is still running after
     * `team_after` batches of 32 instances is parked by its warp and finished by a CTA working as a team (six warps
     * cull and evaluate alphas, two apply them in list order with the reference's recurrence).  0 = library default
     * (off: measured neutral at the benchmark shapes, csrc/blend_forward.cu), < 0 = off. */
    int32_t team_after;
    const float* extra_colors[3];  /* each [P][3] */
    float* extra_out[3];           /* each [3][H][W] */
    /* Tile-row shards only (multi-GPU): != 0 lets the rank skip, BEFORE the per-Gaussian stage, every Gaussian whose
     * splat cannot reach its tile rows (a conservative bound on the screen radius from the trace of the world
     * covariance; one pass over means / scales / rotations, 40 of the 92-236 input bytes per point), and run the
     * per-Gaussian stage, the depth sort and the list passes on the compacted survivors.  Pixels are unchanged;
     * `radii` is then only written for the survivors (0 = "not in this shard" for a caller-zeroed array), so leave it
     * 0 when a backward pass needs the radii of the whole frame. */
    int32_t shard_cull;
    /* Latency mode of the blend (forward only, plain frames: no extra passes).  0 = off: every result is bit-identical
     * to the reference kernels.  > 0: a pixel block whose list walk is still running after `blend_split` batches of 32
     * instances is parked and finished by a whole CTA that cuts the remaining list into segments, walks them in
     * parallel from the neutral state and merges the partial (colour, transmittance) pairs in list order --
     * front-to-back compositing is associative, only the association of the floating-point sums changes (pixels
     * within ~1e-6 of the exact walk; the early-stop rule T (1 - alpha) < 1e-4 is re-applied exactly on the merged
     * state, so n_contrib / final_T stay consistent for gs_backward).  Cuts the single-frame time of silhouette-heavy
     * frames (one warp walking a 19 K-entry list) and is what lets tile-row shards scale; the north star's pixel
     * tolerance (1e-4) holds with two orders of magnitude to spare.  Takes precedence over team_after. */
    int32_t blend_split;
} GsScene;

/* Growable scratch buffer: fn(user, bytes) must return a DEVICE pointer to at least `bytes` bytes that stays
 * valid until the matching gs_backward call (the reference saves the three buffers on the autograd ctx). */
typedef char* (*gs_resize_fn)(void* user, size_t bytes);
typedef struct GsBuffer {
    gs_resize_fn fn;
    void* user;
} GsBuffer;

/* Forward render of one frame.  out_color [3][H][W] and radii [P] (may be NULL) are written by the library;
 * like the reference binding, the caller zero-fills them first (pixels outside a tile-row shard stay untouched).
 * One host synchronisation (read-back of num_rendered) happens inside, exactly as in rasterizer_impl.cu:281.
 * Returns num_rendered (>= 0) or a negative GS_ERR_* code. */
int64_t gs_forward(const GsScene* scene, GsBuffer geometry, GsBuffer binning, GsBuffer image, float* out_color,
                   int32_t* radii, void* stream);

/* Sizes for callers that preallocate (no-sync mode and integration tests). */
size_t gs_geometry_bytes(int32_t P);
size_t gs_image_bytes(int32_t width, int32_t height);
size_t gs_binning_bytes(int64_t num_rendered_capacity, int32_t P, int32_t width, int32_t height);

/* Forward without any host synchronisation: all three buffers are preallocated (sizes from the functions
 * above, binning sized for `num_rendered_capacity`).  The frame's status (num_rendered, error flags) stays on
 * the device; fetch it with gs_read_status whenever the caller next synchronises.  If num_rendered exceeded the
 * capacity the frame is NOT rendered (status.code == GS_ERR_CAPACITY) and must be re-issued with a bigger buffer. */
int32_t gs_forward_nosync(const GsScene* scene, char* geometry, char* binning, int64_t num_rendered_capacity,
                          char* image, float* out_color, int32_t* radii, void* stream);

/* Additional colour pass over the SAME geometry and camera (the reference's caller rasterizes every
 * view four times -- position, RGB, hit-map and normal passes, simple_raw_render.py:411-522 -- and re-does the
 * identical preprocess + sort each time).  geometry/binning/image must hold a frame rendered by gs_forward[_nosync]
 * with the same means/covariances/opacities/camera/raster size; only scene->shs (+ sh_degree/sh_stride) or
 * scene->colors_precomp are read anew.  Recomputes the per-Gaussian colours and re-runs the blend into out_color:
 * bit-identical to a full forward with those colours.  Forward only (the buffers then describe THIS pass). */
int32_t gs_forward_recolor(const GsScene* scene, char* geometry, char* binning, char* image, float* out_color,
                           void* stream);

typedef struct GsStatus {
    int64_t num_rendered;
    int32_t num_visible;
    int32_t code; /* GS_OK, GS_ERR_CAPACITY or GS_ERR_PREFILTERED */
} GsStatus;
/* Asynchronously copies the status block of a geometry buffer into HOST memory `out` (pinned for true async). */
int32_t gs_read_status(const char* geometry, GsStatus* out, void* stream);

/* Backward.  geometry/binning/image are the buffers filled by the forward call of the same frame, `radii` its
 * radii output.  All gradient arrays are caller-zeroed device arrays (rasterize_points.cu:151-159):
 * dL_dmean2D [P][3], dL_dconic [P][2][2], dL_dopacity [P], dL_dcolor [P][3], dL_dmean3D [P][3],
 * dL_dcov3D [P][6], dL_dsh [P][M][3], dL_dscale [P][3], dL_drot [P][4]; dL_dpix is [3][H][W]. */
int32_t gs_backward(const GsScene* scene, int64_t num_rendered, const int32_t* radii, const char* geometry,
                    const char* binning, const char* image, const float* dL_dpix, float* dL_dmean2D,
                    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                    float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream);

/* The two stages of gs_backward as separate calls, for callers that exchange data in between: a tile-row sharded
 * backward runs GS_BWD_BLEND on every rank over its own rows, sums the four per-Gaussian partial arrays
 * dL_dmean2D / dL_dconic / dL_dopacity / dL_dcolor over the ranks (a Gaussian's tiles span ranks, backward.cu:523-554
 * accumulates into (P, .) arrays), then runs GS_BWD_PREPROCESS.  Same arguments as gs_backward; pointers a stage does
 * not touch may be NULL.  gs_backward(...) == gs_backward_stage(..., GS_BWD_BLEND | GS_BWD_PREPROCESS). */
enum { GS_BWD_BLEND = 1, GS_BWD_PREPROCESS = 2 };
int32_t gs_backward_stage(const GsScene* scene, int64_t num_rendered, const int32_t* radii, const char* geometry,
                          const char* binning, const char* image, const float* dL_dpix, float* dL_dmean2D,
                          float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                          float* dL_dsh, float* dL_dscale, float* dL_drot, int32_t stages, void* stream);

/* present[i] = 1 if point i passes the reference's frustum test (near plane only, auxiliary.h:154). */
int32_t gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                        uint8_t* present, void* stream);

/* Camera set-up for a batch of views on the device (replaces the per-view host work of
 * get_rasterize_param_from_camera, simple_raw_render.py:79-112, and Camera.get_H_w2c = inv_homogeneous_tensors,
 * plib/rigid_motion.py:687-703).  c2w: DEVICE array [N][4][4] of row-major camera-to-world matrices (rigid motions).
 * proj4 = {P[0][0], P[1][1], P[2][2], P[2][3]} of getProjectionMatrix (simple_raw_render.py:50-69), computed by the
 * caller.  views: DEVICE array [N][GS_VIEW_STRIDE] floats; per view: [0,16) viewmatrix = transpose(inverse(c2w)),
 * [16,32) projmatrix = viewmatrix * transpose(P), [32,35) campos = translation of c2w -- the three pointers
 * GsScene wants for view k are views + k*GS_VIEW_STRIDE + {0, 16, 32}. */
#define GS_VIEW_STRIDE 48
int32_t gs_make_views(const float* c2w, int32_t N, const float* proj4_host, float* views, void* stream);

/* Network head -> rasterizer inputs in one kernel (replaces the Python list comprehensions of
 * models/model_v2.py:287-375 and the glue of simple_raw_render.py:243-250,390-394).  The flags are the reference's
 * `args` of the same names; feature columns are consumed in the reference's order: rotation (4), scale (3),
 * opacity (1), offset (3), dc offset (3), normal (3), then sh_ac_coeffs x 3 SH AC coefficients.
 *   rotations = use_rotation ? f[0:4] + (1,0,0,0) : (1,0,0,0)
 *   scales    = (use_scale ? clamp(f + 1, min 0) : 1) * radius              radius = sqrt(3)/scale_factor*6
 *   opacities = (use_opacity && enable_opacity) ? clamp(f, 0, 1) : 1
 *   means3D   = ((primitives + (use_offset ? f : 0)) - xyz_offset) / xyz_factor            (pcgc_rescale)
 *   shs       = [(use_dc_offset ? f : 0) + (dc_rgb - 0.5)/C0 , AC...]   [P][1 + sh_ac_coeffs][3]
 *   normals   = est_normal ? (normalize_normal ? f/max(|f|,1e-12) : f) : untouched        (may be NULL)
 * The reference pads shs with 12 zero coefficients and renders with sh_degree 1 (model_v2.py:358-365); zero
 * coefficients contribute exactly nothing (forward.cu:30-62), so the packed [P][1][3] array with sh_degree 0 gives
 * the same frame bit for bit while preprocess reads 12 instead of 48 SH bytes per Gaussian.
 * Divisions by a scalar are evaluated as torch evaluates them on a CUDA tensor (multiplication by the reciprocal),
 * so every output except the normalised normals equals the reference's GPU result bit for bit.
 * All arrays are DEVICE arrays; features is [P][C] row-major, dc_rgb and primitives are [P][3]. */
typedef struct GsHeadLayout {
    int32_t C;                 /* feature columns */
    int32_t use_rotation, use_scale, use_opacity, use_offset, use_dc_offset, est_normal, normalize_normal;
    int32_t sh_ac_coeffs;      /* AC coefficients taken from the remaining columns (0 = DC only) */
    int32_t enable_opacity;    /* the render() argument (simple_raw_render.py:243-247) */
    float radius;              /* multiplies the decoded scales */
    float xyz_offset, xyz_factor;
} GsHeadLayout;
int32_t gs_decode_head(const float* features, const float* dc_rgb, const float* primitives, int32_t P,
                       const GsHeadLayout* layout, float* means3D, float* rotations, float* scales, float* opacities,
                       float* shs, float* normals, void* stream);

/* Introspection for tests / profiling: copies a named internal array of the last forward into HOST memory.
 * names: "records" (P x 12 f32: x y cx cy | cz opacity thr -cy/cz | r g b -cy/cx), "point_list" (R x u32),
 * "ranges" (Tn x 2 u32), "n_contrib" (H*W u32), "final_T" (H*W f32), "sorted_idx" (P u32),
 * "clamped" (P u8, bit c = channel c clamped), "tiles_touched" (P u32).  Returns bytes copied or a negative code. */
int64_t gs_fetch(const GsScene* scene, const char* geometry, const char* binning, const char* image,
                 int64_t num_rendered, const char* name, void* host_dst, int64_t max_bytes, void* stream);

/* Per-stage device timing of the most recent forward on this host thread (CUDA events on the caller's stream):
 * ms4 = {preprocess, depth sort, tile binning, blend forward}.  Off by default; enabling costs 5 event records. */
void gs_profile_enable(int32_t on);
int32_t gs_profile_read(float* ms4);
/* Same for the most recent backward on this host thread: ms2 = {blend backward, preprocess backward}. */
int32_t gs_profile_read_backward(float* ms2);

/* Number of kernels launched by this library since load (bench.py's "gpu_launches"). */
int64_t gs_launch_count(void);
const char* gs_last_error(void);
int32_t gs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H_ */
