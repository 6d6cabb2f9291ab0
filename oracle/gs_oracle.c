/*
 * gs_oracle.c -- CPU ORACLE for the Gaussian-splat rasterizer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (gaussian-pcloud-render_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a from-scratch plain-C restatement of the *semantics* of the reference's CUDA
 * path (diff-gaussian-rasterization, "dgr/" below, paths relative to /root/reference/):
 *   preprocess   dgr/cuda_rasterizer/forward.cu:158-259  (+ :20-71 SH, :74-116 cov2D, :121-155 cov3D,
 *                dgr/cuda_rasterizer/auxiliary.h:41-56 ndc2Pix/getRect, :139-164 in_frustum)
 *   scan         dgr/cuda_rasterizer/rasterizer_impl.cu:277      (inclusive prefix sum)
 *   duplicate    dgr/cuda_rasterizer/rasterizer_impl.cu:70-111   ((tile<<32)|depth_bits keys)
 *   sort         dgr/cuda_rasterizer/rasterizer_impl.cu:300-308  (stable, ascending, bits [0,32+msb))
 *   ranges       dgr/cuda_rasterizer/rasterizer_impl.cu:116-138,310
 *   blend fwd    dgr/cuda_rasterizer/forward.cu:264-377
 *   blend bwd    dgr/cuda_rasterizer/backward.cu:399-557
 *   cov2D bwd    dgr/cuda_rasterizer/backward.cu:144-274
 *   preproc bwd  dgr/cuda_rasterizer/backward.cu:346-396 (+ :20-139 SH bwd, :278-341 cov3D bwd)
 *   markVisible  dgr/cuda_rasterizer/rasterizer_impl.cu:54-66
 * GLM conventions used by the reference (column-major mat3 constructor, product order
 * dgr/third_party/glm/glm/detail/type_mat3x3.inl:486-519) are restated in m3_mul() below.
 *
 * PARITY PIN: the reference ships no tests / golden vectors.  The oracle is
 * pinned against outputs of the UNMODIFIED reference CUDA library (oracle/_ref/libgs_ref.so,
 * built from the sources where they lie under /root/reference by oracle/Makefile) run on a
 * B200; those outputs are committed under tests/golden/ by tests/golden/make_golden.py.
 *
 * Arithmetic: every quantity is computed in `real` (float by default; -DGSO_DOUBLE builds the
 * fp64 flavour used for finite-difference gradient checks).  No FMA contraction (-ffp-contract=off).
 * The only known sub-ulp deviations from the CUDA reference are its FMA contraction and libdevice
 * expf; both are far below the 1e-4 pixel tolerance except at the discontinuities listed in
 * SURVEY.md App. A (alpha<1/255, T<1e-4, ceil(3 sigma)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef GSO_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FMIN fmin
#define R_FMAX fmax
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FMIN fminf
#define R_FMAX fmaxf
#endif
#define RC(x) ((real)(x))

#define TILE_X 16 /* dgr/cuda_rasterizer/config.h:16-17 */
#define TILE_Y 16
#define TILE_PIX (TILE_X * TILE_Y)

/* SH basis constants, dgr/cuda_rasterizer/auxiliary.h:22-39 (values are the standard real-SH ones) */
static const real SH0 = RC(0.28209479177387814f);
static const real SH1 = RC(0.4886025119029199f);
static const real SH2[5] = {RC(1.0925484305920792f), RC(-1.0925484305920792f), RC(0.31539156525252005f),
                            RC(-1.0925484305920792f), RC(0.5462742152960396f)};
static const real SH3[7] = {RC(-0.5900435899266435f), RC(2.890611442640554f), RC(-0.4570457994644658f),
                            RC(0.3731763325901154f), RC(-0.4570457994644658f), RC(1.445305721320277f),
                            RC(-0.5900435899266435f)};

int gso_real_bytes(void) { return (int)sizeof(real); }
int gso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- column-major 3x3, m.c[col][row], like glm::mat3 ---- */
typedef struct { real c[3][3]; } m3;

static m3 m3_cols(real a0, real a1, real a2, real b0, real b1, real b2, real c0, real c1, real c2) {
    m3 m; /* glm::mat3(x0..x8) fills column by column */
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}
static m3 m3_mul(const m3* A, const m3* B) { /* type_mat3x3.inl:486-519: sum over k in order 0,1,2 */
    m3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            r.c[j][i] = A->c[0][i] * B->c[j][0] + A->c[1][i] * B->c[j][1] + A->c[2][i] * B->c[j][2];
    return r;
}
static m3 m3_t(const m3* A) {
    m3 r;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) r.c[j][i] = A->c[i][j];
    return r;
}

/* CUDA float->int conversion saturates and maps NaN to 0 (cvt.rzi.s32.f32). */
static int sat_int(real v) {
    if (v != v) return 0;
    if (v >= RC(2147483647.0)) return 2147483647;
    if (v <= RC(-2147483648.0)) return (-2147483647 - 1);
    return (int)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:41-44 -- the literals there are doubles, so this is evaluated in double */
static real ndc_to_pix(real v, int S) { return (real)((((double)v + 1.0) * S - 1.0) * 0.5); }

/* auxiliary.h:46-56 */
static void tile_rect(real px, real py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    *x0 = imin(gx, imax(0, sat_int((px - (real)max_radius) / (real)TILE_X)));
    *y0 = imin(gy, imax(0, sat_int((py - (real)max_radius) / (real)TILE_Y)));
    *x1 = imin(gx, imax(0, sat_int((px + (real)max_radius + (real)TILE_X - (real)1) / (real)TILE_X)));
    *y1 = imin(gy, imax(0, sat_int((py + (real)max_radius + (real)TILE_Y - (real)1) / (real)TILE_Y)));
}

/* auxiliary.h:58-77; matrices are column-major in memory */
static void xform43(const float* m, real x, real y, real z, real* o) {
    o[0] = RC(m[0]) * x + RC(m[4]) * y + RC(m[8]) * z + RC(m[12]);
    o[1] = RC(m[1]) * x + RC(m[5]) * y + RC(m[9]) * z + RC(m[13]);
    o[2] = RC(m[2]) * x + RC(m[6]) * y + RC(m[10]) * z + RC(m[14]);
}
static void xform44(const float* m, real x, real y, real z, real* o) {
    o[0] = RC(m[0]) * x + RC(m[4]) * y + RC(m[8]) * z + RC(m[12]);
    o[1] = RC(m[1]) * x + RC(m[5]) * y + RC(m[9]) * z + RC(m[13]);
    o[2] = RC(m[2]) * x + RC(m[6]) * y + RC(m[10]) * z + RC(m[14]);
    o[3] = RC(m[3]) * x + RC(m[7]) * y + RC(m[11]) * z + RC(m[15]);
}

/* forward.cu:121-155; quaternion deliberately NOT normalised (:130) */
static void cov3d_from_scale_rot(const float* s, real mod, const float* q, real* cov6) {
    m3 S = m3_cols(mod * RC(s[0]), 0, 0, 0, mod * RC(s[1]), 0, 0, 0, mod * RC(s[2]));
    real r = RC(q[0]), x = RC(q[1]), y = RC(q[2]), z = RC(q[3]);
    m3 Rm = m3_cols(RC(1) - RC(2) * (y * y + z * z), RC(2) * (x * y - r * z), RC(2) * (x * z + r * y),
                    RC(2) * (x * y + r * z), RC(1) - RC(2) * (x * x + z * z), RC(2) * (y * z - r * x),
                    RC(2) * (x * z - r * y), RC(2) * (y * z + r * x), RC(1) - RC(2) * (x * x + y * y));
    m3 M = m3_mul(&S, &Rm);
    m3 Mt = m3_t(&M);
    m3 Sg = m3_mul(&Mt, &M);
    cov6[0] = Sg.c[0][0]; cov6[1] = Sg.c[0][1]; cov6[2] = Sg.c[0][2];
    cov6[3] = Sg.c[1][1]; cov6[4] = Sg.c[1][2]; cov6[5] = Sg.c[2][2];
}

/* shared by forward.cu:74-116 and backward.cu:144-211: builds T = W*J and the dilated 2D covariance */
typedef struct { m3 T, Vrk, W; real tx, ty, tz, txtz, tytz, limx, limy; real a, b, c; } cov2d_ctx;

static void cov2d_eval(real mx, real my, real mz, real fx, real fy, real tanx, real tany, const real* cov6,
                       const float* view, cov2d_ctx* k) {
    real t[3];
    xform43(view, mx, my, mz, t);
    k->limx = RC(1.3f) * tanx;
    k->limy = RC(1.3f) * tany;
    k->txtz = t[0] / t[2];
    k->tytz = t[1] / t[2];
    t[0] = R_FMIN(k->limx, R_FMAX(-k->limx, k->txtz)) * t[2];
    t[1] = R_FMIN(k->limy, R_FMAX(-k->limy, k->tytz)) * t[2];
    k->tx = t[0]; k->ty = t[1]; k->tz = t[2];
    m3 J = m3_cols(fx / t[2], 0, -(fx * t[0]) / (t[2] * t[2]), 0, fy / t[2], -(fy * t[1]) / (t[2] * t[2]), 0, 0, 0);
    k->W = m3_cols(RC(view[0]), RC(view[4]), RC(view[8]), RC(view[1]), RC(view[5]), RC(view[9]), RC(view[2]),
                   RC(view[6]), RC(view[10]));
    k->T = m3_mul(&k->W, &J);
    k->Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
    m3 Tt = m3_t(&k->T), Vt = m3_t(&k->Vrk);
    m3 tmp = m3_mul(&Tt, &Vt);
    m3 cov = m3_mul(&tmp, &k->T);
    k->a = cov.c[0][0] + RC(0.3f); /* forward.cu:111-112: plain +0.3 dilation */
    k->b = cov.c[0][1];
    k->c = cov.c[1][1] + RC(0.3f);
}

/* forward.cu:20-71 */
static void sh_to_rgb(int deg, const float* sh /* [M][3] of this point */, real dx, real dy, real dz, real* rgb,
                      uint8_t* clamped) {
    real len = R_SQRT(dx * dx + dy * dy + dz * dz);
    real x = dx / len, y = dy / len, z = dz / len;
    for (int ch = 0; ch < 3; ch++) {
#define SHC(i) RC(sh[(i) * 3 + ch])
        real r = SH0 * SHC(0);
        if (deg > 0) {
            r = r - SH1 * y * SHC(1) + SH1 * z * SHC(2) - SH1 * x * SHC(3);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH2[0] * xy * SHC(4) + SH2[1] * yz * SHC(5) + SH2[2] * (RC(2) * zz - xx - yy) * SHC(6) +
                    SH2[3] * xz * SHC(7) + SH2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    r = r + SH3[0] * y * (RC(3) * xx - yy) * SHC(9) + SH3[1] * xy * z * SHC(10) +
                        SH3[2] * y * (RC(4) * zz - xx - yy) * SHC(11) +
                        SH3[3] * z * (RC(2) * zz - RC(3) * xx - RC(3) * yy) * SHC(12) +
                        SH3[4] * x * (RC(4) * zz - xx - yy) * SHC(13) + SH3[5] * z * (xx - yy) * SHC(14) +
                        SH3[6] * x * (xx - RC(3) * yy) * SHC(15);
                }
            }
        }
#undef SHC
        r += RC(0.5f);
        clamped[ch] = (r < 0);
        rgb[ch] = R_FMAX(r, RC(0));
    }
}

/* ------------------------------------------------------------------------------------------------
 * Per-Gaussian preprocess.  forward.cu:158-259.  Arrays are caller-allocated, sized by P.
 * Culled Gaussians get radii = tiles = 0 and (unlike the reference, which leaves them
 * uninitialised) zeros in every other output so that comparisons are well defined.
 * Returns 0, or -1 if `prefiltered` was set and a point failed the near-plane test (the reference traps).
 * ---------------------------------------------------------------------------------------------- */
int gso_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                   const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                   const float* colors_precomp, const float* view, const float* proj, const float* campos, int W,
                   int H, float tan_fovx, float tan_fovy, int prefiltered,
                   /* out */ int32_t* radii, real* xy, real* depths, real* cov3D, real* rgb, real* conic_opacity,
                   uint8_t* clamped, uint32_t* tiles_touched) {
    const real fy = (real)H / (RC(2.0f) * RC(tan_fovy)); /* rasterizer_impl.cu:222-223 */
    const real fx = (real)W / (RC(2.0f) * RC(tan_fovx));
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        xy[2 * i] = xy[2 * i + 1] = 0;
        depths[i] = 0;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0; clamped[3 * i + k] = 0; }
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0;

        real mx = RC(means3D[3 * i]), my = RC(means3D[3 * i + 1]), mz = RC(means3D[3 * i + 2]);
        real pv[3], ph[4];
        xform43(view, mx, my, mz, pv);
        if (pv[2] <= RC(0.2f)) { /* auxiliary.h:154 -- near plane only */
            if (prefiltered) bad |= 1;
            continue;
        }
        xform44(proj, mx, my, mz, ph);
        real pw = RC(1.0f) / (ph[3] + RC(0.0000001f));
        real projx = ph[0] * pw, projy = ph[1] * pw;

        real c6[6];
        if (cov3D_precomp) {
            for (int k = 0; k < 6; k++) c6[k] = RC(cov3D_precomp[6 * i + k]);
        } else {
            cov3d_from_scale_rot(scales + 3 * i, RC(scale_modifier), rotations + 4 * i, c6);
            for (int k = 0; k < 6; k++) cov3D[6 * i + k] = c6[k];
        }
        cov2d_ctx k2;
        cov2d_eval(mx, my, mz, fx, fy, RC(tan_fovx), RC(tan_fovy), c6, view, &k2);
        real det = k2.a * k2.c - k2.b * k2.b;
        if (det == RC(0.0f)) continue;
        real det_inv = RC(1.f) / det;
        real con[3] = {k2.c * det_inv, -k2.b * det_inv, k2.a * det_inv};
        real mid = RC(0.5f) * (k2.a + k2.c);
        real l1 = mid + R_SQRT(R_FMAX(RC(0.1f), mid * mid - det));
        real l2 = mid - R_SQRT(R_FMAX(RC(0.1f), mid * mid - det));
        real my_radius = R_CEIL(RC(3.f) * R_SQRT(R_FMAX(l1, l2)));
        real pix_x = ndc_to_pix(projx, W), pix_y = ndc_to_pix(projy, H);
        int x0, y0, x1, y1;
        tile_rect(pix_x, pix_y, sat_int(my_radius), gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        if (!colors_precomp) {
            sh_to_rgb(D, shs + (size_t)i * M * 3, mx - RC(campos[0]), my - RC(campos[1]), mz - RC(campos[2]),
                      rgb + 3 * i, clamped + 3 * i);
        }
        depths[i] = pv[2];
        radii[i] = sat_int(my_radius);
        xy[2 * i] = pix_x;
        xy[2 * i + 1] = pix_y;
        conic_opacity[4 * i + 0] = con[0];
        conic_opacity[4 * i + 1] = con[1];
        conic_opacity[4 * i + 2] = con[2];
        conic_opacity[4 * i + 3] = RC(opacities[i]);
        tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
    return bad ? -1 : 0;
}

/* rasterizer_impl.cu:277 -- inclusive prefix sum; returns the total (= num_rendered, :281) */
uint32_t gso_inclusive_scan(int P, const uint32_t* in, uint32_t* out) {
    uint32_t acc = 0;
    for (int i = 0; i < P; i++) { acc += in[i]; out[i] = acc; }
    return acc;
}

/* rasterizer_impl.cu:35-50 -- smallest b with (n >> b) == 0, found by the reference's bisection */
uint32_t gso_higher_msb(uint32_t n) {
    uint32_t msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* rasterizer_impl.cu:70-111.  The depth key is the *float32* bit pattern of the depth. */
void gso_duplicate_with_keys(int P, int W, int H, const real* xy, const real* depths, const uint32_t* offsets,
                             const int32_t* radii, uint64_t* keys, uint32_t* values) {
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t off = i == 0 ? 0 : offsets[i - 1];
        int x0, y0, x1, y1;
        tile_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        float d32 = (float)depths[i];
        uint32_t dbits;
        memcpy(&dbits, &d32, 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                values[off] = (uint32_t)i;
                off++;
            }
    }
}

/* rasterizer_impl.cu:300-308 -- stable ascending LSD radix sort on key bits [0, end_bit) */
void gso_sort_pairs(uint32_t R, int end_bit, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                    uint32_t* vals_out) {
    if (R == 0) return;
    uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * R);
    memcpy(keys_out, keys_in, sizeof(uint64_t) * R);
    memcpy(vals_out, vals_in, sizeof(uint32_t) * R);
    uint64_t *ka = keys_out, *kbb = kb;
    uint32_t *va = vals_out, *vbb = vb;
    const int RB = 11;
    size_t* hist = (size_t*)malloc(sizeof(size_t) * ((size_t)1 << RB));
    for (int shift = 0; shift < end_bit; shift += RB) {
        int bits = end_bit - shift < RB ? end_bit - shift : RB;
        uint64_t mask = ((uint64_t)1 << bits) - 1;
        memset(hist, 0, sizeof(size_t) * ((size_t)1 << RB));
        for (uint32_t i = 0; i < R; i++) hist[(ka[i] >> shift) & mask]++;
        size_t acc = 0;
        for (size_t d = 0; d <= mask; d++) { size_t c = hist[d]; hist[d] = acc; acc += c; }
        for (uint32_t i = 0; i < R; i++) {
            size_t p = hist[(ka[i] >> shift) & mask]++;
            kbb[p] = ka[i];
            vbb[p] = va[i];
        }
        uint64_t* tk = ka; ka = kbb; kbb = tk;
        uint32_t* tv = va; va = vbb; vbb = tv;
    }
    if (ka != keys_out) {
        memcpy(keys_out, ka, sizeof(uint64_t) * R);
        memcpy(vals_out, va, sizeof(uint32_t) * R);
    }
    free(hist); free(kb); free(vb);
}

/* rasterizer_impl.cu:116-138 + memset :310.  ranges = [Tn][2] */
void gso_tile_ranges(uint32_t R, int num_tiles, const uint64_t* keys, uint32_t* ranges) {
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)num_tiles);
    for (uint32_t i = 0; i < R; i++) {
        uint32_t cur = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = i; ranges[2 * cur] = i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = R;
    }
}

/* forward.cu:264-377.  One tile per work item; pixels of a tile are independent, so each pixel walks the
 * tile's list alone.  The reference's batch-of-256 early exit (:312-314) only skips work no surviving pixel
 * needs, so per-pixel results are identical. */
void gso_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* xy,
                       const real* colors, const real* conic_opacity, const float* bg, real* out_color,
                       real* final_T, uint32_t* n_contrib) {
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; t++) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        for (int ly = 0; ly < TILE_Y; ly++)
            for (int lx = 0; lx < TILE_X; lx++) {
                const int px = tx * TILE_X + lx, py = ty * TILE_Y + ly;
                if (px >= W || py >= H) continue;
                const real pfx = (real)px, pfy = (real)py; /* integer pixel coordinates, no +0.5 (:285) */
                real T = RC(1.0f), C[3] = {0, 0, 0};
                uint32_t contributor = 0, last = 0;
                for (uint32_t e = r0; e < r1; e++) {
                    contributor++;
                    const uint32_t g = point_list[e];
                    const real dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
                    const real* co = conic_opacity + 4 * (size_t)g;
                    const real power = RC(-0.5f) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > RC(0.0f)) continue;
                    const real alpha = R_FMIN(RC(0.99f), co[3] * R_EXP(power));
                    if (alpha < RC(1.0f) / RC(255.0f)) continue;
                    const real test_T = T * (1 - alpha);
                    if (test_T < RC(0.0001f)) break; /* this Gaussian is NOT blended (:343-347) */
                    for (int ch = 0; ch < 3; ch++) C[ch] += colors[3 * (size_t)g + ch] * alpha * T;
                    T = test_T;
                    last = contributor;
                }
                const size_t pid = (size_t)W * py + px;
                final_T[pid] = T;
                n_contrib[pid] = last;
                for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pid] = C[ch] + T * RC(bg[ch]);
            }
    }
}

static void atomic_add_real(real* p, real v) {
#pragma omp atomic
    *p += v;
}

/* backward.cu:399-557.  Gradients are accumulated with `+=` into caller-zeroed arrays:
 * dL_dmean2D [P][3] (z untouched), dL_dconic [P][4] (.x .y .w used, backward.cu:549-551),
 * dL_dopacity [P], dL_dcolors [P][3]. */
void gso_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* xy,
                        const real* colors, const real* conic_opacity, const float* bg, const real* final_T,
                        const uint32_t* n_contrib, const real* dL_dpix, real* dL_dmean2D, real* dL_dconic,
                        real* dL_dopacity, real* dL_dcolors) {
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    const real ddelx_dx = (real)(0.5 * W), ddely_dy = (real)(0.5 * H);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; t++) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        for (int ly = 0; ly < TILE_Y; ly++)
            for (int lx = 0; lx < TILE_X; lx++) {
                const int px = tx * TILE_X + lx, py = ty * TILE_Y + ly;
                if (px >= W || py >= H) continue;
                const size_t pid = (size_t)W * py + px;
                const real pfx = (real)px, pfy = (real)py;
                const real T_final = final_T[pid];
                real T = T_final;
                uint32_t contributor = r1 - r0;
                const uint32_t last_contributor = n_contrib[pid];
                real accum[3] = {0, 0, 0}, dpix[3], last_color[3] = {0, 0, 0}, last_alpha = 0;
                for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[(size_t)ch * H * W + pid];
                for (uint32_t e = r1; e-- > r0;) { /* back to front */
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t g = point_list[e];
                    const real dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
                    const real* co = conic_opacity + 4 * (size_t)g;
                    const real power = RC(-0.5f) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > RC(0.0f)) continue;
                    const real G = R_EXP(power);
                    const real alpha = R_FMIN(RC(0.99f), co[3] * G);
                    if (alpha < RC(1.0f) / RC(255.0f)) continue;
                    T = T / (RC(1.f) - alpha);
                    const real dchannel_dcolor = alpha * T;
                    real dL_dalpha = 0;
                    for (int ch = 0; ch < 3; ch++) {
                        const real c = colors[3 * (size_t)g + ch];
                        accum[ch] = last_alpha * last_color[ch] + (RC(1.f) - last_alpha) * accum[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dpix[ch];
                        atomic_add_real(&dL_dcolors[3 * (size_t)g + ch], dchannel_dcolor * dpix[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    real bg_dot = 0;
                    for (int ch = 0; ch < 3; ch++) bg_dot += RC(bg[ch]) * dpix[ch];
                    dL_dalpha += (-T_final / (RC(1.f) - alpha)) * bg_dot;
                    const real dL_dG = co[3] * dL_dalpha;
                    const real gdx = G * dx, gdy = G * dy;
                    const real dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const real dG_ddely = -gdy * co[2] - gdx * co[1];
                    atomic_add_real(&dL_dmean2D[3 * (size_t)g + 0], dL_dG * dG_ddelx * ddelx_dx);
                    atomic_add_real(&dL_dmean2D[3 * (size_t)g + 1], dL_dG * dG_ddely * ddely_dy);
                    atomic_add_real(&dL_dconic[4 * (size_t)g + 0], RC(-0.5f) * gdx * dx * dL_dG);
                    atomic_add_real(&dL_dconic[4 * (size_t)g + 1], RC(-0.5f) * gdx * dy * dL_dG);
                    atomic_add_real(&dL_dconic[4 * (size_t)g + 3], RC(-0.5f) * gdy * dy * dL_dG);
                    atomic_add_real(&dL_dopacity[g], G * dL_dalpha);
                }
            }
    }
}

/* backward.cu:20-139 */
static void sh_backward(int deg, int M, const float* sh, real dirx, real diry, real dirz, const uint8_t* clamped,
                        const real* dL_dcolor, real* dL_dmean /* += */, real* dL_dsh /* [M][3] = */) {
    (void)M;
    real len = R_SQRT(dirx * dirx + diry * diry + dirz * dirz);
    real x = dirx / len, y = diry / len, z = dirz / len;
    real dRGB[3];
    for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[ch] * (clamped[ch] ? RC(0) : RC(1));
    real ddx[3] = {0, 0, 0}, ddy[3] = {0, 0, 0}, ddz[3] = {0, 0, 0};
#define SHV(i, ch) RC(sh[(i) * 3 + (ch)])
#define DSH(i, w) for (int ch = 0; ch < 3; ch++) dL_dsh[(i) * 3 + ch] = (w) * dRGB[ch]
    DSH(0, SH0);
    if (deg > 0) {
        DSH(1, -SH1 * y); DSH(2, SH1 * z); DSH(3, -SH1 * x);
        for (int ch = 0; ch < 3; ch++) {
            ddx[ch] = -SH1 * SHV(3, ch);
            ddy[ch] = -SH1 * SHV(1, ch);
            ddz[ch] = SH1 * SHV(2, ch);
        }
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, SH2[0] * xy); DSH(5, SH2[1] * yz); DSH(6, SH2[2] * (RC(2.f) * zz - xx - yy));
            DSH(7, SH2[3] * xz); DSH(8, SH2[4] * (xx - yy));
            for (int ch = 0; ch < 3; ch++) {
                ddx[ch] += SH2[0] * y * SHV(4, ch) + SH2[2] * RC(2.f) * -x * SHV(6, ch) + SH2[3] * z * SHV(7, ch) +
                           SH2[4] * RC(2.f) * x * SHV(8, ch);
                ddy[ch] += SH2[0] * x * SHV(4, ch) + SH2[1] * z * SHV(5, ch) + SH2[2] * RC(2.f) * -y * SHV(6, ch) +
                           SH2[4] * RC(2.f) * -y * SHV(8, ch);
                ddz[ch] += SH2[1] * y * SHV(5, ch) + SH2[2] * RC(2.f) * RC(2.f) * z * SHV(6, ch) +
                           SH2[3] * x * SHV(7, ch);
            }
            if (deg > 2) {
                DSH(9, SH3[0] * y * (RC(3.f) * xx - yy)); DSH(10, SH3[1] * xy * z);
                DSH(11, SH3[2] * y * (RC(4.f) * zz - xx - yy));
                DSH(12, SH3[3] * z * (RC(2.f) * zz - RC(3.f) * xx - RC(3.f) * yy));
                DSH(13, SH3[4] * x * (RC(4.f) * zz - xx - yy)); DSH(14, SH3[5] * z * (xx - yy));
                DSH(15, SH3[6] * x * (xx - RC(3.f) * yy));
                for (int ch = 0; ch < 3; ch++) {
                    ddx[ch] += (SH3[0] * SHV(9, ch) * RC(3.f) * RC(2.f) * xy + SH3[1] * SHV(10, ch) * yz +
                                SH3[2] * SHV(11, ch) * RC(-2.f) * xy + SH3[3] * SHV(12, ch) * RC(-3.f) * RC(2.f) * xz +
                                SH3[4] * SHV(13, ch) * (RC(-3.f) * xx + RC(4.f) * zz - yy) +
                                SH3[5] * SHV(14, ch) * RC(2.f) * xz + SH3[6] * SHV(15, ch) * RC(3.f) * (xx - yy));
                    ddy[ch] += (SH3[0] * SHV(9, ch) * RC(3.f) * (xx - yy) + SH3[1] * SHV(10, ch) * xz +
                                SH3[2] * SHV(11, ch) * (RC(-3.f) * yy + RC(4.f) * zz - xx) +
                                SH3[3] * SHV(12, ch) * RC(-3.f) * RC(2.f) * yz + SH3[4] * SHV(13, ch) * RC(-2.f) * xy +
                                SH3[5] * SHV(14, ch) * RC(-2.f) * yz + SH3[6] * SHV(15, ch) * RC(-3.f) * RC(2.f) * xy);
                    ddz[ch] += (SH3[1] * SHV(10, ch) * xy + SH3[2] * SHV(11, ch) * RC(4.f) * RC(2.f) * yz +
                                SH3[3] * SHV(12, ch) * RC(3.f) * (RC(2.f) * zz - xx - yy) +
                                SH3[4] * SHV(13, ch) * RC(4.f) * RC(2.f) * xz + SH3[5] * SHV(14, ch) * (xx - yy));
                }
            }
        }
    }
#undef SHV
#undef DSH
    real dLdx = ddx[0] * dRGB[0] + ddx[1] * dRGB[1] + ddx[2] * dRGB[2];
    real dLdy = ddy[0] * dRGB[0] + ddy[1] * dRGB[1] + ddy[2] * dRGB[2];
    real dLdz = ddz[0] * dRGB[0] + ddz[1] * dRGB[1] + ddz[2] * dRGB[2];
    /* auxiliary.h:108-119 gradient through v/|v| */
    real s2 = dirx * dirx + diry * diry + dirz * dirz;
    real inv = RC(1.0f) / R_SQRT(s2 * s2 * s2);
    dL_dmean[0] += ((+s2 - dirx * dirx) * dLdx - diry * dirx * dLdy - dirz * dirx * dLdz) * inv;
    dL_dmean[1] += (-dirx * diry * dLdx + (s2 - diry * diry) * dLdy - dirz * diry * dLdz) * inv;
    dL_dmean[2] += (-dirx * dirz * dLdx - diry * dirz * dLdy + (s2 - dirz * dirz) * dLdz) * inv;
}

/* backward.cu:278-341 */
static void cov3d_backward(const float* s_in, real mod, const float* q, const real* dc /* dL_dcov3D[6] */,
                           real* dL_dscale, real* dL_drot) {
    real r = RC(q[0]), x = RC(q[1]), y = RC(q[2]), z = RC(q[3]);
    m3 Rm = m3_cols(RC(1) - RC(2) * (y * y + z * z), RC(2) * (x * y - r * z), RC(2) * (x * z + r * y),
                    RC(2) * (x * y + r * z), RC(1) - RC(2) * (x * x + z * z), RC(2) * (y * z - r * x),
                    RC(2) * (x * z - r * y), RC(2) * (y * z + r * x), RC(1) - RC(2) * (x * x + y * y));
    real s[3] = {mod * RC(s_in[0]), mod * RC(s_in[1]), mod * RC(s_in[2])};
    m3 S = m3_cols(s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]);
    m3 M = m3_mul(&S, &Rm);
    m3 dSig = m3_cols(dc[0], RC(0.5f) * dc[1], RC(0.5f) * dc[2], RC(0.5f) * dc[1], dc[3], RC(0.5f) * dc[4],
                      RC(0.5f) * dc[2], RC(0.5f) * dc[4], dc[5]);
    m3 M2; /* 2.0f * M */
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) M2.c[j][i] = RC(2.0f) * M.c[j][i];
    m3 dM = m3_mul(&M2, &dSig);
    m3 Rt = m3_t(&Rm), dMt = m3_t(&dM);
    for (int k = 0; k < 3; k++)
        dL_dscale[k] = Rt.c[k][0] * dMt.c[k][0] + Rt.c[k][1] * dMt.c[k][1] + Rt.c[k][2] * dMt.c[k][2];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) dMt.c[k][i] *= s[k];
#define D(a, b) dMt.c[a][b]
    dL_drot[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
    dL_drot[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) -
                 4 * x * (D(2, 2) + D(1, 1));
    dL_drot[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) -
                 4 * y * (D(2, 2) + D(0, 0));
    dL_drot[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) -
                 4 * z * (D(1, 1) + D(0, 0));
#undef D
    /* no quaternion-normalisation backward (backward.cu:340) */
}

/* backward.cu:144-274 followed by :346-396, per visible Gaussian (radii > 0).
 * In:  dL_dmean2D [P][3], dL_dconic [P][4], dL_dcolor [P][3] (from gso_blend_backward, or the caller's
 *      colour gradient when colours are precomputed), cov3D = cov3D_precomp or the forward's cov3D.
 * Out (caller-zeroed): dL_dmean3D [P][3], dL_dcov3D [P][6], dL_dsh [P][M][3], dL_dscale [P][3], dL_drot [P][4]. */
void gso_preprocess_backward(int P, int D, int M, const float* means3D, const int32_t* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations,
                             float scale_modifier, const real* cov3D, const float* view, const float* proj,
                             const float* campos, int W, int H, float tan_fovx, float tan_fovy,
                             const real* dL_dmean2D, const real* dL_dconic, const real* dL_dcolor,
                             real* dL_dmean3D, real* dL_dcov3D, real* dL_dsh, real* dL_dscale, real* dL_drot) {
    const real h_y = (real)H / (RC(2.0f) * RC(tan_fovy));
    const real h_x = (real)W / (RC(2.0f) * RC(tan_fovx));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        const real mx = RC(means3D[3 * i]), my = RC(means3D[3 * i + 1]), mz = RC(means3D[3 * i + 2]);
        /* ---- computeCov2DCUDA ---- */
        cov2d_ctx k;
        cov2d_eval(mx, my, mz, h_x, h_y, RC(tan_fovx), RC(tan_fovy), cov3D + 6 * (size_t)i, view, &k);
        const real dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
        const real xg = (k.txtz < -k.limx || k.txtz > k.limx) ? RC(0) : RC(1);
        const real yg = (k.tytz < -k.limy || k.tytz > k.limy) ? RC(0) : RC(1);
        const real a = k.a, b = k.b, c = k.c;
        const real denom = a * c - b * b;
        real dL_da = 0, dL_db = 0, dL_dc = 0;
        const real denom2inv = RC(1.0f) / ((denom * denom) + RC(0.0000001f));
        real* dcov = dL_dcov3D + 6 * (size_t)i;
#define TT(col, row) k.T.c[col][row]
#define VV(col, row) k.Vrk.c[col][row]
#define WW(col, row) k.W.c[col][row]
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
            dcov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
            dcov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
            dcov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db +
                      2 * TT(1, 0) * TT(1, 1) * dL_dc;
            dcov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db +
                      2 * TT(1, 0) * TT(1, 2) * dL_dc;
            dcov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db +
                      2 * TT(1, 1) * TT(1, 2) * dL_dc;
        } else {
            for (int q = 0; q < 6; q++) dcov[q] = 0;
        }
        real dT[2][3];
        for (int col = 0; col < 3; col++) {
            dT[0][col] = 2 * (TT(0, 0) * VV(col, 0) + TT(0, 1) * VV(col, 1) + TT(0, 2) * VV(col, 2)) * dL_da +
                         (TT(1, 0) * VV(col, 0) + TT(1, 1) * VV(col, 1) + TT(1, 2) * VV(col, 2)) * dL_db;
            dT[1][col] = 2 * (TT(1, 0) * VV(col, 0) + TT(1, 1) * VV(col, 1) + TT(1, 2) * VV(col, 2)) * dL_dc +
                         (TT(0, 0) * VV(col, 0) + TT(0, 1) * VV(col, 1) + TT(0, 2) * VV(col, 2)) * dL_db;
        }
        const real dJ00 = WW(0, 0) * dT[0][0] + WW(0, 1) * dT[0][1] + WW(0, 2) * dT[0][2];
        const real dJ02 = WW(2, 0) * dT[0][0] + WW(2, 1) * dT[0][1] + WW(2, 2) * dT[0][2];
        const real dJ11 = WW(1, 0) * dT[1][0] + WW(1, 1) * dT[1][1] + WW(1, 2) * dT[1][2];
        const real dJ12 = WW(2, 0) * dT[1][0] + WW(2, 1) * dT[1][1] + WW(2, 2) * dT[1][2];
#undef TT
#undef VV
#undef WW
        const real tz = RC(1.f) / k.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const real dtx = xg * -h_x * tz2 * dJ02;
        const real dty = yg * -h_y * tz2 * dJ12;
        const real dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * k.tx) * tz3 * dJ02 +
                         (2 * h_y * k.ty) * tz3 * dJ12;
        real* dm = dL_dmean3D + 3 * (size_t)i;
        /* transformVec4x3Transpose, assignment (backward.cu:273) */
        dm[0] = RC(view[0]) * dtx + RC(view[1]) * dty + RC(view[2]) * dtz;
        dm[1] = RC(view[4]) * dtx + RC(view[5]) * dty + RC(view[6]) * dtz;
        dm[2] = RC(view[8]) * dtx + RC(view[9]) * dty + RC(view[10]) * dtz;

        /* ---- preprocessCUDA (backward) ---- */
        real mh[4];
        xform44(proj, mx, my, mz, mh);
        const real m_w = RC(1.0f) / (mh[3] + RC(0.0000001f));
        const real mul1 = (RC(proj[0]) * mx + RC(proj[4]) * my + RC(proj[8]) * mz + RC(proj[12])) * m_w * m_w;
        const real mul2 = (RC(proj[1]) * mx + RC(proj[5]) * my + RC(proj[9]) * mz + RC(proj[13])) * m_w * m_w;
        const real g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
        dm[0] += (RC(proj[0]) * m_w - RC(proj[3]) * mul1) * g2x + (RC(proj[1]) * m_w - RC(proj[3]) * mul2) * g2y;
        dm[1] += (RC(proj[4]) * m_w - RC(proj[7]) * mul1) * g2x + (RC(proj[5]) * m_w - RC(proj[7]) * mul2) * g2y;
        dm[2] += (RC(proj[8]) * m_w - RC(proj[11]) * mul1) * g2x + (RC(proj[9]) * m_w - RC(proj[11]) * mul2) * g2y;
        if (shs)
            sh_backward(D, M, shs + (size_t)i * M * 3, mx - RC(campos[0]), my - RC(campos[1]), mz - RC(campos[2]),
                        clamped + 3 * i, dL_dcolor + 3 * (size_t)i, dm, dL_dsh + (size_t)i * M * 3);
        if (scales)
            cov3d_backward(scales + 3 * i, RC(scale_modifier), rotations + 4 * i, dcov, dL_dscale + 3 * (size_t)i,
                           dL_drot + 4 * (size_t)i);
    }
}

/* rasterizer_impl.cu:54-66 -- near-plane test only */
void gso_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present) {
    (void)proj;
    for (int i = 0; i < P; i++) {
        real pv[3];
        xform43(view, RC(means3D[3 * i]), RC(means3D[3 * i + 1]), RC(means3D[3 * i + 2]), pv);
        present[i] = pv[2] > RC(0.2f);
    }
}
