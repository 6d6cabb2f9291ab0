// integration/rasterize_points_b200.cpp -- the reference-side binding a maintainer would compile: the three entry points
// of the reference's pybind module (dgr/rasterize_points.h:19-67, registered in dgr/ext.cpp:15-19 as
// rasterize_gaussians / rasterize_gaussians_backward / mark_visible, same argument order, same return tuples) implemented
// on the C ABI of include/gsplat_b200.h instead of CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (rasterize_points.cu:89-113,163-193,206-213).  Nothing in here is product code: the product's own binding is the ctypes
// module diff_gaussian_rasterization/_C.py; this file exists to show -- and to test on the GPU (tests/test_gpu.py:
// test_compiled_reference_side_binding) -- that the library drops in behind the reference's compiled extension.
// Build: python integration/build.py  (torch.utils.cpp_extension, links -lgsplat_b200).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <tuple>

#include "gsplat_b200.h"

namespace {

// GsBuffer callback == the reference's resizeFunctional (rasterize_points.cu:27-33): grow the byte tensor, hand back
// its device pointer
char* resize_cb(void* user, size_t bytes) {
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({static_cast<long long>(bytes)});
    return reinterpret_cast<char*>(t->data_ptr());
}

// "not provided" arrives as an empty tensor (the Python layer passes torch.Tensor([])) -> NULL for the library
struct Arg {
    torch::Tensor keep;  // contiguous fp32 copy (or the tensor itself), alive until the call returns
    explicit Arg(const torch::Tensor& t) : keep(t.numel() ? t.contiguous() : t) {}
    const float* ptr() const { return keep.numel() ? keep.data_ptr<float>() : nullptr; }
};

void fill_scene(GsScene& s, int P, int degree, int M, int W, int H, float tan_fovx, float tan_fovy, float scale_modifier,
                bool prefiltered, bool debug, const Arg& background, const Arg& means3D, const Arg& sh, const Arg& colors,
                const Arg& opacity, const Arg& scales, const Arg& rotations, const Arg& cov3D_precomp,
                const Arg& viewmatrix, const Arg& projmatrix, const Arg& campos) {
    s.P = P; s.sh_degree = degree; s.sh_stride = M; s.width = W; s.height = H;
    s.tan_fovx = tan_fovx; s.tan_fovy = tan_fovy; s.scale_modifier = scale_modifier;
    s.prefiltered = prefiltered ? 1 : 0; s.debug = debug ? 1 : 0;  // tile_row_begin / end = 0, 0: the whole frame
    s.background = background.ptr(); s.means3D = means3D.ptr(); s.shs = sh.ptr(); s.colors_precomp = colors.ptr();
    s.opacities = opacity.ptr(); s.scales = scales.ptr(); s.rotations = rotations.ptr();
    s.cov3D_precomp = cov3D_precomp.ptr(); s.viewmatrix = viewmatrix.ptr(); s.projmatrix = projmatrix.ptr();
    s.campos = campos.ptr();
}

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizeGaussiansCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
    const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier,
    const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
    const float tan_fovx, const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
    const int degree, const torch::Tensor& campos, const bool prefiltered, const bool debug) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    const int P = static_cast<int>(means3D.size(0)), H = image_height, W = image_width;
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto f32 = means3D.options().dtype(torch::kFloat32);
    torch::Tensor out_color = torch::full({3, H, W}, 0.0, f32);
    torch::Tensor radii = torch::full({P}, 0, means3D.options().dtype(torch::kInt32));
    const auto bytes = torch::TensorOptions().dtype(torch::kByte).device(means3D.device());
    torch::Tensor geomBuffer = torch::empty({0}, bytes), binningBuffer = torch::empty({0}, bytes),
                  imgBuffer = torch::empty({0}, bytes);
    int rendered = 0;
    if (P != 0) {
        const int M = sh.size(0) != 0 ? static_cast<int>(sh.size(1)) : 0;
        const Arg a_bg(background), a_m(means3D), a_sh(sh), a_col(colors), a_op(opacity), a_sc(scales), a_rot(rotations),
            a_cov(cov3D_precomp), a_v(viewmatrix), a_p(projmatrix), a_cam(campos);
        GsScene s{};
        fill_scene(s, P, degree, M, W, H, tan_fovx, tan_fovy, scale_modifier, prefiltered, debug, a_bg, a_m, a_sh, a_col,
                   a_op, a_sc, a_rot, a_cov, a_v, a_p, a_cam);
        const GsBuffer g{resize_cb, &geomBuffer}, b{resize_cb, &binningBuffer}, i{resize_cb, &imgBuffer};
        const int64_t rc = gs_forward(&s, g, b, i, out_color.data_ptr<float>(), radii.data_ptr<int>(),
                                      at::cuda::getCurrentCUDAStream().stream());
        if (rc < 0) AT_ERROR("gs_forward failed (", rc, "): ", gs_last_error());
        rendered = static_cast<int>(rc);
    }
    return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                               const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& sh,
                               const int degree, const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug) {
    const int P = static_cast<int>(means3D.size(0));
    const int H = static_cast<int>(dL_dout_color.size(1)), W = static_cast<int>(dL_dout_color.size(2));
    const int M = sh.size(0) != 0 ? static_cast<int>(sh.size(1)) : 0;
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto o = means3D.options();
    torch::Tensor dL_dmeans3D = torch::zeros({P, 3}, o), dL_dmeans2D = torch::zeros({P, 3}, o),
                  dL_dcolors = torch::zeros({P, 3}, o), dL_dconic = torch::zeros({P, 2, 2}, o),
                  dL_dopacity = torch::zeros({P, 1}, o), dL_dcov3D = torch::zeros({P, 6}, o),
                  dL_dsh = torch::zeros({P, M, 3}, o), dL_dscales = torch::zeros({P, 3}, o),
                  dL_drotations = torch::zeros({P, 4}, o);
    if (P != 0) {
        // (scales / rotations made contiguous here too: the reference passes their raw data_ptr, a latent out-of-bounds
        // read for the stride-0 expand() of Simple_Render, simple_raw_render.py:715-717)
        const Arg a_bg(background), a_m(means3D), a_sh(sh), a_col(colors), a_none{torch::Tensor()}, a_sc(scales),
            a_rot(rotations), a_cov(cov3D_precomp), a_v(viewmatrix), a_p(projmatrix), a_cam(campos);
        GsScene s{};
        fill_scene(s, P, degree, M, W, H, tan_fovx, tan_fovy, scale_modifier, false, debug, a_bg, a_m, a_sh, a_col, a_none,
                   a_sc, a_rot, a_cov, a_v, a_p, a_cam);
        const torch::Tensor rad = radii.contiguous(), dpix = dL_dout_color.contiguous(), gb = geomBuffer.contiguous(),
                            bb = binningBuffer.contiguous(), ib = imageBuffer.contiguous();
        const int32_t rc = gs_backward(&s, R, rad.data_ptr<int>(), reinterpret_cast<const char*>(gb.data_ptr()),
                                       reinterpret_cast<const char*>(bb.data_ptr()),
                                       reinterpret_cast<const char*>(ib.data_ptr()), dpix.data_ptr<float>(),
                                       dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(),
                                       dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
                                       dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(),
                                       dL_dsh.data_ptr<float>(), dL_dscales.data_ptr<float>(),
                                       dL_drotations.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream());
        if (rc < 0) AT_ERROR("gs_backward failed (", rc, "): ", gs_last_error());
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                           dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    const int P = static_cast<int>(means3D.size(0));
    const c10::cuda::CUDAGuard guard(means3D.device());
    torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
    if (P != 0) {
        const torch::Tensor m = means3D.contiguous(), v = viewmatrix.contiguous(), p = projmatrix.contiguous();
        const int32_t rc = gs_mark_visible(P, m.data_ptr<float>(), v.data_ptr<float>(), p.data_ptr<float>(),
                                           reinterpret_cast<uint8_t*>(present.data_ptr<bool>()),
                                           at::cuda::getCurrentCUDAStream().stream());
        if (rc < 0) AT_ERROR("gs_mark_visible failed (", rc, "): ", gs_last_error());
    }
    return present;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {  // the names of dgr/ext.cpp:15-19
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
}
