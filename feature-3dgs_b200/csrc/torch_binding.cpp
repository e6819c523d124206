// torch / pybind11 surface `diff_gaussian_rasterization._C` over the C ABI of libf3dgs_b200.so.
//
// Same three functions, positional signatures and return tuples as the reference module
// (reference ext.cpp:15-19, rasterize_points.h:18-72, rasterize_points.cu:35-236), so the
// reference's Python wrapper logic calls it unchanged:
//   rasterize_gaussians(...)          -> (num_rendered, color, feature_map, depth, radii, geom, binning, img)
//   rasterize_gaussians_backward(...) -> (dL_dmeans2D, dL_dcolors, dL_dsemantic_feature, dL_dopacity,
//                                         dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
//   mark_visible(means3D, viewmatrix, projmatrix) -> bool[P]
// Differences (all permissive): the feature width is read from semantic_feature.size(-1) at run
// time (reference: compile-time NUM_SEMANTIC_CHANNELS, config.h:16); an empty / undefined
// semantic_feature means C = 0; inputs are checked for device/dtype and errors from the C ABI
// are raised as RuntimeError.  This file contains no CUDA code and there is no CPU fallback.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <string>
#include <tuple>
#include <vector>

#include "../../include/f3dgs_b200.h"

namespace {

// reference rasterize_points.cu:27-33 (resizeFunctional): grow a uint8 CUDA tensor on demand
char* resize_tensor(void* ctx, size_t bytes) {
    auto* t = static_cast<torch::Tensor*>(ctx);
    t->resize_({(long long)bytes});
    return reinterpret_cast<char*>(t->data_ptr());
}

const float* fptr(const torch::Tensor& t) {  // empty tensor -> nullptr, as in the reference
    if (!t.defined() || t.numel() == 0) return nullptr;
    return t.data_ptr<float>();
}

torch::Tensor prep(const torch::Tensor& t, const torch::Device& dev, const char* name) {
    if (!t.defined() || t.numel() == 0) return t;
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
    TORCH_CHECK(t.device() == dev, name, " must be on ", dev, " (got ", t.device(), ")");
    return t.contiguous();
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc >= 0, what, " failed (code ", -rc, "): ", f3dgs_last_error());
}

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& semantic_feature, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier,
                       const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered, const bool debug) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
        AT_ERROR("means3D must have dimensions (num_points, 3)");
    }
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto dev = means3D.device();
    const int P = means3D.size(0);
    const int H = image_height, W = image_width;
    // feature width = last dimension (also for an empty cloud [0,1,C]); an absent feature input is a 0-element 1-D tensor
    const int C = (semantic_feature.defined() && semantic_feature.dim() >= 1) ? (int)semantic_feature.size(-1) : 0;
    TORCH_CHECK(!semantic_feature.defined() || semantic_feature.numel() == (int64_t)P * C,
                "semantic_feature must be [P, 1, C]");

    auto float_opts = means3D.options().dtype(torch::kFloat32);
    // every element of the outputs is written by the composite kernel, so no zero-fill pass
    // (reference: torch::full 0, rasterize_points.cu:67-73); P == 0 keeps the reference's zeros
    torch::Tensor out_color = P ? torch::empty({3, H, W}, float_opts) : torch::zeros({3, H, W}, float_opts);
    torch::Tensor out_depth = P ? torch::empty({1, H, W}, float_opts) : torch::zeros({1, H, W}, float_opts);
    torch::Tensor out_feature = P ? torch::empty({C, H, W}, float_opts) : torch::zeros({C, H, W}, float_opts);
    torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));

    auto byte_opts = torch::TensorOptions().dtype(torch::kByte).device(dev);
    torch::Tensor geomBuffer = torch::empty({0}, byte_opts);
    torch::Tensor binningBuffer = torch::empty({0}, byte_opts);
    torch::Tensor imgBuffer = torch::empty({0}, byte_opts);

    int rendered = 0;
    if (P != 0) {
        int M = 0;
        if (sh.defined() && sh.numel() != 0) M = sh.size(1);
        auto bg = prep(background, dev, "bg"), m3 = prep(means3D, dev, "means3D");
        auto col = prep(colors, dev, "colors_precomp"), sf = prep(semantic_feature, dev, "semantic_feature");
        auto op = prep(opacity, dev, "opacities"), sc = prep(scales, dev, "scales");
        auto rot = prep(rotations, dev, "rotations"), cov = prep(cov3D_precomp, dev, "cov3D_precomp");
        auto vm = prep(viewmatrix, dev, "viewmatrix"), pm = prep(projmatrix, dev, "projmatrix");
        auto shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos");
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
        rendered = f3dgs_forward(resize_tensor, &geomBuffer, resize_tensor, &binningBuffer, resize_tensor, &imgBuffer,
                                 P, degree, M, C, fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(sf), fptr(op),
                                 fptr(sc), scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp),
                                 tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(),
                                 C ? out_feature.data_ptr<float>() : nullptr, out_depth.data_ptr<float>(),
                                 radii.data_ptr<int>(), debug ? 1 : 0, (void*)stream);
        check_rc(rendered, "f3dgs_forward");
    }
    return std::make_tuple(rendered, out_color, out_feature, out_depth, radii, geomBuffer, binningBuffer, imgBuffer);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                               const torch::Tensor& radii, const torch::Tensor& colors,
                               const torch::Tensor& semantic_feature, const torch::Tensor& scales,
                               const torch::Tensor& rotations, const float scale_modifier,
                               const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_feature,
                               const torch::Tensor& dL_dout_depth, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                               const bool debug) {
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto dev = means3D.device();
    const int P = means3D.size(0);
    const int H = dL_dout_color.size(1);
    const int W = dL_dout_color.size(2);
    int M = 0;
    if (sh.defined() && sh.numel() != 0) M = sh.size(1);
    const int C = (semantic_feature.defined() && semantic_feature.dim() >= 1) ? (int)semantic_feature.size(-1) : 0;
    const int64_t mid = (semantic_feature.defined() && semantic_feature.dim() == 3) ? semantic_feature.size(1) : 1;

    auto o = means3D.options().dtype(torch::kFloat32);
    torch::Tensor dL_dmeans3D = torch::zeros({P, 3}, o);
    torch::Tensor dL_dmeans2D = torch::zeros({P, 3}, o);
    torch::Tensor dL_dcolors = torch::zeros({P, 3}, o);
    torch::Tensor dL_dsemantic_feature = torch::zeros({P, mid, C}, o);
    torch::Tensor dL_dconic = torch::zeros({P, 2, 2}, o);
    torch::Tensor dL_dopacity = torch::zeros({P, 1}, o);
    torch::Tensor dL_dcov3D = torch::zeros({P, 6}, o);
    torch::Tensor dL_dsh = torch::zeros({P, M, 3}, o);
    torch::Tensor dL_dscales = torch::zeros({P, 3}, o);
    torch::Tensor dL_drotations = torch::zeros({P, 4}, o);
    torch::Tensor dL_dz = torch::zeros({P, 1}, o);

    if (P != 0) {
        auto bg = prep(background, dev, "bg"), m3 = prep(means3D, dev, "means3D");
        auto col = prep(colors, dev, "colors_precomp"), sf = prep(semantic_feature, dev, "semantic_feature");
        auto sc = prep(scales, dev, "scales"), rot = prep(rotations, dev, "rotations");
        auto cov = prep(cov3D_precomp, dev, "cov3D_precomp");
        auto vm = prep(viewmatrix, dev, "viewmatrix"), pm = prep(projmatrix, dev, "projmatrix");
        auto shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos");
        auto gc = prep(dL_dout_color, dev, "grad_out_color");
        auto gd = prep(dL_dout_depth, dev, "grad_out_depth");
        torch::Tensor gf = C ? prep(dL_dout_feature, dev, "grad_out_feature") : dL_dout_feature;
        TORCH_CHECK(radii.scalar_type() == torch::kInt32 && radii.is_cuda(), "radii must be int32 CUDA");
        auto rad = radii.contiguous();
        auto gb = geomBuffer.contiguous(), bb = binningBuffer.contiguous(), ib = imageBuffer.contiguous();
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
        int rc = f3dgs_backward(
            P, degree, M, R, C, fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(sf), fptr(sc), scale_modifier,
            fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy, rad.data_ptr<int>(),
            reinterpret_cast<char*>(gb.data_ptr()), reinterpret_cast<char*>(bb.data_ptr()),
            reinterpret_cast<char*>(ib.data_ptr()), fptr(gc), C ? fptr(gf) : nullptr, fptr(gd),
            dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(),
            dL_dcolors.data_ptr<float>(), C ? dL_dsemantic_feature.data_ptr<float>() : nullptr,
            dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), M ? dL_dsh.data_ptr<float>() : nullptr,
            dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>(), dL_dz.data_ptr<float>(), debug ? 1 : 0,
            (void*)stream);
        check_rc(rc, "f3dgs_backward");
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dsemantic_feature, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
                           dL_dscales, dL_drotations);
}

// Accumulating backward for view batches (additive to the reference module): gradients are ADDED into the tensors the
// caller passes (typically views of one flat gradient buffer, see diff_gaussian_rasterization/parallel.py); nothing is
// allocated besides one cached scratch tensor per device.  Undefined / empty tensors stand for "not an input".
void RasterizeGaussiansBackwardAccumCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier,
    const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
    const float tan_fovx, const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_feature,
    const torch::Tensor& dL_dout_depth, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
    const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
    torch::Tensor scratch, torch::Tensor g_means3D, torch::Tensor g_sh, torch::Tensor g_colors,
    torch::Tensor g_semantic_feature, torch::Tensor g_opacities, torch::Tensor g_scales, torch::Tensor g_rotations,
    torch::Tensor g_cov3D, torch::Tensor g_means2D_out, torch::Tensor grad_accum, torch::Tensor denom,
    const int64_t composite_done_event, const bool debug) {
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto dev = means3D.device();
    const int P = means3D.size(0);
    if (P == 0) return;
    const int H = dL_dout_color.size(1), W = dL_dout_color.size(2);
    int M = 0;
    if (sh.defined() && sh.numel() != 0) M = sh.size(1);
    const int C = (dL_dout_feature.defined() && dL_dout_feature.dim() == 3) ? (int)dL_dout_feature.size(0) : 0;
    auto gcheck = [&](const torch::Tensor& t, int64_t numel, const char* name) -> float* {
        if (!t.defined() || t.numel() == 0) return nullptr;
        TORCH_CHECK(t.is_cuda() && t.device() == dev && t.scalar_type() == torch::kFloat32 && t.is_contiguous() &&
                        t.numel() == numel, name, " must be a contiguous float32 CUDA tensor with ", numel, " elements");
        return t.data_ptr<float>();
    };
    auto bg = prep(background, dev, "bg"), m3 = prep(means3D, dev, "means3D");
    auto col = prep(colors, dev, "colors_precomp");
    auto sc = prep(scales, dev, "scales"), rot = prep(rotations, dev, "rotations");
    auto cov = prep(cov3D_precomp, dev, "cov3D_precomp");
    auto vm = prep(viewmatrix, dev, "viewmatrix"), pm = prep(projmatrix, dev, "projmatrix");
    auto shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos");
    auto gc = prep(dL_dout_color, dev, "grad_out_color"), gd = prep(dL_dout_depth, dev, "grad_out_depth");
    torch::Tensor gf = C ? prep(dL_dout_feature, dev, "grad_out_feature") : dL_dout_feature;
    TORCH_CHECK(radii.scalar_type() == torch::kInt32 && radii.is_cuda(), "radii must be int32 CUDA");
    auto rad = radii.contiguous();
    const size_t need = f3dgs_backward_scratch_bytes(P);
    TORCH_CHECK(scratch.defined() && scratch.is_cuda() && scratch.scalar_type() == torch::kByte &&
                    scratch.is_contiguous() && (size_t)scratch.numel() >= need,
                "scratch must be a contiguous uint8 CUDA tensor of at least ", need, " bytes");
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
    int rc = f3dgs_backward_accum(
        P, degree, M, R, C, fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(sc), scale_modifier, fptr(rot), fptr(cov),
        fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy, rad.data_ptr<int>(),
        reinterpret_cast<char*>(geomBuffer.data_ptr()), reinterpret_cast<char*>(binningBuffer.data_ptr()),
        reinterpret_cast<char*>(imageBuffer.data_ptr()), fptr(gc), C ? fptr(gf) : nullptr, fptr(gd),
        reinterpret_cast<char*>(scratch.data_ptr()), gcheck(g_opacities, P, "g_opacities"),
        gcheck(g_colors, (int64_t)P * 3, "g_colors_precomp"), gcheck(g_semantic_feature, (int64_t)P * C, "g_semantic_feature"),
        gcheck(g_means3D, (int64_t)P * 3, "g_means3D"), gcheck(g_cov3D, (int64_t)P * 6, "g_cov3D_precomp"),
        gcheck(g_sh, (int64_t)P * M * 3, "g_sh"), gcheck(g_scales, (int64_t)P * 3, "g_scales"),
        gcheck(g_rotations, (int64_t)P * 4, "g_rotations"), gcheck(g_means2D_out, (int64_t)P * 3, "g_means2D_out"),
        gcheck(grad_accum, P, "grad_accum"), gcheck(denom, P, "denom"),
        reinterpret_cast<void*>(static_cast<intptr_t>(composite_done_event)), debug ? 1 : 0, (void*)stream);
    check_rc(rc, "f3dgs_backward_accum");
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const auto dev = means3D.device();
    const int P = means3D.size(0);
    torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
    if (P != 0) {
        auto m3 = prep(means3D, dev, "means3D"), vm = prep(viewmatrix, dev, "viewmatrix"),
             pm = prep(projmatrix, dev, "projmatrix");
        cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
        int rc = f3dgs_mark_visible(P, fptr(m3), fptr(vm), fptr(pm), reinterpret_cast<uint8_t*>(present.data_ptr<bool>()),
                                    (void*)stream);
        check_rc(rc, "f3dgs_mark_visible");
    }
    return present;
}

// ---- post-raster feature head (include/f3dgs_b200.h: f3dgs_feature_resize_fwd / _bwd) ----------------------------
// -> (out [C,Hg,Wg], loss_sum [1]); with gt: out = sign(resized - gt) * grad_scale and loss_sum = sum |resized - gt|
std::tuple<torch::Tensor, torch::Tensor> featureResizeFwd(const torch::Tensor& feature_map, const torch::Tensor& gt,
                                                          int64_t Hg, int64_t Wg, double grad_scale) {
    TORCH_CHECK(feature_map.is_cuda() && feature_map.dim() == 3 && feature_map.scalar_type() == torch::kFloat32,
                "feature_map must be a float32 CUDA tensor [C,H,W]");
    const c10::cuda::CUDAGuard guard(feature_map.device());
    auto fm = feature_map.contiguous();
    const int C = fm.size(0), H = fm.size(1), W = fm.size(2);
    const bool has_gt = gt.defined() && gt.numel() > 0;
    torch::Tensor g;
    if (has_gt) {
        TORCH_CHECK(gt.is_cuda() && gt.scalar_type() == torch::kFloat32 && gt.dim() == 3 && gt.size(0) == C &&
                        gt.size(1) == Hg && gt.size(2) == Wg, "gt must be a float32 CUDA tensor [C,Hg,Wg]");
        g = gt.contiguous();
    }
    torch::Tensor out = torch::empty({C, Hg, Wg}, fm.options());
    torch::Tensor loss = torch::zeros({1}, fm.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
    check_rc(f3dgs_feature_resize_fwd(C, H, W, (int)Hg, (int)Wg, fptr(fm), has_gt ? fptr(g) : nullptr, (float)grad_scale,
                                      out.data_ptr<float>(), loss.data_ptr<float>(), (void*)stream),
             "f3dgs_feature_resize_fwd");
    return std::make_tuple(out, loss);
}

torch::Tensor featureResizeBwd(const torch::Tensor& dout, int64_t H, int64_t W) {
    TORCH_CHECK(dout.is_cuda() && dout.dim() == 3 && dout.scalar_type() == torch::kFloat32,
                "dout must be a float32 CUDA tensor [C,Hg,Wg]");
    const c10::cuda::CUDAGuard guard(dout.device());
    auto d = dout.contiguous();
    const int C = d.size(0), Hg = d.size(1), Wg = d.size(2);
    torch::Tensor dfm = torch::empty({C, H, W}, d.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
    check_rc(f3dgs_feature_resize_bwd(C, (int)H, (int)W, Hg, Wg, fptr(d), dfm.data_ptr<float>(), (void*)stream),
             "f3dgs_feature_resize_bwd");
    return dfm;
}

// ---- activation prologue + fused optimizer step (f3dgs_activate / f3dgs_adam_step): in-place on the caller's tensors
void activateParams(const torch::Tensor& raw_opacity, const torch::Tensor& raw_scaling, const torch::Tensor& raw_rotation,
                    const torch::Tensor& f_dc, const torch::Tensor& f_rest, torch::Tensor opacity, torch::Tensor scales,
                    torch::Tensor rotations, torch::Tensor shs) {
    TORCH_CHECK(raw_opacity.is_cuda(), "parameters must be CUDA tensors (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(raw_opacity.device());
    const int P = raw_opacity.size(0);
    const int M = shs.defined() && shs.numel() ? (int)shs.size(1) : 0;
    for (const torch::Tensor* t : std::initializer_list<const torch::Tensor*>{&raw_opacity, &raw_scaling, &raw_rotation, &f_dc, &f_rest, &opacity, &scales, &rotations, &shs})
        TORCH_CHECK(!t->defined() || t->numel() == 0 || (t->is_cuda() && t->is_contiguous() && t->scalar_type() == torch::kFloat32),
                    "activate: tensors must be contiguous float32 CUDA tensors");
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
    check_rc(f3dgs_activate(P, M, fptr(raw_opacity), fptr(raw_scaling), fptr(raw_rotation), fptr(f_dc), fptr(f_rest),
                            const_cast<float*>(fptr(opacity)), const_cast<float*>(fptr(scales)),
                            const_cast<float*>(fptr(rotations)), const_cast<float*>(fptr(shs)), (void*)stream),
             "f3dgs_activate");
}

void adamStep(int64_t kind, torch::Tensor param, const torch::Tensor& grad_activated, torch::Tensor exp_avg,
              torch::Tensor exp_avg_sq, int64_t M, double lr, double beta1, double beta2, double eps, int64_t step) {
    TORCH_CHECK(param.is_cuda(), "parameters must be CUDA tensors (this build has no CPU path)");
    const c10::cuda::CUDAGuard guard(param.device());
    for (const torch::Tensor* t : std::initializer_list<const torch::Tensor*>{&param, &grad_activated, &exp_avg, &exp_avg_sq})
        TORCH_CHECK(t->is_cuda() && t->is_contiguous() && t->scalar_type() == torch::kFloat32,
                    "adam_step: tensors must be contiguous float32 CUDA tensors");
    TORCH_CHECK(exp_avg.numel() == param.numel() && exp_avg_sq.numel() == param.numel(), "adam_step: state shape mismatch");
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream().stream();
    check_rc(f3dgs_adam_step((int)kind, (size_t)param.numel(), (int)M, param.data_ptr<float>(), fptr(grad_activated),
                             exp_avg.data_ptr<float>(), exp_avg_sq.data_ptr<float>(), (float)lr, (float)beta1, (float)beta2,
                             (float)eps, (int)step, (void*)stream),
             "f3dgs_adam_step");
}

// Read-only views into the opaque buffers for the parity harness (not part of the reference API).
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> debugViews(
    const torch::Tensor& geomBuffer, const torch::Tensor& binningBuffer, const torch::Tensor& imgBuffer, int P,
    int W, int H, int R) {
    f3dgs_layout L;
    check_rc(f3dgs_get_layout(P, W, H, R, &L), "f3dgs_get_layout");
    const int64_t tiles = (int64_t)((W + 15) / 16) * ((H + 15) / 16);
    auto dev = geomBuffer.device();
    auto view = [&](const torch::Tensor& buf, size_t off, int64_t count, torch::ScalarType ty, int64_t esize) {
        auto bytes = buf.narrow(0, (int64_t)off, count * esize);
        return bytes.view(ty).clone();
    };
    torch::Tensor point_list = R ? view(binningBuffer, L.bin_point_list, R, torch::kInt32, 4)
                                 : torch::empty({0}, torch::TensorOptions().dtype(torch::kInt32).device(dev));
    torch::Tensor ranges = view(imgBuffer, L.img_ranges, tiles * 2, torch::kInt32, 4).view({tiles, 2});
    torch::Tensor n_contrib = view(imgBuffer, L.img_n_contrib, (int64_t)W * H, torch::kInt32, 4).view({H, W});
    torch::Tensor final_T = view(imgBuffer, L.img_final_T, (int64_t)W * H, torch::kFloat32, 4).view({H, W});
    torch::Tensor rec = view(geomBuffer, L.geom_rec, (int64_t)P * 12, torch::kFloat32, 4).view({P, 12});
    return std::make_tuple(point_list, ranges, n_contrib, final_T, rec);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
    m.def("rasterize_gaussians_backward_accum", &RasterizeGaussiansBackwardAccumCUDA);
    m.def("feature_resize_fwd", &featureResizeFwd);
    m.def("feature_resize_bwd", &featureResizeBwd);
    m.def("activate", &activateParams);
    m.def("adam_step", &adamStep);
    m.def("backward_scratch_bytes", [](int P) { return (unsigned long long)f3dgs_backward_scratch_bytes(P); });
    m.def("debug_views", &debugViews);
    m.def("launch_count", []() { return (unsigned long long)f3dgs_launch_count(); });
    m.def("abi_version", []() { return f3dgs_abi_version(); });
    m.def("profile_enable", [](bool on) { f3dgs_profile_enable(on ? 1 : 0); });
    m.def("profile_read", []() {
        std::vector<double> ms(F3DGS_N_STAGES, 0.0);
        std::vector<unsigned long long> cnt(F3DGS_N_STAGES, 0);
        check_rc(f3dgs_profile_read(ms.data(), cnt.data()), "f3dgs_profile_read");
        return std::make_pair(ms, cnt);
    });
}
