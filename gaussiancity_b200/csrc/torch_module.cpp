// torch_module.cpp -- Seam A as a native module: the pybind module `diff_gaussian_rasterization_ext`
// the reference's Python package imports (DGR/__init__.py:16; DGR/bindings.cpp:15-18), with the
// same three functions, positional signatures and return tuples as RasterizeGaussiansCUDA /
// RasterizeGaussiansBackwardCUDA / markVisible (DGR/rasterize_points.cu:35-173), implemented as
// a thin layer over the C ABI of include/gcr_rasterizer.h.  With gaussiancity_b200/compat on
// sys.path the reference's UNMODIFIED DGR/__init__.py runs on the sm_100a kernels.
//
// This file is host-only C++ (no kernels): tensor checks, output allocation through torch's
// caching allocator, and the pointer hand-off.  Differences from the reference binding, all
// invisible to its callers: kernels run on torch's CURRENT stream of the tensors' device (the
// reference uses the legacy default stream of the current device); the three byte buffers are
// allocated at their final size (no resize_ of an empty tensor); gradient tensors are
// torch::empty, because the fused backward writes every element exactly once.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <tuple>

#include "../../include/gcr_rasterizer.h"

namespace {

constexpr int kChannels = 3;  // DGR/cuda_rasterizer/config.h:15

// allocator callback context: one byte tensor per opaque buffer
struct ByteBuffer {
  torch::Tensor t;
  torch::TensorOptions opts;
};
char* alloc_bytes(void* ctx, size_t n) {
  auto* b = static_cast<ByteBuffer*>(ctx);
  try {
    b->t = torch::empty({static_cast<int64_t>(n)}, b->opts);
  } catch (...) {
    return nullptr;  // surfaces as "allocation failed" from the library
  }
  return static_cast<char*>(b->t.data_ptr());
}

// contiguous fp32 view of an input on `dev` with at least `align`-byte alignment; an empty
// tensor is an absent optional (reference: empty tensor -> nullptr)
struct Arg {
  torch::Tensor keep;
  const float* ptr = nullptr;
};
Arg prep(const torch::Tensor& t, const char* name, const torch::Device& dev, size_t align = 4) {
  Arg a;
  if (t.numel() == 0) return a;
  TORCH_CHECK(t.device() == dev, name, " must be on ", dev, ", got ", t.device());
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
  a.keep = t.contiguous();
  if (reinterpret_cast<uintptr_t>(a.keep.data_ptr()) % align != 0) a.keep = a.keep.clone();
  a.ptr = a.keep.data_ptr<float>();
  return a;
}

void check(int rc, const char* what) {
  TORCH_CHECK(rc >= 0, what, ": ", gcr_last_error());
}

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians(const torch::Tensor& background, const torch::Tensor& means3D,
                    const torch::Tensor& colors, const torch::Tensor& opacity,
                    const torch::Tensor& scales, const torch::Tensor& rotations,
                    const float scale_modifier, const torch::Tensor& cov3D_precomp,
                    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                    const float tan_fovx, const float tan_fovy, const int image_height,
                    const int image_width, const torch::Tensor& sh, const int degree,
                    const torch::Tensor& campos, const bool prefiltered, const bool debug) {
  TORCH_CHECK(means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU path)");
  const torch::Device dev = means3D.device();
  const c10::cuda::CUDAGuard guard(dev);
  const int P = static_cast<int>(means3D.size(0));
  const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
  const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);

  // every pixel is written by the blend when P != 0
  torch::Tensor out_color = P != 0 ? torch::empty({kChannels, image_height, image_width}, f32)
                                   : torch::zeros({kChannels, image_height, image_width}, f32);
  torch::Tensor radii = torch::empty({P}, f32.dtype(torch::kInt32));
  ByteBuffer geom{torch::empty({0}, u8), u8}, binning{torch::empty({0}, u8), u8}, img{torch::empty({0}, u8), u8};
  int rendered = 0;
  if (P != 0) {
    const int M = sh.numel() != 0 ? static_cast<int>(sh.size(1)) : 0;
    const Arg bg = prep(background, "background", dev), m3 = prep(means3D, "means3D", dev);
    const Arg col = prep(colors, "colors_precomp", dev), op = prep(opacity, "opacity", dev);
    const Arg sc = prep(scales, "scales", dev), rot = prep(rotations, "rotations", dev, 16);
    const Arg cov = prep(cov3D_precomp, "cov3D_precomp", dev), vm = prep(viewmatrix, "viewmatrix", dev);
    const Arg pm = prep(projmatrix, "projmatrix", dev), shs = prep(sh, "sh", dev, 32);
    const Arg cp = prep(campos, "campos", dev);
    rendered = gcr_rasterizer_forward(
        alloc_bytes, &geom, alloc_bytes, &binning, alloc_bytes, &img, P, degree, M, bg.ptr, image_width,
        image_height, m3.ptr, shs.ptr, col.ptr, op.ptr, sc.ptr, scale_modifier, rot.ptr, cov.ptr, vm.ptr,
        pm.ptr, cp.ptr, tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(),
        radii.data_ptr<int>(), debug ? 1 : 0, 0, 1, at::cuda::getCurrentCUDAStream(dev.index()).stream());
    check(rendered, "rasterize_gaussians");
  }
  return std::make_tuple(rendered, out_color, radii, geom.t, binning.t, img.t);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
rasterize_gaussians_backward(const torch::Tensor& background, const torch::Tensor& means3D,
                             const torch::Tensor& radii, const torch::Tensor& colors,
                             const torch::Tensor& scales, const torch::Tensor& rotations,
                             const float scale_modifier, const torch::Tensor& cov3D_precomp,
                             const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                             const float tan_fovx, const float tan_fovy,
                             const torch::Tensor& dL_dout_color, const torch::Tensor& sh,
                             const int degree, const torch::Tensor& campos,
                             const torch::Tensor& geomBuffer, const int R,
                             const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                             const bool debug) {
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU path)");
  const torch::Device dev = means3D.device();
  const c10::cuda::CUDAGuard guard(dev);
  const int P = static_cast<int>(means3D.size(0));
  const int H = static_cast<int>(dL_dout_color.size(1)), W = static_cast<int>(dL_dout_color.size(2));
  const int M = sh.numel() != 0 ? static_cast<int>(sh.size(1)) : 0;
  const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
  auto make = [&](std::initializer_list<int64_t> shape) {
    return P != 0 ? torch::empty(shape, f32) : torch::zeros(shape, f32);
  };
  torch::Tensor dL_dmeans3D = make({P, 3}), dL_dmeans2D = make({P, 3}), dL_dcolors = make({P, kChannels});
  torch::Tensor dL_dopacity = make({P, 1}), dL_dcov3D = make({P, 6}), dL_dsh = make({P, M, 3});
  torch::Tensor dL_dscales = make({P, 3}), dL_drotations = make({P, 4});
  if (P != 0) {
    const Arg bg = prep(background, "background", dev), m3 = prep(means3D, "means3D", dev);
    const Arg col = prep(colors, "colors_precomp", dev), sc = prep(scales, "scales", dev);
    const Arg rot = prep(rotations, "rotations", dev, 16), cov = prep(cov3D_precomp, "cov3D_precomp", dev);
    const Arg vm = prep(viewmatrix, "viewmatrix", dev), pm = prep(projmatrix, "projmatrix", dev);
    const Arg shs = prep(sh, "sh", dev, 32), cp = prep(campos, "campos", dev);
    const Arg dpix = prep(dL_dout_color, "dL_dout_color", dev);
    const torch::Tensor rad = radii.contiguous();
    const int rc = gcr_rasterizer_backward(
        P, degree, M, R, bg.ptr, W, H, m3.ptr, shs.ptr, col.ptr, sc.ptr, scale_modifier, rot.ptr, cov.ptr,
        vm.ptr, pm.ptr, cp.ptr, tan_fovx, tan_fovy, rad.numel() ? rad.data_ptr<int>() : nullptr,
        static_cast<char*>(geomBuffer.data_ptr()),
        binningBuffer.numel() ? static_cast<char*>(binningBuffer.data_ptr()) : nullptr,
        static_cast<char*>(imageBuffer.data_ptr()), dpix.ptr, dL_dmeans2D.data_ptr<float>(), nullptr,
        dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(),
        dL_dcov3D.data_ptr<float>(), dL_dsh.numel() ? dL_dsh.data_ptr<float>() : nullptr,
        dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>(), debug ? 1 : 0, 0, 1,
        at::cuda::getCurrentCUDAStream(dev.index()).stream());
    check(rc, "rasterize_gaussians_backward");
  }
  return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                         dL_drotations);
}

torch::Tensor mark_visible(const torch::Tensor& means3D, const torch::Tensor& viewmatrix,
                           const torch::Tensor& projmatrix) {
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU path)");
  const torch::Device dev = means3D.device();
  const c10::cuda::CUDAGuard guard(dev);
  const int P = static_cast<int>(means3D.size(0));
  torch::Tensor present = torch::zeros({P}, torch::TensorOptions().dtype(torch::kBool).device(dev));
  if (P != 0) {
    const Arg m3 = prep(means3D, "means3D", dev), vm = prep(viewmatrix, "viewmatrix", dev);
    const Arg pm = prep(projmatrix, "projmatrix", dev);
    check(gcr_rasterizer_mark_visible(P, m3.ptr, vm.ptr, pm.ptr, static_cast<uint8_t*>(present.data_ptr()),
                                      at::cuda::getCurrentCUDAStream(dev.index()).stream()),
          "mark_visible");
  }
  return present;
}

// ---- adapter fusion: colors_precomp path through a pixel window; an undefined /
// empty opacity or rotation tensor means opaque / identity (gcr_rasterizer_forward_window) --------
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians_window(const torch::Tensor& background, const torch::Tensor& means3D,
                           const torch::Tensor& colors, const torch::Tensor& opacity,
                           const torch::Tensor& scales, const torch::Tensor& rotations,
                           const float scale_modifier, const torch::Tensor& viewmatrix,
                           const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                           const int image_height, const int image_width, const int win_x, const int win_y,
                           const int win_w, const int win_h, const bool debug) {
  TORCH_CHECK(means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU path)");
  const torch::Device dev = means3D.device();
  const c10::cuda::CUDAGuard guard(dev);
  const int P = static_cast<int>(means3D.size(0));
  const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
  const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);
  torch::Tensor out_color = P != 0 ? torch::empty({kChannels, win_h, win_w}, f32)
                                   : torch::zeros({kChannels, win_h, win_w}, f32);
  torch::Tensor radii = torch::empty({P}, f32.dtype(torch::kInt32));
  ByteBuffer geom{torch::empty({0}, u8), u8}, binning{torch::empty({0}, u8), u8}, img{torch::empty({0}, u8), u8};
  int rendered = 0;
  if (P != 0) {
    const Arg bg = prep(background, "background", dev), m3 = prep(means3D, "means3D", dev);
    const Arg col = prep(colors, "colors_precomp", dev), op = prep(opacity, "opacity", dev);
    const Arg sc = prep(scales, "scales", dev), rot = prep(rotations, "rotations", dev, 16);
    const Arg vm = prep(viewmatrix, "viewmatrix", dev), pm = prep(projmatrix, "projmatrix", dev);
    rendered = gcr_rasterizer_forward_window(
        alloc_bytes, &geom, alloc_bytes, &binning, alloc_bytes, &img, P, 0, 0, bg.ptr, image_width, image_height,
        m3.ptr, nullptr, col.ptr, op.ptr, sc.ptr, scale_modifier, rot.ptr, nullptr, vm.ptr, pm.ptr, nullptr,
        tan_fovx, tan_fovy, 0, out_color.data_ptr<float>(), radii.data_ptr<int>(), debug ? 1 : 0, win_x, win_y,
        win_w, win_h, at::cuda::getCurrentCUDAStream(dev.index()).stream());
    check(rendered, "rasterize_gaussians_window");
  }
  return std::make_tuple(rendered, out_color, radii, geom.t, binning.t, img.t);
}

// -> (dL_dmeans3D, dL_dcolors, dL_dscales, dL_dopacity | empty, dL_drotations | empty, dL_dmeans2D)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians_backward_window(const torch::Tensor& background, const torch::Tensor& means3D,
                                    const torch::Tensor& radii, const torch::Tensor& scales,
                                    const torch::Tensor& rotations, const float scale_modifier,
                                    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                                    const float tan_fovx, const float tan_fovy, const int image_height,
                                    const int image_width, const int win_x, const int win_y, const int win_w,
                                    const int win_h, const torch::Tensor& dL_dout_color,
                                    const torch::Tensor& geomBuffer, const int R,
                                    const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                                    const bool want_opacity, const bool debug) {
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (there is no CPU path)");
  const torch::Device dev = means3D.device();
  const c10::cuda::CUDAGuard guard(dev);
  const int P = static_cast<int>(means3D.size(0));
  const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
  auto make = [&](std::initializer_list<int64_t> shape) {
    return P != 0 ? torch::empty(shape, f32) : torch::zeros(shape, f32);
  };
  const bool has_rot = rotations.numel() != 0;
  torch::Tensor dL_dmeans3D = make({P, 3}), dL_dmeans2D = make({P, 3}), dL_dcolors = make({P, kChannels});
  torch::Tensor dL_dcov3D = make({P, 6}), dL_dscales = make({P, 3});
  torch::Tensor dL_dopacity = want_opacity ? make({P, 1}) : torch::empty({0}, f32);
  torch::Tensor dL_drot = has_rot ? make({P, 4}) : torch::empty({0}, f32);
  if (P != 0) {
    const Arg bg = prep(background, "background", dev), m3 = prep(means3D, "means3D", dev);
    const Arg sc = prep(scales, "scales", dev), rot = prep(rotations, "rotations", dev, 16);
    const Arg vm = prep(viewmatrix, "viewmatrix", dev), pm = prep(projmatrix, "projmatrix", dev);
    const Arg dpix = prep(dL_dout_color, "dL_dout_color", dev);
    const torch::Tensor rad = radii.contiguous();
    const int rc = gcr_rasterizer_backward_window(
        P, 0, 0, R, bg.ptr, image_width, image_height, m3.ptr, nullptr, nullptr, sc.ptr, scale_modifier, rot.ptr,
        nullptr, vm.ptr, pm.ptr, nullptr, tan_fovx, tan_fovy, rad.numel() ? rad.data_ptr<int>() : nullptr,
        static_cast<char*>(geomBuffer.data_ptr()),
        binningBuffer.numel() ? static_cast<char*>(binningBuffer.data_ptr()) : nullptr,
        static_cast<char*>(imageBuffer.data_ptr()), dpix.ptr, dL_dmeans2D.data_ptr<float>(), nullptr,
        want_opacity ? dL_dopacity.data_ptr<float>() : nullptr, dL_dcolors.data_ptr<float>(),
        dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), nullptr, dL_dscales.data_ptr<float>(),
        has_rot ? dL_drot.data_ptr<float>() : nullptr, debug ? 1 : 0, win_x, win_y, win_w, win_h,
        at::cuda::getCurrentCUDAStream(dev.index()).stream());
    check(rc, "rasterize_gaussians_backward_window");
  }
  return std::make_tuple(dL_dmeans3D, dL_dcolors, dL_dscales, dL_dopacity, dL_drot, dL_dmeans2D);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native kernels behind the reference's diff_gaussian_rasterization_ext surface";
  m.def("rasterize_gaussians", &rasterize_gaussians);
  m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
  m.def("mark_visible", &mark_visible);
  m.def("rasterize_gaussians_window", &rasterize_gaussians_window);
  m.def("rasterize_gaussians_backward_window", &rasterize_gaussians_backward_window);
  m.def("abi_version", []() { return gcr_abi_version(); });
}
