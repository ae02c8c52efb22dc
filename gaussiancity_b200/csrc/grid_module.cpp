// grid_module.cpp -- the native module `grid_encoder_ext` the reference's
// extensions/grid_encoder/__init__.py imports (:16), with the two functions of
// extensions/grid_encoder/bindings.cpp:35-40 -- forward(inputs, embeddings, offsets, outputs, B, D,
// C, L, S, H, calc_grad_inputs, dy_dx, gridtype, align_corners) and backward(grad, inputs,
// embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_inputs,
// gridtype, align_corners) -- as a thin host layer over the C ABI of include/gcr_grid_encoder.h.
// With gaussiancity_b200/compat on sys.path the reference's UNMODIFIED GridEncoder runs on the
// sm_100a kernels.
//
// Host-only C++.  Checks mirror the reference's CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_*
// (grid_encoder_ext.cu:24-37, 527-543, 564-586) with its message texts.  Differences: fp32
// embeddings only (a half / double table is refused with a message instead of dispatched),
// kernels run on torch's current stream of the tensors' device, and `backward_fused` exists.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/gcr_grid_encoder.h"

namespace {

#define GCR_CHECK_CUDA(x) TORCH_CHECK(x.device().is_cuda(), #x " must be a CUDA tensor")
#define GCR_CHECK_CONTIGUOUS(x) TORCH_CHECK(x.is_contiguous(), #x " must be a contiguous tensor")
#define GCR_CHECK_IS_INT(x) TORCH_CHECK(x.scalar_type() == at::ScalarType::Int, #x " must be an int tensor")
#define GCR_CHECK_IS_FLOATING(x)                                                                                   \
  TORCH_CHECK(x.scalar_type() == at::ScalarType::Float || x.scalar_type() == at::ScalarType::Half ||                \
                  x.scalar_type() == at::ScalarType::Double,                                                        \
              #x " must be a floating tensor")
#define GCR_CHECK_F32(x)                                                                                           \
  TORCH_CHECK(x.scalar_type() == at::ScalarType::Float,                                                             \
              #x " must be float32: the B200 grid encoder implements the fp32 table GaussianCity uses")

void check(int rc, const char* what) { TORCH_CHECK(rc == 0, what, ": ", gcr_grid_last_error()); }

void grid_encode_forward(const at::Tensor inputs, const at::Tensor embeddings, const at::Tensor offsets,
                         at::Tensor outputs, const uint32_t B, const uint32_t D, const uint32_t C, const uint32_t L,
                         const float S, const uint32_t H, const bool calc_grad_inputs, at::Tensor dy_dx,
                         const uint32_t gridtype, const bool align_corners) {
  GCR_CHECK_CUDA(inputs); GCR_CHECK_CUDA(embeddings); GCR_CHECK_CUDA(offsets); GCR_CHECK_CUDA(outputs);
  GCR_CHECK_CUDA(dy_dx);
  GCR_CHECK_CONTIGUOUS(inputs); GCR_CHECK_CONTIGUOUS(embeddings); GCR_CHECK_CONTIGUOUS(offsets);
  GCR_CHECK_CONTIGUOUS(outputs); GCR_CHECK_CONTIGUOUS(dy_dx);
  GCR_CHECK_IS_FLOATING(inputs); GCR_CHECK_IS_FLOATING(embeddings); GCR_CHECK_IS_INT(offsets);
  GCR_CHECK_IS_FLOATING(outputs); GCR_CHECK_IS_FLOATING(dy_dx);
  GCR_CHECK_F32(inputs); GCR_CHECK_F32(embeddings); GCR_CHECK_F32(outputs); GCR_CHECK_F32(dy_dx);
  const c10::cuda::CUDAGuard guard(inputs.device());
  check(gcr_grid_encode_forward(inputs.data_ptr<float>(), embeddings.data_ptr<float>(), offsets.data_ptr<int>(),
                                outputs.data_ptr<float>(), B, D, C, L, S, H, calc_grad_inputs ? 1 : 0,
                                dy_dx.data_ptr<float>(), gridtype, align_corners ? 1 : 0,
                                at::cuda::getCurrentCUDAStream().stream()),
        "grid_encode_forward");
}

void grid_encode_backward(const at::Tensor grad, const at::Tensor inputs, const at::Tensor embeddings,
                          const at::Tensor offsets, at::Tensor grad_embeddings, const uint32_t B, const uint32_t D,
                          const uint32_t C, const uint32_t L, const float S, const uint32_t H,
                          const bool calc_grad_inputs, const at::Tensor dy_dx, at::Tensor grad_inputs,
                          const uint32_t gridtype, const bool align_corners) {
  GCR_CHECK_CUDA(grad); GCR_CHECK_CUDA(inputs); GCR_CHECK_CUDA(embeddings); GCR_CHECK_CUDA(offsets);
  GCR_CHECK_CUDA(grad_embeddings); GCR_CHECK_CUDA(dy_dx); GCR_CHECK_CUDA(grad_inputs);
  GCR_CHECK_CONTIGUOUS(grad); GCR_CHECK_CONTIGUOUS(inputs); GCR_CHECK_CONTIGUOUS(embeddings);
  GCR_CHECK_CONTIGUOUS(offsets); GCR_CHECK_CONTIGUOUS(grad_embeddings); GCR_CHECK_CONTIGUOUS(dy_dx);
  GCR_CHECK_CONTIGUOUS(grad_inputs);
  GCR_CHECK_IS_FLOATING(grad); GCR_CHECK_IS_FLOATING(inputs); GCR_CHECK_IS_FLOATING(embeddings);
  GCR_CHECK_IS_INT(offsets); GCR_CHECK_IS_FLOATING(grad_embeddings); GCR_CHECK_IS_FLOATING(dy_dx);
  GCR_CHECK_IS_FLOATING(grad_inputs);
  GCR_CHECK_F32(grad); GCR_CHECK_F32(inputs); GCR_CHECK_F32(grad_embeddings); GCR_CHECK_F32(dy_dx);
  GCR_CHECK_F32(grad_inputs);
  const c10::cuda::CUDAGuard guard(inputs.device());
  check(gcr_grid_encode_backward(grad.data_ptr<float>(), inputs.data_ptr<float>(), nullptr, offsets.data_ptr<int>(),
                                 grad_embeddings.data_ptr<float>(), B, D, C, L, S, H, calc_grad_inputs ? 1 : 0,
                                 dy_dx.data_ptr<float>(), grad_inputs.data_ptr<float>(), gridtype,
                                 align_corners ? 1 : 0, at::cuda::getCurrentCUDAStream().stream()),
        "grid_encode_backward");
}

// Not in the reference: one launch for both gradients, no dy_dx (gcr_grid_encode_backward_fused).
void grid_encode_backward_fused(const at::Tensor grad, const at::Tensor inputs, const at::Tensor embeddings,
                                const at::Tensor offsets, at::Tensor grad_embeddings, const uint32_t B,
                                const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H,
                                at::Tensor grad_inputs, const uint32_t gridtype, const bool align_corners) {
  GCR_CHECK_CUDA(grad); GCR_CHECK_CUDA(inputs); GCR_CHECK_CUDA(embeddings); GCR_CHECK_CUDA(offsets);
  GCR_CHECK_CUDA(grad_embeddings); GCR_CHECK_CUDA(grad_inputs);
  GCR_CHECK_CONTIGUOUS(grad); GCR_CHECK_CONTIGUOUS(inputs); GCR_CHECK_CONTIGUOUS(embeddings);
  GCR_CHECK_CONTIGUOUS(offsets); GCR_CHECK_CONTIGUOUS(grad_embeddings); GCR_CHECK_CONTIGUOUS(grad_inputs);
  GCR_CHECK_IS_INT(offsets);
  GCR_CHECK_F32(grad); GCR_CHECK_F32(inputs); GCR_CHECK_F32(embeddings); GCR_CHECK_F32(grad_embeddings);
  GCR_CHECK_F32(grad_inputs);
  const c10::cuda::CUDAGuard guard(inputs.device());
  check(gcr_grid_encode_backward_fused(grad.data_ptr<float>(), inputs.data_ptr<float>(),
                                       embeddings.data_ptr<float>(), offsets.data_ptr<int>(),
                                       grad_embeddings.data_ptr<float>(), B, D, C, L, S, H,
                                       grad_inputs.data_ptr<float>(), gridtype, align_corners ? 1 : 0,
                                       at::cuda::getCurrentCUDAStream().stream()),
        "grid_encode_backward_fused");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native kernels behind the reference's grid_encoder_ext surface";
  m.def("forward", &grid_encode_forward, "grid_encode_forward (CUDA)");
  m.def("backward", &grid_encode_backward, "grid_encode_backward (CUDA)");
  m.def("backward_fused", &grid_encode_backward_fused, "grid + input gradients in one launch (CUDA)");
  m.def("abi_version", []() { return gcr_grid_abi_version(); });
}
