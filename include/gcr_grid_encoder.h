/*
 * gcr_grid_encoder.h -- C ABI of the B200-native multi-resolution hash-grid encoder
 * (libgcr_grid_encoder.so).
 *
 * Replaces, argument for argument, the two entry points the reference's pybind module
 * `grid_encoder_ext` exposes (extensions/grid_encoder/bindings.cpp:18-40):
 *
 *   grid_encode_forward   extensions/grid_encoder/grid_encoder_ext.cu:520-552
 *                         (kernel_grid, :95-243)
 *   grid_encode_backward  extensions/grid_encoder/grid_encoder_ext.cu:554-606
 *                         (kernel_grid_backward :245-331, kernel_input_backward :333-360)
 *
 * with at::Tensor replaced by raw DEVICE pointers and a trailing cudaStream_t (passed as
 * void*; NULL = the legacy default stream the reference launches on).  fp32 embeddings only --
 * the reference also instantiates half / double, GaussianCity never uses them
 * (extensions/grid_encoder/__init__.py:158, torch.empty default dtype).
 *
 * Layouts (all contiguous):
 *   inputs        float [B, D]            coordinates in [0, 1]; a point with any coordinate
 *                                         outside [0, 1] encodes to zeros and receives no gradient
 *   embeddings    float [offsets[L], C]   the L level tables back to back
 *   offsets       int32 [L + 1]           first entry of each level, DEVICE memory
 *   outputs       float [L, B, C]         level-major (the Python wrapper permutes to [B, L*C])
 *   dy_dx         float [B, L, D, C]      d outputs / d inputs; written only if calc_grad_inputs
 *   grad          float [L, B, C]
 *   grad_embeddings float [offsets[L], C] ACCUMULATED INTO (caller zero-fills, as the
 *                                         reference's torch.zeros_like does, __init__.py:93)
 *   grad_inputs   float [B, D]            written (not accumulated) only if calc_grad_inputs
 *
 *   D in {2,3,4,5}; C in {1,2,4,8}; S = log2(per_level_scale); H = base resolution;
 *   gridtype 0 = hash, 1 = tiled; align_corners as the reference.
 *
 * Every function returns 0 on success, non-zero on failure with a message in
 * gcr_grid_last_error() (thread-local).  Unsupported D / C fail with the reference's own
 * message text ("GridEncoding: C must be 1, 2, 4, or 8.", grid_encoder_ext.cu:392,429).
 * Launches are asynchronous on `stream`; nothing is synchronised.
 */
#ifndef GCR_GRID_ENCODER_H
#define GCR_GRID_ENCODER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCR_GRID_ABI_VERSION 1

#if defined(__GNUC__)
#define GCR_GRID_API __attribute__((visibility("default")))
#else
#define GCR_GRID_API
#endif

GCR_GRID_API int gcr_grid_abi_version(void);
GCR_GRID_API const char *gcr_grid_last_error(void);

/* grid_encode_forward (grid_encoder_ext.cu:520-552). */
GCR_GRID_API int gcr_grid_encode_forward(
    const float *inputs, const float *embeddings, const int *offsets, float *outputs,
    uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
    int calc_grad_inputs, float *dy_dx, uint32_t gridtype, int align_corners, void *stream);

/* grid_encode_backward (grid_encoder_ext.cu:554-606).  `embeddings` is accepted and ignored,
 * like the reference (the backward never reads the table). */
GCR_GRID_API int gcr_grid_encode_backward(
    const float *grad, const float *inputs, const float *embeddings, const int *offsets,
    float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
    int calc_grad_inputs, const float *dy_dx, float *grad_inputs, uint32_t gridtype,
    int align_corners, void *stream);

/* Fused variant for callers that own the autograd node (gaussiancity_b200.grid_encoder): one
 * launch computes grad_embeddings AND grad_inputs, the latter by re-reading the 2^D corner rows
 * the forward read instead of a [B, L, D, C] dy_dx tensor written by the forward and read back
 * here (2 * 4*L*D*C bytes per point never touch HBM).  grad_inputs must be zero-filled by the
 * caller (levels accumulate into it).  Same arithmetic per (point, level) as
 * kernel_grid :196-241 followed by kernel_input_backward :333-360, summed over levels in a
 * different order (float atomics), hence not bit-identical to them. */
GCR_GRID_API int gcr_grid_encode_backward_fused(
    const float *grad, const float *inputs, const float *embeddings, const int *offsets,
    float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
    float *grad_inputs, uint32_t gridtype, int align_corners, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GCR_GRID_ENCODER_H */
