/* gcr_rasterizer.h -- C ABI of the B200-native differentiable Gaussian rasterizer.
 *
 * This is the drop-in boundary for the hot path of hzxie/GaussianCity's
 * extensions/diff_gaussian_rasterization ("DGR").  The three entry points replace, one for one,
 * the static methods of CudaRasterizer::Rasterizer that the reference's torch binding calls:
 *
 *   gcr_rasterizer_forward       <- Rasterizer::forward     DGR/cuda_rasterizer/rasterizer.h:24-36
 *                                   (called from RasterizeGaussiansCUDA, DGR/rasterize_points.cu:77-90)
 *   gcr_rasterizer_backward      <- Rasterizer::backward    DGR/cuda_rasterizer/rasterizer.h:38-50
 *                                   (called from RasterizeGaussiansBackwardCUDA, rasterize_points.cu:128-152)
 *   gcr_rasterizer_mark_visible  <- Rasterizer::markVisible DGR/cuda_rasterizer/rasterizer.h:21-22
 *                                   (called from markVisible, rasterize_points.cu:157-173)
 *
 * Same argument order and meaning as those methods; std::function<char*(size_t)> becomes a
 * (function pointer, context) pair; bool becomes int; three trailing extensions are added:
 * the owner's tile-row shard (rank, count) for multi-GPU screen-tile sharding and the CUDA
 * stream to launch on (the reference uses the legacy default stream).  All pointers are DEVICE
 * pointers to contiguous fp32/int32 arrays; optional inputs are NULL exactly where the
 * reference receives nullptr (empty tensors).  No torch types appear here.
 *
 * Error behaviour: the reference throws std::runtime_error (CHECK_CUDA in debug mode, the
 * "For non-RGB" check); a C ABI cannot throw, so functions return a negative value and the
 * message is available from gcr_last_error().  The host bindings turn that into RuntimeError.
 */
#ifndef GCR_RASTERIZER_H_INCLUDED
#define GCR_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCR_ABI_VERSION 1

#if defined(__GNUC__)
#define GCR_API __attribute__((visibility("default")))
#else
#define GCR_API
#endif

/* Resizable-buffer callback: must return a device pointer to at least `bytes` bytes (>= 128 B
 * aligned) that stays valid until the matching backward call has completed
 * (replaces std::function<char*(size_t)>, DGR/rasterize_points.cu:27-33). */
typedef char* (*gcr_alloc_fn)(void* ctx, size_t bytes);

GCR_API int gcr_abi_version(void);
GCR_API const char* gcr_last_error(void);

/* Returns num_rendered (>= 0), or < 0 on error.  radii may be NULL (an internal array is used).
 * With shard_count > 1 only tile rows r with r % shard_count == shard_rank are binned and
 * blended (other pixels of out_color are left untouched); radii are global either way. */
GCR_API int gcr_rasterizer_forward(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                           gcr_alloc_fn binningBuffer, void* binning_ctx,
                           gcr_alloc_fn imageBuffer, void* image_ctx,
                           int P, int D, int M, const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                           float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                           int* radii, int debug,
                           int shard_rank, int shard_count, void* cuda_stream);

/* Returns 0, or < 0 on error.  Every element of every gradient array is written (zeros for
 * culled Gaussians): callers need not pre-zero them.  dL_dconic ([P,2,2]), dL_dsh, dL_dscale,
 * dL_drot may be NULL.  With shard_count > 1 the per-Gaussian blend gradients are partial sums
 * over this rank's tile rows: call gcr_rasterizer_backward_blend on every rank, reduce the
 * [P,12] accumulators, then gcr_rasterizer_backward_geometry. */
GCR_API int gcr_rasterizer_backward(int P, int D, int M, int R, const float* background, int width,
                            int height, const float* means3D, const float* shs,
                            const float* colors_precomp, const float* scales,
                            float scale_modifier, const float* rotations,
                            const float* cov3D_precomp, const float* viewmatrix,
                            const float* projmatrix, const float* campos, float tan_fovx,
                            float tan_fovy, const int* radii, char* geom_buffer,
                            char* binning_buffer, char* image_buffer, const float* dL_dpix,
                            float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                            float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                            float* dL_dscale, float* dL_drot, int debug,
                            int shard_rank, int shard_count, void* cuda_stream);

/* The two halves of gcr_rasterizer_backward, split at the only cross-tile coupling so that a
 * tile-sharded job can reduce between them.  grad_acc is [P,12] fp32 (48 B per Gaussian:
 * dmean2D.xy, dconic.xyw, dopacity, dcolor.rgb, 3 pad), zeroed by _blend before accumulation.
 * _geometry processes Gaussians [range_start, range_start + range_count) only (all pointers are
 * bases of the full [P,...] arrays; range_count < 0 means all): after a reduce-scatter each rank
 * finishes its own slice. */
GCR_API int gcr_rasterizer_backward_blend(int P, int R, const float* background, int width, int height,
                                  char* binning_buffer, char* image_buffer, const float* dL_dpix,
                                  float* grad_acc, int debug, int shard_rank, int shard_count,
                                  void* cuda_stream);
GCR_API int gcr_rasterizer_backward_geometry(int P, int D, int M, const float* means3D, const float* shs,
                                     const float* scales, float scale_modifier,
                                     const float* rotations, const float* cov3D_precomp,
                                     const float* viewmatrix, const float* projmatrix,
                                     const float* campos, int width, int height, float tan_fovx,
                                     float tan_fovy, const int* radii, char* geom_buffer,
                                     const float* grad_acc, float* dL_dmean2D, float* dL_dconic,
                                     float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                                     float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                                     float* dL_drot, int debug, int range_start, int range_count,
                                     void* cuda_stream);

/* present: P bytes (bool). Returns 0 or < 0. */
GCR_API int gcr_rasterizer_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, void* cuda_stream);

/* ---- introspection (tests / parity harness only) ------------------------------------------
 * Byte offsets of the internal arrays inside the three opaque buffers, so a test can view
 * them without the library exposing its layout as API.  `which` values below; returns
 * (size_t)-1 for an unknown id. */
enum {
  GCR_GEOM_DEPTH_SORTED_KEYS = 0, /* u32[P]  depth keys in sorted order (invisible = 0xFFFFFFFF) */
  GCR_GEOM_TILES_TOUCHED = 1,     /* u32[P]  */
  GCR_GEOM_RECORDS = 2,           /* 48 B[P] {x,y,A,B | C,o,r,g | b,idx,2ln(255o),0} */
  GCR_GEOM_CLAMPED = 3,           /* u8[P]   bit ch set if SH colour channel was clamped */
  GCR_GEOM_SORTED_GAUSS = 4,      /* u32[P]  Gaussian indices in depth order */
  GCR_GEOM_OFFSETS = 5,           /* u32[P]  inclusive scan of tiles_touched in depth order */
  GCR_GEOM_GRAD_ACC = 6,          /* 48 B[P] backward accumulator */
  GCR_GEOM_RADII = 7,             /* i32[P]  internal radii (used when radii == NULL) */
  GCR_GEOM_TOTAL_BYTES = 100,
  GCR_BIN_POINT_LIST = 200,       /* u32[R]  sorted instance -> Gaussian index */
  GCR_BIN_TILE_KEYS = 201,        /* u32[R]  sorted tile ids */
  GCR_BIN_INSTANCES = 202,        /* 48 B[R] gathered records in sorted order */
  GCR_BIN_TOTAL_BYTES = 300,
  GCR_IMG_FINAL_T = 400,          /* f32[H*W] */
  GCR_IMG_N_CONTRIB = 401,        /* u32[H*W] */
  GCR_IMG_RANGES = 402,           /* uint2[tiles] */
  GCR_IMG_TOTAL_BYTES = 500
};
GCR_API size_t gcr_debug_offset(int which, int P, int R, int width, int height);
/* test hook: the NEXT gcr_rasterizer_forward on this thread also stores the per-Gaussian 3D
 * covariance ([P,6] fp32, device) it computed (the reference keeps it in geomBuffer). */
GCR_API void gcr_debug_set_cov3d_out(float* cov3d);

/* ---- per-stage device timing (bench.py) ------------------------------------------------------
 * When enabled, every pipeline stage of the next forward / backward call is bracketed by CUDA
 * events recorded on the caller's stream (no synchronisation is added).  After the caller has
 * synchronised, gcr_profile_stage_ms(i) is the device time of stage i of the most recent call
 * (-1 if the stage did not run).  Stage names: gcr_profile_stage_name(i), i < stage_count. */
GCR_API void gcr_profile_enable(int on);
GCR_API int gcr_profile_stage_count(void);
GCR_API const char* gcr_profile_stage_name(int stage);
GCR_API float gcr_profile_stage_ms(int stage);

#ifdef __cplusplus
}
#endif
#endif /* GCR_RASTERIZER_H_INCLUDED */
