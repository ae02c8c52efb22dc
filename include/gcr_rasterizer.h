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
 * (function pointer, context) pair; bool becomes int; trailing extensions are added: the tile-row
 * stripe of a multi-GPU frame and the CUDA stream to launch on (the reference uses the legacy
 * default stream).  All pointers are DEVICE pointers to contiguous fp32/int32 arrays unless
 * stated otherwise; optional inputs are NULL exactly where the reference receives nullptr
 * (empty tensors).  No torch types appear here.
 *
 * Error behaviour: the reference throws std::runtime_error (CHECK_CUDA in debug mode, the
 * "For non-RGB" check); a C ABI cannot throw, so functions return a negative value and the
 * message is available from gcr_last_error().  The host bindings turn that into RuntimeError.
 *
 * Threading: every function may be called concurrently from different host threads (the
 * reference is called from the caller's thread in forward and from the autograd engine's device
 * thread in backward).  gcr_last_error() is per thread.  The profiler (gcr_profile_*) is a
 * single-threaded bench facility.
 */
#ifndef GCR_RASTERIZER_H_INCLUDED
#define GCR_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCR_ABI_VERSION 2
#define GCR_MAX_SHARDS 16          /* tile-row stripes (GPUs) of one frame */
#define GCR_PEER_HANDLE_BYTES 64   /* opaque inter-process memory handle */

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
 * M (SH coefficients per Gaussian) must be <= 16 on the SH path.
 *
 * Tile-row stripes (shard_count > 1): the frame's tile rows are cut into shard_count contiguous
 * stripes and this call bins and blends only stripe `shard_rank` (other pixels of out_color are
 * left untouched; radii are global either way; num_rendered counts this stripe's instances).
 * stripe_bounds is a DEVICE array of shard_count + 1 non-decreasing tile-row indices
 * (bounds[0] = 0, bounds[shard_count] = ceil(height / 16)), e.g. from gcr_stripe_partition.  With
 * stripe_bounds == NULL the stripes are equal-height (balanced == 0; gcr_rasterizer_forward is
 * this form) or BALANCED (balanced != 0): the projection pass counts the tile instances of every
 * tile row and cuts the rows into stripes of about equal instance count on the device, at no
 * extra pass over the inputs.  The cut is deterministic, so ranks that render the same Gaussians
 * obtain the same bounds without communicating.  The bounds used travel inside geom_buffer. */
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
GCR_API int gcr_rasterizer_forward_striped(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                           gcr_alloc_fn binningBuffer, void* binning_ctx,
                           gcr_alloc_fn imageBuffer, void* image_ctx,
                           int P, int D, int M, const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                           float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                           int* radii, int debug,
                           int shard_rank, int shard_count, const int* stripe_bounds,
                           int balanced, void* cuda_stream);

/* Crop folded into the rasterizer (replaces render-then-slice in
 * utils/helpers.py:250-270).  Only the tiles that intersect the pixel window [win_x, win_x+win_w) x
 * [win_y, win_y+win_h) are binned and blended; out_color is [3, win_h, win_w].  The tile grid, the
 * camera and every Gaussian's tile rect are those of the FULL width x height frame, so each pixel of
 * the window has exactly the value the full render would give it.  num_rendered counts the window's
 * tile instances.  Two more conveniences for GaussianCity's call pattern, valid in every forward:
 * opacities == NULL renders every Gaussian opaque (opacity 1) and rotations == NULL (with scales)
 * uses the identity rotation -- bit-identical to passing ones / (1,0,0,0), without materialising
 * them (utils/helpers.py:226-247).  gcr_rasterizer_backward_window is the matching backward:
 * dL_dpix is [3, win_h, win_w]; dL_dopacity / dL_drot may be NULL. */
GCR_API int gcr_rasterizer_forward_window(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                           gcr_alloc_fn binningBuffer, void* binning_ctx,
                           gcr_alloc_fn imageBuffer, void* image_ctx,
                           int P, int D, int M, const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                           float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                           int* radii, int debug,
                           int win_x, int win_y, int win_w, int win_h, void* cuda_stream);
GCR_API int gcr_rasterizer_backward_window(int P, int D, int M, int R, const float* background, int width,
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
                            int win_x, int win_y, int win_w, int win_h, void* cuda_stream);

/* Returns 0, or < 0 on error.  Every element of every gradient array is written (zeros for
 * culled Gaussians): callers need not pre-zero them.  dL_dconic ([P,2,2]), dL_dsh, dL_dscale,
 * dL_drot may be NULL.  Single-stripe frames only (shard_count must be 1): a striped frame goes
 * through the two halves below. */
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

/* The two halves of the backward, split at the only cross-tile coupling (the per-Gaussian sums
 * over tiles).  An accumulator is [P,12] fp32 (48 B per Gaussian: dmean2D.xy, dconic.xyw,
 * dopacity, dcolor.rgb, 3 pad), 16-byte aligned.
 *
 * _blend adds the gradients of this rank's stripe into accumulators.  `accumulators` is a HOST
 * array of n_accumulators device pointers:
 *   n_accumulators == 1            every Gaussian's sums go to accumulators[0] (single GPU, or a
 *                                  striped frame whose partial [P,12] sums are reduced by a
 *                                  collective afterwards);
 *   n_accumulators == shard_count  accumulators[r] is rank r's accumulator, peer-mapped into this
 *                                  process (gcr_peer_open): each Gaussian's sums go straight to
 *                                  the rank that OWNS it (the stripe holding its centre row), over
 *                                  NVLink, inside the blend kernel.  Owners then need only a
 *                                  gcr_peer_barrier before _geometry -- no reduce-scatter.
 * The stripe bounds of the forward travel inside geom_buffer.  window: NULL, or a HOST array
 * {x, y, w, h} = the pixel window of a gcr_rasterizer_forward_window frame (dL_dpix is then [3,h,w]).
 * zero_first != 0 zeroes accumulators[n == 1 ? 0 : shard_rank] before accumulating (not wanted for
 * peer-shared accumulators: _geometry(clear_accumulator) leaves them zeroed instead).
 *
 * _geometry turns accumulator rows into the public gradients for the Gaussians this rank owns
 * within [range_start, range_start + range_count) (range_count < 0: all).  striped == 0: rows of
 * the other Gaussians are written as zeros (reference semantics); striped != 0: they are left
 * untouched (their owners write them).  dL_packed (optional, [P,24] fp32, 32-byte aligned): when
 * given, everything except dL_dsh / dL_dconic goes into ONE 96-byte row per Gaussian instead of the
 * seven separate arrays (floats: mean3D 0:3 | opacity 3 | scale 4:7 | mean2D 7:10 | colour 10:13 |
 * cov3D 13:19 | pad | rotation 20:24) -- on striped frames a rank's Gaussians are scattered over
 * the index range, and one row of whole 32-byte sectors replaces seven partial-sector writes. */
GCR_API int gcr_rasterizer_backward_blend(int P, int R, const float* background, int width, int height,
                                  char* geom_buffer, char* binning_buffer, char* image_buffer,
                                  const float* dL_dpix, float* const* accumulators,
                                  int n_accumulators, int zero_first, int remote_scalar_atomics,
                                  int debug, int shard_rank, int shard_count, const int* window,
                                  void* cuda_stream);
GCR_API int gcr_rasterizer_backward_geometry(int P, int D, int M, const float* means3D, const float* shs,
                                     const float* scales, float scale_modifier,
                                     const float* rotations, const float* cov3D_precomp,
                                     const float* viewmatrix, const float* projmatrix,
                                     const float* campos, int width, int height, float tan_fovx,
                                     float tan_fovy, const int* radii, char* geom_buffer,
                                     float* accumulator, float* dL_dmean2D, float* dL_dconic,
                                     float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                                     float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                                     float* dL_drot, int debug, int range_start, int range_count,
                                     int shard_rank, int striped, int clear_accumulator,
                                     float* dL_packed, void* cuda_stream);

/* present: P bytes (bool). Returns 0 or < 0. */
GCR_API int gcr_rasterizer_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, void* cuda_stream);

/* ---- tile-row stripes of a multi-GPU frame ----------------------------------------
 * Balanced partition: a geometry-only pass over all Gaussians counts tile instances per tile
 * row and cuts the rows into shard_count contiguous stripes of about equal instance count.
 * Deterministic: every rank that runs it on the same inputs obtains the same bounds, so no
 * communication is needed.  workspace: (ceil(height/16) + 1) * 4 bytes, device;
 * stripe_bounds_out: shard_count + 1 ints, device.  Nothing is copied to the host. */
GCR_API int gcr_stripe_partition(int P, const float* means3D, const float* scales, float scale_modifier,
                         const float* rotations, const float* cov3D_precomp,
                         const float* viewmatrix, const float* projmatrix, int width, int height,
                         float tan_fovx, float tan_fovy, int shard_count, void* workspace,
                         int* stripe_bounds_out, void* cuda_stream);

/* ---- peer memory: accumulators other GPUs of the node add into over NVLink --------------------
 * gcr_peer_alloc: device allocation (zero-filled) + an inter-process handle to publish to the
 * other ranks; gcr_peer_open maps another rank's allocation into this process (the device must
 * be able to access the peer: NVLink / NVSwitch);  gcr_peer_barrier: all `world` ranks meet on
 * the stream -- flag_arrays is a HOST array of `world` device pointers, entry r = rank r's flag
 * array (>= world uint32, zero-initialised, peer-mapped); epochs must increase by 1 per call. */
GCR_API int gcr_peer_alloc(size_t bytes, void** device_ptr, unsigned char* handle_out);
GCR_API int gcr_peer_open(const unsigned char* handle, void** device_ptr);
GCR_API int gcr_peer_close(void* device_ptr);
GCR_API int gcr_peer_free(void* device_ptr);
GCR_API int gcr_peer_barrier(void* const* flag_arrays, int rank, int world, unsigned int epoch,
                     void* cuda_stream);

/* ---- introspection (tests / parity harness only) ------------------------------------------
 * Byte offsets of the internal arrays inside the three opaque buffers, so a test can view
 * them without the library exposing its layout as API.  `which` values below; returns
 * (size_t)-1 for an unknown id. */
enum {
  GCR_GEOM_DEPTH_SORTED_KEYS = 0, /* u32[n_vis] depth keys of the rendered Gaussians, sorted */
  GCR_GEOM_TILES_TOUCHED = 1,     /* u32[P]  tiles inside this rank's stripe */
  GCR_GEOM_RECORDS = 2,           /* 48 B[P] {x,y,A,B | C,o,2ln(255o),- | r,g,b,-} */
  GCR_GEOM_CLAMPED = 3,           /* u8[P]   bit ch set if SH colour channel was clamped */
  GCR_GEOM_SORTED_GAUSS = 4,      /* u32[n_vis] Gaussian indices in depth order */
  GCR_GEOM_OFFSETS = 5,           /* u32[n_vis] inclusive scan of tiles_touched in depth order */
  GCR_GEOM_GRAD_ACC = 6,          /* 48 B[P] backward accumulator (single-GPU backward) */
  GCR_GEOM_RADII = 7,             /* i32[P]  internal radii (used when radii == NULL) */
  GCR_GEOM_OWNER = 8,             /* u8[P]   owning rank, 0xFF = culled */
  GCR_GEOM_COUNTERS = 9,          /* u32[..] [0] n_vis, [1] num_rendered, [2] overflow flag,
                                     [8..8+shard_count] tile-row stripe bounds of this frame */
  GCR_GEOM_TOTAL_BYTES = 100,
  GCR_BIN_POINT_LIST = 200,       /* u32[R]  sorted instance -> Gaussian index */
  GCR_BIN_TILE_KEYS = 201,        /* u32[R]  sorted tile ids */
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

/* ---- programmatic dependent launch ------------------------------------------------------------
 * The dependent kernels of a frame (sort passes, scan + emit, tile ranges, both blends, geometry
 * backward) can be launched with CUDA's programmatic stream serialization: each starts with
 * griddepcontrol.wait, so results are unchanged, and the launch latency between them overlaps
 * the predecessor's tail -- what matters for frames of GaussianCity's own size (<= 16 384 points,
 * 13 kernels of 4-10 us).  Process-wide; on = 1 / off = 0; returns the previous setting.  The
 * initial value is GCR_PDL from the environment when set (0 / 1), else on. */
GCR_API int gcr_set_programmatic_launch(int on);

#ifdef __cplusplus
}
#endif
#endif /* GCR_RASTERIZER_H_INCLUDED */
