// gcr_kernels.h -- internal launcher declarations shared by the .cu translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

struct GcrRecord;
struct GcrGradAcc;

struct GcrPreprocessArgs {
  int P, D, M;
  const float* means3D;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* cov3D_precomp;
  const float* colors_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  int W, H;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int grid_x, grid_y;
  int shard_rank, shard_count;
  bool prefiltered;
  // outputs
  int* radii;
  uint32_t* tiles_touched;
  uint32_t* depth_keys;
  GcrRecord* records;
  uint8_t* clamped;
  float* dbg_cov3D;  // optional [P,6]
};

void gcr_launch_preprocess_fwd(const GcrPreprocessArgs& a, cudaStream_t stream);
void gcr_launch_check_frustum(int P, const float* means3D, const float* viewmatrix, bool* present,
                              cudaStream_t stream);

// ---- scan / sort / binning (binning.cu) ------------------------------------------------------
// Inclusive scan of in[0..n) (optionally gathered through `gather`: in[gather[i]]) -> out.
size_t gcr_scan_workspace_bytes(size_t n);
void gcr_launch_inclusive_scan(const uint32_t* in, const uint32_t* gather, uint32_t* out, size_t n,
                               void* workspace, cudaStream_t stream);

// Stable LSD radix sort of (key,value) u32 pairs on key bits [0, end_bit). Ping-pongs between
// (keys_a, vals_a) and (keys_b, vals_b); returns 0 if the result is in the a buffers, 1 if in b.
// When vals_iota is true the values of the first pass are generated as 0..n-1 (vals_a unread).
size_t gcr_sort_workspace_bytes(size_t n);
int gcr_launch_radix_sort(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                          size_t n, int end_bit, bool vals_iota, void* workspace,
                          cudaStream_t stream);

// Emit (tile id, gaussian index) pairs in depth-sorted Gaussian order (duplicateWithKeys,
// rasterizer_impl.cu:66-99, restricted to owned tile rows).
void gcr_launch_emit_pairs(int P, const uint32_t* sorted_gauss, const uint32_t* offsets_incl,
                           const uint32_t* tiles_touched, const GcrRecord* records,
                           const int* radii, int grid_x, int grid_y, int shard_rank,
                           int shard_count, uint32_t* tile_keys, uint32_t* gauss_vals,
                           cudaStream_t stream);

// ranges[tile] = [start,end) from sorted tile keys (identifyTileRanges, rasterizer_impl.cu:104-124);
// also gathers the per-instance records: inst[i] = records[point_list[i]].
void gcr_launch_ranges_and_gather(size_t R, const uint32_t* sorted_tile_keys,
                                  const uint32_t* point_list, const GcrRecord* records,
                                  uint2* ranges, GcrRecord* inst, cudaStream_t stream);

// ---- blend (blend_fwd.cu / blend_bwd.cu) -----------------------------------------------------
struct GcrBlendArgs {
  int W, H, grid_x, grid_y;
  int shard_rank, shard_count;
  const uint2* ranges;
  const GcrRecord* inst;
  const float* bg;  // [3]
  float* final_T;          // [H*W]
  uint32_t* n_contrib;     // [H*W]
  float* out_color;        // [3,H,W]
  // backward only
  const float* dL_dpix;    // [3,H,W]
  GcrGradAcc* grad_acc;    // [P]
};
void gcr_launch_blend_fwd(const GcrBlendArgs& a, cudaStream_t stream);
void gcr_launch_blend_bwd(const GcrBlendArgs& a, cudaStream_t stream);

// ---- backward preprocess (preprocess_bwd.cu) --------------------------------------------------
struct GcrPreprocessBwdArgs {
  int P, D, M;
  int range_start, range_count;  // Gaussians [range_start, range_start+range_count) are processed
  const float* means3D;
  const int* radii;
  const float* shs;
  const uint8_t* clamped;
  const float* scales;
  const float* rotations;
  float scale_modifier;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float focal_x, focal_y, tan_fovx, tan_fovy;
  const GcrGradAcc* grad_acc;
  // outputs (every element written exactly once; no pre-zeroing required)
  float* dL_dmean2D;   // [P,3]
  float* dL_dconic;    // [P,4] or null
  float* dL_dopacity;  // [P]
  float* dL_dcolor;    // [P,3]
  float* dL_dmean3D;   // [P,3]
  float* dL_dcov3D;    // [P,6]
  float* dL_dsh;       // [P,M,3] or null
  float* dL_dscale;    // [P,3] or null
  float* dL_drot;      // [P,4] or null
};
void gcr_launch_preprocess_bwd(const GcrPreprocessBwdArgs& a, cudaStream_t stream);
