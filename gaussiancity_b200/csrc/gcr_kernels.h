// gcr_kernels.h -- internal launcher declarations shared by the .cu translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

struct GcrRecord;
struct GcrGradAcc;

// Device-side counters of one forward call (uint32 words at the start of a 256-byte block in
// the geometry buffer): nothing downstream of the preprocess needs a count on the host.
enum { GCR_CNT_NVIS = 0, GCR_CNT_R = 1, GCR_CNT_OVERFLOW = 2, GCR_CNT_BWD_TICKET = 3,
       GCR_CNT_TOTAL_TILES64 = 4 /* 2 words */,
       GCR_CNT_STRIPE_BOUNDS = 8 /* int[GCR_MAX_RANKS + 1] */, GCR_CNT_WORDS = 64 };

constexpr int GCR_MAX_RANKS = 16;          // tile-row stripes / peers of one frame
constexpr uint8_t GCR_NO_OWNER = 0xFF;     // owner byte of a Gaussian nobody renders

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
  // tile-row stripes: rank r owns rows [bounds[r], bounds[r+1]); bounds == nullptr (shard_count
  // must be 1) means one stripe covering every row.  DEVICE pointer: no host round trip.
  int shard_rank, shard_count;
  const int* stripe_bounds;
  // tile window [win_col0, win_col1) x [win_row0, win_row1): only tiles inside it are binned (a
  // crop folded into the rasterizer; the full grid by default)
  int win_col0, win_col1, win_row0, win_row1;
  bool prefiltered;
  // outputs
  int* radii;               // global screen-space radius (0 = culled everywhere)
  uint32_t* tiles_touched;  // tiles inside this rank's stripe
  uint32_t* depth_keys;     // 0xFFFFFFFF unless the Gaussian touches this rank's stripe
  GcrRecord* records;
  uint2* rects;             // stripe-clipped tile rect, packed (written when tiles_touched > 0)
  uint8_t* clamped;
  uint8_t* owner;           // rank whose stripe holds the centre row (GCR_NO_OWNER if culled)
  unsigned long long* total_tiles;  // zeroed; += sum of tiles_touched (= num_rendered)
  // balanced stripes (deferred mode) / standalone partition
  unsigned long long* packed_rects;  // [P] global rect + centre row, 0 = culled
  uint32_t* depth_in;                // [P] depth keys in index order (deferred mode: compacted by the select stage
  uint32_t* sorted_init;             //     into depth_keys / sorted_init = the depth sort's input pairs)
  uint32_t* n_vis_out;               // number of Gaussians that reach this rank's stripe
  uint32_t* select_ticket;           // zeroed
  unsigned long long* select_status; // [ceil(P / 1024)] zeroed
  uint32_t* row_hist;                // [grid_y + 1] zeroed: instances per tile row, +1 = done-CTA ticket
  int* stripe_bounds_out;            // [shard_count + 1]
  float* dbg_cov3D;  // optional [P,6]
};

// Forward preprocessing.  deferred == false: one kernel (projection + colour of the rendered
// Gaussians).  deferred == true (balanced stripes): projection of every Gaussian + row histogram
// + stripe cut, then gcr_launch_stripe_select: owner / clipped rect / colours of this stripe's
// Gaussians.
cudaError_t gcr_launch_project(const GcrPreprocessArgs& a, bool deferred, cudaStream_t stream);
cudaError_t gcr_launch_stripe_select(const GcrPreprocessArgs& a, cudaStream_t stream);
cudaError_t gcr_launch_check_frustum(int P, const float* means3D, const float* viewmatrix, bool* present,
                                     cudaStream_t stream);

// Balanced tile-row stripes as a standalone pass: a geometry-only sweep over all
// Gaussians (44 B each: no colour) accumulates the number of tile instances per tile row; the last
// CTA cuts the rows into `shard_count` contiguous stripes of about equal instance count and writes
// a.stripe_bounds_out[0..shard_count] (device).  a.row_hist is grid_y + 1 zeroed words.
cudaError_t gcr_launch_stripe_partition(const GcrPreprocessArgs& a, cudaStream_t stream);

// ---- sort / emit / ranges (binning.cu) --------------------------------------------------------
// Stable LSD radix sort of (key,value) u32 pairs on key bits [0, end_bit): ceil(end_bit/8)
// passes ping-pong from (keys_in, vals_in) to (keys_out, vals_out) and back, so the result is in
// the `in` pair after an even number of passes and in the `out` pair after an odd number.
// n_max = host-known upper bound of the element count, n_ptr = device count (null: n_max).
// depth_mode: values of the first pass are generated as 0..n-1 (vals_in unread), elements with
// key 0xFFFFFFFF are dropped by the first pass, and *n_out (device, required) receives the number
// of survivors, which is what the remaining passes sort.  hist_done: the global digit
// histograms in the workspace were already accumulated by the producer of the keys.
// The workspace must be zeroed before the histograms are accumulated.
size_t gcr_sort_workspace_bytes(size_t n_max);
uint32_t* gcr_sort_ghist(void* sort_workspace);   // [4][256] digit histograms inside it
cudaError_t gcr_launch_radix_sort(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                                  uint32_t* vals_out, size_t n_max, const uint32_t* n_ptr,
                                  int end_bit, bool depth_mode, bool hist_done, uint32_t* n_out,
                                  void* workspace, cudaStream_t stream);

// Inclusive scan of the tile counts in depth order fused with the emission of (tile id, gaussian)
// pairs (duplicateWithKeys, rasterizer_impl.cu:66-99 + :228-231), clipped to the rank's stripe;
// accumulates the tile sort's digit histograms and writes counters[GCR_CNT_R / _OVERFLOW].
struct GcrEmitLaunch {
  uint32_t n_max;
  const uint32_t* n_vis;
  const uint32_t* sorted_gauss;
  const uint2* rects;      // packed stripe-clipped tile rects from the preprocess
  int grid_x;
  uint32_t* tile_keys;
  uint32_t* gauss_vals;
  uint32_t cap;
  void* workspace;         // gcr_emit_workspace_bytes(n_max), zeroed
  uint32_t* ghist_tile;    // gcr_sort_ghist(tile sort workspace), zeroed
  int tile_end_bit;
  uint32_t* counters;
  uint32_t* offsets_out;   // optional
};
size_t gcr_emit_workspace_bytes(size_t n_max);
cudaError_t gcr_launch_emit_scan(const GcrEmitLaunch& l, cudaStream_t stream);

// ranges[tile] = [start,end) from sorted tile keys (identifyTileRanges, rasterizer_impl.cu:104-124)
cudaError_t gcr_launch_tile_ranges(uint32_t n_max, const uint32_t* n_ptr, const uint32_t* sorted_keys,
                                   uint2* ranges, cudaStream_t stream);

// ---- blend (blend_fwd.cu / blend_bwd.cu) -----------------------------------------------------
struct GcrBlendArgs {
  int W, H, grid_x, grid_y;
  // one CTA per tile of the host-known tile window (origin tile_x0, tile_y0; tiles_x x tiles_y),
  // or, with `stripe`, per tile of rows [stripe[0], stripe[1]) (grid covers every row)
  int tile_x0, tile_y0, tiles_x, tiles_y;
  // pixel window: out_color / dL_dpix are [3, ph, pw] images holding pixels [px0, px0+pw) x [py0, py0+ph)
  int px0, py0, pw, ph;
  const int* stripe;           // device {row0,row1} or null
  const uint2* ranges;
  const uint32_t* point_list;  // sorted instance -> Gaussian index
  const GcrRecord* records;    // per-Gaussian records, gathered through point_list
  const uint8_t* owner;        // per-Gaussian owner rank (backward with n_acc > 1)
  const float* bg;  // [3]
  float* final_T;          // [H*W]
  uint32_t* n_contrib;     // [H*W]
  float* out_color;        // [3,H,W]
  // backward only
  const float* dL_dpix;    // [3,H,W]
  // per-Gaussian accumulators, indexed by the owner rank stored in the record: acc[owner][gidx].
  // Single GPU: acc[0] only.  Tile-sharded: peer-mapped buffers, reductions travel over NVLink.
  GcrGradAcc* acc[GCR_MAX_RANKS];
  int n_acc;
  int self_rank;
  int remote_scalar;       // 1: scalar atomics for accumulators of other ranks
};
cudaError_t gcr_launch_blend_fwd(const GcrBlendArgs& a, cudaStream_t stream);
cudaError_t gcr_launch_blend_bwd(const GcrBlendArgs& a, cudaStream_t stream);

// ---- backward preprocess (preprocess_bwd.cu) --------------------------------------------------
struct GcrPreprocessBwdArgs {
  int P, D, M;
  int range_start, range_count;  // Gaussians [range_start, range_start+range_count) are processed
  const float* means3D;
  const uint8_t* owner;          // from the forward
  int my_rank;                   // Gaussians with owner == my_rank are differentiated
  bool zero_unowned;             // true: write zeros for the others (reference semantics at N = 1);
                                 // false: leave their rows untouched (tile-sharded: other ranks own them)
  bool clear_acc;                // zero each accumulator entry after reading it (persistent buffers)
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
  GcrGradAcc* grad_acc;
  uint32_t* ticket;              // chunk dispenser (one word; the launcher zeroes it)
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
  float* packed_out;   // [P,24] or null: all of the above except dL_dsh / dL_dconic in one 96-byte row
};
cudaError_t gcr_launch_preprocess_bwd(const GcrPreprocessBwdArgs& a, cudaStream_t stream);

// ---- peer memory (peer.cu) ---------------------------------------------------------------------
// All ranks of a frame meet: rank `rank` stores `epoch` into slot [rank] of every peer's flag
// array (system-scope release) and waits until its own slots [0..world) have reached `epoch`.
struct GcrPeerFlags { uint32_t* p[GCR_MAX_RANKS]; };
cudaError_t gcr_launch_peer_barrier(const GcrPeerFlags& flags, int rank, int world, uint32_t epoch,
                                    cudaStream_t stream);
