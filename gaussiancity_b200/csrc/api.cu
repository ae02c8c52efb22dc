// api.cu -- C ABI (include/gcr_rasterizer.h) and host orchestration of the sm_100a pipeline.
//
// Forward  (replaces Rasterizer::forward, DGR/cuda_rasterizer/rasterizer_impl.cu:178-283):
//   preprocess -> depth radix sort of the P Gaussians -> scan of tile counts in depth order
//   -> [host reads R] -> emit (tile, gaussian) pairs -> stable radix split on tile id
//   -> tile ranges + gather of per-instance records -> TMA-staged per-tile blend.
// Backward (replaces Rasterizer::backward, rasterizer_impl.cu:287-339):
//   zero the 48 B/Gaussian accumulator -> TMA-staged per-tile gradient blend -> fused
//   per-Gaussian geometry backward that writes every output exactly once.
// Everything is launched on the caller's stream; the only host synchronisation is the 4-byte
// read of R that sizes the binning buffer (the reference has the same one, :235-238).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/gcr_rasterizer.h"
#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

thread_local std::string g_last_error;
thread_local float* g_dbg_cov3d = nullptr;  // test hook: next forward also writes cov3D here

// ---- optional per-stage timing with CUDA events recorded on the caller's stream (no host
// sync is added; bench.py reads the elapsed times after it has synchronised) -----------------
enum Stage {
  ST_PREPROCESS = 0, ST_DEPTH_SORT, ST_SCAN, ST_EMIT, ST_TILE_SORT, ST_RANGES_GATHER, ST_BLEND_FWD,
  ST_BWD_ZERO, ST_BLEND_BWD, ST_GEOM_BWD, ST_COUNT
};
const char* const kStageNames[ST_COUNT] = {"preprocess_fwd", "depth_sort", "scan", "emit_pairs",
                                           "tile_sort", "ranges_gather", "blend_fwd", "bwd_zero",
                                           "blend_bwd", "geometry_bwd"};
constexpr int kProfSlots = 64;  // ring of profiled forward(+backward) calls
struct Profiler {
  bool enabled = false;
  bool created = false;
  int slot = -1;  // advanced by every forward call
  cudaEvent_t ev[kProfSlots][ST_COUNT][2] = {};
  bool have[kProfSlots][ST_COUNT] = {};
};
Profiler g_prof;  // process-wide; profiling is a single-threaded bench facility

void prof_next_call() {
  if (!g_prof.enabled) return;
  g_prof.slot = (g_prof.slot + 1) % kProfSlots;
  for (int i = 0; i < ST_COUNT; ++i) g_prof.have[g_prof.slot][i] = false;
}

void prof_mark(int stage, int which, cudaStream_t stream) {
  if (!g_prof.enabled) return;
  if (!g_prof.created) {
    for (int k = 0; k < kProfSlots; ++k)
      for (int i = 0; i < ST_COUNT; ++i)
        for (int j = 0; j < 2; ++j) cudaEventCreate(&g_prof.ev[k][i][j]);
    g_prof.created = true;
  }
  if (g_prof.slot < 0) g_prof.slot = 0;
  cudaEventRecord(g_prof.ev[g_prof.slot][stage][which], stream);
  if (which == 1) g_prof.have[g_prof.slot][stage] = true;
}

// host-side phase timing to stderr when GCR_HOST_TIMING=1 (diagnostics only)
struct HostTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  std::string log;
  HostTimer() : on(getenv("GCR_HOST_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof(buf), " %s=%.3f", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    log += buf;
    t0 = t1;
  }
  ~HostTimer() { if (on) fprintf(stderr, "[gcr host ms]%s\n", log.c_str()); }
};

int fail(const std::string& msg) {
  g_last_error = msg;
  return -1;
}

#define GCR_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

#define GCR_CHECK_LAUNCH(what, debug, stream)                                              \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ == cudaSuccess && (debug)) e__ = cudaStreamSynchronize(stream);                \
    if (e__ != cudaSuccess)                                                                \
      return fail(std::string("[CUDA ERROR] in ") + (what) + ": " + cudaGetErrorString(e__)); \
  } while (0)

int tile_bits(int tiles) {
  int b = 0;
  while ((1ll << b) < (long long)tiles) ++b;
  return b;  // ids 0..tiles-1 fit in b bits
}

// ---- deterministic carving of the three opaque buffers (cf. *State::fromChunk,
// rasterizer_impl.cu:134-174; the layout itself is ours) -----------------------------------
struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    off = gcr_align_up(off, 256);
    const size_t o = off;
    off += bytes;
    return o;
  }
};

struct GeomLayout {
  size_t keys_a, keys_b, vals_a, vals_b, tiles, records, clamped, offsets, grad_acc, radii,
      scan_ws, sort_ws, total;
  explicit GeomLayout(size_t P) {
    Carver c;
    keys_a = c.take(4 * P);
    keys_b = c.take(4 * P);
    vals_a = c.take(4 * P);
    vals_b = c.take(4 * P);
    tiles = c.take(4 * P);
    records = c.take(48 * P);
    clamped = c.take(P);
    offsets = c.take(4 * P);
    grad_acc = c.take(48 * P);
    radii = c.take(4 * P);
    scan_ws = c.take(gcr_scan_workspace_bytes(P));
    sort_ws = c.take(gcr_sort_workspace_bytes(P));
    total = gcr_align_up(c.off, 256) + 256;
  }
};

struct BinLayout {
  size_t keys_a, keys_b, vals_a, vals_b, inst, sort_ws, total;
  explicit BinLayout(size_t R) {
    Carver c;
    keys_a = c.take(4 * R);
    keys_b = c.take(4 * R);
    vals_a = c.take(4 * R);
    vals_b = c.take(4 * R);
    inst = c.take(48 * R);
    sort_ws = c.take(gcr_sort_workspace_bytes(R));
    total = gcr_align_up(c.off, 256) + 256;
  }
};

struct ImgLayout {
  size_t final_T, n_contrib, ranges, total;
  ImgLayout(size_t npix, size_t tiles) {
    Carver c;
    final_T = c.take(4 * npix);
    n_contrib = c.take(4 * npix);
    ranges = c.take(8 * tiles);
    total = gcr_align_up(c.off, 256) + 256;
  }
};

char* align256(char* p) {
  return reinterpret_cast<char*>(gcr_align_up(reinterpret_cast<uintptr_t>(p), 256));
}

// number of LSD passes the tile sort performs -> which ping-pong side holds the result
int tile_sort_result_side(int tiles) {
  int bits = tile_bits(tiles);
  if (bits <= 0) bits = 1;
  return ((bits + 7) / 8) & 1;
}

}  // namespace

extern "C" {

int gcr_abi_version(void) { return GCR_ABI_VERSION; }
const char* gcr_last_error(void) { return g_last_error.c_str(); }

int gcr_rasterizer_forward(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                           gcr_alloc_fn binningBuffer, void* binning_ctx,
                           gcr_alloc_fn imageBuffer, void* image_ctx, int P, int D, int M,
                           const float* background, int width, int height, const float* means3D,
                           const float* shs, const float* colors_precomp, const float* opacities,
                           const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix,
                           const float* projmatrix, const float* cam_pos, float tan_fovx,
                           float tan_fovy, int prefiltered, float* out_color, int* radii,
                           int debug, int shard_rank, int shard_count, void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  prof_next_call();
  if (width <= 0 || height <= 0) return fail("image size must be positive");
  if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
    return fail("invalid tile-row shard (rank, count)");
  if (colors_precomp == nullptr && shs == nullptr)
    return fail("provide either SHs or precomputed colours");
  if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
    return fail("provide either scale/rotation or a precomputed 3D covariance");
  if (D < 0 || D > 3 || (colors_precomp == nullptr && (D + 1) * (D + 1) > M))
    return fail("SH degree does not fit the coefficient count");
  if (geometryBuffer == nullptr || binningBuffer == nullptr || imageBuffer == nullptr)
    return fail("buffer callbacks must not be NULL");
  if (means3D == nullptr || opacities == nullptr || viewmatrix == nullptr || projmatrix == nullptr ||
      background == nullptr || out_color == nullptr)
    return fail("means3D, opacities, viewmatrix, projmatrix, background and out_color must not be NULL");
  if (colors_precomp == nullptr && cam_pos == nullptr) return fail("cam_pos must not be NULL with SHs");
  // vector loads: rotations are read as float4, SH rows as 32-byte (M = 16) or 16-byte words
  if (rotations != nullptr && (reinterpret_cast<uintptr_t>(rotations) & 15u) != 0)
    return fail("rotations must be 16-byte aligned");
  if (colors_precomp == nullptr && ((M * 3) % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(shs) & (((M * 3) % 8) == 0 ? 31u : 15u)) != 0)
    return fail("shs must be 32-byte aligned (16-byte when 3*M is not a multiple of 8)");

  const int grid_x = (width + GCR_TILE_X - 1) / GCR_TILE_X;
  const int grid_y = (height + GCR_TILE_Y - 1) / GCR_TILE_Y;
  const int tiles = grid_x * grid_y;
  const size_t npix = (size_t)width * height;

  HostTimer ht;
  const GeomLayout gl((size_t)P);
  char* gptr = geometryBuffer(geometry_ctx, gl.total);
  if (gptr == nullptr) return fail("geometry buffer allocation failed");
  gptr = align256(gptr);
  const ImgLayout il(npix, (size_t)tiles);
  char* iptr = imageBuffer(image_ctx, il.total);
  if (iptr == nullptr) return fail("image buffer allocation failed");
  iptr = align256(iptr);

  ht.lap("alloc_geom_img");
  uint32_t* keys_a = reinterpret_cast<uint32_t*>(gptr + gl.keys_a);
  uint32_t* keys_b = reinterpret_cast<uint32_t*>(gptr + gl.keys_b);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(gptr + gl.vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(gptr + gl.vals_b);
  uint32_t* tiles_touched = reinterpret_cast<uint32_t*>(gptr + gl.tiles);
  GcrRecord* records = reinterpret_cast<GcrRecord*>(gptr + gl.records);
  uint8_t* clamped = reinterpret_cast<uint8_t*>(gptr + gl.clamped);
  uint32_t* offsets = reinterpret_cast<uint32_t*>(gptr + gl.offsets);
  if (radii == nullptr) radii = reinterpret_cast<int*>(gptr + gl.radii);

  // 1. per-Gaussian preprocessing
  GcrPreprocessArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.P = P; pa.D = D; pa.M = M;
  pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier;
  pa.rotations = rotations; pa.opacities = opacities; pa.shs = shs;
  pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
  pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix; pa.campos = cam_pos;
  pa.W = width; pa.H = height; pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy;
  pa.focal_y = height / (2.0f * tan_fovy);   // rasterizer_impl.cu:189-190
  pa.focal_x = width / (2.0f * tan_fovx);
  pa.grid_x = grid_x; pa.grid_y = grid_y;
  pa.shard_rank = shard_rank; pa.shard_count = shard_count;
  pa.prefiltered = prefiltered != 0;
  pa.radii = radii; pa.tiles_touched = tiles_touched; pa.depth_keys = keys_a;
  pa.records = records; pa.clamped = clamped; pa.dbg_cov3D = g_dbg_cov3d;
  g_dbg_cov3d = nullptr;
  prof_mark(ST_PREPROCESS, 0, stream);
  gcr_launch_preprocess_fwd(pa, stream);
  GCR_CHECK_LAUNCH("preprocess_fwd", debug, stream);
  prof_mark(ST_PREPROCESS, 1, stream);

  // 2. stable depth sort of the Gaussians (4 passes: result back in the a buffers)
  prof_mark(ST_DEPTH_SORT, 0, stream);
  const int side = gcr_launch_radix_sort(keys_a, vals_a, keys_b, vals_b, (size_t)P, 32, true,
                                         gptr + gl.sort_ws, stream);
  GCR_CHECK_LAUNCH("depth sort", debug, stream);
  prof_mark(ST_DEPTH_SORT, 1, stream);
  uint32_t* sorted_gauss = side ? vals_b : vals_a;

  // 3. offsets of each depth-ordered Gaussian's instances
  prof_mark(ST_SCAN, 0, stream);
  gcr_launch_inclusive_scan(tiles_touched, sorted_gauss, offsets, (size_t)P, gptr + gl.scan_ws,
                            stream);
  GCR_CHECK_LAUNCH("tile-count scan", debug, stream);
  prof_mark(ST_SCAN, 1, stream);

  // 4. R to the host (sizes the binning buffer; same sync as rasterizer_impl.cu:235-238)
  ht.lap("launch_pre");
  static thread_local uint32_t* pinned_R = nullptr;   // pinned: the D2H copy is a true async DMA
  if (pinned_R == nullptr) GCR_CUDA_OK(cudaHostAlloc(&pinned_R, sizeof(uint32_t), cudaHostAllocDefault));
  GCR_CUDA_OK(cudaMemcpyAsync(pinned_R, offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              stream));
  GCR_CUDA_OK(cudaStreamSynchronize(stream));
  const uint32_t num_rendered_u = *pinned_R;
  ht.lap("sync_R");
  if (num_rendered_u > 0x7fffffffu) return fail("num_rendered exceeds int32");
  const size_t R = num_rendered_u;

  const BinLayout bl(R);
  char* bptr = binningBuffer(binning_ctx, bl.total);
  if (bptr == nullptr) return fail("binning buffer allocation failed");
  bptr = align256(bptr);
  ht.lap("alloc_bin");
  uint32_t* tk_a = reinterpret_cast<uint32_t*>(bptr + bl.keys_a);
  uint32_t* tk_b = reinterpret_cast<uint32_t*>(bptr + bl.keys_b);
  uint32_t* tv_a = reinterpret_cast<uint32_t*>(bptr + bl.vals_a);
  uint32_t* tv_b = reinterpret_cast<uint32_t*>(bptr + bl.vals_b);
  GcrRecord* inst = reinterpret_cast<GcrRecord*>(bptr + bl.inst);
  uint2* ranges = reinterpret_cast<uint2*>(iptr + il.ranges);

  GCR_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)tiles, stream));
  if (R > 0) {
    // 5. emit (tile, gaussian) pairs in depth order, 6. stable split by tile id
    prof_mark(ST_EMIT, 0, stream);
    gcr_launch_emit_pairs(P, sorted_gauss, offsets, tiles_touched, records, radii, grid_x, grid_y,
                          shard_rank, shard_count, tk_a, tv_a, stream);
    GCR_CHECK_LAUNCH("emit pairs", debug, stream);
    prof_mark(ST_EMIT, 1, stream);
    prof_mark(ST_TILE_SORT, 0, stream);
    const int tside = gcr_launch_radix_sort(tk_a, tv_a, tk_b, tv_b, R, tile_bits(tiles), false,
                                            bptr + bl.sort_ws, stream);
    GCR_CHECK_LAUNCH("tile sort", debug, stream);
    prof_mark(ST_TILE_SORT, 1, stream);
    const uint32_t* sorted_keys = tside ? tk_b : tk_a;
    const uint32_t* point_list = tside ? tv_b : tv_a;
    // 7. tile ranges + contiguous per-instance records
    prof_mark(ST_RANGES_GATHER, 0, stream);
    gcr_launch_ranges_and_gather(R, sorted_keys, point_list, records, ranges, inst, stream);
    GCR_CHECK_LAUNCH("ranges + gather", debug, stream);
    prof_mark(ST_RANGES_GATHER, 1, stream);
  }

  // 8. per-tile blend
  GcrBlendArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.W = width; ba.H = height; ba.grid_x = grid_x; ba.grid_y = grid_y;
  ba.shard_rank = shard_rank; ba.shard_count = shard_count;
  ba.ranges = ranges; ba.inst = inst; ba.bg = background;
  ba.final_T = reinterpret_cast<float*>(iptr + il.final_T);
  ba.n_contrib = reinterpret_cast<uint32_t*>(iptr + il.n_contrib);
  ba.out_color = out_color;
  prof_mark(ST_BLEND_FWD, 0, stream);
  gcr_launch_blend_fwd(ba, stream);
  GCR_CHECK_LAUNCH("blend_fwd", debug, stream);
  prof_mark(ST_BLEND_FWD, 1, stream);
  ht.lap("launch_post");
  return (int)R;
}

int gcr_rasterizer_backward_blend(int P, int R, const float* background, int width, int height,
                                  char* binning_buffer, char* image_buffer, const float* dL_dpix,
                                  float* grad_acc, int debug, int shard_rank, int shard_count,
                                  void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
    return fail("invalid tile-row shard (rank, count)");
  prof_mark(ST_BWD_ZERO, 0, stream);
  GCR_CUDA_OK(cudaMemsetAsync(grad_acc, 0, sizeof(GcrGradAcc) * (size_t)P, stream));
  prof_mark(ST_BWD_ZERO, 1, stream);
  if (R <= 0) return 0;
  const int grid_x = (width + GCR_TILE_X - 1) / GCR_TILE_X;
  const int grid_y = (height + GCR_TILE_Y - 1) / GCR_TILE_Y;
  const BinLayout bl((size_t)R);
  const ImgLayout il((size_t)width * height, (size_t)grid_x * grid_y);
  char* bptr = align256(binning_buffer);
  char* iptr = align256(image_buffer);
  GcrBlendArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.W = width; ba.H = height; ba.grid_x = grid_x; ba.grid_y = grid_y;
  ba.shard_rank = shard_rank; ba.shard_count = shard_count;
  ba.ranges = reinterpret_cast<const uint2*>(iptr + il.ranges);
  ba.inst = reinterpret_cast<const GcrRecord*>(bptr + bl.inst);
  ba.bg = background;
  ba.final_T = reinterpret_cast<float*>(iptr + il.final_T);
  ba.n_contrib = reinterpret_cast<uint32_t*>(iptr + il.n_contrib);
  ba.dL_dpix = dL_dpix;
  ba.grad_acc = reinterpret_cast<GcrGradAcc*>(grad_acc);
  prof_mark(ST_BLEND_BWD, 0, stream);
  gcr_launch_blend_bwd(ba, stream);
  GCR_CHECK_LAUNCH("blend_bwd", debug, stream);
  prof_mark(ST_BLEND_BWD, 1, stream);
  return 0;
}

int gcr_rasterizer_backward_geometry(int P, int D, int M, const float* means3D, const float* shs,
                                     const float* scales, float scale_modifier,
                                     const float* rotations, const float* cov3D_precomp,
                                     const float* viewmatrix, const float* projmatrix,
                                     const float* campos, int width, int height, float tan_fovx,
                                     float tan_fovy, const int* radii, char* geom_buffer,
                                     const float* grad_acc, float* dL_dmean2D, float* dL_dconic,
                                     float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                                     float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                                     float* dL_drot, int debug, int range_start, int range_count,
                                     void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  if (range_count < 0) { range_start = 0; range_count = P; }
  if (range_start < 0 || range_start + range_count > P) return fail("Gaussian range out of bounds");
  if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
    return fail("provide either scale/rotation or a precomputed 3D covariance");
  if (shs != nullptr && M > 0 && dL_dsh == nullptr) return fail("dL_dsh must not be NULL with SHs");
  if (rotations != nullptr && (reinterpret_cast<uintptr_t>(rotations) & 15u) != 0)
    return fail("rotations must be 16-byte aligned");
  if (shs != nullptr && M > 0 && ((M * 3) % 4) == 0) {
    const uintptr_t mask = ((M * 3) % 8) == 0 ? 31u : 15u;
    if ((reinterpret_cast<uintptr_t>(shs) & mask) != 0 || (reinterpret_cast<uintptr_t>(dL_dsh) & mask) != 0)
      return fail("shs / dL_dsh must be 32-byte aligned (16-byte when 3*M is not a multiple of 8)");
  }
  const GeomLayout gl((size_t)P);
  char* gptr = align256(geom_buffer);
  GcrPreprocessBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P; a.D = D; a.M = M;
  a.range_start = range_start; a.range_count = range_count;
  a.means3D = means3D;
  a.radii = radii != nullptr ? radii : reinterpret_cast<const int*>(gptr + gl.radii);
  a.shs = (M > 0) ? shs : nullptr;
  a.clamped = reinterpret_cast<const uint8_t*>(gptr + gl.clamped);
  a.scales = scales; a.rotations = rotations; a.scale_modifier = scale_modifier;
  a.cov3D_precomp = cov3D_precomp;
  a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos;
  a.focal_y = height / (2.0f * tan_fovy);   // rasterizer_impl.cu:306-307
  a.focal_x = width / (2.0f * tan_fovx);
  a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
  a.grad_acc = reinterpret_cast<const GcrGradAcc*>(grad_acc);
  a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dopacity = dL_dopacity;
  a.dL_dcolor = dL_dcolor; a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D;
  a.dL_dsh = dL_dsh; a.dL_dscale = (scales != nullptr) ? dL_dscale : nullptr;
  a.dL_drot = (scales != nullptr) ? dL_drot : nullptr;
  prof_mark(ST_GEOM_BWD, 0, stream);
  gcr_launch_preprocess_bwd(a, stream);
  GCR_CHECK_LAUNCH("preprocess_bwd", debug, stream);
  prof_mark(ST_GEOM_BWD, 1, stream);
  // reference semantics: with a precomputed covariance the scale/rotation gradients are zeros
  if (scales == nullptr) {
    if (dL_dscale)
      GCR_CUDA_OK(cudaMemsetAsync(dL_dscale + 3 * (size_t)range_start, 0,
                                  sizeof(float) * 3 * (size_t)range_count, stream));
    if (dL_drot)
      GCR_CUDA_OK(cudaMemsetAsync(dL_drot + 4 * (size_t)range_start, 0,
                                  sizeof(float) * 4 * (size_t)range_count, stream));
  }
  return 0;
}

int gcr_rasterizer_backward(int P, int D, int M, int R, const float* background, int width,
                            int height, const float* means3D, const float* shs,
                            const float* colors_precomp, const float* scales,
                            float scale_modifier, const float* rotations,
                            const float* cov3D_precomp, const float* viewmatrix,
                            const float* projmatrix, const float* campos, float tan_fovx,
                            float tan_fovy, const int* radii, char* geom_buffer,
                            char* binning_buffer, char* image_buffer, const float* dL_dpix,
                            float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                            float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                            float* dL_dscale, float* dL_drot, int debug, int shard_rank,
                            int shard_count, void* cuda_stream) {
  (void)colors_precomp;  // colours live in the per-instance records since the forward
  if (P <= 0) return 0;
  const GeomLayout gl((size_t)P);
  float* grad_acc = reinterpret_cast<float*>(align256(geom_buffer) + gl.grad_acc);
  int rc = gcr_rasterizer_backward_blend(P, R, background, width, height, binning_buffer,
                                         image_buffer, dL_dpix, grad_acc, debug, shard_rank,
                                         shard_count, cuda_stream);
  if (rc < 0) return rc;
  return gcr_rasterizer_backward_geometry(P, D, M, means3D, shs, scales, scale_modifier, rotations,
                                          cov3D_precomp, viewmatrix, projmatrix, campos, width,
                                          height, tan_fovx, tan_fovy, radii, geom_buffer, grad_acc,
                                          dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                          dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug,
                                          0, -1, cuda_stream);
}

int gcr_rasterizer_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, void* cuda_stream) {
  (void)projmatrix;  // the reference's test only uses the view-space depth (auxiliary.h:145)
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  gcr_launch_check_frustum(P, means3D, viewmatrix, reinterpret_cast<bool*>(present), stream);
  GCR_CHECK_LAUNCH("check_frustum", 0, stream);
  return 0;
}

void gcr_debug_set_cov3d_out(float* cov3d) { g_dbg_cov3d = cov3d; }

void gcr_profile_enable(int on) {
  g_prof.enabled = on != 0;
  g_prof.slot = -1;
  for (int k = 0; k < kProfSlots; ++k)
    for (int i = 0; i < ST_COUNT; ++i) g_prof.have[k][i] = false;
}
int gcr_profile_stage_count(void) { return ST_COUNT; }
const char* gcr_profile_stage_name(int stage) {
  return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : "";
}
// mean device time of `stage` over the profiled calls since gcr_profile_enable(1)
float gcr_profile_stage_ms(int stage) {
  if (stage < 0 || stage >= ST_COUNT || !g_prof.created) return -1.f;
  double sum = 0.0;
  int n = 0;
  for (int k = 0; k < kProfSlots; ++k) {
    if (!g_prof.have[k][stage]) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.ev[k][stage][0], g_prof.ev[k][stage][1]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    sum += ms;
    ++n;
  }
  return n ? (float)(sum / n) : -1.f;
}

size_t gcr_debug_offset(int which, int P, int R, int width, int height) {
  const int grid_x = (width + GCR_TILE_X - 1) / GCR_TILE_X;
  const int grid_y = (height + GCR_TILE_Y - 1) / GCR_TILE_Y;
  const int tiles = grid_x * grid_y;
  if (which < 200) {
    const GeomLayout gl((size_t)(P > 0 ? P : 0));
    switch (which) {
      case GCR_GEOM_DEPTH_SORTED_KEYS: return gl.keys_a;  // 4 passes -> back in a
      case GCR_GEOM_TILES_TOUCHED: return gl.tiles;
      case GCR_GEOM_RECORDS: return gl.records;
      case GCR_GEOM_CLAMPED: return gl.clamped;
      case GCR_GEOM_SORTED_GAUSS: return gl.vals_a;
      case GCR_GEOM_OFFSETS: return gl.offsets;
      case GCR_GEOM_GRAD_ACC: return gl.grad_acc;
      case GCR_GEOM_RADII: return gl.radii;
      case GCR_GEOM_TOTAL_BYTES: return gl.total;
      default: return (size_t)-1;
    }
  } else if (which < 400) {
    const BinLayout bl((size_t)(R > 0 ? R : 0));
    const int side = tile_sort_result_side(tiles);
    switch (which) {
      case GCR_BIN_POINT_LIST: return side ? bl.vals_b : bl.vals_a;
      case GCR_BIN_TILE_KEYS: return side ? bl.keys_b : bl.keys_a;
      case GCR_BIN_INSTANCES: return bl.inst;
      case GCR_BIN_TOTAL_BYTES: return bl.total;
      default: return (size_t)-1;
    }
  } else {
    const ImgLayout il((size_t)width * height, (size_t)tiles);
    switch (which) {
      case GCR_IMG_FINAL_T: return il.final_T;
      case GCR_IMG_N_CONTRIB: return il.n_contrib;
      case GCR_IMG_RANGES: return il.ranges;
      case GCR_IMG_TOTAL_BYTES: return il.total;
      default: return (size_t)-1;
    }
  }
}

}  // extern "C"
