// api.cu -- C ABI (include/gcr_rasterizer.h) and host orchestration of the sm_100a pipeline.
//
// Forward  (replaces Rasterizer::forward, DGR/cuda_rasterizer/rasterizer_impl.cu:178-283):
//   preprocess (+ per-CTA sum of tile counts = num_rendered) -> depth radix sort of the visible
//   Gaussians (culled ones dropped by its first pass) -> fused scan + emission of (tile,
//   gaussian) pairs -> stable radix split on tile id -> tile ranges -> per-tile blend that
//   gathers its records through point_list.
// Backward (replaces Rasterizer::backward, rasterizer_impl.cu:287-339):
//   zero the 48 B/Gaussian accumulator -> per-tile gradient blend -> fused per-Gaussian geometry
//   backward that writes every output exactly once.
// Everything is launched on the caller's stream.  The host needs ONE number, num_rendered, to
// size the binning buffer (the reference reads it with a blocking cudaMemcpy after its scan,
// :235-238).  Here the preprocess kernel produces it, a 8-byte copy to pinned memory + an event
// follow it on the stream, the depth sort is enqueued behind them, and only then does the host
// wait -- on the event, not the stream -- so the GPU keeps sorting while the host allocates the
// binning buffer and enqueues the rest: no pipeline bubble.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/gcr_rasterizer.h"
#include "gcr_common.cuh"
#include "gcr_kernels.h"

static_assert(GCR_MAX_SHARDS == GCR_MAX_RANKS, "public and internal stripe limits must agree");

// Programmatic dependent launch of the kernel chain (gcr_common.cuh): process-wide switch, initial
// value from GCR_PDL (0 / 1) when set.  ON by default: measured on a B200 (profiles/r02_small_scenes.md)
// forward of a 16 384-point frame 0.108 -> 0.093 ms, fwd+bwd 0.205 -> 0.196 ms, 100 k: -2.4 %, 5 M: -0.4 %;
// which of the backward's two launches carry the attribute makes no measurable difference (all do).
#ifndef GCR_PDL_DEFAULT
#define GCR_PDL_DEFAULT 1
#endif
static std::atomic<int>& pdl_flag() {
  static std::atomic<int> flag{[] {
    const char* e = getenv("GCR_PDL");
    return e != nullptr ? (e[0] != '0' ? 1 : 0) : GCR_PDL_DEFAULT;
  }()};
  return flag;
}
int gcr_pdl_edges() {
  return pdl_flag().load(std::memory_order_relaxed) != 0 ? (GCR_EDGE_FWD | GCR_EDGE_BLEND_BWD | GCR_EDGE_GEOM_BWD) : 0;
}


namespace {

thread_local std::string g_last_error;
thread_local float* g_dbg_cov3d = nullptr;  // test hook: next forward also writes cov3D here

// ---- optional per-stage timing with CUDA events recorded on the caller's stream (no host
// sync is added; bench.py reads the elapsed times after it has synchronised) -----------------
enum Stage {
  ST_PREPROCESS = 0, ST_SELECT, ST_DEPTH_SORT, ST_EMIT, ST_TILE_SORT, ST_RANGES, ST_BLEND_FWD,
  ST_BWD_ZERO, ST_BLEND_BWD, ST_GEOM_BWD, ST_PARTITION, ST_COUNT
};
const char* const kStageNames[ST_COUNT] = {"preprocess_fwd", "stripe_select", "depth_sort", "scan_emit", "tile_sort",
                                           "tile_ranges", "blend_fwd", "bwd_zero", "blend_bwd",
                                           "geometry_bwd", "stripe_partition"};
constexpr int kProfSlots = 64;  // ring of profiled forward(+backward) calls
struct Profiler {
  std::mutex mu;   // guards creation; marking itself is a single-threaded bench facility
  bool enabled = false;
  bool created = false;
  int slot = -1;  // advanced by every forward call
  cudaEvent_t ev[kProfSlots][ST_COUNT][2] = {};
  bool have[kProfSlots][ST_COUNT] = {};
};
Profiler g_prof;

void prof_next_call() {
  if (!g_prof.enabled) return;
  g_prof.slot = (g_prof.slot + 1) % kProfSlots;
  for (int i = 0; i < ST_COUNT; ++i) g_prof.have[g_prof.slot][i] = false;
}

void prof_mark(int stage, int which, cudaStream_t stream) {
  if (!g_prof.enabled) return;
  {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (!g_prof.created) {
      for (int k = 0; k < kProfSlots; ++k)
        for (int i = 0; i < ST_COUNT; ++i)
          for (int j = 0; j < 2; ++j) cudaEventCreate(&g_prof.ev[k][i][j]);
      g_prof.created = true;
    }
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

// launchers return the launch error; in debug mode the stream is synchronised too
#define GCR_LAUNCH(what, expr, debug, stream)                                              \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
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
  size_t keys_a, keys_b, vals_a, vals_b, tiles, records, rects, packed_rects, clamped, owner, offsets, grad_acc,
      radii, zero_begin, counters, row_hist, select_ws, sort_ws, emit_ws, zero_end, total;
  explicit GeomLayout(size_t P) {
    Carver c;
    keys_a = c.take(4 * P);
    keys_b = c.take(4 * P);
    vals_a = c.take(4 * P);
    vals_b = c.take(4 * P);
    tiles = c.take(4 * P);
    records = c.take(48 * P);
    rects = c.take(8 * P);
    packed_rects = c.take(8 * P);
    clamped = c.take(P);
    owner = c.take(P);
    offsets = c.take(4 * P);
    grad_acc = c.take(48 * P);
    radii = c.take(4 * P);
    // one contiguous block zeroed by a single memset per forward: counters + both workspaces
    counters = zero_begin = c.take(GCR_CNT_WORDS * sizeof(uint32_t));
    row_hist = c.take((4096 + 1) * sizeof(uint32_t));   // instances per tile row (+1: done-CTA ticket)
    select_ws = c.take(256 + ((P + 1023) / 1024 + 1) * sizeof(unsigned long long));   // ticket | status
    sort_ws = c.take(gcr_sort_workspace_bytes(P));
    emit_ws = c.take(gcr_emit_workspace_bytes(P));
    zero_end = gcr_align_up(c.off, 256);
    total = zero_end + 256;
  }
};

// point_list lands in vals_a = offset 0: the backward needs no count to find it
struct BinLayout {
  size_t vals_a, keys_a, vals_b, keys_b, sort_ws, total;
  explicit BinLayout(size_t R) {
    Carver c;
    vals_a = c.take(4 * R);
    keys_a = c.take(4 * R);
    vals_b = c.take(4 * R);
    keys_b = c.take(4 * R);
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

// pinned landing zone + event for the one number the host needs per forward; per host thread
// and per device, released when the thread exits
struct HostMailbox {
  unsigned long long* pinned[64] = {};
  cudaEvent_t ev[64] = {};
  ~HostMailbox() {
    for (int d = 0; d < 64; ++d) {
      if (pinned[d] != nullptr) cudaFreeHost(pinned[d]);
      if (ev[d] != nullptr) cudaEventDestroy(ev[d]);
    }
  }
};
thread_local HostMailbox g_mailbox;

bool valid_shard(int rank, int count) { return count >= 1 && count <= GCR_MAX_SHARDS && rank >= 0 && rank < count; }

int forward_impl(gcr_alloc_fn geometryBuffer, void* geometry_ctx, gcr_alloc_fn binningBuffer,
                 void* binning_ctx, gcr_alloc_fn imageBuffer, void* image_ctx, int P, int D, int M,
                 const float* background, int width, int height, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* opacities,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                 float* out_color, int* radii, int debug, int shard_rank, int shard_count,
                 const int* stripe_bounds, int balanced, const int* window, void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  prof_next_call();
  if (width <= 0 || height <= 0) return fail("image size must be positive");
  if (!valid_shard(shard_rank, shard_count)) return fail("invalid tile-row shard (rank, count)");
  if (colors_precomp == nullptr && shs == nullptr)
    return fail("provide either SHs or precomputed colours");
  if (cov3D_precomp == nullptr && scales == nullptr)   // rotations == NULL alone: identity rotation
    return fail("provide either scale/rotation or a precomputed 3D covariance");
  if (D < 0 || D > 3 || (colors_precomp == nullptr && (D + 1) * (D + 1) > M))
    return fail("SH degree does not fit the coefficient count");
  if (colors_precomp == nullptr && M > 16)
    return fail("at most 16 SH coefficients per Gaussian (degree 3) are supported");
  if (geometryBuffer == nullptr || binningBuffer == nullptr || imageBuffer == nullptr)
    return fail("buffer callbacks must not be NULL");
  if (means3D == nullptr || viewmatrix == nullptr || projmatrix == nullptr || background == nullptr ||
      out_color == nullptr)   // opacities == NULL: every Gaussian opaque (opacity 1)
    return fail("means3D, viewmatrix, projmatrix, background and out_color must not be NULL");
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
  if (grid_x > 4095 || grid_y > 4095) return fail("images above 65520 pixels a side are not supported");
  // pixel window (a crop folded into the rasterizer): tiles outside it are neither binned nor blended
  int win[4] = {0, 0, width, height};
  if (window != nullptr) {
    if (shard_count != 1) return fail("a pixel window cannot be combined with tile-row stripes");
    for (int k = 0; k < 4; ++k) win[k] = window[k];
    if (win[0] < 0 || win[1] < 0 || win[2] <= 0 || win[3] <= 0 || win[0] + win[2] > width || win[1] + win[3] > height)
      return fail("pixel window must lie inside the image");
  }
  const int wc0 = win[0] / GCR_TILE_X, wc1 = (win[0] + win[2] + GCR_TILE_X - 1) / GCR_TILE_X;
  const int wr0 = win[1] / GCR_TILE_Y, wr1 = (win[1] + win[3] + GCR_TILE_Y - 1) / GCR_TILE_Y;
  int dev = 0;
  GCR_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail("device ordinal out of range");

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
  uint32_t* counters = reinterpret_cast<uint32_t*>(gptr + gl.counters);
  uint2* ranges = reinterpret_cast<uint2*>(iptr + il.ranges);
  if (radii == nullptr) radii = reinterpret_cast<int*>(gptr + gl.radii);

  GCR_CUDA_OK(cudaMemsetAsync(gptr + gl.zero_begin, 0, gl.zero_end - gl.zero_begin, stream));
  GCR_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)tiles, stream));
  // the stripe bounds travel with the geometry buffer (the backward reads them from there)
  int* bounds_dev = nullptr;
  bool deferred = false;   // balanced stripes cut on the device from this frame's own rects
  if (shard_count > 1) {
    bounds_dev = reinterpret_cast<int*>(counters + GCR_CNT_STRIPE_BOUNDS);
    if (stripe_bounds != nullptr) {
      GCR_CUDA_OK(cudaMemcpyAsync(bounds_dev, stripe_bounds, sizeof(int) * (shard_count + 1),
                                  cudaMemcpyDeviceToDevice, stream));
    } else if (balanced) {
      deferred = true;
    } else {
      int eq[GCR_MAX_SHARDS + 1];
      for (int k = 0; k <= shard_count; ++k) eq[k] = (int)((long long)grid_y * k / shard_count);
      GCR_CUDA_OK(cudaMemcpyAsync(bounds_dev, eq, sizeof(int) * (shard_count + 1), cudaMemcpyHostToDevice,
                                  stream));
    }
  }
  const int* stripe = bounds_dev != nullptr ? bounds_dev + shard_rank : nullptr;

  // 1. per-Gaussian projection (+ stripe selection)
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
  pa.shard_rank = shard_rank; pa.shard_count = shard_count; pa.stripe_bounds = bounds_dev;
  pa.win_col0 = wc0; pa.win_col1 = wc1; pa.win_row0 = wr0; pa.win_row1 = wr1;
  pa.prefiltered = prefiltered != 0;
  pa.radii = radii; pa.tiles_touched = tiles_touched; pa.depth_keys = keys_a;
  pa.records = records; pa.rects = reinterpret_cast<uint2*>(gptr + gl.rects);
  pa.clamped = reinterpret_cast<uint8_t*>(gptr + gl.clamped);
  pa.owner = reinterpret_cast<uint8_t*>(gptr + gl.owner);
  pa.total_tiles = reinterpret_cast<unsigned long long*>(counters + GCR_CNT_TOTAL_TILES64);
  pa.packed_rects = reinterpret_cast<unsigned long long*>(gptr + gl.packed_rects);
  pa.row_hist = reinterpret_cast<uint32_t*>(gptr + gl.row_hist);
  pa.stripe_bounds_out = bounds_dev;
  // balanced stripes: the projection leaves index-ordered depths in the b side of the sort's key
  // ping-pong; the select stage compacts this stripe's (key, index) pairs into the a side
  pa.depth_in = keys_b;
  pa.sorted_init = vals_a;
  pa.n_vis_out = counters + GCR_CNT_NVIS;
  pa.select_ticket = reinterpret_cast<uint32_t*>(gptr + gl.select_ws);
  pa.select_status = reinterpret_cast<unsigned long long*>(gptr + gl.select_ws + 256);
  pa.dbg_cov3D = g_dbg_cov3d;
  g_dbg_cov3d = nullptr;
  prof_mark(ST_PREPROCESS, 0, stream);
  GCR_LAUNCH("project", gcr_launch_project(pa, deferred, stream), debug, stream);
  prof_mark(ST_PREPROCESS, 1, stream);
  if (deferred) {
    prof_mark(ST_SELECT, 0, stream);
    GCR_LAUNCH("stripe select", gcr_launch_stripe_select(pa, stream), debug, stream);
    prof_mark(ST_SELECT, 1, stream);
  }

  // num_rendered on its way to the host; the wait comes after the depth sort is enqueued
  HostMailbox& mb = g_mailbox;
  if (mb.pinned[dev] == nullptr) {
    GCR_CUDA_OK(cudaHostAlloc(&mb.pinned[dev], sizeof(unsigned long long), cudaHostAllocDefault));
    GCR_CUDA_OK(cudaEventCreateWithFlags(&mb.ev[dev], cudaEventDisableTiming));
  }
  GCR_CUDA_OK(cudaMemcpyAsync(mb.pinned[dev], pa.total_tiles, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, stream));
  GCR_CUDA_OK(cudaEventRecord(mb.ev[dev], stream));

  // 2. stable depth sort of the Gaussians that touch this stripe (4 passes: result in the a pair)
  prof_mark(ST_DEPTH_SORT, 0, stream);
  // direct mode: culled Gaussians are dropped (and indices generated) by the sort's first pass;
  // balanced stripes: the select stage already compacted this stripe's pairs, count on the device
  GCR_LAUNCH("depth sort",
             deferred ? gcr_launch_radix_sort(keys_a, vals_a, keys_b, vals_b, (size_t)P, counters + GCR_CNT_NVIS,
                                              32, false, false, nullptr, gptr + gl.sort_ws, stream)
                      : gcr_launch_radix_sort(keys_a, vals_a, keys_b, vals_b, (size_t)P, nullptr, 32, true, false,
                                              counters + GCR_CNT_NVIS, gptr + gl.sort_ws, stream),
             debug, stream);
  prof_mark(ST_DEPTH_SORT, 1, stream);
  const uint32_t* sorted_gauss = vals_a;
  ht.lap("launch_pre");

  // 3. R (sizes the binning buffer; the GPU is busy sorting meanwhile)
  GCR_CUDA_OK(cudaEventSynchronize(mb.ev[dev]));
  const unsigned long long R64 = *mb.pinned[dev];
  ht.lap("wait_R");
  if (R64 > 0x7fffffffull) return fail("num_rendered exceeds int32");
  const size_t R = (size_t)R64;

  const BinLayout bl(R);
  char* bptr = binningBuffer(binning_ctx, bl.total);
  if (bptr == nullptr) return fail("binning buffer allocation failed");
  bptr = align256(bptr);
  ht.lap("alloc_bin");
  uint32_t* tk_a = reinterpret_cast<uint32_t*>(bptr + bl.keys_a);
  uint32_t* tk_b = reinterpret_cast<uint32_t*>(bptr + bl.keys_b);
  uint32_t* tv_a = reinterpret_cast<uint32_t*>(bptr + bl.vals_a);
  uint32_t* tv_b = reinterpret_cast<uint32_t*>(bptr + bl.vals_b);

  if (R > 0) {
    const int tbits = tile_bits(tiles) > 0 ? tile_bits(tiles) : 1;
    const int tpass = (tbits + 7) / 8;
    // an odd number of passes starts from the b pair, so point_list always ends up in vals_a
    uint32_t* k_in = (tpass & 1) ? tk_b : tk_a;
    uint32_t* v_in = (tpass & 1) ? tv_b : tv_a;
    uint32_t* k_out = (tpass & 1) ? tk_a : tk_b;
    uint32_t* v_out = (tpass & 1) ? tv_a : tv_b;
    void* tile_ws = bptr + bl.sort_ws;
    GCR_CUDA_OK(cudaMemsetAsync(tile_ws, 0, gcr_sort_workspace_bytes(R), stream));
    // 4. scan + emit (tile, gaussian) pairs in depth order, 5. stable split by tile id
    GcrEmitLaunch el;
    memset(&el, 0, sizeof(el));
    el.n_max = (uint32_t)P; el.n_vis = counters + GCR_CNT_NVIS; el.sorted_gauss = sorted_gauss;
    el.rects = reinterpret_cast<const uint2*>(gptr + gl.rects); el.grid_x = grid_x;
    el.tile_keys = k_in; el.gauss_vals = v_in; el.cap = (uint32_t)R;
    el.workspace = gptr + gl.emit_ws; el.ghist_tile = gcr_sort_ghist(tile_ws); el.tile_end_bit = tbits;
    el.counters = counters; el.offsets_out = reinterpret_cast<uint32_t*>(gptr + gl.offsets);
    prof_mark(ST_EMIT, 0, stream);
    GCR_LAUNCH("scan + emit pairs", gcr_launch_emit_scan(el, stream), debug, stream);
    prof_mark(ST_EMIT, 1, stream);
    prof_mark(ST_TILE_SORT, 0, stream);
    GCR_LAUNCH("tile sort",
               gcr_launch_radix_sort(k_in, v_in, k_out, v_out, R, counters + GCR_CNT_R, tbits, false, true,
                                     nullptr, tile_ws, stream),
               debug, stream);
    prof_mark(ST_TILE_SORT, 1, stream);
    // 6. tile ranges
    prof_mark(ST_RANGES, 0, stream);
    GCR_LAUNCH("tile ranges", gcr_launch_tile_ranges((uint32_t)R, counters + GCR_CNT_R, tk_a, ranges, stream),
               debug, stream);
    prof_mark(ST_RANGES, 1, stream);
  }

  // 7. per-tile blend
  GcrBlendArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.W = width; ba.H = height; ba.grid_x = grid_x; ba.grid_y = grid_y;
  ba.tile_x0 = wc0; ba.tile_y0 = wr0; ba.tiles_x = wc1 - wc0; ba.tiles_y = wr1 - wr0;
  ba.px0 = win[0]; ba.py0 = win[1]; ba.pw = win[2]; ba.ph = win[3];
  ba.stripe = stripe;
  ba.ranges = ranges; ba.point_list = tv_a; ba.records = records; ba.bg = background;
  ba.final_T = reinterpret_cast<float*>(iptr + il.final_T);
  ba.n_contrib = reinterpret_cast<uint32_t*>(iptr + il.n_contrib);
  ba.out_color = out_color;
  prof_mark(ST_BLEND_FWD, 0, stream);
  GCR_LAUNCH("blend_fwd", gcr_launch_blend_fwd(ba, stream), debug, stream);
  prof_mark(ST_BLEND_FWD, 1, stream);
  ht.lap("launch_post");
  return (int)R;
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
  return forward_impl(geometryBuffer, geometry_ctx, binningBuffer, binning_ctx, imageBuffer, image_ctx,
                      P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
                      scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                      tan_fovx, tan_fovy, prefiltered, out_color, radii, debug, shard_rank, shard_count,
                      nullptr, 0, nullptr, cuda_stream);
}

int gcr_rasterizer_forward_striped(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                                   gcr_alloc_fn binningBuffer, void* binning_ctx,
                                   gcr_alloc_fn imageBuffer, void* image_ctx, int P, int D, int M,
                                   const float* background, int width, int height,
                                   const float* means3D, const float* shs,
                                   const float* colors_precomp, const float* opacities,
                                   const float* scales, float scale_modifier,
                                   const float* rotations, const float* cov3D_precomp,
                                   const float* viewmatrix, const float* projmatrix,
                                   const float* cam_pos, float tan_fovx, float tan_fovy,
                                   int prefiltered, float* out_color, int* radii, int debug,
                                   int shard_rank, int shard_count, const int* stripe_bounds,
                                   int balanced, void* cuda_stream) {
  return forward_impl(geometryBuffer, geometry_ctx, binningBuffer, binning_ctx, imageBuffer, image_ctx,
                      P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
                      scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                      tan_fovx, tan_fovy, prefiltered, out_color, radii, debug, shard_rank, shard_count,
                      stripe_bounds, balanced, nullptr, cuda_stream);
}

int gcr_rasterizer_forward_window(gcr_alloc_fn geometryBuffer, void* geometry_ctx,
                                  gcr_alloc_fn binningBuffer, void* binning_ctx,
                                  gcr_alloc_fn imageBuffer, void* image_ctx, int P, int D, int M,
                                  const float* background, int width, int height,
                                  const float* means3D, const float* shs,
                                  const float* colors_precomp, const float* opacities,
                                  const float* scales, float scale_modifier,
                                  const float* rotations, const float* cov3D_precomp,
                                  const float* viewmatrix, const float* projmatrix,
                                  const float* cam_pos, float tan_fovx, float tan_fovy,
                                  int prefiltered, float* out_color, int* radii, int debug,
                                  int win_x, int win_y, int win_w, int win_h, void* cuda_stream) {
  const int window[4] = {win_x, win_y, win_w, win_h};
  return forward_impl(geometryBuffer, geometry_ctx, binningBuffer, binning_ctx, imageBuffer, image_ctx,
                      P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
                      scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                      tan_fovx, tan_fovy, prefiltered, out_color, radii, debug, 0, 1, nullptr, 0, window,
                      cuda_stream);
}

int gcr_rasterizer_backward_blend(int P, int R, const float* background, int width, int height,
                                  char* geom_buffer, char* binning_buffer, char* image_buffer,
                                  const float* dL_dpix, float* const* accumulators,
                                  int n_accumulators, int zero_first, int remote_scalar_atomics,
                                  int debug, int shard_rank, int shard_count, const int* window,
                                  void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  if (!valid_shard(shard_rank, shard_count)) return fail("invalid tile-row shard (rank, count)");
  int win[4] = {0, 0, width, height};
  if (window != nullptr) {
    if (shard_count != 1) return fail("a pixel window cannot be combined with tile-row stripes");
    for (int k = 0; k < 4; ++k) win[k] = window[k];
    if (win[0] < 0 || win[1] < 0 || win[2] <= 0 || win[3] <= 0 || win[0] + win[2] > width || win[1] + win[3] > height)
      return fail("pixel window must lie inside the image");
  }
  if (width <= 0 || height <= 0) return fail("image size must be positive");
  if (accumulators == nullptr || (n_accumulators != 1 && n_accumulators != shard_count))
    return fail("accumulators: pass 1 pointer, or one per rank of the striped frame");
  for (int k = 0; k < n_accumulators; ++k) {
    if (accumulators[k] == nullptr) return fail("accumulator pointers must not be NULL");
    if ((reinterpret_cast<uintptr_t>(accumulators[k]) & 15u) != 0)
      return fail("accumulators must be 16-byte aligned");
  }
  if (geom_buffer == nullptr || image_buffer == nullptr || dL_dpix == nullptr || background == nullptr)
    return fail("geom_buffer, image_buffer, dL_dpix and background must not be NULL");
  if (R > 0 && binning_buffer == nullptr) return fail("binning_buffer must not be NULL when R > 0");
  if (n_accumulators > 1 && P >= (1 << 27))
    return fail("per-rank accumulators support up to 2^27 Gaussians (owner rank is packed into the id)");
  float* local = accumulators[n_accumulators == 1 ? 0 : shard_rank];
  if (zero_first) {
    prof_mark(ST_BWD_ZERO, 0, stream);
    GCR_CUDA_OK(cudaMemsetAsync(local, 0, sizeof(GcrGradAcc) * (size_t)P, stream));
    prof_mark(ST_BWD_ZERO, 1, stream);
  }
  if (R <= 0) return 0;
  const int grid_x = (width + GCR_TILE_X - 1) / GCR_TILE_X;
  const int grid_y = (height + GCR_TILE_Y - 1) / GCR_TILE_Y;
  const GeomLayout gl((size_t)P);
  const ImgLayout il((size_t)width * height, (size_t)grid_x * grid_y);
  char* gptr = align256(geom_buffer);
  char* bptr = align256(binning_buffer);
  char* iptr = align256(image_buffer);
  const uint32_t* counters = reinterpret_cast<const uint32_t*>(gptr + gl.counters);
  GcrBlendArgs ba;
  memset(&ba, 0, sizeof(ba));
  ba.W = width; ba.H = height; ba.grid_x = grid_x; ba.grid_y = grid_y;
  ba.tile_x0 = win[0] / GCR_TILE_X; ba.tile_y0 = win[1] / GCR_TILE_Y;
  ba.tiles_x = (win[0] + win[2] + GCR_TILE_X - 1) / GCR_TILE_X - ba.tile_x0;
  ba.tiles_y = (win[1] + win[3] + GCR_TILE_Y - 1) / GCR_TILE_Y - ba.tile_y0;
  ba.px0 = win[0]; ba.py0 = win[1]; ba.pw = win[2]; ba.ph = win[3];
  ba.stripe = shard_count > 1
                  ? reinterpret_cast<const int*>(counters + GCR_CNT_STRIPE_BOUNDS) + shard_rank : nullptr;
  ba.ranges = reinterpret_cast<const uint2*>(iptr + il.ranges);
  ba.point_list = reinterpret_cast<const uint32_t*>(bptr);   // BinLayout::vals_a == 0
  ba.records = reinterpret_cast<const GcrRecord*>(gptr + gl.records);
  ba.owner = reinterpret_cast<const uint8_t*>(gptr + gl.owner);
  ba.bg = background;
  ba.final_T = reinterpret_cast<float*>(iptr + il.final_T);
  ba.n_contrib = reinterpret_cast<uint32_t*>(iptr + il.n_contrib);
  ba.dL_dpix = dL_dpix;
  for (int k = 0; k < n_accumulators; ++k) ba.acc[k] = reinterpret_cast<GcrGradAcc*>(accumulators[k]);
  ba.n_acc = n_accumulators;
  ba.self_rank = n_accumulators == 1 ? 0 : shard_rank;
  ba.remote_scalar = remote_scalar_atomics != 0;
  prof_mark(ST_BLEND_BWD, 0, stream);
  GCR_LAUNCH("blend_bwd", gcr_launch_blend_bwd(ba, stream), debug, stream);
  prof_mark(ST_BLEND_BWD, 1, stream);
  return 0;
}

int gcr_rasterizer_backward_geometry(int P, int D, int M, const float* means3D, const float* shs,
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
                                     float* dL_packed, void* cuda_stream) {
  (void)radii;  // ownership (which implies visibility) is recorded in the geometry buffer
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  if (range_count < 0) { range_start = 0; range_count = P; }
  if (range_start < 0 || range_start + range_count > P) return fail("Gaussian range out of bounds");
  if (shard_rank < 0 || shard_rank >= GCR_MAX_SHARDS) return fail("invalid tile-row shard (rank, count)");
  if (width <= 0 || height <= 0) return fail("image size must be positive");
  if (cov3D_precomp == nullptr && scales == nullptr)   // rotations == NULL alone: identity rotation
    return fail("provide either scale/rotation or a precomputed 3D covariance");
  if (means3D == nullptr || viewmatrix == nullptr || projmatrix == nullptr || geom_buffer == nullptr ||
      accumulator == nullptr)
    return fail("means3D, viewmatrix, projmatrix, geom_buffer and accumulator must not be NULL");
  if (dL_packed == nullptr &&
      (dL_dmean2D == nullptr || dL_dcolor == nullptr || dL_dmean3D == nullptr || dL_dcov3D == nullptr))
    return fail("dL_dmean2D, dL_dcolor, dL_dmean3D and dL_dcov3D must not be NULL");
  if (dL_packed != nullptr && (reinterpret_cast<uintptr_t>(dL_packed) & 31u) != 0)
    return fail("dL_packed must be 32-byte aligned");
  if ((reinterpret_cast<uintptr_t>(accumulator) & 15u) != 0) return fail("accumulators must be 16-byte aligned");
  if (shs != nullptr && M > 16) return fail("at most 16 SH coefficients per Gaussian (degree 3) are supported");
  if (shs != nullptr && M > 0 && dL_dsh == nullptr) return fail("dL_dsh must not be NULL with SHs");
  if (shs != nullptr && M > 0 && campos == nullptr) return fail("cam_pos must not be NULL with SHs");
  if (rotations != nullptr && (reinterpret_cast<uintptr_t>(rotations) & 15u) != 0)
    return fail("rotations must be 16-byte aligned");
  if (dL_drot != nullptr && (reinterpret_cast<uintptr_t>(dL_drot) & 15u) != 0)
    return fail("dL_drot must be 16-byte aligned");
  if (dL_dconic != nullptr && (reinterpret_cast<uintptr_t>(dL_dconic) & 15u) != 0)
    return fail("dL_dconic must be 16-byte aligned");
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
  a.owner = reinterpret_cast<const uint8_t*>(gptr + gl.owner);
  a.my_rank = shard_rank; a.zero_unowned = striped == 0; a.clear_acc = clear_accumulator != 0;
  a.shs = (M > 0) ? shs : nullptr;
  a.clamped = reinterpret_cast<const uint8_t*>(gptr + gl.clamped);
  a.scales = scales; a.rotations = rotations; a.scale_modifier = scale_modifier;
  a.cov3D_precomp = cov3D_precomp;
  a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos;
  a.focal_y = height / (2.0f * tan_fovy);   // rasterizer_impl.cu:306-307
  a.focal_x = width / (2.0f * tan_fovx);
  a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
  a.grad_acc = reinterpret_cast<GcrGradAcc*>(accumulator);
  a.ticket = reinterpret_cast<uint32_t*>(gptr + gl.counters) + GCR_CNT_BWD_TICKET;
  a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dopacity = dL_dopacity;
  a.dL_dcolor = dL_dcolor; a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D;
  a.dL_dsh = dL_dsh; a.dL_dscale = (scales != nullptr) ? dL_dscale : nullptr;
  a.dL_drot = (scales != nullptr && rotations != nullptr) ? dL_drot : nullptr;
  a.packed_out = dL_packed;
  prof_mark(ST_GEOM_BWD, 0, stream);
  GCR_LAUNCH("preprocess_bwd", gcr_launch_preprocess_bwd(a, stream), debug, stream);
  prof_mark(ST_GEOM_BWD, 1, stream);
  // reference semantics: with a precomputed covariance the scale/rotation gradients are zeros
  if (scales == nullptr && striped == 0 && dL_packed == nullptr) {
    if (dL_dscale)
      GCR_CUDA_OK(cudaMemsetAsync(dL_dscale + 3 * (size_t)range_start, 0,
                                  sizeof(float) * 3 * (size_t)range_count, stream));
    if (dL_drot)
      GCR_CUDA_OK(cudaMemsetAsync(dL_drot + 4 * (size_t)range_start, 0,
                                  sizeof(float) * 4 * (size_t)range_count, stream));
  }
  return 0;
}

namespace {
int backward_impl(int P, int D, int M, int R, const float* background, int width, int height,
                  const float* means3D, const float* shs, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                  const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                  float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                  float* dL_drot, int debug, const int* window, void* cuda_stream) {
  if (P <= 0) return 0;
  if (geom_buffer == nullptr) return fail("geom_buffer must not be NULL");
  const GeomLayout gl((size_t)P);
  float* grad_acc = reinterpret_cast<float*>(align256(geom_buffer) + gl.grad_acc);
  int rc = gcr_rasterizer_backward_blend(P, R, background, width, height, geom_buffer, binning_buffer,
                                         image_buffer, dL_dpix, &grad_acc, 1, 1, 0, debug, 0, 1, window,
                                         cuda_stream);
  if (rc < 0) return rc;
  return gcr_rasterizer_backward_geometry(P, D, M, means3D, shs, scales, scale_modifier, rotations,
                                          cov3D_precomp, viewmatrix, projmatrix, campos, width,
                                          height, tan_fovx, tan_fovy, radii, geom_buffer, grad_acc,
                                          dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                          dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug,
                                          0, -1, 0, 0, 0, nullptr, cuda_stream);
}
}  // namespace

int gcr_rasterizer_backward_window(int P, int D, int M, int R, const float* background, int width,
                                   int height, const float* means3D, const float* shs,
                                   const float* colors_precomp, const float* scales,
                                   float scale_modifier, const float* rotations,
                                   const float* cov3D_precomp, const float* viewmatrix,
                                   const float* projmatrix, const float* campos, float tan_fovx,
                                   float tan_fovy, const int* radii, char* geom_buffer,
                                   char* binning_buffer, char* image_buffer, const float* dL_dpix,
                                   float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                   float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                                   float* dL_dsh, float* dL_dscale, float* dL_drot, int debug,
                                   int win_x, int win_y, int win_w, int win_h, void* cuda_stream) {
  (void)colors_precomp;
  const int window[4] = {win_x, win_y, win_w, win_h};
  return backward_impl(P, D, M, R, background, width, height, means3D, shs, scales, scale_modifier, rotations,
                       cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                       binning_buffer, image_buffer, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                       dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, window, cuda_stream);
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
  (void)colors_precomp;  // colours live in the per-Gaussian records since the forward
  if (P <= 0) return 0;
  if (shard_count != 1 || shard_rank != 0)
    return fail("gcr_rasterizer_backward handles single-stripe frames; a striped frame goes through "
                "gcr_rasterizer_backward_blend / _geometry");
  return backward_impl(P, D, M, R, background, width, height, means3D, shs, scales, scale_modifier, rotations,
                       cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                       binning_buffer, image_buffer, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                       dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, nullptr, cuda_stream);
}

int gcr_rasterizer_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, uint8_t* present, void* cuda_stream) {
  (void)projmatrix;  // the reference's test only uses the view-space depth (auxiliary.h:145)
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P <= 0) return 0;
  if (means3D == nullptr || viewmatrix == nullptr || present == nullptr)
    return fail("means3D, viewmatrix and present must not be NULL");
  GCR_LAUNCH("check_frustum",
             gcr_launch_check_frustum(P, means3D, viewmatrix, reinterpret_cast<bool*>(present), stream), 0,
             stream);
  return 0;
}

int gcr_stripe_partition(int P, const float* means3D, const float* scales, float scale_modifier,
                         const float* rotations, const float* cov3D_precomp,
                         const float* viewmatrix, const float* projmatrix, int width, int height,
                         float tan_fovx, float tan_fovy, int shard_count, void* workspace,
                         int* stripe_bounds_out, void* cuda_stream) {
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (width <= 0 || height <= 0) return fail("image size must be positive");
  if (shard_count < 1 || shard_count > GCR_MAX_SHARDS) return fail("invalid tile-row shard (rank, count)");
  if (workspace == nullptr || stripe_bounds_out == nullptr)
    return fail("workspace and stripe_bounds_out must not be NULL");
  const int grid_x = (width + GCR_TILE_X - 1) / GCR_TILE_X;
  const int grid_y = (height + GCR_TILE_Y - 1) / GCR_TILE_Y;
  if (grid_x > 4095 || grid_y > 4095) return fail("images above 65520 pixels a side are not supported");
  if (P <= 0) {
    int eq[GCR_MAX_SHARDS + 1];
    for (int k = 0; k <= shard_count; ++k) eq[k] = (int)((long long)grid_y * k / shard_count);
    GCR_CUDA_OK(cudaMemcpyAsync(stripe_bounds_out, eq, sizeof(int) * (shard_count + 1), cudaMemcpyHostToDevice,
                                stream));
    return 0;
  }
  if (means3D == nullptr || viewmatrix == nullptr || projmatrix == nullptr)
    return fail("means3D, viewmatrix and projmatrix must not be NULL");
  if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
    return fail("provide either scale/rotation or a precomputed 3D covariance");
  if (rotations != nullptr && (reinterpret_cast<uintptr_t>(rotations) & 15u) != 0)
    return fail("rotations must be 16-byte aligned");
  GcrPreprocessArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.P = P;
  pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier;
  pa.rotations = rotations; pa.cov3D_precomp = cov3D_precomp;
  pa.viewmatrix = viewmatrix; pa.projmatrix = projmatrix;
  pa.W = width; pa.H = height; pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy;
  pa.focal_y = height / (2.0f * tan_fovy);
  pa.focal_x = width / (2.0f * tan_fovx);
  pa.grid_x = grid_x; pa.grid_y = grid_y;
  pa.shard_rank = 0; pa.shard_count = shard_count;
  pa.row_hist = static_cast<uint32_t*>(workspace);
  pa.stripe_bounds_out = stripe_bounds_out;
  GCR_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(uint32_t) * (size_t)(grid_y + 1), stream));
  prof_mark(ST_PARTITION, 0, stream);
  GCR_LAUNCH("stripe partition", gcr_launch_stripe_partition(pa, stream), 0, stream);
  prof_mark(ST_PARTITION, 1, stream);
  return 0;
}

// ---- peer memory ----------------------------------------------------------------------------
int gcr_peer_alloc(size_t bytes, void** device_ptr, unsigned char* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) <= GCR_PEER_HANDLE_BYTES, "handle size");
  if (device_ptr == nullptr || handle_out == nullptr || bytes == 0) return fail("gcr_peer_alloc: bad arguments");
  void* p = nullptr;
  GCR_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(std::string("gcr_peer_alloc: ") + cudaGetErrorString(e));
  }
  memset(handle_out, 0, GCR_PEER_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  *device_ptr = p;
  return 0;
}

int gcr_peer_open(const unsigned char* handle, void** device_ptr) {
  if (handle == nullptr || device_ptr == nullptr) return fail("gcr_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  GCR_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *device_ptr = p;
  return 0;
}

int gcr_peer_close(void* device_ptr) {
  if (device_ptr != nullptr) GCR_CUDA_OK(cudaIpcCloseMemHandle(device_ptr));
  return 0;
}

int gcr_peer_free(void* device_ptr) {
  if (device_ptr != nullptr) GCR_CUDA_OK(cudaFree(device_ptr));
  return 0;
}

int gcr_peer_barrier(void* const* flag_arrays, int rank, int world, unsigned int epoch,
                     void* cuda_stream) {
  if (!valid_shard(rank, world) || flag_arrays == nullptr) return fail("gcr_peer_barrier: bad arguments");
  GcrPeerFlags f;
  memset(&f, 0, sizeof(f));
  for (int k = 0; k < world; ++k) {
    if (flag_arrays[k] == nullptr) return fail("gcr_peer_barrier: NULL flag array");
    f.p[k] = static_cast<uint32_t*>(flag_arrays[k]);
  }
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  GCR_LAUNCH("peer barrier", gcr_launch_peer_barrier(f, rank, world, epoch, stream), 0, stream);
  return 0;
}

void gcr_debug_set_cov3d_out(float* cov3d) { g_dbg_cov3d = cov3d; }

int gcr_set_programmatic_launch(int on) { return pdl_flag().exchange(on != 0 ? 1 : 0); }

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
      case GCR_GEOM_OWNER: return gl.owner;
      case GCR_GEOM_COUNTERS: return gl.counters;
      case GCR_GEOM_TOTAL_BYTES: return gl.total;
      default: return (size_t)-1;
    }
  } else if (which < 400) {
    const BinLayout bl((size_t)(R > 0 ? R : 0));
    switch (which) {
      case GCR_BIN_POINT_LIST: return bl.vals_a;
      case GCR_BIN_TILE_KEYS: return bl.keys_a;
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
