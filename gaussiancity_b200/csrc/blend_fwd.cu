// blend_fwd.cu -- per-tile front-to-back alpha composition for sm_100a.
//
// Behavioural spec: DGR/cuda_rasterizer/forward.cu:238-346 (renderCUDA): per pixel, walk the
// tile's depth-sorted instance list; d = mean2D - pix; power = -1/2 (A dx^2 + C dy^2) - B dx dy;
// skip power > 0; alpha = min(0.99, o exp(power)); skip alpha < 1/255; stop when
// T (1 - alpha) < 1e-4; C += rgb alpha T.  Outputs colour (CHW) = C + T bg, final_T, n_contrib
// (1-based list position of the last contributor).
//
// B200 design (not the reference's): a tile's list is walked in batches of 256 instances through
// a 2-stage shared-memory ring.  Each thread fetches ONE instance of the next batch straight from
// the per-Gaussian record array -- its id from point_list one batch ahead, then three 16-byte
// asynchronous copies (cp.async -> LDGSTS, no register staging) whose completion is signalled
// on the stage's mbarrier (cp.async.mbarrier.arrive) -- while the current batch is blended.
// Only batches a tile actually reaches are ever fetched: tiles saturate after ~20 % of their
// list on the headline workload, so this replaced a 48 B x R materialisation pass that wrote
// (and re-read) 2.5x more than the blend consumes (profiles/r02_*).  Each warp owns an 8x4
// pixel sub-rectangle: lanes first test 32 records at a time against that sub-rectangle
// (exact-conservative ellipse/rect test, blend_common.cuh), ballot, and only the surviving
// records are evaluated per pixel.  A warp whose 32 pixels have all saturated (T < 1e-4) stops
// evaluating; the CTA leaves when all eight have (block vote once per batch).  Arithmetic per
// pixel is pinned to the reference's SASS order (gcr_power, expf, fma order of the colour
// accumulation), so final_T / n_contrib are bit-identical.
#include "blend_common.cuh"
#include "gcr_kernels.h"

namespace {

__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(GcrBlendArgs a) {
  __shared__ __align__(128) GcrRecord stage[kBlendStages][kBlendBatch];
  __shared__ __align__(8) uint64_t full_bar[kBlendStages];
  __shared__ int vote[3];
  gcr_pdl_wait();
  gcr_pdl_trigger();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = a.tile_x0 + (int)blockIdx.x;
  int tile_y = a.tile_y0 + (int)blockIdx.y;
  if (a.stripe != nullptr) {
    tile_y = a.stripe[0] + (int)blockIdx.y;
    if (tile_y >= a.stripe[1]) return;   // grid covers every row; the stripe is device-side
  }
  const uint2 range = a.ranges[tile_y * a.grid_x + tile_x];
  const int n = (int)(range.y - range.x);
  const int nb = (n + kBlendBatch - 1) / kBlendBatch;

  const int sub_x0 = tile_x * GCR_TILE_X + (warp & 1) * 8;
  const int sub_y0 = tile_y * GCR_TILE_Y + (warp >> 1) * 4;
  const int pix_x = sub_x0 + (lane & 7);
  const int pix_y = sub_y0 + (lane >> 3);
  // pixels of the frame that are also inside the pixel window (the whole frame by default)
  const bool inside = pix_x < a.W && pix_y < a.H && pix_x >= a.px0 && pix_x < a.px0 + a.pw && pix_y >= a.py0 &&
                      pix_y < a.py0 + a.ph;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float rx0 = (float)sub_x0, rx1 = (float)(sub_x0 + 7);
  const float ry0 = (float)sub_y0, ry1 = (float)(sub_y0 + 3);

  if (tid == 0) {
    gcr_mbar_init(&full_bar[0], kBlendThreads);
    gcr_mbar_init(&full_bar[1], kBlendThreads);
    gcr_mbar_fence_init();
    vote[0] = vote[1] = vote[2] = 0;
  }
  __syncthreads();

  const uint32_t* __restrict__ ids = a.point_list + range.x;
  // batch 0 is fetched now; the id of this thread's instance in batch 1 rides in a register
  if (nb > 0) gcr_gather_record(&stage[0][tid], a.records, tid < n ? (int)ids[tid] : -1, &full_bar[0]);
  int next_id = (kBlendBatch + tid < n) ? (int)ids[kBlendBatch + tid] : -1;
  int fetched = nb > 0 ? 1 : 0;   // batches whose copies this thread has issued

  float T = 1.0f;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last_contributor = 0;
  bool done = !inside;

  int b = 0;
  for (; b < nb; ++b) {
    const int s = b & 1;
    if (b + 1 < nb) {
      // stage s^1 was last read in iteration b-1; the vote barrier at its end ordered those reads
      gcr_gather_record(&stage[s ^ 1][tid], a.records, next_id, &full_bar[s ^ 1]);
      fetched = b + 2;
      const int j = (b + 2) * kBlendBatch + tid;
      next_id = j < n ? (int)ids[j] : -1;
    }
    gcr_mbar_wait(&full_bar[s], (uint32_t)((b >> 1) & 1));

    const int cnt = min(kBlendBatch, n - b * kBlendBatch);
    const GcrRecord* __restrict__ st = stage[s];
    if (__ballot_sync(0xffffffffu, !done) != 0u) {
      for (int g0 = 0; g0 < cnt; g0 += 32) {
        const int j = g0 + lane;
        bool touch = false;
        if (j < cnt) {
          const float4 q0 = st[j].q0;   // x, y, A, B
          const float4 q1 = st[j].q1;   // C, o, 2 ln(255 o), -
          touch = gcr_subrect_touch(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rx0, rx1, ry0, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, touch);
        const GcrRecord* __restrict__ stc = st + g0;        // chunk-invariant parts hoisted
        const uint32_t pos1 = (uint32_t)(b * kBlendBatch + g0 + 1);
        while (mask) {
          const int bit = __ffs(mask) - 1;
          mask &= mask - 1;
          const GcrRecord* __restrict__ rec = stc + bit;
          const float4 r0 = rec->q0;   // x, y, A, B   (broadcast LDS.128)
          const float2 r1 = *reinterpret_cast<const float2*>(&rec->q1);   // C, o
          // straight-line evaluation, state updates predicated: same arithmetic as the reference
          // on every contributing lane, no per-test branches (the warp is issue-bound)
          const float dx = __fsub_rn(r0.x, pxf);
          const float dy = __fsub_rn(r0.y, pyf);
          const float power = gcr_power(dx, dy, r0.z, r0.w, r1.x);
          const float alpha = fminf(0.99f, __fmul_rn(r1.y, expf(power)));
          const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
          const bool ok = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          const bool sat = ok && (test_T < 0.0001f);
          done = done || sat;
          if (ok && !sat) {
            const float4 c = rec->q2;    // r, g, b, -
            C0 = __fmaf_rn(T, __fmul_rn(alpha, c.x), C0);
            C1 = __fmaf_rn(T, __fmul_rn(alpha, c.y), C1);
            C2 = __fmaf_rn(T, __fmul_rn(alpha, c.z), C2);
            T = test_T;
            last_contributor = pos1 + (uint32_t)bit;
          }
        }
        if (__ballot_sync(0xffffffffu, !done) == 0u) break;  // warp saturated
      }
    }
    const bool warp_done_all = __ballot_sync(0xffffffffu, !done) == 0u;
    // CTA vote: everyone saturated -> leave.  The barrier also fences this batch's shared-memory
    // reads before the stage is refilled.  (A plain bar.sync plus a rotating flag rather than
    // __syncthreads_count: same cost here, and compute-sanitizer's racecheck understands it.)
    if (lane == 0 && !warp_done_all) vote[b % 3] = 1;
    __syncthreads();
    const bool any_active = vote[b % 3] != 0;
    if (tid == 0) vote[(b + 2) % 3] = 0;   // used two iterations from now; ordered by the next barrier
    if (!any_active) {
      ++b;
      break;
    }
  }
  // never exit with this thread's asynchronous copies still in flight into shared memory
  if (fetched > b) gcr_cp_async_wait_all();

  if (inside) {
    const int pix_id = a.W * pix_y + pix_x;
    a.final_T[pix_id] = T;
    a.n_contrib[pix_id] = last_contributor;
    const size_t out_id = (size_t)a.pw * (pix_y - a.py0) + (pix_x - a.px0);
    const size_t plane = (size_t)a.ph * a.pw;
    a.out_color[out_id] = __fmaf_rn(T, a.bg[0], C0);
    a.out_color[plane + out_id] = __fmaf_rn(T, a.bg[1], C1);
    a.out_color[2 * plane + out_id] = __fmaf_rn(T, a.bg[2], C2);
  }
}

}  // namespace

cudaError_t gcr_launch_blend_fwd(const GcrBlendArgs& a, cudaStream_t stream) {
  if (a.tiles_x <= 0 || a.tiles_y <= 0) return cudaSuccess;
  dim3 grid(a.tiles_x, a.stripe != nullptr ? a.grid_y : a.tiles_y, 1);
  return gcr_launch_chain(blend_fwd_kernel, grid, dim3(kBlendThreads), 0, stream, a);
}
