// blend_bwd.cu -- per-tile back-to-front gradient blend for sm_100a.
//
// Behavioural spec: DGR/cuda_rasterizer/backward.cu:428-581 (renderCUDA backward): per pixel,
// walk the tile's instance list from its last contributor backwards, rebuild T by division,
// accumulate dL/dcolor, dL/dalpha (recursive accum_rec + background term), and from it
// dL/dmean2D (x,y), dL/dconic (x,y,w), dL/dopacity for every contributing (pixel, Gaussian)
// pair.  The reference issues 9 scalar atomicAdd per pair.
//
// B200 design: same 2-stage asynchronous gather ring as the forward (cp.async + mbarrier; batches
// taken from the END of the tile's list), same per-warp sub-rectangle cull, instances beyond the
// CTA's furthest last-contributor are never fetched.  Transposed reduction: phase A (lane =
// pixel) parks three scalars per (record, pixel) in a warp-private shared-memory panel, phase B
// (lane = record x pixel row) accumulates the 9 gradient terms per record in registers and emits
// two 16-byte vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4) + one scalar into
// the 48-byte per-Gaussian accumulator GcrGradAcc OF THE RANK THAT OWNS THE GAUSSIAN: on one GPU
// that is local memory; under tile-row sharding the accumulators are peer-mapped, so the
// reductions of a Gaussian that straddles a stripe boundary travel over NVLink to its owner
// inside this kernel -- no [P,12] reduce-scatter afterwards.
#include "blend_common.cuh"
#include "gcr_kernels.h"

namespace {

// ------------------------------------------------------------------------------------------
// Transposed reduction.  Phase A (lane = pixel) walks the surviving records back to front
// exactly like the reference but, instead of reducing 9 gradient terms across the warp per record, parks
// three scalars per (record, pixel) -- alpha*T, dL/dalpha, G -- in a warp-private shared-memory
// panel of kSlots records.  Phase B (lane = record x pixel-row) then accumulates the 9 terms of
// each record over the pixels in registers: no shuffles per record, full lane utilisation, one
// quarter-combine (2 shuffle levels) per panel, and the sums leave as two 16-byte vector
// reductions + one scalar per record.
// ------------------------------------------------------------------------------------------
constexpr int kSlots = 8;                       // records per panel
constexpr int kPanelStride = 33;                // padded pixel stride (bank-conflict free)
constexpr int kWarpPanelFloats = 3 * kSlots * kPanelStride;            // 792
constexpr int kWarpScratchBytes = 3776;         // panel 3168 + slot_j 32 + pixel consts 576
// stage ring | 8 warp scratch areas | barriers + s_max_last (64 B) | per-stage Gaussian ids.
// 56 896 B: four CTAs per SM still fit (4 x (56 896 + 1 024 reserved) <= 228 KB).
constexpr int kBwdSmemBytes = kBlendStages * kBlendBatch * (int)sizeof(GcrRecord) + 8 * kWarpScratchBytes + 64 +
                              kBlendStages * kBlendBatch * (int)sizeof(uint32_t);
constexpr int kOwnerShift = 27;   // striped frames: owner rank in the top 5 bits of the staged id

// Math: the gradients only have to meet the 1e-4 bar, so exp is ex2.approx(power * log2 e)
// (2 instructions instead of ~10) and 1/(1-alpha) one MUFU.RCP (instead of the IEEE reciprocal's
// ~8): measured -9 % kernel time, gradients still <= 1e-6 norm-relative vs the reference
// (profiles/r02_variants_ab.md).  The set of contributors must still be EXACTLY the forward's (a
// flipped alpha >= 1/255 decision would rescale the rest of that pixel's T chain), so any alpha
// within 1e-7 of the threshold -- 30x the worst-case error of the approximation there -- is
// re-evaluated with the exact expf.
__global__ void __launch_bounds__(kBlendThreads, 4)   // 64 registers: 4 CTAs/SM (shared memory allows 4)
blend_bwd_kernel(GcrBlendArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  gcr_pdl_wait();
  gcr_pdl_trigger();
  GcrRecord (*stage)[kBlendBatch] = reinterpret_cast<GcrRecord (*)[kBlendBatch]>(smem_raw);
  unsigned char* scratch0 = smem_raw + kBlendStages * kBlendBatch * sizeof(GcrRecord);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch0 + 8 * kWarpScratchBytes);
  int* s_max_last_p = reinterpret_cast<int*>(full_bar + kBlendStages);
  // which Gaussian each staged record belongs to (= its accumulator row); on striped frames with
  // per-rank accumulators the owner rank rides in the top 5 bits (api.cu bounds P accordingly)
  uint32_t (*stage_id)[kBlendBatch] =
      reinterpret_cast<uint32_t (*)[kBlendBatch]>(scratch0 + 8 * kWarpScratchBytes + 64);
  const bool striped = a.n_acc > 1;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* panel = reinterpret_cast<float*>(scratch0 + warp * kWarpScratchBytes);
  int* slot_j = reinterpret_cast<int*>(panel + kWarpPanelFloats);
  float4* pxc = reinterpret_cast<float4*>(slot_j + kSlots);

  const int tile_x = a.tile_x0 + (int)blockIdx.x;
  int tile_y = a.tile_y0 + (int)blockIdx.y;
  if (a.stripe != nullptr) {
    tile_y = a.stripe[0] + (int)blockIdx.y;
    if (tile_y >= a.stripe[1]) return;   // grid covers every row; the stripe is device-side
  }
  const uint2 range = a.ranges[tile_y * a.grid_x + tile_x];
  const int n = (int)(range.y - range.x);

  const int sub_x0 = tile_x * GCR_TILE_X + (warp & 1) * 8;
  const int sub_y0 = tile_y * GCR_TILE_Y + (warp >> 1) * 4;
  const int pix_x = sub_x0 + (lane & 7);
  const int pix_y = sub_y0 + (lane >> 3);
  const bool inside = pix_x < a.W && pix_y < a.H && pix_x >= a.px0 && pix_x < a.px0 + a.pw && pix_y >= a.py0 &&
                      pix_y < a.py0 + a.ph;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float rx0 = (float)sub_x0, rx1 = (float)(sub_x0 + 7);
  const float ry0 = (float)sub_y0, ry1 = (float)(sub_y0 + 3);
  const int pix_id = a.W * pix_y + pix_x;
  const size_t win_id = (size_t)a.pw * (pix_y - a.py0) + (pix_x - a.px0);   // index into dL_dpix
  const size_t plane = (size_t)a.ph * a.pw;

  const float T_final = inside ? a.final_T[pix_id] : 0.f;
  const int last_contributor = inside ? (int)a.n_contrib[pix_id] : 0;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  if (inside) {
    dLp0 = a.dL_dpix[win_id];
    dLp1 = a.dL_dpix[plane + win_id];
    dLp2 = a.dL_dpix[2 * plane + win_id];
  }
  pxc[(lane >> 3) * 9 + (lane & 7)] = make_float4(dLp0, dLp1, dLp2, 0.f);  // row stride 9: no bank conflicts
  const float bg_dot_dpixel = a.bg[0] * dLp0 + a.bg[1] * dLp1 + a.bg[2] * dLp2;

  if (tid == 0) {
    *s_max_last_p = 0;
    gcr_mbar_init(&full_bar[0], kBlendThreads);
    gcr_mbar_init(&full_bar[1], kBlendThreads);
    gcr_mbar_fence_init();
  }
  __syncthreads();
  int warp_last = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
  if (lane == 0 && warp_last > 0) atomicMax(s_max_last_p, warp_last);
  __syncthreads();
  const int m = min(n, *s_max_last_p);
  const int nb = (m + kBlendBatch - 1) / kBlendBatch;

  const uint32_t* __restrict__ ids = a.point_list + range.x;
  // batch bb (bb = 0 is the LAST one) covers list indices [lo, hi), hi = m - bb*BATCH; thread t
  // fetches index lo + t.  The id of the next batch's instance rides in a register.
  auto batch_id = [&](int bb) -> int {
    const int hi = m - bb * kBlendBatch;
    if (hi <= 0) return -1;
    const int j = max(0, hi - kBlendBatch) + tid;
    return j < hi ? (int)ids[j] : -1;
  };
  if (nb > 0) {
    const int id0 = batch_id(0);
    gcr_gather_record(&stage[0][tid], a.records, id0, &full_bar[0]);
    uint32_t tag = (uint32_t)id0;
    if (striped && id0 >= 0) tag |= (uint32_t)a.owner[id0] << kOwnerShift;
    stage_id[0][tid] = tag;
  }
  int next_id = batch_id(1);
  __syncthreads();   // stage_id / stage_owner of batch 0

  float T = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  float last_alpha = 0.f;
  float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
  const float ddelx_dx = 0.5f * a.W;
  const float ddely_dy = 0.5f * a.H;

  // phase-B roles: record slot r, pixel row q of the 8x4 block (pixels q*8 .. q*8+7)
  const int rB = lane & 7, qB = lane >> 3;
  const float pyB = (float)(sub_y0 + qB);

  for (int b = 0; b < nb; ++b) {
    const int s = b & 1;
    uint32_t fetched_tag = 0;
    if (b + 1 < nb) {
      gcr_gather_record(&stage[s ^ 1][tid], a.records, next_id, &full_bar[s ^ 1]);
      fetched_tag = (uint32_t)next_id;
      // consumed at the end of this batch: the load's latency hides behind the batch's work
      if (striped && next_id >= 0) fetched_tag |= (uint32_t)a.owner[next_id] << kOwnerShift;
      next_id = batch_id(b + 2);
    }
    gcr_mbar_wait(&full_bar[s], (uint32_t)((b >> 1) & 1));

    const int hi = m - b * kBlendBatch, lo = max(0, hi - kBlendBatch);
    const int cnt = hi - lo;
    const GcrRecord* __restrict__ st = stage[s];
    int nslot = 0;

    // ---- phase B: accumulate the panel's records over the warp's 32 pixels, emit ----
    auto flush = [&](int count) {
      __syncwarp();
      float g_mx = 0.f, g_my = 0.f, g_ca = 0.f, g_cb = 0.f, g_cc = 0.f, g_op = 0.f;
      float g_r = 0.f, g_g = 0.f, g_b = 0.f;
      uint32_t gidx = 0;
      int owner = 0;
      if (rB < count) {
        const int jj = slot_j[rB];
        const float4 r0 = st[jj].q0;
        const float2 r1 = *reinterpret_cast<const float2*>(&st[jj].q1);
        gidx = stage_id[s][jj];
        if (striped) {
          owner = (int)(gidx >> kOwnerShift);
          gidx &= (1u << kOwnerShift) - 1u;
        }
        const float* wa = panel + (0 * kSlots + rB) * kPanelStride + qB * 8;
        const float* wb = panel + (1 * kSlots + rB) * kPanelStride + qB * 8;
        const float* wg = panel + (2 * kSlots + rB) * kPanelStride + qB * 8;
        const float dy = r0.y - pyB;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float aT = wa[i], dA = wb[i], G = wg[i];
          const float4 dl = pxc[qB * 9 + i];
          const float dx = r0.x - (float)(sub_x0 + i);
          g_r = fmaf(aT, dl.x, g_r);
          g_g = fmaf(aT, dl.y, g_g);
          g_b = fmaf(aT, dl.z, g_b);
          const float dL_dG = r1.y * dA;
          const float gdx = G * dx, gdy = G * dy;
          const float dGx = -gdx * r0.z - gdy * r0.w;
          const float dGy = -gdy * r1.x - gdx * r0.w;
          g_mx = fmaf(dL_dG, dGx, g_mx);
          g_my = fmaf(dL_dG, dGy, g_my);
          const float t = gdx * dL_dG;
          g_ca = fmaf(t, dx, g_ca);
          g_cb = fmaf(t, dy, g_cb);
          g_cc = fmaf(gdy * dy, dL_dG, g_cc);
          g_op = fmaf(G, dA, g_op);
        }
      }
      // combine the four pixel rows (lanes r, r+8, r+16, r+24)
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        g_mx += __shfl_xor_sync(0xffffffffu, g_mx, o);
        g_my += __shfl_xor_sync(0xffffffffu, g_my, o);
        g_ca += __shfl_xor_sync(0xffffffffu, g_ca, o);
        g_cb += __shfl_xor_sync(0xffffffffu, g_cb, o);
        g_cc += __shfl_xor_sync(0xffffffffu, g_cc, o);
        g_op += __shfl_xor_sync(0xffffffffu, g_op, o);
        g_r += __shfl_xor_sync(0xffffffffu, g_r, o);
        g_g += __shfl_xor_sync(0xffffffffu, g_g, o);
        g_b += __shfl_xor_sync(0xffffffffu, g_b, o);
      }
      if (qB == 0 && rB < count) {
        GcrGradAcc* dst = a.acc[owner] + gidx;
        if (a.remote_scalar && owner != a.self_rank) {
          // fallback for fabrics that do not take 16-byte vector reductions on peer memory
          float* f = reinterpret_cast<float*>(dst);
          atomicAdd(f + 0, g_mx * ddelx_dx); atomicAdd(f + 1, g_my * ddely_dy);
          atomicAdd(f + 2, -0.5f * g_ca); atomicAdd(f + 3, -0.5f * g_cb);
          atomicAdd(f + 4, -0.5f * g_cc); atomicAdd(f + 5, g_op);
          atomicAdd(f + 6, g_r); atomicAdd(f + 7, g_g); atomicAdd(f + 8, g_b);
        } else {
          gcr_red_add_v4(&dst->g0, g_mx * ddelx_dx, g_my * ddely_dy, -0.5f * g_ca, -0.5f * g_cb);
          gcr_red_add_v4(&dst->g1, -0.5f * g_cc, g_op, g_r, g_g);
          atomicAdd(&dst->g2.x, g_b);
        }
      }
      __syncwarp();
    };

    if (warp_last > lo) {
      for (int g0 = ((cnt - 1) >> 5) << 5; g0 >= 0; g0 -= 32) {
        if (lo + g0 >= warp_last) continue;
        const int j = g0 + lane;
        bool touch = false;
        if (j < cnt && lo + j < warp_last) {
          const float4 q0 = st[j].q0;   // x, y, A, B
          const float4 q1 = st[j].q1;   // C, o, 2 ln(255 o), -
          touch = gcr_subrect_touch(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rx0, rx1, ry0, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, touch);
        while (mask) {
          const int bit = 31 - __clz(mask);
          mask &= ~(1u << bit);
          const int jj = g0 + bit;
          const float4 r0 = st[jj].q0;
          const float2 r1 = *reinterpret_cast<const float2*>(&st[jj].q1);   // C, o
          const float4 col = st[jj].q2;                                     // r, g, b, -
          // straight-line evaluation (the warp is issue-bound: no per-test branches); the
          // per-pixel recurrences are only committed on contributing lanes
          const float dx = __fsub_rn(r0.x, pxf);
          const float dy = __fsub_rn(r0.y, pyf);
          const float power = gcr_power(dx, dy, r0.z, r0.w, r1.x);
          float G = gcr_ex2_approx(power * 1.4426950408889634f);
          float alpha = fminf(0.99f, r1.y * G);
          if (fabsf(alpha - 1.0f / 255.0f) < 1e-7f) {   // borderline: decide exactly like the forward
            G = expf(power);
            alpha = fminf(0.99f, __fmul_rn(r1.y, G));
          }
          // reference: contributor-- ; if (contributor >= last_contributor) continue; power > 0 and
          // alpha < 1/255 skip as well
          const bool contrib = (lo + jj < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          float wA = 0.f, wB = 0.f, wG = 0.f;
          if (contrib) {
            const float inv_1ma = gcr_rcp_approx(1.f - alpha);
            T = T * inv_1ma;
            const float one_m_la = 1.f - last_alpha;
            acc0 = fmaf(last_alpha, lc0, one_m_la * acc0);
            acc1 = fmaf(last_alpha, lc1, one_m_la * acc1);
            acc2 = fmaf(last_alpha, lc2, one_m_la * acc2);
            lc0 = col.x; lc1 = col.y; lc2 = col.z;
            float dL_dalpha = (col.x - acc0) * dLp0;
            dL_dalpha = fmaf(col.y - acc1, dLp1, dL_dalpha);
            dL_dalpha = fmaf(col.z - acc2, dLp2, dL_dalpha);
            last_alpha = alpha;
            wB = fmaf(dL_dalpha, T, (-T_final * inv_1ma) * bg_dot_dpixel);
            wA = alpha * T;
            wG = G;
          }
          if (__ballot_sync(0xffffffffu, contrib) != 0u) {
            panel[(0 * kSlots + nslot) * kPanelStride + lane] = wA;
            panel[(1 * kSlots + nslot) * kPanelStride + lane] = wB;
            panel[(2 * kSlots + nslot) * kPanelStride + lane] = wG;
            if (lane == 0) slot_j[nslot] = jj;
            if (++nslot == kSlots) {
              flush(kSlots);
              nslot = 0;
            }
          }
        }
      }
    }
    if (nslot > 0) flush(nslot);   // stage s is recycled after this batch's barrier
    stage_id[s ^ 1][tid] = fetched_tag;   // nobody reads side s^1 during this batch
    __syncthreads();
  }
}

}  // namespace

cudaError_t gcr_launch_blend_bwd(const GcrBlendArgs& a, cudaStream_t stream) {
  if (a.tiles_x <= 0 || a.tiles_y <= 0) return cudaSuccess;
  dim3 grid(a.tiles_x, a.stripe != nullptr ? a.grid_y : a.tiles_y, 1);
  static std::atomic<unsigned long long> configured{0ull};
  cudaError_t e = gcr_set_dynamic_smem_once(blend_bwd_kernel, kBwdSmemBytes, configured);
  if (e != cudaSuccess) return e;
  return gcr_launch_chain<GCR_EDGE_BLEND_BWD>(blend_bwd_kernel, grid, dim3(kBlendThreads), kBwdSmemBytes, stream, a);
}
