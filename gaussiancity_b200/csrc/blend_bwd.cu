// blend_bwd.cu -- per-tile back-to-front gradient blend for sm_100a.
//
// Behavioural spec: DGR/cuda_rasterizer/backward.cu:428-581 (renderCUDA backward): per pixel,
// walk the tile's instance list from its last contributor backwards, rebuild T by division,
// accumulate dL/dcolor, dL/dalpha (recursive accum_rec + background term), and from it
// dL/dmean2D (x,y), dL/dconic (x,y,w), dL/dopacity for every contributing (pixel, Gaussian)
// pair.  The reference issues 9 scalar atomicAdd per pair.
//
// B200 design: same TMA-staged 2-stage ring as the forward (batches taken from the END of
// the tile's contiguous record slice), same per-warp sub-rectangle cull, records beyond the
// CTA's furthest last-contributor are never staged.  Two kernels:
//   blend_bwd_kernel_v2 (default): transposed reduction -- phase A (lane = pixel) parks three
//     scalars per (record, pixel) in a warp-private shared-memory panel, phase B (lane = record
//     x pixel row) accumulates the 9 gradient terms per record in registers and emits two
//     16-byte vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4) + one scalar.
//   blend_bwd_kernel (GCR_BLEND_BWD=v1): per record, a value-halving shuffle butterfly
//     (8 values in 9 shuffles + 1 in 5) and one predicated RED.ADD.F32 from 9 lanes.
// Both write into the 48-byte per-Gaussian accumulator GcrGradAcc.
#include <cstdlib>

#include "blend_common.cuh"
#include "gcr_kernels.h"

namespace {

__forceinline__ __device__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum 8 per-lane values across the warp with 4+2+1+1+1 = 9 shuffles (instead of 8 x 5): at
// each of the first three butterfly levels a lane keeps half of its values and ships the other
// half to its partner.  Afterwards lane L with L % 4 == 0 holds the warp total of value
//   q(L) = 4*bit4(L) + 2*bit3(L) + bit2(L).
__forceinline__ __device__ float warp_sum8_scatter(float a0, float a1, float a2, float a3,
                                                   float a4, float a5, float a6, float a7,
                                                   int lane) {
  const bool u16 = (lane & 16) != 0;
  // level 16: lower half keeps a0..a3, upper half keeps a4..a7
  float k0 = u16 ? a4 : a0, s0 = u16 ? a0 : a4;
  float k1 = u16 ? a5 : a1, s1 = u16 ? a1 : a5;
  float k2 = u16 ? a6 : a2, s2 = u16 ? a2 : a6;
  float k3 = u16 ? a7 : a3, s3 = u16 ? a3 : a7;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  k2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  k3 += __shfl_xor_sync(0xffffffffu, s3, 16);
  // level 8: keep (k0,k1) or (k2,k3)
  const bool u8 = (lane & 8) != 0;
  float m0 = u8 ? k2 : k0, t0 = u8 ? k0 : k2;
  float m1 = u8 ? k3 : k1, t1 = u8 ? k1 : k3;
  m0 += __shfl_xor_sync(0xffffffffu, t0, 8);
  m1 += __shfl_xor_sync(0xffffffffu, t1, 8);
  // level 4: keep m0 or m1
  const bool u4 = (lane & 4) != 0;
  float v = u4 ? m1 : m0, w = u4 ? m0 : m1;
  v += __shfl_xor_sync(0xffffffffu, w, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(GcrBlendArgs a) {
  __shared__ __align__(128) GcrRecord stage[kBlendStages][kBlendBatch];
  __shared__ __align__(8) uint64_t full_bar[kBlendStages];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x;
  const int tile_y = a.shard_rank + (int)blockIdx.y * a.shard_count;
  const uint2 range = a.ranges[tile_y * a.grid_x + tile_x];
  const int n = (int)(range.y - range.x);

  const int sub_x0 = tile_x * GCR_TILE_X + (warp & 1) * 8;
  const int sub_y0 = tile_y * GCR_TILE_Y + (warp >> 1) * 4;
  const int pix_x = sub_x0 + (lane & 7);
  const int pix_y = sub_y0 + (lane >> 3);
  const bool inside = pix_x < a.W && pix_y < a.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float rx0 = (float)sub_x0, rx1 = (float)(sub_x0 + 7);
  const float ry0 = (float)sub_y0, ry1 = (float)(sub_y0 + 3);
  const int pix_id = a.W * pix_y + pix_x;
  const size_t plane = (size_t)a.H * a.W;

  const float T_final = inside ? a.final_T[pix_id] : 0.f;
  const int last_contributor = inside ? (int)a.n_contrib[pix_id] : 0;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  if (inside) {
    dLp0 = a.dL_dpix[pix_id];
    dLp1 = a.dL_dpix[plane + pix_id];
    dLp2 = a.dL_dpix[2 * plane + pix_id];
  }
  const float bg_dot_dpixel = a.bg[0] * dLp0 + a.bg[1] * dLp1 + a.bg[2] * dLp2;

  // Only list positions <= the furthest last contributor of the CTA matter at all.
  __shared__ int s_max_last;
  if (tid == 0) {
    s_max_last = 0;
    gcr_mbar_init(&full_bar[0], 1);
    gcr_mbar_init(&full_bar[1], 1);
    gcr_mbar_fence_init();
  }
  __syncthreads();
  int warp_last = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
  if (lane == 0 && warp_last > 0) atomicMax(&s_max_last, warp_last);
  __syncthreads();
  const int m = min(n, s_max_last);  // records [0, m) of the tile slice are relevant
  const int nb = (m + kBlendBatch - 1) / kBlendBatch;

  const GcrRecord* __restrict__ src = a.inst + range.x;
  // batch b (b = 0 is the LAST one) covers list indices [lo_b, hi_b), hi_b = m - b*BATCH
  if (tid == 0 && nb > 0) {
    const int hi = m, lo = max(0, hi - kBlendBatch);
    const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(GcrRecord);
    gcr_mbar_expect_tx(&full_bar[0], bytes);
    gcr_bulk_g2s(&stage[0][0], src + lo, bytes, &full_bar[0]);
  }

  float T = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;       // accum_rec
  float last_alpha = 0.f;
  float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;          // last_color
  const float ddelx_dx = 0.5f * a.W;
  const float ddely_dy = 0.5f * a.H;

  for (int b = 0; b < nb; ++b) {
    const int s = b & 1;
    if (tid == 0 && b + 1 < nb) {
      const int hi = m - (b + 1) * kBlendBatch, lo = max(0, hi - kBlendBatch);
      const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(GcrRecord);
      gcr_mbar_expect_tx(&full_bar[s ^ 1], bytes);
      gcr_bulk_g2s(&stage[s ^ 1][0], src + lo, bytes, &full_bar[s ^ 1]);
    }
    gcr_mbar_wait(&full_bar[s], (uint32_t)((b >> 1) & 1));

    const int hi = m - b * kBlendBatch, lo = max(0, hi - kBlendBatch);
    const int cnt = hi - lo;
    const GcrRecord* __restrict__ st = stage[s];
    // list index of st[j] is lo + j; 1-based position = lo + j + 1
    if (warp_last > lo) {
      for (int g0 = ((cnt - 1) >> 5) << 5; g0 >= 0; g0 -= 32) {
        if (lo + g0 >= warp_last) continue;  // whole group behind every pixel's last contributor
        const int j = g0 + lane;
        bool touch = false;
        if (j < cnt && lo + j < warp_last) {
          const float4 q0 = st[j].q0;
          const float2 q1 = *reinterpret_cast<const float2*>(&st[j].q1);
          const float twoL = st[j].q2.z;
          touch = gcr_subrect_touch(q0.x, q0.y, q0.z, q0.w, q1.x, twoL, rx0, rx1, ry0, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, touch);
        while (mask) {
          const int bit = 31 - __clz(mask);
          mask &= ~(1u << bit);
          const int jj = g0 + bit;
          const float4 r0 = st[jj].q0;
          const float4 r1 = st[jj].q1;
          const float4 r2 = st[jj].q2;
          float g_mx = 0.f, g_my = 0.f, g_ca = 0.f, g_cb = 0.f, g_cc = 0.f, g_op = 0.f;
          float g_r = 0.f, g_g = 0.f, g_b = 0.f;
          bool contrib = false;
          // reference: contributor-- ; if (contributor >= last_contributor) continue;
          if (lo + jj < last_contributor) {
            const float dx = __fsub_rn(r0.x, pxf);
            const float dy = __fsub_rn(r0.y, pyf);
            const float power = gcr_power(dx, dy, r0.z, r0.w, r1.x);
            if (!(power > 0.0f)) {
              const float G = expf(power);
              const float alpha = fminf(0.99f, __fmul_rn(r1.y, G));
              if (!(alpha < 1.0f / 255.0f)) {
                contrib = true;
                // one IEEE reciprocal serves both divisions by (1 - alpha) (<= 1 ulp from x / y)
                const float inv_1ma = __frcp_rn(1.f - alpha);
                T = T * inv_1ma;
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.f;
                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                lc0 = r1.z;
                dL_dalpha += (r1.z - acc0) * dLp0;
                g_r = dchannel_dcolor * dLp0;
                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                lc1 = r1.w;
                dL_dalpha += (r1.w - acc1) * dLp1;
                g_g = dchannel_dcolor * dLp1;
                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                lc2 = r2.x;
                dL_dalpha += (r2.x - acc2) * dLp2;
                g_b = dchannel_dcolor * dLp2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final * inv_1ma) * bg_dot_dpixel;

                const float dL_dG = r1.y * dL_dalpha;
                const float gdx = G * dx;
                const float gdy = G * dy;
                const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
                const float dG_ddely = -gdy * r1.x - gdx * r0.w;
                g_mx = dL_dG * dG_ddelx * ddelx_dx;
                g_my = dL_dG * dG_ddely * ddely_dy;
                g_ca = -0.5f * gdx * dx * dL_dG;
                g_cb = -0.5f * gdx * dy * dL_dG;
                g_cc = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
              }
            }
          }
          if (__ballot_sync(0xffffffffu, contrib) != 0u) {
            // accumulator float layout: 0 mean2D.x, 1 mean2D.y, 2 conic.x, 3 conic.y,
            // 4 conic.w, 5 opacity, 6 colour.r, 7 colour.g, 8 colour.b
            const float tot = warp_sum8_scatter(g_mx, g_my, g_ca, g_cb, g_cc, g_op, g_r, g_g, lane);
            const float tot_b = warp_sum(g_b);
            if ((lane & 3) == 0 || lane == 1) {
              const int q = (lane == 1) ? 8 : (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1));
              float* dst = reinterpret_cast<float*>(a.grad_acc + __float_as_uint(r2.y)) + q;
              atomicAdd(dst, (lane == 1) ? tot_b : tot);
            }
          }
        }
      }
    }
    __syncthreads();  // stage s may be refilled at the top of iteration b+2
  }
}


// ------------------------------------------------------------------------------------------
// v2: transposed reduction.  Phase A (lane = pixel) walks the surviving records back to front
// exactly like v1 but, instead of reducing 9 gradient terms across the warp per record, parks
// three scalars per (record, pixel) -- alpha*T, dL/dalpha, G -- in a warp-private shared-memory
// panel of kSlots records.  Phase B (lane = record x pixel-row) then accumulates the 9 terms of
// each record over the pixels in registers: no shuffles per record, full lane utilisation, one
// quarter-combine (2 shuffle levels) per panel, and the sums leave as two 16-byte vector
// reductions + one scalar per record.
// ------------------------------------------------------------------------------------------
constexpr int kSlots = 8;                       // records per panel
constexpr int kPanelStride = 33;                // padded pixel stride (bank-conflict free)
constexpr int kWarpPanelFloats = 3 * kSlots * kPanelStride;            // 792
constexpr int kWarpScratchBytes = 3840;         // panel 3168 + slot_j 32 + pixel consts 576, padded
constexpr int kBwdV2SmemBytes = kBlendStages * kBlendBatch * (int)sizeof(GcrRecord) + 8 * kWarpScratchBytes + 64;

struct BwdRec { float x, y, A, B, C, o; uint32_t gidx; };

//
// kApprox (experimental, env GCR_BWD_MATH=approx; not yet measured on hardware): the gradients
// only have to meet the 1e-4 bar, so exp becomes ex2.approx(power * log2 e) (2 instructions instead
// of ~10) and 1/(1-alpha) one MUFU.RCP (instead of the IEEE reciprocal's ~8).  The set of
// contributors must still be EXACTLY the forward's (a flipped alpha >= 1/255 decision would
// rescale the rest of that pixel's T chain), so any alpha within 1e-7 of the threshold -- 30x the
// worst-case error of the approximation there -- is re-evaluated with the exact expf.
template <bool kApprox>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel_v2(GcrBlendArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GcrRecord (*stage)[kBlendBatch] = reinterpret_cast<GcrRecord (*)[kBlendBatch]>(smem_raw);
  unsigned char* scratch0 = smem_raw + kBlendStages * kBlendBatch * sizeof(GcrRecord);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch0 + 8 * kWarpScratchBytes);
  int* s_max_last_p = reinterpret_cast<int*>(full_bar + kBlendStages);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* panel = reinterpret_cast<float*>(scratch0 + warp * kWarpScratchBytes);
  int* slot_j = reinterpret_cast<int*>(panel + kWarpPanelFloats);
  float4* pxc = reinterpret_cast<float4*>(slot_j + kSlots);

  const int tile_x = blockIdx.x;
  const int tile_y = a.shard_rank + (int)blockIdx.y * a.shard_count;
  const uint2 range = a.ranges[tile_y * a.grid_x + tile_x];
  const int n = (int)(range.y - range.x);

  const int sub_x0 = tile_x * GCR_TILE_X + (warp & 1) * 8;
  const int sub_y0 = tile_y * GCR_TILE_Y + (warp >> 1) * 4;
  const int pix_x = sub_x0 + (lane & 7);
  const int pix_y = sub_y0 + (lane >> 3);
  const bool inside = pix_x < a.W && pix_y < a.H;
  const float pxf = (float)pix_x, pyf = (float)pix_y;
  const float rx0 = (float)sub_x0, rx1 = (float)(sub_x0 + 7);
  const float ry0 = (float)sub_y0, ry1 = (float)(sub_y0 + 3);
  const int pix_id = a.W * pix_y + pix_x;
  const size_t plane = (size_t)a.H * a.W;

  const float T_final = inside ? a.final_T[pix_id] : 0.f;
  const int last_contributor = inside ? (int)a.n_contrib[pix_id] : 0;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  if (inside) {
    dLp0 = a.dL_dpix[pix_id];
    dLp1 = a.dL_dpix[plane + pix_id];
    dLp2 = a.dL_dpix[2 * plane + pix_id];
  }
  pxc[(lane >> 3) * 9 + (lane & 7)] = make_float4(dLp0, dLp1, dLp2, 0.f);  // row stride 9: no bank conflicts
  const float bg_dot_dpixel = a.bg[0] * dLp0 + a.bg[1] * dLp1 + a.bg[2] * dLp2;

  if (tid == 0) {
    *s_max_last_p = 0;
    gcr_mbar_init(&full_bar[0], 1);
    gcr_mbar_init(&full_bar[1], 1);
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

  const GcrRecord* __restrict__ src = a.inst + range.x;
  if (tid == 0 && nb > 0) {
    const int hi = m, lo = max(0, hi - kBlendBatch);
    const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(GcrRecord);
    gcr_mbar_expect_tx(&full_bar[0], bytes);
    gcr_bulk_g2s(&stage[0][0], src + lo, bytes, &full_bar[0]);
  }

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
    if (tid == 0 && b + 1 < nb) {
      const int hi = m - (b + 1) * kBlendBatch, lo = max(0, hi - kBlendBatch);
      const uint32_t bytes = (uint32_t)(hi - lo) * (uint32_t)sizeof(GcrRecord);
      gcr_mbar_expect_tx(&full_bar[s ^ 1], bytes);
      gcr_bulk_g2s(&stage[s ^ 1][0], src + lo, bytes, &full_bar[s ^ 1]);
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
      if (rB < count) {
        const int jj = slot_j[rB];
        const float4 r0 = st[jj].q0;
        const float2 r1 = *reinterpret_cast<const float2*>(&st[jj].q1);
        gidx = __float_as_uint(st[jj].q2.y);
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
        GcrGradAcc* dst = a.grad_acc + gidx;
        gcr_red_add_v4(&dst->g0, g_mx * ddelx_dx, g_my * ddely_dy, -0.5f * g_ca, -0.5f * g_cb);
        gcr_red_add_v4(&dst->g1, -0.5f * g_cc, g_op, g_r, g_g);
        atomicAdd(&dst->g2.x, g_b);
      }
      __syncwarp();
    };

    if (warp_last > lo) {
      for (int g0 = ((cnt - 1) >> 5) << 5; g0 >= 0; g0 -= 32) {
        if (lo + g0 >= warp_last) continue;
        const int j = g0 + lane;
        bool touch = false;
        if (j < cnt && lo + j < warp_last) {
          const float4 q0 = st[j].q0;
          const float2 q1 = *reinterpret_cast<const float2*>(&st[j].q1);
          const float twoL = st[j].q2.z;
          touch = gcr_subrect_touch(q0.x, q0.y, q0.z, q0.w, q1.x, twoL, rx0, rx1, ry0, ry1);
        }
        unsigned mask = __ballot_sync(0xffffffffu, touch);
        while (mask) {
          const int bit = 31 - __clz(mask);
          mask &= ~(1u << bit);
          const int jj = g0 + bit;
          const float4 r0 = st[jj].q0;
          const float4 r1 = st[jj].q1;
          const float cb = st[jj].q2.x;
          // straight-line evaluation (the warp is issue-bound: no per-test branches); the
          // per-pixel recurrences are only committed on contributing lanes
          const float dx = __fsub_rn(r0.x, pxf);
          const float dy = __fsub_rn(r0.y, pyf);
          const float power = gcr_power(dx, dy, r0.z, r0.w, r1.x);
          float G, alpha;
          if (kApprox) {
            G = gcr_ex2_approx(power * 1.4426950408889634f);
            alpha = fminf(0.99f, r1.y * G);
            if (fabsf(alpha - 1.0f / 255.0f) < 1e-7f) {   // borderline: decide exactly like the forward
              G = expf(power);
              alpha = fminf(0.99f, __fmul_rn(r1.y, G));
            }
          } else {
            G = expf(power);
            alpha = fminf(0.99f, __fmul_rn(r1.y, G));
          }
          // reference: contributor-- ; if (contributor >= last_contributor) continue; power > 0 and
          // alpha < 1/255 skip as well
          const bool contrib = (lo + jj < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          float wA = 0.f, wB = 0.f, wG = 0.f;
          if (contrib) {
            const float inv_1ma = kApprox ? gcr_rcp_approx(1.f - alpha) : __frcp_rn(1.f - alpha);
            T = T * inv_1ma;
            const float one_m_la = 1.f - last_alpha;
            acc0 = fmaf(last_alpha, lc0, one_m_la * acc0);
            acc1 = fmaf(last_alpha, lc1, one_m_la * acc1);
            acc2 = fmaf(last_alpha, lc2, one_m_la * acc2);
            lc0 = r1.z; lc1 = r1.w; lc2 = cb;
            float dL_dalpha = (r1.z - acc0) * dLp0;
            dL_dalpha = fmaf(r1.w - acc1, dLp1, dL_dalpha);
            dL_dalpha = fmaf(cb - acc2, dLp2, dL_dalpha);
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
    __syncthreads();
  }
}

}  // namespace

void gcr_launch_blend_bwd(const GcrBlendArgs& a, cudaStream_t stream) {
  const int rows = (a.grid_y - a.shard_rank + a.shard_count - 1) / a.shard_count;
  if (rows <= 0 || a.grid_x <= 0) return;
  dim3 grid(a.grid_x, rows, 1);
  static const bool use_v1 = [] {
    const char* e = getenv("GCR_BLEND_BWD");
    return e != nullptr && e[0] == 'v' && e[1] == '1';
  }();
  if (use_v1) {
    blend_bwd_kernel<<<grid, kBlendThreads, 0, stream>>>(a);
    return;
  }
  static const bool approx = [] {
    const char* e = getenv("GCR_BWD_MATH");
    return e != nullptr && e[0] == 'a';
  }();
  static bool configured[64] = {};   // the opt-in shared-memory size is a per-device attribute
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(blend_bwd_kernel_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kBwdV2SmemBytes);
    cudaFuncSetAttribute(blend_bwd_kernel_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kBwdV2SmemBytes);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (approx)
    blend_bwd_kernel_v2<true><<<grid, kBlendThreads, kBwdV2SmemBytes, stream>>>(a);
  else
    blend_bwd_kernel_v2<false><<<grid, kBlendThreads, kBwdV2SmemBytes, stream>>>(a);
}
