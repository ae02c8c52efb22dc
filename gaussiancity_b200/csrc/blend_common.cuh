// blend_common.cuh -- pieces shared by the forward and backward per-tile blend kernels.
#pragma once
#include "gcr_common.cuh"

// One CTA = one 16x16 tile, 8 warps; warp w owns the 8x4 pixel sub-rectangle
//   x in [tile_x*16 + (w&1)*8, +8),  y in [tile_y*16 + (w>>1)*4, +4)
// and lane l the pixel (l&7, l>>3) inside it.
constexpr int kBlendThreads = 256;
constexpr int kBlendBatch = 256;   // tile-instances per TMA stage (12 KB)
constexpr int kBlendStages = 2;

// Conservative sub-rectangle cull.  Returns false only if NO pixel centre in
// [rx0,rx1]x[ry0,ry1] can reach alpha >= 1/255 for this splat, i.e. if the minimum of
//   q(d) = A dx^2 + 2 B dx dy + C dy^2   over the rectangle exceeds 2 ln(255 o) (=twoL)
// by a margin that covers fp32 rounding of both this test and the reference's per-pixel
// `power` (DGR forward.cu:307-318).  Culled splats are exactly those every lane of the warp
// would have skipped through `alpha < 1/255`, so results (colour, final_T, n_contrib) are
// unchanged.  Any NaN / non-positive-definite conic => not culled (evaluated exactly).
__forceinline__ __device__ bool gcr_subrect_touch(float mx, float my, float A, float B, float C,
                                                  float twoL, float rx0, float rx1, float ry0,
                                                  float ry1) {
  const float dx0 = rx0 - mx, dx1 = rx1 - mx;
  const float dy0 = ry0 - my, dy1 = ry1 - my;
  const float ex = fminf(fmaxf(0.f, dx0), dx1);  // nearest x offset of the rect (0 if inside)
  const float ey = fminf(fmaxf(0.f, dy0), dy1);
  // facing vertical edge (dx = ex): minimise over dy in [dy0,dy1]
  const float yv = fminf(fmaxf(__fdividef(-B * ex, C), dy0), dy1);
  const float qv = A * ex * ex + 2.f * B * ex * yv + C * yv * yv;
  // facing horizontal edge (dy = ey): minimise over dx in [dx0,dx1]
  const float xh = fminf(fmaxf(__fdividef(-B * ey, A), dx0), dx1);
  const float qh = A * xh * xh + 2.f * B * xh * ey + C * ey * ey;
  float qmin;
  if (ex != 0.f)
    qmin = (ey != 0.f) ? fminf(qv, qh) : qv;
  else
    qmin = (ey != 0.f) ? qh : 0.f;
  const float ax = fmaxf(fabsf(dx0), fabsf(dx1));
  const float ay = fmaxf(fabsf(dy0), fabsf(dy1));
  const float mag = A * ax * ax + 2.f * fabsf(B) * ax * ay + C * ay * ay;
  // non-finite centre (inf - inf = NaN, NaN - NaN = NaN) must be evaluated exactly
  const bool finite_c = (mx - mx == 0.f) && (my - my == 0.f);
  const bool pd = finite_c && (A > 0.f) && (C > 0.f) && (A * C - B * B > 0.f);
  const bool cull = pd && (qmin > twoL + 1e-5f * mag + 1e-3f);
  return !cull;
}

// Approximate SFU forms for paths that only carry a tolerance (never the forward):
// ex2.approx.ftz (<= 2 ulp) and rcp.approx.ftz (<= 1 ulp), one MUFU each.
__forceinline__ __device__ float gcr_ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__forceinline__ __device__ float gcr_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
