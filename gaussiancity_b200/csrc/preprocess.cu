// preprocess.cu -- per-Gaussian forward preprocessing for sm_100a.
//
// Behavioural spec: DGR/cuda_rasterizer/forward.cu:147-233 (preprocessCUDA) with
// in_frustum (auxiliary.h:132-156), computeCov3D (forward.cu:110-144), computeCov2D
// (forward.cu:69-105), computeColorFromSH (forward.cu:20-66), ndc2Pix/getRect
// (auxiliary.h:32-46).  radii / tiles_touched / depth / mean2D / conic feed integer work that
// must be bit-exact with the reference extension, so every rounding on that path is pinned
// with explicit _rn intrinsics in the order nvcc 12.9 contracts the reference expressions
// (a*b + c*d + e*f  ->  fma(e,f, fma(a,b, rn(c*d))); see DESIGN.md "bit-exactness").
//
// Layout differences from the reference (free to choose): outputs go to one
// 48-byte record per Gaussian (GcrRecord) + a depth key array + a packed clamp byte; cov3D is
// not stored (the backward recomputes it) unless a debug pointer is given.
#include <cstdio>
#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

// rn(a*b + c*d + e*f) in the contraction order nvcc uses for GLM's mat3 products.
__forceinline__ __device__ float dot3(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
// m[k]*x + m[4+k]*y + m[8+k]*z (+ m[12+k]) as auxiliary.h:48-66 compiles.
__forceinline__ __device__ float xf3(const float* __restrict__ m, int k, float x, float y,
                                     float z) {
  return __fmaf_rn(z, m[8 + k], __fmaf_rn(x, m[k], __fmul_rn(y, m[4 + k])));
}

// Projection of one Gaussian: everything of preprocessCUDA (forward.cu:147-215) that does not
// involve colour.  Shared by the forward kernel and the stripe-partition pre-pass.
struct GcrProjected {
  float vz, pix_x, pix_y, conic_x, conic_y, conic_z;
  int radius;
  uint2 rmin, rmax;
};

__forceinline__ __device__ bool gcr_project(const GcrPreprocessArgs& a, int idx, float px, float py,
                                            float pz, GcrProjected& o) {
  const float* __restrict__ V = a.viewmatrix;
  const float* __restrict__ PM = a.projmatrix;

  // in_frustum: view-space depth test (auxiliary.h:141-154)
  const float vz = __fadd_rn(xf3(V, 2, px, py, pz), V[14]);
  const bool visible = vz > 0.2f;
  if (!visible) {
    if (a.prefiltered) {
      printf("Point is filtered although prefiltered is set. This shouldn't happen!");
      __trap();
    }
    return false;
  }
  // clip-space projection (forward.cu:179-181)
  const float hx = __fadd_rn(xf3(PM, 0, px, py, pz), PM[12]);
  const float hy = __fadd_rn(xf3(PM, 1, px, py, pz), PM[13]);
  const float hw = __fadd_rn(xf3(PM, 3, px, py, pz), PM[15]);
  const float p_w = 1.0f / __fadd_rn(hw, 0.0000001f);
  const float ndc_x = __fmul_rn(hx, p_w);
  const float ndc_y = __fmul_rn(hy, p_w);

  // 3D covariance (forward.cu:110-144) -- quaternion deliberately NOT normalised.
  float c0, c1, c2, c3, c4, c5;
  if (a.cov3D_precomp != nullptr) {
    const float* c = a.cov3D_precomp + 6 * (size_t)idx;
    c0 = c[0]; c1 = c[1]; c2 = c[2]; c3 = c[3]; c4 = c[4]; c5 = c[5];
  } else {
    const float s0 = __fmul_rn(a.scale_modifier, a.scales[3 * idx + 0]);
    const float s1 = __fmul_rn(a.scale_modifier, a.scales[3 * idx + 1]);
    const float s2 = __fmul_rn(a.scale_modifier, a.scales[3 * idx + 2]);
    // rotations == nullptr: identity quaternion (GaussianCity's own call pattern,
    // utils/helpers.py:238-244) -- same arithmetic, so the result is what (1,0,0,0) gives
    const float4 q = a.rotations != nullptr ? reinterpret_cast<const float4*>(a.rotations)[idx]
                                            : make_float4(1.f, 0.f, 0.f, 0.f);
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    // R columns (GLM column-major constructor order)
    // (decoded from the reference SASS: shared products are materialised, the other one fused;
    //  yy + zz is a plain add, xz +- ry fuses r*y onto rn(x*z))
    const float R00 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
    const float R01 = __fmul_rn(2.f, __fmaf_rn(x, y, -__fmul_rn(r, z)));
    const float R02 = __fmul_rn(2.f, __fmaf_rn(r, y, __fmul_rn(x, z)));
    const float R10 = __fmul_rn(2.f, __fmaf_rn(x, y, __fmul_rn(r, z)));
    const float R11 = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, __fmul_rn(z, z))));
    const float R12 = __fmul_rn(2.f, __fmaf_rn(y, z, -__fmul_rn(r, x)));
    const float R20 = __fmul_rn(2.f, __fmaf_rn(-r, y, __fmul_rn(x, z)));
    const float R21 = __fmul_rn(2.f, __fmaf_rn(y, z, __fmul_rn(r, x)));
    const float R22 = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, __fmul_rn(y, y))));
    // M = S * R  ->  M[i][j] = s_j * R[i][j]   (column i, row j)
    const float M00 = __fmul_rn(s0, R00), M01 = __fmul_rn(s1, R01), M02 = __fmul_rn(s2, R02);
    const float M10 = __fmul_rn(s0, R10), M11 = __fmul_rn(s1, R11), M12 = __fmul_rn(s2, R12);
    const float M20 = __fmul_rn(s0, R20), M21 = __fmul_rn(s1, R21), M22 = __fmul_rn(s2, R22);
    // Sigma = M^T M : Sigma[i][j] = M[j][0]*M[i][0] + M[j][1]*M[i][1] + M[j][2]*M[i][2]
    c0 = dot3(M00, M00, M01, M01, M02, M02);
    c1 = dot3(M10, M00, M11, M01, M12, M02);
    c2 = dot3(M20, M00, M21, M01, M22, M02);
    c3 = dot3(M10, M10, M11, M11, M12, M12);
    c4 = dot3(M20, M10, M21, M11, M22, M12);
    c5 = dot3(M20, M20, M21, M21, M22, M22);
  }
  if (a.dbg_cov3D != nullptr) {
    float* c = a.dbg_cov3D + 6 * (size_t)idx;
    c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; c[4] = c4; c[5] = c5;
  }

  // 2D covariance (forward.cu:69-105), EWA splatting with clamped view-space x/z, y/z.
  float tx = __fadd_rn(xf3(V, 0, px, py, pz), V[12]);
  float ty = __fadd_rn(xf3(V, 1, px, py, pz), V[13]);
  const float tz = vz;
  const float limx = __fmul_rn(1.3f, a.tan_fovx);
  const float limy = __fmul_rn(1.3f, a.tan_fovy);
  const float txtz = tx / tz;
  const float tytz = ty / tz;
  tx = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), tz);
  ty = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), tz);
  const float tz2 = __fmul_rn(tz, tz);
  const float J00 = a.focal_x / tz;
  const float J02 = -__fmul_rn(a.focal_x, tx) / tz2;
  const float J11 = a.focal_y / tz;
  const float J12 = -__fmul_rn(a.focal_y, ty) / tz2;
  // T = W * J with W[k][j] = V[4*j + k]; T[0][j] = fma(W2j,J02, rn(W0j*J00)),
  // T[1][j] = fma(W2j,J12, rn(W1j*J11)), T[2][j] = 0.
  const float T00 = __fmaf_rn(V[2], J02, __fmul_rn(V[0], J00));
  const float T01 = __fmaf_rn(V[6], J02, __fmul_rn(V[4], J00));
  const float T02 = __fmaf_rn(V[10], J02, __fmul_rn(V[8], J00));
  const float T10 = __fmaf_rn(V[2], J12, __fmul_rn(V[1], J11));
  const float T11 = __fmaf_rn(V[6], J12, __fmul_rn(V[5], J11));
  const float T12 = __fmaf_rn(V[10], J12, __fmul_rn(V[9], J11));
  // X = T^T * Vrk^T : X[i][j] = T[j][0]*S(i,0) + T[j][1]*S(i,1) + T[j][2]*S(i,2)
  const float X00 = dot3(T00, c0, T01, c1, T02, c2);
  const float X10 = dot3(T00, c1, T01, c3, T02, c4);
  const float X20 = dot3(T00, c2, T01, c4, T02, c5);
  const float X01 = dot3(T10, c0, T11, c1, T12, c2);
  const float X11 = dot3(T10, c1, T11, c3, T12, c4);
  const float X21 = dot3(T10, c2, T11, c4, T12, c5);
  // cov = X * T : cov[i][j] = X[0][j]*T[i][0] + X[1][j]*T[i][1] + X[2][j]*T[i][2]
  float cov_x = dot3(X00, T00, X10, T01, X20, T02);
  const float cov_y = dot3(X01, T00, X11, T01, X21, T02);
  float cov_z = dot3(X01, T10, X11, T11, X21, T12);
  cov_x = __fadd_rn(cov_x, 0.3f);
  cov_z = __fadd_rn(cov_z, 0.3f);

  // conic = inverse (forward.cu:196-201)
  const float det = __fmaf_rn(cov_x, cov_z, -__fmul_rn(cov_y, cov_y));
  if (det == 0.0f) return false;
  const float det_inv = 1.f / det;
  o.conic_x = __fmul_rn(cov_z, det_inv);
  o.conic_y = __fmul_rn(cov_y, -det_inv);
  o.conic_z = __fmul_rn(cov_x, det_inv);

  // screen-space extent (forward.cu:203-215)
  const float mid = __fmul_rn(0.5f, __fadd_rn(cov_x, cov_z));
  const float sq = sqrtf(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));
  const float lambda1 = __fadd_rn(mid, sq);
  const float lambda2 = __fsub_rn(mid, sq);
  const float my_radius = ceilf(__fmul_rn(3.f, sqrtf(fmaxf(lambda1, lambda2))));
  o.pix_x = gcr_ndc2pix(ndc_x, a.W);
  o.pix_y = gcr_ndc2pix(ndc_y, a.H);
  o.radius = (int)my_radius;
  o.vz = vz;
  gcr_get_rect(o.pix_x, o.pix_y, o.radius, a.grid_x, a.grid_y, o.rmin, o.rmax);
  return (o.rmax.x - o.rmin.x) * (o.rmax.y - o.rmin.y) != 0;
}

// ---- stripes ------------------------------------------------------------------------------------
// A Gaussian's global tile rect and centre tile row in one 64-bit word (12 bits each: images up to
// 65 536 px a side), 0 = culled: what the stripe selection needs once the bounds are known.
__forceinline__ __device__ unsigned long long pack_rect(const GcrProjected& g) {
  // centre row, clamped into the rows the rect touches (it always is one of them: radius >= 1)
  int cr = (int)floorf(g.pix_y * (1.0f / GCR_TILE_Y));
  cr = min(max(cr, (int)g.rmin.y), (int)g.rmax.y - 1);
  return (unsigned long long)g.rmin.x | ((unsigned long long)g.rmax.x << 12) |
         ((unsigned long long)g.rmin.y << 24) | ((unsigned long long)g.rmax.y << 36) |
         ((unsigned long long)cr << 48) | (1ull << 63);
}

// What one rank makes of a visible Gaussian once the stripe bounds are known: the owner (rank
// whose stripe [bounds[r], bounds[r+1]) holds the centre row), the rect clipped to this rank's
// stripe in the form the emit stage reads (x0 | y0 << 16, width | rows << 16) and its tile count.
struct StripeSelect {
  uint32_t tiles;
  uint2 rect;
  uint8_t owner;
};
__forceinline__ __device__ StripeSelect stripe_select(unsigned long long packed, const GcrPreprocessArgs& a) {
  const int rx0 = (int)(packed & 0xFFFu), rx1 = (int)((packed >> 12) & 0xFFFu);
  const int ry0 = (int)((packed >> 24) & 0xFFFu), ry1 = (int)((packed >> 36) & 0xFFFu);
  const int cr = (int)((packed >> 48) & 0xFFFu);
  int row0 = a.win_row0, row1 = a.win_row1, owner = 0;
  if (a.stripe_bounds != nullptr) {
    row0 = max(row0, a.stripe_bounds[a.shard_rank]);
    row1 = min(row1, a.stripe_bounds[a.shard_rank + 1]);
    for (int r = 1; r < a.shard_count; ++r)
      if (cr >= a.stripe_bounds[r]) owner = r;
  }
  const int x0 = max(rx0, a.win_col0), x1 = min(rx1, a.win_col1);
  const int y0 = max(ry0, row0), y1 = min(ry1, row1);
  StripeSelect o;
  o.owner = (uint8_t)owner;
  o.tiles = (y1 > y0 && x1 > x0) ? (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0) : 0u;
  o.rect = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)(x1 - x0) | ((uint32_t)(y1 - y0) << 16));
  return o;
}

// SH colour of one Gaussian (computeColorFromSH, forward.cu:20-66): rgb clamped at 0, returns
// the clamp flags (bit ch set when channel ch was negative before the clamp).
__forceinline__ __device__ uint8_t gcr_sh_colour(const GcrPreprocessArgs& a, const int idx, const float px,
                                                  const float py, const float pz, float& cr, float& cg,
                                                  float& cb) {
    const float dx = __fsub_rn(px, a.campos[0]);
    const float dy = __fsub_rn(py, a.campos[1]);
    const float dz = __fsub_rn(pz, a.campos[2]);
    const float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
    const float x = dx / len, y = dy / len, z = dz / len;
    const float4* __restrict__ sh4 =
        reinterpret_cast<const float4*>(a.shs + (size_t)idx * a.M * 3);
    // coefficients are read as float4 (M*3 floats = 3M/4 float4 when M in {1,4,9,16}:
    // M=1 -> 3 floats (not a float4 multiple) so fall back to scalar loads there).
    float sh[48];
    const int nfl = a.M * 3;
    if ((nfl & 7) == 0) {
      const float* __restrict__ shp = a.shs + (size_t)idx * a.M * 3;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (k * 8 < nfl) {
          float v[8];
          gcr_ldg_nc_v8(shp + 8 * k, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) sh[8 * k + i] = v[i];
        }
      }
    } else if ((nfl & 3) == 0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        if (k * 4 < nfl) {
          const float4 v = __ldg(sh4 + k);
          sh[4 * k + 0] = v.x; sh[4 * k + 1] = v.y; sh[4 * k + 2] = v.z; sh[4 * k + 3] = v.w;
        }
      }
    } else {
      const float* __restrict__ shf = a.shs + (size_t)idx * a.M * 3;
#pragma unroll
      for (int k = 0; k < 48; ++k)
        if (k < nfl) sh[k] = __ldg(shf + k);
    }
#define SHC(k, ch) sh[3 * (k) + (ch)]
    // Basis scalars and the accumulation chain are pinned to the contraction nvcc emits for
    // the reference's glm::vec3 expression (decoded from its SASS, forward.cu:20-66): every
    // basis scalar is built with plain mul/add except 3xx-yy, 4zz-xx, 2zz-3xx-3yy, xx-3yy
    // (fma), then res = fma(scalar, sh[k], res) in coefficient order.
    float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f, s8 = 0.f;
    float s9 = 0.f, s10 = 0.f, s11 = 0.f, s12 = 0.f, s13 = 0.f, s14 = 0.f, s15 = 0.f;
    if (a.D > 0) {
      s1 = __fmul_rn(y, GCR_SH_C1);
      s2 = __fmul_rn(z, GCR_SH_C1);
      s3 = __fmul_rn(x, GCR_SH_C1);
      if (a.D > 1) {
        const float xy = __fmul_rn(y, x), zy = __fmul_rn(z, y), zx = __fmul_rn(z, x);
        const float zz = __fmul_rn(z, z), xx = __fmul_rn(x, x), yy = __fmul_rn(y, y);
        const float zz2 = __fadd_rn(zz, zz);
        const float xx_yy = __fsub_rn(xx, yy);
        s4 = __fmul_rn(xy, GCR_SH_C2[0]);
        s5 = __fmul_rn(zy, GCR_SH_C2[1]);
        s6 = __fmul_rn(__fsub_rn(__fsub_rn(zz2, xx), yy), GCR_SH_C2[2]);
        s7 = __fmul_rn(zx, GCR_SH_C2[3]);
        s8 = __fmul_rn(xx_yy, GCR_SH_C2[4]);
        if (a.D > 2) {
          const float q = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);                 // 4zz - xx - yy
          const float u = __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2));          // 2zz - 3xx - 3yy
          s9 = __fmul_rn(__fmul_rn(y, GCR_SH_C3[0]), __fmaf_rn(xx, 3.0f, -yy));
          s10 = __fmul_rn(__fmul_rn(xy, GCR_SH_C3[1]), z);
          s11 = __fmul_rn(__fmul_rn(y, GCR_SH_C3[2]), q);
          s12 = __fmul_rn(__fmul_rn(z, GCR_SH_C3[3]), u);
          s13 = __fmul_rn(q, __fmul_rn(x, GCR_SH_C3[4]));
          s14 = __fmul_rn(xx_yy, __fmul_rn(z, GCR_SH_C3[5]));
          s15 = __fmul_rn(__fmul_rn(x, GCR_SH_C3[6]), __fmaf_rn(yy, -3.0f, xx));
        }
      }
    }
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = __fmul_rn(GCR_SH_C0, SHC(0, ch));
      if (a.D > 0) {
        v = __fmaf_rn(-s1, SHC(1, ch), v);
        v = __fmaf_rn(s2, SHC(2, ch), v);
        v = __fmaf_rn(-s3, SHC(3, ch), v);
        if (a.D > 1) {
          v = __fmaf_rn(s4, SHC(4, ch), v);
          v = __fmaf_rn(s5, SHC(5, ch), v);
          v = __fmaf_rn(s6, SHC(6, ch), v);
          v = __fmaf_rn(s7, SHC(7, ch), v);
          v = __fmaf_rn(s8, SHC(8, ch), v);
          if (a.D > 2) {
            v = __fmaf_rn(s9, SHC(9, ch), v);
            v = __fmaf_rn(s10, SHC(10, ch), v);
            v = __fmaf_rn(s11, SHC(11, ch), v);
            v = __fmaf_rn(s12, SHC(12, ch), v);
            v = __fmaf_rn(s13, SHC(13, ch), v);
            v = __fmaf_rn(s14, SHC(14, ch), v);
            v = __fmaf_rn(s15, SHC(15, ch), v);
          }
        }
      }
      res[ch] = __fadd_rn(v, 0.5f);
    }
#undef SHC
    const uint8_t cl = (res[0] < 0 ? 1 : 0) | (res[1] < 0 ? 2 : 0) | (res[2] < 0 ? 4 : 0);
    cr = fmaxf(res[0], 0.0f);
    cg = fmaxf(res[1], 0.0f);
    cb = fmaxf(res[2], 0.0f);
  return cl;
}

// ---- kernel 1: projection -----------------------------------------------------------------------
// One thread per Gaussian, geometry only (44 B read; the colour comes later and only for the
// Gaussians that are rendered).  Writes radii, the depth key, the geometry half of the record
// and -- on the colors_precomp path -- its colour.
//   kDeferred == false: the stripe (if any) is known: tiles_touched / clipped rect / owner are
//     finished here and the CTA adds its tile count to num_rendered.
//   kDeferred == true : balanced stripes are wanted and can only be cut once every Gaussian's
//     rect is known: the kernel stores the packed global rect, accumulates the tile-instance
//     count of every tile row, and the LAST CTA cuts the rows into shard_count stripes of about
//     equal instance count (deterministic: every rank computes the same bounds from the same
//     inputs, so the ranks never have to talk).  stripe_select_kernel then finishes the job.
constexpr int kPartRowsSmem = 1024;

__device__ void cut_stripes(volatile const uint32_t* h, int grid_y, int n, int* __restrict__ bounds) {
  unsigned long long total = 0;
  for (int r = 0; r < grid_y; ++r) total += h[r];
  bounds[0] = 0;
  bounds[n] = grid_y;
  if (total == 0) {
    for (int k = 1; k < n; ++k) bounds[k] = (int)((long long)grid_y * k / n);
    return;
  }
  // bounds[k] = the row boundary whose instance prefix is nearest to k/n of the total
  unsigned long long prefix = 0;   // instances in rows < r
  int r = 0;
  for (int k = 1; k < n; ++k) {
    const unsigned long long target = total * (unsigned long long)k / (unsigned long long)n;
    while (r < grid_y && prefix + h[r] <= target) prefix += h[r++];
    // boundary at r (prefix <= target) or r + 1 (prefix + h[r] > target): take the nearer
    if (r < grid_y && (prefix + h[r] - target) < (target - prefix)) prefix += h[r++];
    bounds[k] = r;
  }
}

template <bool kHasSH, bool kDeferred>
__global__ void __launch_bounds__(256, (kHasSH && !kDeferred) ? 3 : 6)
project_kernel(GcrPreprocessArgs a) {
  __shared__ uint32_t hist[kDeferred ? kPartRowsSmem : 1];
  __shared__ uint32_t s_sum;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const bool use_smem = a.grid_y <= kPartRowsSmem;
  if (tid == 0) s_sum = 0;
  if (kDeferred && use_smem)
    for (int r = tid; r < a.grid_y; r += blockDim.x) hist[r] = 0;
  __syncthreads();
  uint32_t my_tiles = 0;
  // deferred: persistent CTAs (one row-histogram flush each); direct: one Gaussian per thread
  const int stride = kDeferred ? (int)(gridDim.x * blockDim.x) : a.P;
  for (int idx = blockIdx.x * blockDim.x + tid; idx < a.P; idx += stride) {
    int radius_out = 0;
    uint32_t depth_key = 0xFFFFFFFFu;  // dropped by the first pass of the depth sort
    unsigned long long packed = 0ull;
    uint32_t tiles = 0;
    uint8_t owner_out = GCR_NO_OWNER;
    bool want_colour = false;
    const float px = a.means3D[3 * idx + 0];
    const float py = a.means3D[3 * idx + 1];
    const float pz = a.means3D[3 * idx + 2];
    GcrProjected g;
    if (gcr_project(a, idx, px, py, pz, g)) {
      radius_out = g.radius;
      packed = pack_rect(g);
      bool rendered = true;
      if (kDeferred) {
        const uint32_t w = g.rmax.x - g.rmin.x;
        for (uint32_t r = g.rmin.y; r < g.rmax.y; ++r) atomicAdd(use_smem ? &hist[r] : &a.row_hist[r], w);
      } else {
        const StripeSelect sel = stripe_select(packed, a);
        owner_out = sel.owner;
        tiles = sel.tiles;
        rendered = tiles != 0;
        if (rendered) a.rects[idx] = sel.rect;
      }
      if (rendered) {
        const float opacity = a.opacities != nullptr ? a.opacities[idx] : 1.0f;   // nullptr: opaque
        GcrRecord* rec = a.records + idx;
        rec->q0 = make_float4(g.pix_x, g.pix_y, g.conic_x, g.conic_y);
        // cull threshold 2*ln(255*opacity): a pixel can only reach alpha >= 1/255 when
        // A dx^2 + 2B dx dy + C dy^2 <= this value (used with a safety margin, blend_*.cu).
        rec->q1 = make_float4(g.conic_z, opacity, 2.0f * logf(255.0f * opacity), 0.f);
        if (!kHasSH) {
          rec->q2 = make_float4(a.colors_precomp[3 * idx + 0], a.colors_precomp[3 * idx + 1],
                                a.colors_precomp[3 * idx + 2], 0.f);
        }
        depth_key = __float_as_uint(g.vz);
        want_colour = kHasSH && !kDeferred;
      }
    }
    a.radii[idx] = radius_out;
    if (kDeferred) {
      a.depth_in[idx] = depth_key;       // compacted into depth_keys by stripe_select_kernel
      a.packed_rects[idx] = packed;
    } else {
      a.depth_keys[idx] = depth_key;
      a.tiles_touched[idx] = tiles;
      a.owner[idx] = owner_out;
      my_tiles += tiles;
    }
    if (want_colour) {
      // the stripe is known, so only Gaussians that are rendered get here.  Last on purpose: the
      // 48 coefficient registers meet as few live values as possible.
      float cr, cg, cb;
      a.clamped[idx] = gcr_sh_colour(a, idx, px, py, pz, cr, cg, cb);
      a.records[idx].q2 = make_float4(cr, cg, cb, 0.f);
    }
  }
  if (!kDeferred) {
    // num_rendered = sum of the tile counts: one reduction per CTA, so the host can size the
    // binning buffer while the depth sort is still running (api.cu)
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, my_tiles);
    if ((tid & 31) == 0 && wsum != 0) atomicAdd(&s_sum, wsum);
    __syncthreads();
    if (tid == 0 && s_sum != 0) atomicAdd(a.total_tiles, (unsigned long long)s_sum);
    return;
  }
  __syncthreads();
  if (use_smem)
    for (int r = tid; r < a.grid_y; r += blockDim.x)
      if (hist[r] != 0) atomicAdd(&a.row_hist[r], hist[r]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&a.row_hist[a.grid_y], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || tid != 0) return;
  __threadfence();
  cut_stripes(a.row_hist, a.grid_y, a.shard_count, a.stripe_bounds_out);
}

// ---- kernel 2 (balanced stripes only): finish what needed the bounds ------------------------------
// Per Gaussian, in index order: owner, clipped rect, tile count (a few integer operations on the
// packed rect).  Then (SH path) the Gaussians of the chunk that reach this rank's stripe -- ~1/N
// of them -- are compacted (ballot + warp prefix) and only they load their 12*M bytes of
// coefficients and evaluate the colour, with full warps.  Index order matters: the same
// coefficient rows fetched in depth order run at a quarter of the bandwidth (measured, DRAM page
// locality: profiles/r02_*).
constexpr int kSelItems = 4;                      // Gaussians per thread and chunk
constexpr int kSelChunk = 256 * kSelItems;        // 1024: four independent loads in flight per thread

#define kSelAgg (1ull << 62)
#define kSelIncl (2ull << 62)
#define kSelVal ((1ull << 62) - 1ull)

template <bool kHasSH>
__global__ void __launch_bounds__(256, 3)
stripe_select_kernel(GcrPreprocessArgs a) {
  __shared__ uint16_t list[kSelChunk];
  __shared__ uint32_t wcount[kSelItems][8];
  __shared__ uint32_t s_sum;
  __shared__ int s_chunk;
  __shared__ uint32_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_sum = 0;
  uint32_t my_tiles = 0;
  const int nchunks = (a.P + kSelChunk - 1) / kSelChunk;
  volatile unsigned long long* status = a.select_status;
  while (true) {
    // chunks in ticket order: a chunk's predecessors have always started (look-back cannot deadlock)
    __syncthreads();   // s_chunk / s_excl / list / wcount of the previous chunk are no longer read
    if (tid == 0) s_chunk = (int)atomicAdd(a.select_ticket, 1u);
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk >= nchunks) break;
    const int base = chunk * kSelChunk;
    unsigned long long packed[kSelItems];
#pragma unroll
    for (int k = 0; k < kSelItems; ++k) {
      const int idx = base + k * 256 + tid;
      packed[k] = idx < a.P ? a.packed_rects[idx] : 0ull;
    }
    bool need[kSelItems];
    unsigned bal[kSelItems];
#pragma unroll
    for (int k = 0; k < kSelItems; ++k) {
      const int idx = base + k * 256 + tid;
      uint32_t tiles = 0;
      if (idx < a.P) {
        uint8_t owner_out = GCR_NO_OWNER;
        if (packed[k] != 0ull) {
          const StripeSelect sel = stripe_select(packed[k], a);
          owner_out = sel.owner;
          tiles = sel.tiles;
          if (tiles != 0) a.rects[idx] = sel.rect;
        }
        a.tiles_touched[idx] = tiles;
        a.owner[idx] = owner_out;
        my_tiles += tiles;
      }
      need[k] = tiles != 0;
      bal[k] = __ballot_sync(0xffffffffu, need[k]);
      if (lane == 0) wcount[k][warp] = __popc(bal[k]);
    }
    __syncthreads();
    // positions in index order: sub-chunk k, then warp, then lane
    uint32_t count = 0;
#pragma unroll
    for (int k = 0; k < kSelItems; ++k) {
      uint32_t before = count;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const uint32_t c = wcount[k][w];
        if (w < warp) before += c;
        count += c;
      }
      if (need[k]) list[before + __popc(bal[k] & ((1u << lane) - 1))] = (uint16_t)(k * 256 + tid);
    }
    // publish this chunk's member count before the colour work: successors look back late
    if (tid == 0) status[chunk] = (chunk == 0 ? kSelIncl : kSelAgg) | (unsigned long long)count;
    __syncthreads();   // list complete
    if (kHasSH) {
      for (uint32_t j = tid; j < count; j += 256) {
        const int g = base + (int)list[j];
        float cr, cg, cb;
        a.clamped[g] = gcr_sh_colour(a, g, a.means3D[3 * g + 0], a.means3D[3 * g + 1], a.means3D[3 * g + 2],
                                     cr, cg, cb);
        a.records[g].q2 = make_float4(cr, cg, cb, 0.f);
      }
    }
    // global offset of the chunk: decoupled look-back (warp 0, 32 predecessors at a time)
    if (warp == 0) {
      unsigned long long excl = 0;
      if (chunk != 0) {
        long long t = (long long)chunk - 1;
        while (true) {
          const long long i = t - lane;
          unsigned long long v;
          do {
            v = i >= 0 ? status[i] : kSelIncl;
          } while (__any_sync(0xffffffffu, (v & ~kSelVal) == 0ull));
          const unsigned incl = __ballot_sync(0xffffffffu, (v & ~kSelVal) == kSelIncl);
          const int first = __ffs(incl) - 1;
          unsigned long long c = (incl == 0u || lane <= first) ? (v & kSelVal) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
          excl += c;
          if (incl != 0u) break;
          t -= 32;
        }
        if (lane == 0) status[chunk] = kSelIncl | (excl + count);
      }
      if (lane == 0) {
        s_excl = (uint32_t)excl;
        if (chunk == nchunks - 1) *a.n_vis_out = (uint32_t)(excl + count);
      }
    }
    __syncthreads();
    const uint32_t excl = s_excl;
    // the compacted, still index-ordered (key, index) list the depth sort works on
    for (uint32_t j = tid; j < count; j += 256) {
      const uint32_t g = (uint32_t)base + list[j];
      a.depth_keys[excl + j] = a.depth_in[g];
      a.sorted_init[excl + j] = g;
    }
  }
  const uint32_t wsum = __reduce_add_sync(0xffffffffu, my_tiles);
  if (lane == 0 && wsum != 0) atomicAdd(&s_sum, wsum);
  __syncthreads();
  if (tid == 0 && s_sum != 0) atomicAdd(a.total_tiles, (unsigned long long)s_sum);
}

// Standalone partition (gcr_stripe_partition): the same row histogram + cut, nothing else written.
__global__ void __launch_bounds__(256)
stripe_partition_kernel(GcrPreprocessArgs a) {
  __shared__ uint32_t hist[kPartRowsSmem];
  __shared__ bool s_last;
  const bool use_smem = a.grid_y <= kPartRowsSmem;
  if (use_smem)
    for (int r = threadIdx.x; r < a.grid_y; r += blockDim.x) hist[r] = 0;
  __syncthreads();
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < a.P; idx += gridDim.x * blockDim.x) {
    const float px = a.means3D[3 * idx + 0];
    const float py = a.means3D[3 * idx + 1];
    const float pz = a.means3D[3 * idx + 2];
    GcrProjected g;
    if (gcr_project(a, idx, px, py, pz, g)) {
      const uint32_t w = g.rmax.x - g.rmin.x;
      for (uint32_t r = g.rmin.y; r < g.rmax.y; ++r) atomicAdd(use_smem ? &hist[r] : &a.row_hist[r], w);
    }
  }
  __syncthreads();
  if (use_smem)
    for (int r = threadIdx.x; r < a.grid_y; r += blockDim.x)
      if (hist[r] != 0) atomicAdd(&a.row_hist[r], hist[r]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&a.row_hist[a.grid_y], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  cut_stripes(a.row_hist, a.grid_y, a.shard_count, a.stripe_bounds_out);
}

// mark_visible (rasterizer_impl.cu:52-62): z_view > 0.2
__global__ void check_frustum_kernel(int P, const float* __restrict__ means3D,
                                     const float* __restrict__ V, bool* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
  const float vz = __fadd_rn(xf3(V, 2, px, py, pz), V[14]);
  present[idx] = vz > 0.2f;
}

}  // namespace

namespace {
inline unsigned persistent_grid(int n, int per_sm) {
  const int blocks = (n + 255) / 256;
  return (unsigned)(blocks < 148 * per_sm ? blocks : 148 * per_sm);
}
}  // namespace

cudaError_t gcr_launch_project(const GcrPreprocessArgs& a, bool deferred, cudaStream_t stream) {
  if (a.P <= 0) return cudaSuccess;
  const bool sh = a.colors_precomp == nullptr;
  const unsigned grid = deferred ? persistent_grid(a.P, 6) : (unsigned)((a.P + 255) / 256);
  if (deferred) {
    if (sh) project_kernel<true, true><<<grid, 256, 0, stream>>>(a);
    else project_kernel<false, true><<<grid, 256, 0, stream>>>(a);
  } else {
    if (sh) project_kernel<true, false><<<grid, 256, 0, stream>>>(a);
    else project_kernel<false, false><<<grid, 256, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t gcr_launch_stripe_select(const GcrPreprocessArgs& a, cudaStream_t stream) {
  if (a.P <= 0) return cudaSuccess;
  const int chunks = (a.P + kSelChunk - 1) / kSelChunk;
  const unsigned grid = (unsigned)(chunks < 148 * 12 ? chunks : 148 * 12);
  if (a.colors_precomp == nullptr) stripe_select_kernel<true><<<grid, 256, 0, stream>>>(a);
  else stripe_select_kernel<false><<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t gcr_launch_stripe_partition(const GcrPreprocessArgs& a, cudaStream_t stream) {
  if (a.P <= 0) return cudaSuccess;
  stripe_partition_kernel<<<persistent_grid(a.P, 8), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t gcr_launch_check_frustum(int P, const float* means3D, const float* viewmatrix, bool* present,
                                     cudaStream_t stream) {
  if (P <= 0) return cudaSuccess;
  check_frustum_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
  return cudaGetLastError();
}
