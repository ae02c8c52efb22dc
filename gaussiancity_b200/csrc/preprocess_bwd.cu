// preprocess_bwd.cu -- per-Gaussian backward for sm_100a: one fused kernel for what the
// reference runs as computeCov2DCUDA + preprocessCUDA (DGR/cuda_rasterizer/backward.cu:143-293,
// 378-425, with the SH backward :20-138 and the scale/rotation backward :297-373).
//
// Input: the 48-byte per-Gaussian accumulator filled by blend_bwd (dL/dmean2D, dL/dconic,
// dL/dopacity, dL/dcolor).  Output: every public gradient tensor, each element written exactly
// once (zeros for culled Gaussians), so the host does not have to memset (108+12M)*P bytes the
// way the reference's 9 torch::zeros do.  cov3D is recomputed from scale/rotation instead of
// being stored by the forward.  Quaternions are NOT normalised and no normalisation Jacobian
// is applied (backward.cu:301,369-372), as in the reference.
#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

// 3x3 matrices use GLM indexing m[col][row]; prod follows glm::operator*(mat3, mat3).
struct M3 {
  float m[3][3];
};
__forceinline__ __device__ M3 m3_mul(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R.m[i][j] = A.m[0][j] * B.m[i][0] + A.m[1][j] * B.m[i][1] + A.m[2][j] * B.m[i][2];
  return R;
}

// One Gaussian: `visible` = this rank owns it (on one GPU: it is visible) -> full backward;
// otherwise every output row is written as zeros.
template <bool kHasSH>
__forceinline__ __device__ void bwd_one(const GcrPreprocessBwdArgs& a, const int idx, const bool visible) {
  float o_m2x = 0.f, o_m2y = 0.f, o_op = 0.f, o_cr = 0.f, o_cg = 0.f, o_cb = 0.f;
  float o_ca = 0.f, o_cbb = 0.f, o_cc = 0.f;
  float dmean[3] = {0.f, 0.f, 0.f};
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dscale[3] = {0.f, 0.f, 0.f};
  float drot[4] = {0.f, 0.f, 0.f, 0.f};

  const size_t shbase = (size_t)idx * a.M * 3;

  if (visible) {
    const float4 g0 = a.grad_acc[idx].g0;
    const float4 g1 = a.grad_acc[idx].g1;
    const float g2x = a.grad_acc[idx].g2.x;
    o_m2x = g0.x; o_m2y = g0.y; o_ca = g0.z; o_cbb = g0.w; o_cc = g1.x; o_op = g1.y;
    o_cr = g1.z; o_cg = g1.w; o_cb = g2x;

    const float mx = a.means3D[3 * idx], my = a.means3D[3 * idx + 1], mz = a.means3D[3 * idx + 2];
    const float* __restrict__ V = a.viewmatrix;
    const float* __restrict__ PM = a.projmatrix;

    // ---- scale / rotation -> M = S R, cov3D (recomputed; forward.cu:110-144) ----
    float c[6];
    M3 Rm, Mm;
    float s[3] = {0.f, 0.f, 0.f};
    float qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
    const bool has_scale = a.scales != nullptr;
    if (has_scale) {
      s[0] = a.scale_modifier * a.scales[3 * idx + 0];
      s[1] = a.scale_modifier * a.scales[3 * idx + 1];
      s[2] = a.scale_modifier * a.scales[3 * idx + 2];
      const float4 q = a.rotations != nullptr ? reinterpret_cast<const float4*>(a.rotations)[idx]
                                              : make_float4(1.f, 0.f, 0.f, 0.f);
      qr = q.x; qx = q.y; qy = q.z; qz = q.w;
      Rm.m[0][0] = 1.f - 2.f * (qy * qy + qz * qz);
      Rm.m[0][1] = 2.f * (qx * qy - qr * qz);
      Rm.m[0][2] = 2.f * (qx * qz + qr * qy);
      Rm.m[1][0] = 2.f * (qx * qy + qr * qz);
      Rm.m[1][1] = 1.f - 2.f * (qx * qx + qz * qz);
      Rm.m[1][2] = 2.f * (qy * qz - qr * qx);
      Rm.m[2][0] = 2.f * (qx * qz - qr * qy);
      Rm.m[2][1] = 2.f * (qy * qz + qr * qx);
      Rm.m[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Mm.m[i][j] = s[j] * Rm.m[i][j];
    }
    if (a.cov3D_precomp != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; ++k) c[k] = a.cov3D_precomp[6 * (size_t)idx + k];
    } else {
      // Sigma[i][j] = sum_k M[j][k] M[i][k]
      c[0] = Mm.m[0][0] * Mm.m[0][0] + Mm.m[0][1] * Mm.m[0][1] + Mm.m[0][2] * Mm.m[0][2];
      c[1] = Mm.m[1][0] * Mm.m[0][0] + Mm.m[1][1] * Mm.m[0][1] + Mm.m[1][2] * Mm.m[0][2];
      c[2] = Mm.m[2][0] * Mm.m[0][0] + Mm.m[2][1] * Mm.m[0][1] + Mm.m[2][2] * Mm.m[0][2];
      c[3] = Mm.m[1][0] * Mm.m[1][0] + Mm.m[1][1] * Mm.m[1][1] + Mm.m[1][2] * Mm.m[1][2];
      c[4] = Mm.m[2][0] * Mm.m[1][0] + Mm.m[2][1] * Mm.m[1][1] + Mm.m[2][2] * Mm.m[1][2];
      c[5] = Mm.m[2][0] * Mm.m[2][0] + Mm.m[2][1] * Mm.m[2][1] + Mm.m[2][2] * Mm.m[2][2];
    }

    // ---- 2D covariance backward (backward.cu:143-293) ----
    float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
    float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
    const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
    const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

    M3 J, Wm, Vrk;
    J.m[0][0] = a.focal_x / tz; J.m[0][1] = 0.f; J.m[0][2] = -(a.focal_x * tx) / (tz * tz);
    J.m[1][0] = 0.f; J.m[1][1] = a.focal_y / tz; J.m[1][2] = -(a.focal_y * ty) / (tz * tz);
    J.m[2][0] = 0.f; J.m[2][1] = 0.f; J.m[2][2] = 0.f;
    Wm.m[0][0] = V[0]; Wm.m[0][1] = V[4]; Wm.m[0][2] = V[8];
    Wm.m[1][0] = V[1]; Wm.m[1][1] = V[5]; Wm.m[1][2] = V[9];
    Wm.m[2][0] = V[2]; Wm.m[2][1] = V[6]; Wm.m[2][2] = V[10];
    Vrk.m[0][0] = c[0]; Vrk.m[0][1] = c[1]; Vrk.m[0][2] = c[2];
    Vrk.m[1][0] = c[1]; Vrk.m[1][1] = c[3]; Vrk.m[1][2] = c[4];
    Vrk.m[2][0] = c[2]; Vrk.m[2][1] = c[4]; Vrk.m[2][2] = c[5];
    const M3 Tm = m3_mul(Wm, J);
    // cov2D = T^T Vrk^T T ; only the upper-left 2x2 is needed
    M3 Tt, X;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Tt.m[i][j] = Tm.m[j][i];
    X = m3_mul(m3_mul(Tt, Vrk), Tm);  // Vrk symmetric
    const float ca = X.m[0][0] + 0.3f;
    const float cb = X.m[0][1];
    const float cc = X.m[1][1] + 0.3f;
    const float denom = ca * cc - cb * cb;
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    const float dcx = o_ca, dcy = o_cbb, dcz = o_cc;  // dL/dconic (x, y, w)
    if (denom2inv != 0.f) {
      dL_da = denom2inv * (-cc * cc * dcx + 2.f * cb * cc * dcy + (denom - ca * cc) * dcz);
      dL_dc = denom2inv * (-ca * ca * dcz + 2.f * ca * cb * dcy + (denom - ca * cc) * dcx);
      dL_db = denom2inv * 2.f * (cb * cc * dcx - (denom + 2.f * cb * cb) * dcy + ca * cb * dcz);
      const float T00 = Tm.m[0][0], T01 = Tm.m[0][1], T02 = Tm.m[0][2];
      const float T10 = Tm.m[1][0], T11 = Tm.m[1][1], T12 = Tm.m[1][2];
      dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
      dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
      dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
      dcov[1] = 2.f * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2.f * T10 * T11 * dL_dc;
      dcov[2] = 2.f * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2.f * T10 * T12 * dL_dc;
      dcov[4] = 2.f * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2.f * T11 * T12 * dL_dc;
    }
    {
      const float T00 = Tm.m[0][0], T01 = Tm.m[0][1], T02 = Tm.m[0][2];
      const float T10 = Tm.m[1][0], T11 = Tm.m[1][1], T12 = Tm.m[1][2];
      // row-vectors u_k = T[0][:] . Vrk[k][:], w_k = T[1][:] . Vrk[k][:]
      float u[3], w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        u[k] = T00 * Vrk.m[k][0] + T01 * Vrk.m[k][1] + T02 * Vrk.m[k][2];
        w[k] = T10 * Vrk.m[k][0] + T11 * Vrk.m[k][1] + T12 * Vrk.m[k][2];
      }
      const float dT00 = 2.f * u[0] * dL_da + w[0] * dL_db;
      const float dT01 = 2.f * u[1] * dL_da + w[1] * dL_db;
      const float dT02 = 2.f * u[2] * dL_da + w[2] * dL_db;
      const float dT10 = 2.f * w[0] * dL_dc + u[0] * dL_db;
      const float dT11 = 2.f * w[1] * dL_dc + u[1] * dL_db;
      const float dT12 = 2.f * w[2] * dL_dc + u[2] * dL_db;
      const float dJ00 = Wm.m[0][0] * dT00 + Wm.m[0][1] * dT01 + Wm.m[0][2] * dT02;
      const float dJ02 = Wm.m[2][0] * dT00 + Wm.m[2][1] * dT01 + Wm.m[2][2] * dT02;
      const float dJ11 = Wm.m[1][0] * dT10 + Wm.m[1][1] * dT11 + Wm.m[1][2] * dT12;
      const float dJ12 = Wm.m[2][0] * dT10 + Wm.m[2][1] * dT11 + Wm.m[2][2] * dT12;
      const float itz = 1.f / tz;
      const float itz2 = itz * itz;
      const float itz3 = itz2 * itz;
      const float dtx = x_grad_mul * -a.focal_x * itz2 * dJ02;
      const float dty = y_grad_mul * -a.focal_y * itz2 * dJ12;
      const float dtz = -a.focal_x * itz2 * dJ00 - a.focal_y * itz2 * dJ11 +
                        (2.f * a.focal_x * tx) * itz3 * dJ02 + (2.f * a.focal_y * ty) * itz3 * dJ12;
      // dL/dmean (covariance path) = W^T-transposed transform (transformVec4x3Transpose)
      dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
      dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
      dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    }

    // ---- screen-space mean gradient -> 3D mean (backward.cu:389-413) ----
    {
      const float hw = PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15];
      const float m_w = 1.0f / (hw + 0.0000001f);
      const float mul1 = (PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12]) * m_w * m_w;
      const float mul2 = (PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13]) * m_w * m_w;
      dmean[0] += (PM[0] * m_w - PM[3] * mul1) * o_m2x + (PM[1] * m_w - PM[3] * mul2) * o_m2y;
      dmean[1] += (PM[4] * m_w - PM[7] * mul1) * o_m2x + (PM[5] * m_w - PM[7] * mul2) * o_m2y;
      dmean[2] += (PM[8] * m_w - PM[11] * mul1) * o_m2x + (PM[9] * m_w - PM[11] * mul2) * o_m2y;
    }

    // ---- SH backward (backward.cu:20-138) ----
    if (kHasSH) {
      const float dox = mx - a.campos[0], doy = my - a.campos[1], doz = mz - a.campos[2];
      const float len = sqrtf(dox * dox + doy * doy + doz * doz);
      const float x = dox / len, y = doy / len, z = doz / len;
      const uint8_t cl = a.clamped[idx];
      float dRGB[3] = {(cl & 1) ? 0.f : o_cr, (cl & 2) ? 0.f : o_cg, (cl & 4) ? 0.f : o_cb};
      // Basis b[k] and its partial derivatives (x, y, z treated as independent, exactly the
      // reference's dRGBd{x,y,z} terms); bands above the active degree are zeroed so their
      // coefficients receive zero gradient (the reference leaves torch::zeros there).
      float b[16], bx[16], by[16], bz[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) { b[k] = 0.f; bx[k] = 0.f; by[k] = 0.f; bz[k] = 0.f; }
      b[0] = GCR_SH_C0;
      if (a.D > 0) {
        b[1] = -GCR_SH_C1 * y; by[1] = -GCR_SH_C1;
        b[2] = GCR_SH_C1 * z;  bz[2] = GCR_SH_C1;
        b[3] = -GCR_SH_C1 * x; bx[3] = -GCR_SH_C1;
        if (a.D > 1) {
          const float xx = x * x, yy = y * y, zz = z * z;
          const float xy = x * y, yz = y * z, xz = x * z;
          b[4] = GCR_SH_C2[0] * xy; bx[4] = GCR_SH_C2[0] * y; by[4] = GCR_SH_C2[0] * x;
          b[5] = GCR_SH_C2[1] * yz; by[5] = GCR_SH_C2[1] * z; bz[5] = GCR_SH_C2[1] * y;
          b[6] = GCR_SH_C2[2] * (2.f * zz - xx - yy);
          bx[6] = GCR_SH_C2[2] * 2.f * -x; by[6] = GCR_SH_C2[2] * 2.f * -y; bz[6] = GCR_SH_C2[2] * 2.f * 2.f * z;
          b[7] = GCR_SH_C2[3] * xz; bx[7] = GCR_SH_C2[3] * z; bz[7] = GCR_SH_C2[3] * x;
          b[8] = GCR_SH_C2[4] * (xx - yy); bx[8] = GCR_SH_C2[4] * 2.f * x; by[8] = GCR_SH_C2[4] * 2.f * -y;
          if (a.D > 2) {
            b[9] = GCR_SH_C3[0] * y * (3.f * xx - yy);
            bx[9] = GCR_SH_C3[0] * 3.f * 2.f * xy; by[9] = GCR_SH_C3[0] * 3.f * (xx - yy);
            b[10] = GCR_SH_C3[1] * xy * z;
            bx[10] = GCR_SH_C3[1] * yz; by[10] = GCR_SH_C3[1] * xz; bz[10] = GCR_SH_C3[1] * xy;
            b[11] = GCR_SH_C3[2] * y * (4.f * zz - xx - yy);
            bx[11] = GCR_SH_C3[2] * -2.f * xy; by[11] = GCR_SH_C3[2] * (-3.f * yy + 4.f * zz - xx);
            bz[11] = GCR_SH_C3[2] * 4.f * 2.f * yz;
            b[12] = GCR_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
            bx[12] = GCR_SH_C3[3] * -3.f * 2.f * xz; by[12] = GCR_SH_C3[3] * -3.f * 2.f * yz;
            bz[12] = GCR_SH_C3[3] * 3.f * (2.f * zz - xx - yy);
            b[13] = GCR_SH_C3[4] * x * (4.f * zz - xx - yy);
            bx[13] = GCR_SH_C3[4] * (-3.f * xx + 4.f * zz - yy); by[13] = GCR_SH_C3[4] * -2.f * xy;
            bz[13] = GCR_SH_C3[4] * 4.f * 2.f * xz;
            b[14] = GCR_SH_C3[5] * z * (xx - yy);
            bx[14] = GCR_SH_C3[5] * 2.f * xz; by[14] = GCR_SH_C3[5] * -2.f * yz; bz[14] = GCR_SH_C3[5] * (xx - yy);
            b[15] = GCR_SH_C3[6] * x * (xx - 3.f * yy);
            bx[15] = GCR_SH_C3[6] * 3.f * (xx - yy); by[15] = GCR_SH_C3[6] * -3.f * 2.f * xy;
          }
        }
      }
      // One pass over the 3*M coefficients: dL/dsh[f] = b[f/3] dRGB[f%3] and
      // dL/ddir += (db/ddir)[f/3] * sh[f] * dRGB[f%3].  16-byte loads/stores when 3*M % 4 == 0
      // (M = 4, 16): a quarter of the L2 transactions of the per-float access pattern.
      float ddx = 0.f, ddy = 0.f, ddz = 0.f;
      const int nfl = a.M * 3;
      if ((nfl & 7) == 0) {
        // 32-byte accesses (M = 16): every thread reads/writes whole sectors
        const float* __restrict__ shp = a.shs + shbase;
        float* __restrict__ dshp = a.dL_dsh + shbase;
        // one sector at a time (load -> 8 outputs -> store): measured faster than issuing all six
        // loads first, which costs 176 registers and drops the kernel to one CTA per SM
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          if (8 * q < nfl) {
            float in[8], out[8];
            gcr_ldg_nc_v8(shp + 8 * q, in);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int f = 8 * q + i, k = f / 3, ch = f % 3;
              out[i] = b[k] * dRGB[ch];
              const float t = in[i] * dRGB[ch];
              ddx += bx[k] * t; ddy += by[k] * t; ddz += bz[k] * t;
            }
            gcr_stg_v8(dshp + 8 * q, out);
          }
        }
      } else if ((nfl & 3) == 0) {
        const float4* __restrict__ sh4 = reinterpret_cast<const float4*>(a.shs + shbase);
        float4* __restrict__ dsh4 = reinterpret_cast<float4*>(a.dL_dsh + shbase);
#pragma unroll
        for (int q = 0; q < 12; ++q) {
          if (4 * q < nfl) {
            const float4 v = __ldg(sh4 + q);
            const float in[4] = {v.x, v.y, v.z, v.w};
            float out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int f = 4 * q + i, k = f / 3, ch = f % 3;
              out[i] = b[k] * dRGB[ch];
              const float t = in[i] * dRGB[ch];
              ddx += bx[k] * t; ddy += by[k] * t; ddz += bz[k] * t;
            }
            dsh4[q] = make_float4(out[0], out[1], out[2], out[3]);
          }
        }
      } else {
        const float* __restrict__ sh = a.shs + shbase;
        float* __restrict__ dsh = a.dL_dsh + shbase;
#pragma unroll
        for (int f = 0; f < 48; ++f) {
          if (f < nfl) {
            const int k = f / 3, ch = f % 3;
            dsh[f] = b[k] * dRGB[ch];
            const float t = __ldg(sh + f) * dRGB[ch];
            ddx += bx[k] * t; ddy += by[k] * t; ddz += bz[k] * t;
          }
        }
      }
      // normalisation Jacobian (dnormvdv, auxiliary.h:95-112)
      const float sum2 = dox * dox + doy * doy + doz * doz;
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean[0] += ((sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
      dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
      dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
    }

    // ---- scale / rotation backward (backward.cu:297-373) ----
    if (has_scale) {
      M3 dS;
      dS.m[0][0] = dcov[0]; dS.m[0][1] = 0.5f * dcov[1]; dS.m[0][2] = 0.5f * dcov[2];
      dS.m[1][0] = 0.5f * dcov[1]; dS.m[1][1] = dcov[3]; dS.m[1][2] = 0.5f * dcov[4];
      dS.m[2][0] = 0.5f * dcov[2]; dS.m[2][1] = 0.5f * dcov[4]; dS.m[2][2] = dcov[5];
      M3 dM = m3_mul(Mm, dS);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dM.m[i][j] *= 2.0f;
      // dL_dMt[a][b] = dM[b][a];  Rt[a][b] = R[b][a]
      dscale[0] = Rm.m[0][0] * dM.m[0][0] + Rm.m[1][0] * dM.m[1][0] + Rm.m[2][0] * dM.m[2][0];
      dscale[1] = Rm.m[0][1] * dM.m[0][1] + Rm.m[1][1] * dM.m[1][1] + Rm.m[2][1] * dM.m[2][1];
      dscale[2] = Rm.m[0][2] * dM.m[0][2] + Rm.m[1][2] * dM.m[1][2] + Rm.m[2][2] * dM.m[2][2];
      float t[3][3];  // t[a][b] = dL_dMt[a][b] * s_a
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) t[p][q] = dM.m[q][p] * s[p];
      drot[0] = 2.f * qz * (t[0][1] - t[1][0]) + 2.f * qy * (t[2][0] - t[0][2]) +
                2.f * qx * (t[1][2] - t[2][1]);
      drot[1] = 2.f * qy * (t[1][0] + t[0][1]) + 2.f * qz * (t[2][0] + t[0][2]) +
                2.f * qr * (t[1][2] - t[2][1]) - 4.f * qx * (t[2][2] + t[1][1]);
      drot[2] = 2.f * qx * (t[1][0] + t[0][1]) + 2.f * qr * (t[2][0] - t[0][2]) +
                2.f * qz * (t[1][2] + t[2][1]) - 4.f * qy * (t[2][2] + t[0][0]);
      drot[3] = 2.f * qr * (t[0][1] - t[1][0]) + 2.f * qx * (t[2][0] + t[0][2]) +
                2.f * qy * (t[1][2] + t[2][1]) - 4.f * qz * (t[1][1] + t[0][0]);
    }
  } else if (kHasSH) {
    const int nfl = a.M * 3;
    if ((nfl & 7) == 0) {
      const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < nfl / 8; ++q) gcr_stg_v8(a.dL_dsh + shbase + 8 * q, z8);
    } else if ((nfl & 3) == 0) {
      float4* __restrict__ dsh4 = reinterpret_cast<float4*>(a.dL_dsh + shbase);
      for (int q = 0; q < nfl / 4; ++q) dsh4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      float* __restrict__ dsh = a.dL_dsh + shbase;
      for (int k = 0; k < nfl; ++k) dsh[k] = 0.f;
    }
  }

  if (a.packed_out != nullptr) {
    // striped frames: a rank's Gaussians are scattered over the index range, so seven small rows
    // per Gaussian would be seven partial-sector writes (each a read-modify-write in DRAM); one
    // 96-byte row = three whole sectors.  Layout (floats): mean3D 0:3 | opacity 3 | scale 4:7 |
    // mean2D 7:10 | colour 10:13 | cov3D 13:19 | pad 19 | rotation 20:24 (ext.py hands out views).
    float4* o = reinterpret_cast<float4*>(a.packed_out + 24 * (size_t)idx);
    o[0] = make_float4(dmean[0], dmean[1], dmean[2], o_op);
    o[1] = make_float4(dscale[0], dscale[1], dscale[2], o_m2x);
    o[2] = make_float4(o_m2y, 0.f, o_cr, o_cg);
    o[3] = make_float4(o_cb, dcov[0], dcov[1], dcov[2]);
    o[4] = make_float4(dcov[3], dcov[4], dcov[5], 0.f);
    o[5] = make_float4(drot[0], drot[1], drot[2], drot[3]);
    if (visible && a.clear_acc) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      a.grad_acc[idx].g0 = z; a.grad_acc[idx].g1 = z; a.grad_acc[idx].g2 = z;
    }
    return;
  }
  a.dL_dmean2D[3 * idx + 0] = o_m2x;
  a.dL_dmean2D[3 * idx + 1] = o_m2y;
  a.dL_dmean2D[3 * idx + 2] = 0.f;
  if (a.dL_dconic != nullptr)
    reinterpret_cast<float4*>(a.dL_dconic)[idx] = make_float4(o_ca, o_cbb, 0.f, o_cc);
  if (a.dL_dopacity != nullptr) a.dL_dopacity[idx] = o_op;
  a.dL_dcolor[3 * idx + 0] = o_cr;
  a.dL_dcolor[3 * idx + 1] = o_cg;
  a.dL_dcolor[3 * idx + 2] = o_cb;
#pragma unroll
  for (int k = 0; k < 3; ++k) a.dL_dmean3D[3 * idx + k] = dmean[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
  if (a.dL_dscale != nullptr) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a.dL_dscale[3 * idx + k] = dscale[k];
  }
  if (a.dL_drot != nullptr)
    reinterpret_cast<float4*>(a.dL_drot)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
  // persistent (peer-mapped) accumulator: leave the entry zeroed for the frame after next.  Last
  // statement on purpose: a store into grad_acc earlier would order itself against every later
  // load of this thread (measured: +0.6 ms at 5 M Gaussians when it followed the loads directly).
  if (visible && a.clear_acc) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    a.grad_acc[idx].g0 = z; a.grad_acc[idx].g1 = z; a.grad_acc[idx].g2 = z;
  }
}

// The Gaussians a CTA differentiates are compacted first (ballot + warp prefix, index order
// kept), 2048 at a time: under tile-row sharding a rank owns ~1/N of them, scattered over the index
// range.  Without compaction every warp would run the whole backward with 1/N of its lanes, and
// with one Gaussian per thread most warps of a 120-register CTA would hold their registers for
// nothing (measured at N = 8: 0.39 ms for 1/8 of the work of 0.68 ms).  The others get their zeros
// (reference semantics on one GPU) or are left alone (striped: their owner writes them).
// kItems Gaussians per thread and chunk: 8 (2048-Gaussian chunks) for large inputs, 1 for small
// ones, where 2048-wide chunks would leave most SMs without a CTA.  Chunks are handed out by an
// atomic ticket (zeroed by the host): with ~25 us of work per chunk a static split would leave a
// 10 % tail.
template <bool kHasSH, int kItems>
__global__ void __launch_bounds__(256)
preprocess_bwd_kernel(GcrPreprocessBwdArgs a) {
  constexpr int kChunk = 256 * kItems;
  // owned Gaussians (global indices) waiting to be differentiated: whatever does not fill a whole
  // 256-thread batch is carried over to the next chunk, so every batch but a CTA's last runs with
  // full warps (at N = 8 a 2048-Gaussian chunk holds ~270 owned ones: 256 + a 14-thread straggler
  // batch otherwise)
  __shared__ uint32_t queue[kChunk + 256];
  __shared__ uint32_t wcount[kItems][8];
  __shared__ int s_chunk;
  gcr_pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (a.range_count + kChunk - 1) / kChunk;
  uint32_t pending = 0;   // block-uniform: entries in the queue
  while (true) {
    if (tid == 0) s_chunk = (int)atomicAdd(a.ticket, 1u);
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk >= nchunks) break;
    const int loc0 = chunk * kChunk;
    const int base = a.range_start + loc0;
    bool owned[kItems];
    unsigned bal[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int t = k * 256 + tid;
      const bool in_range = loc0 + t < a.range_count;
      owned[k] = in_range && a.owner[base + t] == (uint8_t)a.my_rank;
      bal[k] = __ballot_sync(0xffffffffu, owned[k]);
      if (lane == 0) wcount[k][warp] = __popc(bal[k]);
    }
    if (a.zero_unowned) {
#pragma unroll 1
      for (int k = 0; k < kItems; ++k) {
        const int t = k * 256 + tid;
        if (loc0 + t < a.range_count && !owned[k]) bwd_one<kHasSH>(a, base + t, false);
      }
    }
    __syncthreads();
    uint32_t count = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      uint32_t before = count;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const uint32_t c = wcount[k][w];
        if (w < warp) before += c;
        count += c;
      }
      if (owned[k]) queue[pending + before + __popc(bal[k] & ((1u << lane) - 1))] = (uint32_t)(base + k * 256 + tid);
    }
    pending += count;
    __syncthreads();
    uint32_t head = 0;
    while (pending - head >= 256u) {
      bwd_one<kHasSH>(a, (int)queue[head + tid], true);
      head += 256;
    }
    const uint32_t rest = pending - head;   // < 256
    uint32_t carry = 0;
    if ((uint32_t)tid < rest) carry = queue[head + tid];
    __syncthreads();
    if ((uint32_t)tid < rest) queue[tid] = carry;
    pending = rest;
    // the next chunk's first barrier (ticket) orders these writes before anything reads the queue
  }
  if ((uint32_t)tid < pending) bwd_one<kHasSH>(a, (int)queue[tid], true);
}

}  // namespace

cudaError_t gcr_launch_preprocess_bwd(const GcrPreprocessBwdArgs& a, cudaStream_t stream) {
  if (a.P <= 0 || a.range_count <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  const bool sh = a.shs != nullptr && a.M > 0;
  // 2048-Gaussian chunks pay off when a rank owns a fraction of the range (striped frames) and the
  // range is long enough to give every resident CTA several; a dense single-GPU range is best served
  // one Gaussian per thread (measured: 0.68 vs 0.74 ms at 5 M)
  const bool big = !a.zero_unowned && a.range_count >= 148 * 2 * 2048;
  const int chunk = big ? 2048 : 256;
  const int chunks = (a.range_count + chunk - 1) / chunk;
  const int blocks = chunks < 148 * 4 ? chunks : 148 * 4;
  void (*kernel)(GcrPreprocessBwdArgs) =
      big ? (sh ? preprocess_bwd_kernel<true, 8> : preprocess_bwd_kernel<false, 8>)
          : (sh ? preprocess_bwd_kernel<true, 1> : preprocess_bwd_kernel<false, 1>);
  return gcr_launch_chain<GCR_EDGE_GEOM_BWD>(kernel, dim3(blocks), dim3(256), 0, stream, a);
}
