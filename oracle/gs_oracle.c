/* gs_oracle.c -- CPU restatement of the reference Gaussian rasterizer.  TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle for the B200 kernels in gaussiancity_b200/csrc.  It may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; the product package never touches it.
 *
 * It restates, in plain C, the algorithm of hzxie/GaussianCity's
 * extensions/diff_gaussian_rasterization ("DGR"):
 *   preprocess      DGR/cuda_rasterizer/forward.cu:147-233  (+ auxiliary.h:32-46,132-156,
 *                   computeCov3D forward.cu:110-144, computeCov2D :69-105, SH :20-66)
 *   binning         DGR/cuda_rasterizer/rasterizer_impl.cu:66-124 (duplicateWithKeys, stable
 *                   sort of tile|depth keys, identifyTileRanges)
 *   blend           forward.cu:238-346
 *   blend backward  backward.cu:428-581
 *   cov2D backward  backward.cu:143-293
 *   geometry bwd    backward.cu:378-425 (+ SH :20-138, scale/rotation :297-373)
 * One source, two builds: -DGSO_DOUBLE=0 is an fp32 restatement (compiled with
 * -ffp-contract=off: every operation rounds to float, no FMA), -DGSO_DOUBLE=1 evaluates the same
 * formulas in fp64 and is the arbiter for gradient noise.  Inputs are always fp32 arrays.
 *
 * Pinning: the reference ships no tests or golden vectors.  The pin is the
 * reference extension itself: tests/golden/ holds outputs of the UNMODIFIED reference
 * (oracle/_ref, built by oracle/build_ref.py) generated on a B200 by tools/make_golden.py;
 * tests/test_oracle_golden.py checks this restatement against them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if GSO_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FMIN fmin
#define R_FMAX fmax
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FMIN fminf
#define R_FMAX fmaxf
#endif

#define TILE 16

static const real C0 = (real)0.28209479177387814f;
static const real C1 = (real)0.4886025119029199f;
static const real C2[5] = {(real)1.0925484305920792f, (real)-1.0925484305920792f,
                           (real)0.31539156525252005f, (real)-1.0925484305920792f,
                           (real)0.5462742152960396f};
static const real C3[7] = {(real)-0.5900435899266435f, (real)2.890611442640554f,
                           (real)-0.4570457994644658f, (real)0.3731763325901154f,
                           (real)-0.4570457994644658f, (real)1.445305721320277f,
                           (real)-0.5900435899266435f};

int gso_is_double(void) { return GSO_DOUBLE; }

/* Test-only "smooth" mode: removes the reference's discontinuities (alpha < 1/255 skip,
 * T < 1e-4 stop, 3-sigma bounding square) so finite differences can validate the analytic
 * backward formulas. Never used for parity. */
static real g_alpha_min = (real)(1.0f / 255.0f), g_T_min = (real)0.0001f, g_radius_k = (real)3.0f;
void gso_set_smooth(int on) {
  g_alpha_min = on ? (real)0 : (real)(1.0f / 255.0f);
  g_T_min = on ? (real)0 : (real)0.0001f;
  g_radius_k = on ? (real)12.0f : (real)3.0f;
}

/* column-major 3x3 helpers: m[c][r], product as glm::operator*(mat3,mat3) */
typedef struct { real m[3][3]; } M3;
static M3 m3mul(const M3* A, const M3* B) {
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R.m[i][j] = A->m[0][j] * B->m[i][0] + A->m[1][j] * B->m[i][1] + A->m[2][j] * B->m[i][2];
  return R;
}
static M3 m3t(const M3* A) {
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = A->m[j][i];
  return R;
}

static void quat_to_R(const float* q, M3* R) {
  real r = q[0], x = q[1], y = q[2], z = q[3]; /* NOT normalised (forward.cu:119) */
  R->m[0][0] = 1 - 2 * (y * y + z * z); R->m[0][1] = 2 * (x * y - r * z); R->m[0][2] = 2 * (x * z + r * y);
  R->m[1][0] = 2 * (x * y + r * z); R->m[1][1] = 1 - 2 * (x * x + z * z); R->m[1][2] = 2 * (y * z - r * x);
  R->m[2][0] = 2 * (x * z - r * y); R->m[2][1] = 2 * (y * z + r * x); R->m[2][2] = 1 - 2 * (x * x + y * y);
}

static void cov3d_from_scale_rot(const float* scale, real mod, const float* q, real* c, M3* Rout, M3* Mout, real* s) {
  M3 R, S, M;
  quat_to_R(q, &R);
  memset(&S, 0, sizeof(S));
  s[0] = mod * scale[0]; s[1] = mod * scale[1]; s[2] = mod * scale[2];
  S.m[0][0] = s[0]; S.m[1][1] = s[1]; S.m[2][2] = s[2];
  M = m3mul(&S, &R);
  M3 Mt = m3t(&M);
  M3 Sg = m3mul(&Mt, &M);
  c[0] = Sg.m[0][0]; c[1] = Sg.m[0][1]; c[2] = Sg.m[0][2];
  c[3] = Sg.m[1][1]; c[4] = Sg.m[1][2]; c[5] = Sg.m[2][2];
  if (Rout) *Rout = R;
  if (Mout) *Mout = M;
}

typedef struct {
  real tx, ty, tz, txtz, tytz, limx, limy;
  M3 J, W, T, Vrk;
  real a, b, c; /* cov2D entries incl. +0.3 */
} Cov2D;

static void cov2d_eval(const real* mean, real fx, real fy, real tanx, real tany, const real* c3,
                       const float* V, Cov2D* o) {
  real tx = V[0] * mean[0] + V[4] * mean[1] + V[8] * mean[2] + V[12];
  real ty = V[1] * mean[0] + V[5] * mean[1] + V[9] * mean[2] + V[13];
  real tz = V[2] * mean[0] + V[6] * mean[1] + V[10] * mean[2] + V[14];
  o->limx = (real)1.3f * tanx; o->limy = (real)1.3f * tany;
  o->txtz = tx / tz; o->tytz = ty / tz;
  tx = R_FMIN(o->limx, R_FMAX(-o->limx, o->txtz)) * tz;
  ty = R_FMIN(o->limy, R_FMAX(-o->limy, o->tytz)) * tz;
  o->tx = tx; o->ty = ty; o->tz = tz;
  memset(&o->J, 0, sizeof(M3));
  o->J.m[0][0] = fx / tz; o->J.m[0][2] = -(fx * tx) / (tz * tz);
  o->J.m[1][1] = fy / tz; o->J.m[1][2] = -(fy * ty) / (tz * tz);
  o->W.m[0][0] = V[0]; o->W.m[0][1] = V[4]; o->W.m[0][2] = V[8];
  o->W.m[1][0] = V[1]; o->W.m[1][1] = V[5]; o->W.m[1][2] = V[9];
  o->W.m[2][0] = V[2]; o->W.m[2][1] = V[6]; o->W.m[2][2] = V[10];
  o->T = m3mul(&o->W, &o->J);
  o->Vrk.m[0][0] = c3[0]; o->Vrk.m[0][1] = c3[1]; o->Vrk.m[0][2] = c3[2];
  o->Vrk.m[1][0] = c3[1]; o->Vrk.m[1][1] = c3[3]; o->Vrk.m[1][2] = c3[4];
  o->Vrk.m[2][0] = c3[2]; o->Vrk.m[2][1] = c3[4]; o->Vrk.m[2][2] = c3[5];
  M3 Tt = m3t(&o->T), Vt = m3t(&o->Vrk);
  M3 X = m3mul(&Tt, &Vt);
  M3 cov = m3mul(&X, &o->T);
  o->a = cov.m[0][0] + (real)0.3f;
  o->b = cov.m[0][1];
  o->c = cov.m[1][1] + (real)0.3f;
}

static real ndc2pix(real v, int S) { return (real)((((double)v + 1.0) * S - 1.0) * 0.5); }

static void get_rect(real px, real py, int rad, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
  /* float arithmetic + truncation (auxiliary.h:36-46); the fp64 build keeps float here so the
   * integer tile work is the same in both builds given the same inputs */
  float fpx = (float)px, fpy = (float)py;
  int a;
  a = (int)((fpx - rad) / TILE); *x0 = a < 0 ? 0 : (a > gx ? gx : a);
  a = (int)((fpy - rad) / TILE); *y0 = a < 0 ? 0 : (a > gy ? gy : a);
  a = (int)((fpx + rad + TILE - 1) / TILE); *x1 = a < 0 ? 0 : (a > gx ? gx : a);
  a = (int)((fpy + rad + TILE - 1) / TILE); *y1 = a < 0 ? 0 : (a > gy ? gy : a);
}

/* ---- stage 1: per-Gaussian preprocessing.  Returns num_rendered (sum of tiles_touched). ---- */
long gso_preprocess(int P, int D, int M, int W, int H, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* opacities, const float* scales,
                    float scale_modifier, const float* rotations, const float* cov3D_precomp,
                    const float* V, const float* PM, const float* campos, float tan_fovx,
                    float tan_fovy,
                    /* out, caller-allocated */
                    int* radii, real* depths, real* means2D, real* cov3D, real* conic_opacity,
                    real* rgb, unsigned char* clamped, uint32_t* tiles_touched) {
  const real fy = H / ((real)2.0f * tan_fovy), fx = W / ((real)2.0f * tan_fovx);
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  long total = 0;
  for (int i = 0; i < P; ++i) {
    radii[i] = 0; tiles_touched[i] = 0;
    const real p[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    const real vz = V[2] * p[0] + V[6] * p[1] + V[10] * p[2] + V[14];
    if (!(vz > (real)0.2f)) continue;
    const real hx = PM[0] * p[0] + PM[4] * p[1] + PM[8] * p[2] + PM[12];
    const real hy = PM[1] * p[0] + PM[5] * p[1] + PM[9] * p[2] + PM[13];
    const real hw = PM[3] * p[0] + PM[7] * p[1] + PM[11] * p[2] + PM[15];
    const real pw = (real)1.0f / (hw + (real)0.0000001f);
    const real ndcx = hx * pw, ndcy = hy * pw;
    real c3[6];
    if (cov3D_precomp) {
      for (int k = 0; k < 6; ++k) c3[k] = cov3D_precomp[6 * (size_t)i + k];
    } else {
      real s[3];
      cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, c3, 0, 0, s);
      for (int k = 0; k < 6; ++k) cov3D[6 * (size_t)i + k] = c3[k];
    }
    Cov2D cv;
    cov2d_eval(p, fx, fy, tan_fovx, tan_fovy, c3, V, &cv);
    const real det = cv.a * cv.c - cv.b * cv.b;
    if (det == 0) continue;
    const real di = (real)1.0f / det;
    const real cx = cv.c * di, cy = -cv.b * di, cz = cv.a * di;
    const real mid = (real)0.5f * (cv.a + cv.c);
    const real sq = R_SQRT(R_FMAX((real)0.1f, mid * mid - det));
    const real l1 = mid + sq, l2 = mid - sq;
    const real rad = R_CEIL(g_radius_k * R_SQRT(R_FMAX(l1, l2)));
    const real pxi = ndc2pix(ndcx, W), pyi = ndc2pix(ndcy, H);
    int x0, y0, x1, y1;
    get_rect(pxi, pyi, (int)rad, gx, gy, &x0, &y0, &x1, &y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (!colors_precomp) {
      real dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
      real len = R_SQRT(dx * dx + dy * dy + dz * dz);
      real x = dx / len, y = dy / len, z = dz / len;
      const float* sh = shs + (size_t)i * M * 3;
      for (int ch = 0; ch < 3; ++ch) {
#define SH(k) ((real)sh[3 * (k) + ch])
        real v = C0 * SH(0);
        if (D > 0) {
          v = v - C1 * y * SH(1) + C1 * z * SH(2) - C1 * x * SH(3);
          if (D > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            v = v + C2[0] * xy * SH(4) + C2[1] * yz * SH(5) + C2[2] * (2 * zz - xx - yy) * SH(6) +
                C2[3] * xz * SH(7) + C2[4] * (xx - yy) * SH(8);
            if (D > 2) {
              v = v + C3[0] * y * (3 * xx - yy) * SH(9) + C3[1] * xy * z * SH(10) +
                  C3[2] * y * (4 * zz - xx - yy) * SH(11) +
                  C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * SH(12) +
                  C3[4] * x * (4 * zz - xx - yy) * SH(13) + C3[5] * z * (xx - yy) * SH(14) +
                  C3[6] * x * (xx - 3 * yy) * SH(15);
            }
          }
        }
#undef SH
        v += (real)0.5f;
        clamped[3 * i + ch] = v < 0;
        rgb[3 * i + ch] = v < 0 ? 0 : v;
      }
    }
    depths[i] = vz;
    radii[i] = (int)rad;
    means2D[2 * i] = pxi; means2D[2 * i + 1] = pyi;
    conic_opacity[4 * i] = cx; conic_opacity[4 * i + 1] = cy; conic_opacity[4 * i + 2] = cz;
    conic_opacity[4 * i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    total += tiles_touched[i];
  }
  return total;
}

/* ---- stage 2: duplicate with keys, stable sort, tile ranges ------------------------------- */
typedef struct { uint64_t key; uint32_t pos; uint32_t val; } Pair;
static int pair_cmp(const void* a, const void* b) {
  const Pair* x = (const Pair*)a; const Pair* y = (const Pair*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

int gso_bin(int P, int W, int H, const int* radii, const real* depths, const real* means2D,
            const uint32_t* tiles_touched, long R, uint64_t* keys_sorted, uint32_t* point_list,
            uint32_t* ranges /* [2*tiles], zero-filled here */) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  Pair* pr = (Pair*)malloc(sizeof(Pair) * (size_t)(R > 0 ? R : 1));
  if (!pr) return -1;
  size_t off = 0;
  for (int i = 0; i < P; ++i) {
    if (radii[i] <= 0) continue;
    int x0, y0, x1, y1;
    get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
    float d = (float)depths[i];
    uint32_t dbits; memcpy(&dbits, &d, 4);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        pr[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
        pr[off].pos = (uint32_t)off; pr[off].val = (uint32_t)i; ++off;
      }
    (void)tiles_touched;
  }
  if ((long)off != R) { free(pr); return -2; }
  qsort(pr, (size_t)R, sizeof(Pair), pair_cmp);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
  for (long i = 0; i < R; ++i) {
    keys_sorted[i] = pr[i].key; point_list[i] = pr[i].val;
    uint32_t cur = (uint32_t)(pr[i].key >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(pr[i - 1].key >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
  free(pr);
  return 0;
}

/* ---- stage 3: per-pixel front-to-back blend ------------------------------------------------- */
void gso_render(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                const real* means2D, const real* colors /* [P,3] */, const real* conic_opacity,
                const float* bg, real* final_T, uint32_t* n_contrib, real* out_color) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < gx * gy; ++t) {
    const int tx = t % gx, ty = t / gx;
    const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
    for (int py = ty * TILE; py < (ty + 1) * TILE && py < H; ++py)
      for (int px = tx * TILE; px < (tx + 1) * TILE && px < W; ++px) {
        real T = 1, C[3] = {0, 0, 0};
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = r0; k < r1; ++k) {
          ++contributor;
          const uint32_t g = point_list[k];
          const real dx = means2D[2 * g] - (real)px, dy = means2D[2 * g + 1] - (real)py;
          const real* co = conic_opacity + 4 * (size_t)g;
          const real power = (real)-0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0) continue;
          const real alpha = R_FMIN((real)0.99f, co[3] * R_EXP(power));
          if (alpha < g_alpha_min) continue;
          const real test_T = T * (1 - alpha);
          if (test_T < g_T_min) break;
          for (int ch = 0; ch < 3; ++ch) C[ch] += colors[3 * (size_t)g + ch] * alpha * T;
          T = test_T; last = contributor;
        }
        const size_t pid = (size_t)W * py + px;
        final_T[pid] = T; n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
      }
  }
}

/* ---- stage 4: backward blend (backward.cu:428-581), sequential accumulation per tile -------- */
void gso_render_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                         const float* bg, const real* means2D, const real* conic_opacity,
                         const real* colors, const real* final_T, const uint32_t* n_contrib,
                         const real* dL_dpix /* [3,H,W] */,
                         real* dL_dmean2D /* [P,2] */, real* dL_dconic /* [P,3] (x,y,w) */,
                         real* dL_dopacity /* [P] */, real* dL_dcolor /* [P,3] */) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const real ddelx_dx = (real)0.5f * W, ddely_dy = (real)0.5f * H;
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < gx * gy; ++t) {
    const int tx = t % gx, ty = t / gx;
    const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
    for (int py = ty * TILE; py < (ty + 1) * TILE && py < H; ++py)
      for (int px = tx * TILE; px < (tx + 1) * TILE && px < W; ++px) {
        const size_t pid = (size_t)W * py + px;
        const real T_final = final_T[pid];
        real T = T_final;
        const uint32_t last_contributor = n_contrib[pid];
        real accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0, dLp[3];
        for (int ch = 0; ch < 3; ++ch) dLp[ch] = dL_dpix[(size_t)ch * H * W + pid];
        real bg_dot = 0;
        for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dLp[ch];
        for (uint32_t k = r1; k-- > r0;) {
          const uint32_t contributor = k - r0; /* 0-based position in the tile list */
          if (contributor >= last_contributor) continue;
          const uint32_t g = point_list[k];
          const real dx = means2D[2 * g] - (real)px, dy = means2D[2 * g + 1] - (real)py;
          const real* co = conic_opacity + 4 * (size_t)g;
          const real power = (real)-0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0) continue;
          const real G = R_EXP(power);
          const real alpha = R_FMIN((real)0.99f, co[3] * G);
          if (alpha < g_alpha_min) continue;
          T = T / (1 - alpha);
          const real dchannel_dcolor = alpha * T;
          real dL_dalpha = 0;
          for (int ch = 0; ch < 3; ++ch) {
            const real c = colors[3 * (size_t)g + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + (1 - last_alpha) * accum_rec[ch];
            last_color[ch] = c;
            dL_dalpha += (c - accum_rec[ch]) * dLp[ch];
            const real v = dchannel_dcolor * dLp[ch];
#pragma omp atomic
            dL_dcolor[3 * (size_t)g + ch] += v;
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1 - alpha)) * bg_dot;
          const real dL_dG = co[3] * dL_dalpha;
          const real gdx = G * dx, gdy = G * dy;
          const real dG_ddelx = -gdx * co[0] - gdy * co[1];
          const real dG_ddely = -gdy * co[2] - gdx * co[1];
          const real v0 = dL_dG * dG_ddelx * ddelx_dx, v1 = dL_dG * dG_ddely * ddely_dy;
          const real v2 = (real)-0.5f * gdx * dx * dL_dG, v3 = (real)-0.5f * gdx * dy * dL_dG;
          const real v4 = (real)-0.5f * gdy * dy * dL_dG, v5 = G * dL_dalpha;
#pragma omp atomic
          dL_dmean2D[2 * (size_t)g] += v0;
#pragma omp atomic
          dL_dmean2D[2 * (size_t)g + 1] += v1;
#pragma omp atomic
          dL_dconic[3 * (size_t)g] += v2;
#pragma omp atomic
          dL_dconic[3 * (size_t)g + 1] += v3;
#pragma omp atomic
          dL_dconic[3 * (size_t)g + 2] += v4;
#pragma omp atomic
          dL_dopacity[g] += v5;
        }
      }
  }
}

/* ---- stage 5: per-Gaussian geometry backward (backward.cu:143-293, 378-425, 20-138, 297-373) */
void gso_geometry_backward(int P, int D, int M, int W, int H, const float* means3D, const int* radii,
                           const float* shs, const unsigned char* clamped, const float* scales,
                           const float* rotations, float scale_modifier, const float* cov3D_precomp,
                           const float* V, const float* PM, const float* campos, float tan_fovx,
                           float tan_fovy, const real* dL_dmean2D /* [P,2] */,
                           const real* dL_dconic /* [P,3] */, const real* dL_dcolor /* [P,3] */,
                           real* dL_dmean3D /* [P,3] */, real* dL_dcov3D /* [P,6] */,
                           real* dL_dsh /* [P,M,3] or NULL */, real* dL_dscale /* [P,3] or NULL */,
                           real* dL_drot /* [P,4] or NULL */) {
  const real fy = H / ((real)2.0f * tan_fovy), fx = W / ((real)2.0f * tan_fovx);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; ++i) {
    if (!(radii[i] > 0)) continue;
    const real p[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    real c3[6], s[3] = {0, 0, 0};
    M3 Rm, Mm;
    if (scales) cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, c3, &Rm, &Mm, s);
    if (cov3D_precomp)
      for (int k = 0; k < 6; ++k) c3[k] = cov3D_precomp[6 * (size_t)i + k];
    Cov2D cv;
    cov2d_eval(p, fx, fy, tan_fovx, tan_fovy, c3, V, &cv);
    const real xg = (cv.txtz < -cv.limx || cv.txtz > cv.limx) ? 0 : 1;
    const real yg = (cv.tytz < -cv.limy || cv.tytz > cv.limy) ? 0 : 1;
    const real a = cv.a, b = cv.b, c = cv.c;
    const real denom = a * c - b * b;
    real da = 0, db = 0, dc = 0;
    const real d2i = (real)1.0f / ((denom * denom) + (real)0.0000001f);
    const real gx_ = dL_dconic[3 * (size_t)i], gy_ = dL_dconic[3 * (size_t)i + 1], gz_ = dL_dconic[3 * (size_t)i + 2];
    real* dcov = dL_dcov3D + 6 * (size_t)i;
    const M3* T = &cv.T;
    if (d2i != 0) {
      da = d2i * (-c * c * gx_ + 2 * b * c * gy_ + (denom - a * c) * gz_);
      dc = d2i * (-a * a * gz_ + 2 * a * b * gy_ + (denom - a * c) * gx_);
      db = d2i * 2 * (b * c * gx_ - (denom + 2 * b * b) * gy_ + a * b * gz_);
      dcov[0] = T->m[0][0] * T->m[0][0] * da + T->m[0][0] * T->m[1][0] * db + T->m[1][0] * T->m[1][0] * dc;
      dcov[3] = T->m[0][1] * T->m[0][1] * da + T->m[0][1] * T->m[1][1] * db + T->m[1][1] * T->m[1][1] * dc;
      dcov[5] = T->m[0][2] * T->m[0][2] * da + T->m[0][2] * T->m[1][2] * db + T->m[1][2] * T->m[1][2] * dc;
      dcov[1] = 2 * T->m[0][0] * T->m[0][1] * da + (T->m[0][0] * T->m[1][1] + T->m[0][1] * T->m[1][0]) * db + 2 * T->m[1][0] * T->m[1][1] * dc;
      dcov[2] = 2 * T->m[0][0] * T->m[0][2] * da + (T->m[0][0] * T->m[1][2] + T->m[0][2] * T->m[1][0]) * db + 2 * T->m[1][0] * T->m[1][2] * dc;
      dcov[4] = 2 * T->m[0][2] * T->m[0][1] * da + (T->m[0][1] * T->m[1][2] + T->m[0][2] * T->m[1][1]) * db + 2 * T->m[1][1] * T->m[1][2] * dc;
    } else {
      for (int k = 0; k < 6; ++k) dcov[k] = 0;
    }
    real dT[2][3];
    for (int k = 0; k < 3; ++k) {
      real u = T->m[0][0] * cv.Vrk.m[k][0] + T->m[0][1] * cv.Vrk.m[k][1] + T->m[0][2] * cv.Vrk.m[k][2];
      real w = T->m[1][0] * cv.Vrk.m[k][0] + T->m[1][1] * cv.Vrk.m[k][1] + T->m[1][2] * cv.Vrk.m[k][2];
      dT[0][k] = 2 * u * da + w * db;
      dT[1][k] = 2 * w * dc + u * db;
    }
    const real dJ00 = cv.W.m[0][0] * dT[0][0] + cv.W.m[0][1] * dT[0][1] + cv.W.m[0][2] * dT[0][2];
    const real dJ02 = cv.W.m[2][0] * dT[0][0] + cv.W.m[2][1] * dT[0][1] + cv.W.m[2][2] * dT[0][2];
    const real dJ11 = cv.W.m[1][0] * dT[1][0] + cv.W.m[1][1] * dT[1][1] + cv.W.m[1][2] * dT[1][2];
    const real dJ12 = cv.W.m[2][0] * dT[1][0] + cv.W.m[2][1] * dT[1][1] + cv.W.m[2][2] * dT[1][2];
    const real tz = (real)1.0f / cv.tz, tz2 = tz * tz, tz3 = tz2 * tz;
    const real dtx = xg * -fx * tz2 * dJ02;
    const real dty = yg * -fy * tz2 * dJ12;
    const real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * cv.tx) * tz3 * dJ02 + (2 * fy * cv.ty) * tz3 * dJ12;
    real dm[3];
    dm[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dm[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dm[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    /* projected-mean path (backward.cu:389-413) */
    {
      const real hw = PM[3] * p[0] + PM[7] * p[1] + PM[11] * p[2] + PM[15];
      const real mw = (real)1.0f / (hw + (real)0.0000001f);
      const real mul1 = (PM[0] * p[0] + PM[4] * p[1] + PM[8] * p[2] + PM[12]) * mw * mw;
      const real mul2 = (PM[1] * p[0] + PM[5] * p[1] + PM[9] * p[2] + PM[13]) * mw * mw;
      const real d2x = dL_dmean2D[2 * (size_t)i], d2y = dL_dmean2D[2 * (size_t)i + 1];
      dm[0] += (PM[0] * mw - PM[3] * mul1) * d2x + (PM[1] * mw - PM[3] * mul2) * d2y;
      dm[1] += (PM[4] * mw - PM[7] * mul1) * d2x + (PM[5] * mw - PM[7] * mul2) * d2y;
      dm[2] += (PM[8] * mw - PM[11] * mul1) * d2x + (PM[9] * mw - PM[11] * mul2) * d2y;
    }
    if (shs && dL_dsh) {
      const real ox = p[0] - campos[0], oy = p[1] - campos[1], oz = p[2] - campos[2];
      const real len = R_SQRT(ox * ox + oy * oy + oz * oz);
      const real x = ox / len, y = oy / len, z = oz / len;
      const float* sh = shs + (size_t)i * M * 3;
      real* dsh = dL_dsh + (size_t)i * M * 3;
      real dRGB[3], ddx = 0, ddy = 0, ddz = 0;
      for (int ch = 0; ch < 3; ++ch) dRGB[ch] = clamped[3 * i + ch] ? 0 : dL_dcolor[3 * (size_t)i + ch];
      real basis[16];
      int nb = 1;
      basis[0] = C0;
      real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      if (D > 0) { basis[1] = -C1 * y; basis[2] = C1 * z; basis[3] = -C1 * x; nb = 4; }
      if (D > 1) {
        basis[4] = C2[0] * xy; basis[5] = C2[1] * yz; basis[6] = C2[2] * (2 * zz - xx - yy);
        basis[7] = C2[3] * xz; basis[8] = C2[4] * (xx - yy); nb = 9;
      }
      if (D > 2) {
        basis[9] = C3[0] * y * (3 * xx - yy); basis[10] = C3[1] * xy * z;
        basis[11] = C3[2] * y * (4 * zz - xx - yy); basis[12] = C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
        basis[13] = C3[4] * x * (4 * zz - xx - yy); basis[14] = C3[5] * z * (xx - yy);
        basis[15] = C3[6] * x * (xx - 3 * yy); nb = 16;
      }
      for (int k = 0; k < nb; ++k)
        for (int ch = 0; ch < 3; ++ch) dsh[3 * k + ch] = basis[k] * dRGB[ch];
      for (int ch = 0; ch < 3; ++ch) {
#define SH(k) ((real)sh[3 * (k) + ch])
        real gxv = 0, gyv = 0, gzv = 0;
        if (D > 0) {
          gxv = -C1 * SH(3); gyv = -C1 * SH(1); gzv = C1 * SH(2);
          if (D > 1) {
            gxv += C2[0] * y * SH(4) + C2[2] * 2 * -x * SH(6) + C2[3] * z * SH(7) + C2[4] * 2 * x * SH(8);
            gyv += C2[0] * x * SH(4) + C2[1] * z * SH(5) + C2[2] * 2 * -y * SH(6) + C2[4] * 2 * -y * SH(8);
            gzv += C2[1] * y * SH(5) + C2[2] * 2 * 2 * z * SH(6) + C2[3] * x * SH(7);
            if (D > 2) {
              gxv += C3[0] * SH(9) * 3 * 2 * xy + C3[1] * SH(10) * yz + C3[2] * SH(11) * -2 * xy +
                     C3[3] * SH(12) * -3 * 2 * xz + C3[4] * SH(13) * (-3 * xx + 4 * zz - yy) +
                     C3[5] * SH(14) * 2 * xz + C3[6] * SH(15) * 3 * (xx - yy);
              gyv += C3[0] * SH(9) * 3 * (xx - yy) + C3[1] * SH(10) * xz +
                     C3[2] * SH(11) * (-3 * yy + 4 * zz - xx) + C3[3] * SH(12) * -3 * 2 * yz +
                     C3[4] * SH(13) * -2 * xy + C3[5] * SH(14) * -2 * yz + C3[6] * SH(15) * -3 * 2 * xy;
              gzv += C3[1] * SH(10) * xy + C3[2] * SH(11) * 4 * 2 * yz +
                     C3[3] * SH(12) * 3 * (2 * zz - xx - yy) + C3[4] * SH(13) * 4 * 2 * xz +
                     C3[5] * SH(14) * (xx - yy);
            }
          }
        }
#undef SH
        ddx += gxv * dRGB[ch]; ddy += gyv * dRGB[ch]; ddz += gzv * dRGB[ch];
      }
      const real sum2 = ox * ox + oy * oy + oz * oz;
      const real inv = (real)1.0f / R_SQRT(sum2 * sum2 * sum2);
      dm[0] += ((sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv;
      dm[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv;
      dm[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv;
    }
    for (int k = 0; k < 3; ++k) dL_dmean3D[3 * (size_t)i + k] = dm[k];
    if (scales && dL_dscale && dL_drot) {
      M3 dS, dM;
      dS.m[0][0] = dcov[0]; dS.m[0][1] = (real)0.5f * dcov[1]; dS.m[0][2] = (real)0.5f * dcov[2];
      dS.m[1][0] = (real)0.5f * dcov[1]; dS.m[1][1] = dcov[3]; dS.m[1][2] = (real)0.5f * dcov[4];
      dS.m[2][0] = (real)0.5f * dcov[2]; dS.m[2][1] = (real)0.5f * dcov[4]; dS.m[2][2] = dcov[5];
      M3 M2 = Mm;
      for (int u = 0; u < 3; ++u) for (int v = 0; v < 3; ++v) M2.m[u][v] *= 2;
      dM = m3mul(&M2, &dS);
      M3 Rt = m3t(&Rm), dMt = m3t(&dM);
      real* ds = dL_dscale + 3 * (size_t)i;
      for (int k = 0; k < 3; ++k)
        ds[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
      for (int k = 0; k < 3; ++k) for (int v = 0; v < 3; ++v) dMt.m[k][v] *= s[k];
      const real r = rotations[4 * i], x = rotations[4 * i + 1], y = rotations[4 * i + 2], z = rotations[4 * i + 3];
      real* dq = dL_drot + 4 * (size_t)i;
      dq[0] = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
      dq[1] = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
      dq[2] = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
      dq[3] = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
    }
  }
}
