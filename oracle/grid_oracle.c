/*
 * grid_oracle.c -- CPU restatement of the reference's multi-resolution hash-grid encoder
 * (extensions/grid_encoder/grid_encoder_ext.cu).  TEST INFRASTRUCTURE ONLY: imported by tests/,
 * bench_grid_encoder.py's cpu_baseline leg and nothing else; gaussiancity_b200 never
 * links, loads or calls it.
 *
 * What it follows, function by function:
 *   ggo_table_index     get_grid_index + fast_hash         grid_encoder_ext.cu:51-96
 *   ggo_forward         kernel_grid                        grid_encoder_ext.cu:98-243
 *   ggo_backward_grid   kernel_grid_backward               grid_encoder_ext.cu:245-331
 *   ggo_backward_input  kernel_input_backward              grid_encoder_ext.cu:333-360
 *
 * Arithmetic: fp32, compiled with -ffp-contract=off; fmaf() is written exactly where nvcc 12.9
 * fused a multiply-add in the reference build (decoded from the SASS of
 * oracle/_ref/grid_encoder_ext*.so): the level scale, the position, the corner accumulation
 * and the derivative accumulation.  The one operation a CPU cannot reproduce bit for bit is
 * exp2f(level * S): the GPU evaluates it with MUFU.EX2 (within 2 ulp), libm rounds correctly; the
 * `scales` argument takes the level scale from outside for that reason.
 *
 * Pinning: tests/golden/grid/*.npz are outputs of the unmodified reference extension on a B200
 * (tests/golden/make_golden_grid.py, 5 cases: hashed / dense / tiled levels, align_corners,
 * D = 2..5, C = 1..8, out-of-range points).  With the device's exp2f value (recovered per level
 * among the 5 floats within 2 ulp of libm's -- exactly one reproduces the level's outputs) the
 * oracle's outputs, dy_dx and grad_inputs equal the reference's BIT FOR BIT on every case; with
 * libm's own value they agree to 1e-5 (one ulp of scale at resolution 2000 moves a position by
 * 1e-4 of a cell).  -> parity is pinned.
 *
 * Sums that the reference forms with atomics (embedding gradients) are formed here in point
 * order per level -- the reference's own order is nondeterministic.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GGO_MAX_D 7

static const uint32_t GGO_PRIMES[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                       2097192037u, 1434869437u, 2165219737u};

/* get_grid_index (:73-96): dense index over the leading dimensions while the running stride
 * still fits the level table; hashed when it does not fit and gridtype == 0. */
uint32_t ggo_table_index(uint32_t D, uint32_t gridtype, int align_corners, uint32_t hashmap_size,
                         uint32_t resolution, const uint32_t *pos_grid) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
        for (uint32_t d = 0; d < D; d++) index ^= pos_grid[d] * GGO_PRIMES[d];
    }
    return index % hashmap_size;
}

float ggo_level_scale(uint32_t level, float S, uint32_t H) {
    return fmaf(exp2f((float)level * S), (float)H, -1.0f);
}

typedef struct {
    float frac[GGO_MAX_D];
    uint32_t cell[GGO_MAX_D];
    float scale;
    uint32_t resolution, hashmap_size;
    int oob;
} ggo_site;

/* `scales` (optional, one float per level) replaces ggo_level_scale(): it lets a test hand in the
 * value the GPU's exp2f produced, the only operation of the path libm does not reproduce bit for bit. */
static void ggo_locate(ggo_site *s, const float *x, const int *offsets, uint32_t D, uint32_t level, float S,
                       uint32_t H, int align_corners, const float *scales) {
    s->oob = 0;
    for (uint32_t d = 0; d < D; d++)
        if (x[d] < 0 || x[d] > 1) s->oob = 1;
    s->hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    s->scale = scales ? scales[level] : ggo_level_scale(level, S, H);
    s->resolution = (uint32_t)ceilf(s->scale) + 1;
    for (uint32_t d = 0; d < D; d++) {
        float pos = fmaf(x[d], s->scale, align_corners ? 0.0f : 0.5f);
        float fl = floorf(pos);
        s->cell[d] = (uint32_t)fl;
        s->frac[d] = pos - (float)s->cell[d];
    }
}

static uint32_t ggo_corner(const ggo_site *s, uint32_t D, uint32_t corner, uint32_t gridtype, int align_corners,
                           float *w_out) {
    uint32_t pg[GGO_MAX_D];
    float w = 1;
    for (uint32_t d = 0; d < D; d++) {
        if ((corner & (1u << d)) == 0) {
            w *= 1 - s->frac[d];
            pg[d] = s->cell[d];
        } else {
            w *= s->frac[d];
            pg[d] = s->cell[d] + 1;
        }
    }
    *w_out = w;
    return ggo_table_index(D, gridtype, align_corners, s->hashmap_size, s->resolution, pg);
}

/* kernel_grid: outputs [L,B,C]; dy_dx [B,L,D,C] when calc_grad_inputs. */
void ggo_forward(const float *inputs, const float *embeddings, const int *offsets, float *outputs, uint32_t B,
                 uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float *dy_dx,
                 uint32_t gridtype, int align_corners, const float *scales) {
    for (uint32_t level = 0; level < L; level++) {
        const float *table = embeddings + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; b++) {
            float *out = outputs + ((size_t)level * B + b) * C;
            float *dd = calc_grad_inputs ? dy_dx + ((size_t)b * L + level) * D * C : 0;
            ggo_site s;
            ggo_locate(&s, inputs + (size_t)b * D, offsets, D, level, S, H, align_corners, scales);
            if (s.oob) {
                memset(out, 0, sizeof(float) * C);
                if (dd) memset(dd, 0, sizeof(float) * D * C);
                continue;
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                float w;
                uint32_t idx = ggo_corner(&s, D, corner, gridtype, align_corners, &w);
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = fmaf(table[(size_t)idx * C + ch], w, out[ch]);
            }
            if (!dd) continue;
            for (uint32_t gd = 0; gd < D; gd++) {
                for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = 0;
                for (uint32_t k = 0; k < (1u << (D - 1)); k++) {
                    float w = s.scale;
                    uint32_t pg[GGO_MAX_D];
                    for (uint32_t nd = 0; nd < D - 1; nd++) {
                        uint32_t d = nd >= gd ? nd + 1 : nd;
                        if ((k & (1u << nd)) == 0) {
                            w *= 1 - s.frac[d];
                            pg[d] = s.cell[d];
                        } else {
                            w *= s.frac[d];
                            pg[d] = s.cell[d] + 1;
                        }
                    }
                    pg[gd] = s.cell[gd];
                    uint32_t left = ggo_table_index(D, gridtype, align_corners, s.hashmap_size, s.resolution, pg);
                    pg[gd] = s.cell[gd] + 1;
                    uint32_t right = ggo_table_index(D, gridtype, align_corners, s.hashmap_size, s.resolution, pg);
                    for (uint32_t ch = 0; ch < C; ch++)
                        dd[gd * C + ch] = fmaf(w, table[(size_t)right * C + ch] - table[(size_t)left * C + ch],
                                               dd[gd * C + ch]);
                }
            }
        }
    }
}

/* kernel_grid_backward: grad [L,B,C] scattered into grad_embeddings (accumulated, caller zeroes). */
void ggo_backward_grid(const float *grad, const float *inputs, const int *offsets, float *grad_embeddings,
                       uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                       int align_corners, const float *scales) {
    for (uint32_t level = 0; level < L; level++) {
        float *table = grad_embeddings + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; b++) {
            const float *g = grad + ((size_t)level * B + b) * C;
            ggo_site s;
            ggo_locate(&s, inputs + (size_t)b * D, offsets, D, level, S, H, align_corners, scales);
            if (s.oob) continue;
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                float w;
                uint32_t idx = ggo_corner(&s, D, corner, gridtype, align_corners, &w);
                for (uint32_t ch = 0; ch < C; ch++) table[(size_t)idx * C + ch] += w * g[ch];
            }
        }
    }
}

/* kernel_input_backward: grad_inputs[b,d] = sum_l sum_ch grad[l,b,ch] * dy_dx[b,l,d,ch]. */
void ggo_backward_input(const float *grad, const float *dy_dx, float *grad_inputs, uint32_t B, uint32_t D,
                        uint32_t C, uint32_t L) {
    for (uint32_t b = 0; b < B; b++)
        for (uint32_t d = 0; d < D; d++) {
            float r = 0;
            for (uint32_t l = 0; l < L; l++)
                for (uint32_t ch = 0; ch < C; ch++)
                    r = fmaf(grad[((size_t)l * B + b) * C + ch], dy_dx[(((size_t)b * L + l) * D + d) * C + ch], r);
            grad_inputs[(size_t)b * D + d] = r;
        }
}

/* Test helper: the table row (within its level) of every corner of every (level, point):
 * rows [L,B,2^D] uint32, weights [L,B,2^D] -- lets a test compare integer work exactly. */
void ggo_corner_rows(const float *inputs, const int *offsets, uint32_t *rows, float *weights, uint32_t B,
                     uint32_t D, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                     const float *scales) {
    for (uint32_t level = 0; level < L; level++)
        for (uint32_t b = 0; b < B; b++) {
            ggo_site s;
            ggo_locate(&s, inputs + (size_t)b * D, offsets, D, level, S, H, align_corners, scales);
            for (uint32_t corner = 0; corner < (1u << D); corner++) {
                size_t o = (((size_t)level * B + b) << D) + corner;
                rows[o] = ggo_corner(&s, D, corner, gridtype, align_corners, &weights[o]);
            }
        }
}
