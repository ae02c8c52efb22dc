// B200-native multi-resolution hash-grid encoder: forward gather + trilinear
// (2^D-linear) interpolation, embedding-gradient scatter, input gradient.  C ABI in
// include/gcr_grid_encoder.h.
//
// What the reference does (extensions/grid_encoder/grid_encoder_ext.cu) and what changes here:
//
//  * kernel_grid (:95-243) -- one thread per (point, level) walks the 2^D corners, C scalar loads
//    each, and for calc_grad_inputs re-reads every corner D more times (left/right differences).
//    Here a thread owns CPT = min(C, 4) channels of a (point, level): one 16-byte load per corner,
//    all 2^D of them issued before the first use (the kernel is latency / sector bound: random
//    32-byte rows of a 16.8 MB table per level), adjacent lanes covering the two halves of a
//    32-byte row; the input-derivative sums are formed from the SAME registers (no second visit).
//    Levels are the slow grid dimension so one level table at a time is live in the 126 MB L2.
//  * kernel_grid_backward (:245-331) -- 2 channels per thread, one scalar atomicAdd per channel
//    and corner.  Here: one `red.global.add.v4.f32` per (corner, 4 channels): a quarter of the
//    reduction traffic into L2 for C = 8.
//  * kernel_input_backward (:333-360) -- kept as a separate pass over dy_dx for the drop-in entry
//    point (vector loads); the fused entry point recomputes the derivative instead of storing it.
//
// Bit-exactness of the forward and of dy_dx against the reference is by construction: per
// (point, level, channel) the same operations in the same order, with the FMA contraction nvcc
// 12.9 chose for the reference decoded from its SASS (oracle/_ref/grid_encoder_ext*.so,
// kernel_grid<float,2,1>) and pinned with intrinsics:
//      scale = fma(exp2f(level * S), (float)H, -1)          pos = fma(x, scale, 0.5 | 0)
//      acc   = fma(v, w, acc)  corner 0 .. 2^D-1            w   = ((1 * w0) * w1) * ...
//      dacc  = fma(scale * w.., v_right - v_left, dacc)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gcr_grid_encoder.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return 1;
}

int fail_cuda(const char *what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return 2;
}

constexpr int kThreads = 128;

// fast_hash primes (grid_encoder_ext.cu:60-61)
__device__ __forceinline__ constexpr uint32_t prime(int d) {
    return d == 0 ? 1u : d == 1 ? 2654435761u : d == 2 ? 805459861u : d == 3 ? 3674653429u
         : d == 4 ? 2097192037u : d == 5 ? 1434869437u : 2165219737u;
}

// Per-level constants; identical for every thread of a block (level = blockIdx.y).
template <int D>
struct Level {
    uint32_t hashmap_size;
    uint32_t stride[D];    // 0 for the dimensions the reference's index loop never reaches
    float scale;
    bool hashed;           // gridtype == 0 && the dense index would not fit (get_grid_index :91)
    bool pow2;
};

template <int D>
__device__ __forceinline__ Level<D> level_setup(const int *__restrict__ offsets, uint32_t level, float S,
                                                uint32_t H, uint32_t gridtype, bool align_corners) {
    Level<D> lv;
    lv.hashmap_size = (uint32_t)(__ldg(offsets + level + 1) - __ldg(offsets + level));
    // exp2f(level * S) * H - 1.0f contracts to one FMA in the reference build
    lv.scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);
    const uint32_t resolution = __float2uint_ru(lv.scale) + 1u;
    const uint32_t step = align_corners ? resolution : resolution + 1u;
    uint32_t stride = 1;
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (stride <= lv.hashmap_size) {        // the reference's loop condition (:83), d-prefix
            lv.stride[d] = stride;
            stride *= step;
        } else {
            lv.stride[d] = 0;
        }
    }
    lv.hashed = gridtype == 0 && stride > lv.hashmap_size;
    lv.pow2 = (lv.hashmap_size & (lv.hashmap_size - 1u)) == 0u;
    return lv;
}

template <int D>
struct Cell {
    float frac[D];
    uint32_t lo[D], hi[D];     // per-dimension index contributions of the low / high corner
};

// pos = x * scale + (0.5 | 0), split into cell and fraction (:147-151); contributions of each
// dimension to the table index (hash: coordinate * prime, dense: coordinate * stride).
template <int D>
__device__ __forceinline__ Cell<D> cell_setup(const float (&x)[D], const Level<D> &lv, bool align_corners) {
    Cell<D> c;
    const float half = align_corners ? 0.0f : 0.5f;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const float pos = __fmaf_rn(x[d], lv.scale, half);
        const uint32_t g = __float2uint_rd(pos);
        c.frac[d] = __fsub_rn(pos, (float)g);
        const uint32_t m = lv.hashed ? prime(d) : lv.stride[d];
        c.lo[d] = g * m;
        c.hi[d] = c.lo[d] + m;
    }
    return c;
}

template <int D>
__device__ __forceinline__ uint32_t corner_index(const Cell<D> &c, const Level<D> &lv, int corner) {
    uint32_t idx = (corner & 1) ? c.hi[0] : c.lo[0];
    if (lv.hashed) {
#pragma unroll
        for (int d = 1; d < D; d++) idx ^= ((corner >> d) & 1) ? c.hi[d] : c.lo[d];
    } else {
#pragma unroll
        for (int d = 1; d < D; d++) idx += ((corner >> d) & 1) ? c.hi[d] : c.lo[d];
    }
    return lv.pow2 ? (idx & (lv.hashmap_size - 1u)) : (idx % lv.hashmap_size);
}

template <int D>
__device__ __forceinline__ float corner_weight(const Cell<D> &c, int corner) {
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < D; d++)
        w = __fmul_rn(w, ((corner >> d) & 1) ? c.frac[d] : __fsub_rn(1.0f, c.frac[d]));
    return w;
}

template <int N> struct Vec;
template <> struct Vec<1> { using T = float; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

template <int N>
__device__ __forceinline__ void load_row(float (&v)[N], const float *p) {
    typename Vec<N>::T t = __ldg(reinterpret_cast<const typename Vec<N>::T *>(p));
    memcpy(v, &t, sizeof(t));
}

template <int N>
__device__ __forceinline__ void store_row(float *p, const float (&v)[N]) {
    typename Vec<N>::T t;
    memcpy(&t, v, sizeof(t));
    *reinterpret_cast<typename Vec<N>::T *>(p) = t;
}

template <int N>
__device__ __forceinline__ void red_add_row(float *p, const float (&v)[N]) {
    if constexpr (N == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                     "f"(v[3])
                     : "memory");
    } else if constexpr (N == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
    }
}

template <int D>
__device__ __forceinline__ bool load_point(float (&x)[D], const float *__restrict__ inputs, uint32_t b) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; d++) {
        x[d] = __ldg(inputs + (size_t)b * D + d);
        oob |= (x[d] < 0.0f) || (x[d] > 1.0f);      // NaN passes, as in the reference (:117-121)
    }
    return oob;
}

// ---- forward ---------------------------------------------------------------------------------------
// thread = (level, point, channel group); grid (ceil(B * C/CPT / 128), L).
template <int D, int C, bool GRAD>
__global__ void __launch_bounds__(kThreads)
grid_fwd_kernel(const float *__restrict__ inputs, const float *__restrict__ grid, const int *__restrict__ offsets,
                float *__restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, float *__restrict__ dy_dx,
                uint32_t gridtype, bool align_corners) {
    constexpr int CPT = C < 4 ? C : 4;
    constexpr int TPP = C / CPT;
    constexpr int NC = 1 << D;
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    const uint32_t b = t / TPP;
    const uint32_t ch = (t % TPP) * CPT;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;

    float x[D];
    const bool oob = load_point<D>(x, inputs, b);
    float *out = outputs + ((size_t)level * B + b) * C + ch;
    float *dout = GRAD ? dy_dx + ((size_t)b * L + level) * (D * C) + ch : nullptr;
    if (oob) {                                           // zeros, and zero derivative (:123-138)
        float z[CPT] = {};
        store_row<CPT>(out, z);
        if (GRAD) {
#pragma unroll
            for (int d = 0; d < D; d++) store_row<CPT>(dout + d * C, z);
        }
        return;
    }
    const Level<D> lv = level_setup<D>(offsets, level, S, H, gridtype, align_corners);
    const Cell<D> cell = cell_setup<D>(x, lv, align_corners);
    const float *table = grid + (size_t)(uint32_t)__ldg(offsets + level) * C + ch;

    float v[NC][CPT];
#pragma unroll
    for (int k = 0; k < NC; k++) load_row<CPT>(v[k], table + (size_t)corner_index<D>(cell, lv, k) * C);

    float acc[CPT] = {};
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const float w = corner_weight<D>(cell, k);
#pragma unroll
        for (int c = 0; c < CPT; c++) acc[c] = __fmaf_rn(v[k][c], w, acc[c]);
    }
    store_row<CPT>(out, acc);

    if (GRAD) {                                          // d out / d x_gd (:196-241), from registers
#pragma unroll
        for (int gd = 0; gd < D; gd++) {
            float dacc[CPT] = {};
#pragma unroll
            for (int k = 0; k < NC / 2; k++) {
                float w = lv.scale;
                int left = 0;
#pragma unroll
                for (int nd = 0; nd < D - 1; nd++) {
                    const int d = nd >= gd ? nd + 1 : nd;
                    const bool up = (k >> nd) & 1;
                    w = __fmul_rn(w, up ? cell.frac[d] : __fsub_rn(1.0f, cell.frac[d]));
                    left |= up ? (1 << d) : 0;
                }
                const int right = left | (1 << gd);
#pragma unroll
                for (int c = 0; c < CPT; c++)
                    dacc[c] = __fmaf_rn(w, __fsub_rn(v[right][c], v[left][c]), dacc[c]);
            }
            store_row<CPT>(dout + gd * C, dacc);
        }
    }
}

// ---- backward: embedding gradient ------------------------------------------------------------------
// Same thread mapping; one vector reduction per (corner, channel group).  FUSED additionally
// forms this level's contribution to grad_inputs from the corner rows (dy_dx never stored).
template <int D, int C, bool FUSED>
__global__ void __launch_bounds__(kThreads)
grid_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ inputs, const float *__restrict__ grid,
                const int *__restrict__ offsets, float *__restrict__ grad_grid, uint32_t B, uint32_t L, float S,
                uint32_t H, float *__restrict__ grad_inputs, uint32_t gridtype, bool align_corners) {
    constexpr int CPT = C < 4 ? C : 4;
    constexpr int TPP = C / CPT;
    constexpr int NC = 1 << D;
    const uint32_t t = blockIdx.x * kThreads + threadIdx.x;
    const uint32_t b = t / TPP;
    const uint32_t ch = (t % TPP) * CPT;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;

    float x[D];
    if (load_point<D>(x, inputs, b)) return;             // gradient buffers are zero-initialised (:276-281)
    const Level<D> lv = level_setup<D>(offsets, level, S, H, gridtype, align_corners);
    const Cell<D> cell = cell_setup<D>(x, lv, align_corners);
    const size_t base = (size_t)(uint32_t)__ldg(offsets + level) * C + ch;

    float g[CPT];
    load_row<CPT>(g, grad + ((size_t)level * B + b) * C + ch);

    if (!FUSED) {
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const float w = corner_weight<D>(cell, k);
            float wg[CPT];
#pragma unroll
            for (int c = 0; c < CPT; c++) wg[c] = __fmul_rn(w, g[c]);
            red_add_row<CPT>(grad_grid + base + (size_t)corner_index<D>(cell, lv, k) * C, wg);
        }
    } else {
        float v[NC][CPT];                                // the rows the forward read, again
        uint32_t idx[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            idx[k] = corner_index<D>(cell, lv, k);
            load_row<CPT>(v[k], grid + base + (size_t)idx[k] * C);
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const float w = corner_weight<D>(cell, k);
            float wg[CPT];
#pragma unroll
            for (int c = 0; c < CPT; c++) wg[c] = __fmul_rn(w, g[c]);
            red_add_row<CPT>(grad_grid + base + (size_t)idx[k] * C, wg);
        }
#pragma unroll
        for (int gd = 0; gd < D; gd++) {
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < NC / 2; k++) {
                float w = lv.scale;
                int left = 0;
#pragma unroll
                for (int nd = 0; nd < D - 1; nd++) {
                    const int d = nd >= gd ? nd + 1 : nd;
                    const bool up = (k >> nd) & 1;
                    w *= up ? cell.frac[d] : 1.0f - cell.frac[d];
                    left |= up ? (1 << d) : 0;
                }
                const int right = left | (1 << gd);
                float t = 0.0f;                          // <row difference, upstream gradient>
#pragma unroll
                for (int c = 0; c < CPT; c++) t = fmaf(v[right][c] - v[left][c], g[c], t);
                s = fmaf(w, t, s);
            }
            // the TPP lanes of a point are adjacent: fold them before the atomic
#pragma unroll
            for (int o = 1; o < TPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (ch == 0) atomicAdd(grad_inputs + (size_t)b * D + gd, s);
        }
    }
}

// ---- backward: input gradient from the stored dy_dx (kernel_input_backward :333-360) ---------------
// thread = (point, dimension); sum over levels then channels in the reference's order, one FMA each.
template <int D, int C>
__global__ void __launch_bounds__(256)
grid_input_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ dy_dx,
                      float *__restrict__ grad_inputs, uint32_t B, uint32_t L) {
    constexpr int CPT = C < 4 ? C : 4;
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float *dd = dy_dx + (size_t)b * L * (D * C) + d * C;
    float result = 0.0f;
    for (uint32_t l = 0; l < L; l++) {
        const float *gp = grad + ((size_t)l * B + b) * C;
        const float *dp = dd + (size_t)l * (D * C);
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += CPT) {
            float gv[CPT], dv[CPT];
            load_row<CPT>(gv, gp + c0);
            load_row<CPT>(dv, dp + c0);
#pragma unroll
            for (int c = 0; c < CPT; c++) result = __fmaf_rn(gv[c], dv[c], result);
        }
    }
    grad_inputs[t] = result;
}

// ---- host dispatch ---------------------------------------------------------------------------------
struct Args {
    const float *grad, *inputs, *embeddings;
    const int *offsets;
    float *outputs, *dy_dx, *grad_embeddings, *grad_inputs;
    const float *dy_dx_in;
    uint32_t B, L, H, gridtype;
    float S;
    bool calc, align;
    cudaStream_t stream;
};

enum Op { FWD, BWD, BWD_FUSED };

template <int D, int C>
cudaError_t launch(Op op, const Args &a) {
    constexpr int CPT = C < 4 ? C : 4;
    constexpr int TPP = C / CPT;
    const uint64_t threads = (uint64_t)a.B * TPP;
    const dim3 gridDim((unsigned)((threads + kThreads - 1) / kThreads), a.L, 1);
    if (op == FWD) {
        if (a.calc)
            grid_fwd_kernel<D, C, true><<<gridDim, kThreads, 0, a.stream>>>(
                a.inputs, a.embeddings, a.offsets, a.outputs, a.B, a.L, a.S, a.H, a.dy_dx, a.gridtype, a.align);
        else
            grid_fwd_kernel<D, C, false><<<gridDim, kThreads, 0, a.stream>>>(
                a.inputs, a.embeddings, a.offsets, a.outputs, a.B, a.L, a.S, a.H, nullptr, a.gridtype, a.align);
    } else if (op == BWD) {
        grid_bwd_kernel<D, C, false><<<gridDim, kThreads, 0, a.stream>>>(
            a.grad, a.inputs, nullptr, a.offsets, a.grad_embeddings, a.B, a.L, a.S, a.H, nullptr, a.gridtype, a.align);
        if (a.calc) {
            const uint64_t n = (uint64_t)a.B * D;
            grid_input_bwd_kernel<D, C><<<(unsigned)((n + 255) / 256), 256, 0, a.stream>>>(
                a.grad, a.dy_dx_in, a.grad_inputs, a.B, a.L);
        }
    } else {
        grid_bwd_kernel<D, C, true><<<gridDim, kThreads, 0, a.stream>>>(
            a.grad, a.inputs, a.embeddings, a.offsets, a.grad_embeddings, a.B, a.L, a.S, a.H, a.grad_inputs,
            a.gridtype, a.align);
    }
    return cudaGetLastError();
}

template <int D>
int dispatch_c(Op op, uint32_t C, const Args &a, cudaError_t *e) {
    switch (C) {
    case 1: *e = launch<D, 1>(op, a); return 0;
    case 2: *e = launch<D, 2>(op, a); return 0;
    case 4: *e = launch<D, 4>(op, a); return 0;
    case 8: *e = launch<D, 8>(op, a); return 0;
    default: return 1;
    }
}

int dispatch(Op op, uint32_t D, uint32_t C, const Args &a) {
    cudaError_t e = cudaSuccess;
    int bad;
    switch (D) {
    case 2: bad = dispatch_c<2>(op, C, a, &e); break;
    case 3: bad = dispatch_c<3>(op, C, a, &e); break;
    case 4: bad = dispatch_c<4>(op, C, a, &e); break;
    case 5: bad = dispatch_c<5>(op, C, a, &e); break;
    default: bad = 1;
    }
    if (bad) return fail("GridEncoding: C must be 1, 2, 4, or 8.");
    if (e != cudaSuccess) return fail_cuda("grid encoder launch", e);
    return 0;
}

bool misaligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

int check_common(const float *inputs, const int *offsets, uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    // the reference throws this text for an unsupported C and, verbatim, for an unsupported D
    // (grid_encoder_ext.cu:392,429)
    if (D < 2 || D > 5 || !(C == 1 || C == 2 || C == 4 || C == 8)) return fail("GridEncoding: C must be 1, 2, 4, or 8.");
    if (!inputs || !offsets) return fail("gcr_grid: inputs / offsets must not be NULL");
    if (L == 0 || L > 65535u) return fail("gcr_grid: L must be in [1, 65535]");
    // thread indices are 32-bit: B * C / min(C, 4) in the gather / scatter kernels, B * D in the input gradient
    if ((uint64_t)B * (C > D ? C : D) >= (1ull << 32)) return fail("gcr_grid: B * max(C, D) must be below 2^32");
    return 0;
}

}  // namespace

extern "C" {

int gcr_grid_abi_version(void) { return GCR_GRID_ABI_VERSION; }

const char *gcr_grid_last_error(void) { return g_err; }

int gcr_grid_encode_forward(const float *inputs, const float *embeddings, const int *offsets, float *outputs,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            int calc_grad_inputs, float *dy_dx, uint32_t gridtype, int align_corners, void *stream) {
    if (B == 0) return 0;
    if (int r = check_common(inputs, offsets, B, D, C, L)) return r;
    if (!embeddings || !outputs) return fail("gcr_grid_encode_forward: embeddings / outputs must not be NULL");
    if (calc_grad_inputs && !dy_dx) return fail("gcr_grid_encode_forward: dy_dx must not be NULL with calc_grad_inputs");
    const size_t al = 4 * (C < 4 ? C : 4);
    if (misaligned(embeddings, al) || misaligned(outputs, al) || (calc_grad_inputs && misaligned(dy_dx, al)))
        return fail("gcr_grid_encode_forward: embeddings / outputs / dy_dx must be aligned to min(C, 4) floats");
    Args a{};
    a.inputs = inputs; a.embeddings = embeddings; a.offsets = offsets; a.outputs = outputs; a.dy_dx = dy_dx;
    a.B = B; a.L = L; a.H = H; a.S = S; a.gridtype = gridtype; a.calc = calc_grad_inputs != 0;
    a.align = align_corners != 0; a.stream = (cudaStream_t)stream;
    return dispatch(FWD, D, C, a);
}

int gcr_grid_encode_backward(const float *grad, const float *inputs, const float *embeddings, const int *offsets,
                             float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                             uint32_t H, int calc_grad_inputs, const float *dy_dx, float *grad_inputs,
                             uint32_t gridtype, int align_corners, void *stream) {
    (void)embeddings;
    if (B == 0) return 0;
    if (int r = check_common(inputs, offsets, B, D, C, L)) return r;
    if (!grad || !grad_embeddings) return fail("gcr_grid_encode_backward: grad / grad_embeddings must not be NULL");
    if (calc_grad_inputs && (!dy_dx || !grad_inputs))
        return fail("gcr_grid_encode_backward: dy_dx / grad_inputs must not be NULL with calc_grad_inputs");
    const size_t al = 4 * (C < 4 ? C : 4);
    if (misaligned(grad, al) || misaligned(grad_embeddings, al) || (calc_grad_inputs && misaligned(dy_dx, al)))
        return fail("gcr_grid_encode_backward: grad / grad_embeddings / dy_dx must be aligned to min(C, 4) floats");
    Args a{};
    a.grad = grad; a.inputs = inputs; a.offsets = offsets; a.grad_embeddings = grad_embeddings;
    a.dy_dx_in = dy_dx; a.grad_inputs = grad_inputs;
    a.B = B; a.L = L; a.H = H; a.S = S; a.gridtype = gridtype; a.calc = calc_grad_inputs != 0;
    a.align = align_corners != 0; a.stream = (cudaStream_t)stream;
    return dispatch(BWD, D, C, a);
}

int gcr_grid_encode_backward_fused(const float *grad, const float *inputs, const float *embeddings,
                                   const int *offsets, float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                                   uint32_t L, float S, uint32_t H, float *grad_inputs, uint32_t gridtype,
                                   int align_corners, void *stream) {
    if (B == 0) return 0;
    if (int r = check_common(inputs, offsets, B, D, C, L)) return r;
    if (!grad || !grad_embeddings || !embeddings || !grad_inputs)
        return fail("gcr_grid_encode_backward_fused: grad / embeddings / grad_embeddings / grad_inputs must not be NULL");
    const size_t al = 4 * (C < 4 ? C : 4);
    if (misaligned(grad, al) || misaligned(grad_embeddings, al) || misaligned(embeddings, al))
        return fail("gcr_grid_encode_backward_fused: grad / embeddings / grad_embeddings must be aligned to min(C, 4) floats");
    Args a{};
    a.grad = grad; a.inputs = inputs; a.embeddings = embeddings; a.offsets = offsets;
    a.grad_embeddings = grad_embeddings; a.grad_inputs = grad_inputs;
    a.B = B; a.L = L; a.H = H; a.S = S; a.gridtype = gridtype; a.calc = true;
    a.align = align_corners != 0; a.stream = (cudaStream_t)stream;
    return dispatch(BWD_FUSED, D, C, a);
}

}  // extern "C"
