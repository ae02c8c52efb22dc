// gcr_common.cuh -- shared definitions for the B200 (sm_100a) Gaussian rasterizer kernels.
//
// Domain vocabulary follows the reference (extensions/diff_gaussian_rasterization, "DGR"):
// Gaussians (P of them), tiles (16x16 px), tile-instances (R = num_rendered duplicated
// Gaussian/tile pairs), point_list (sorted instance -> Gaussian index), ranges (per tile
// [start,end) into point_list).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define GCR_TILE_X 16           // DGR/cuda_rasterizer/config.h:16
#define GCR_TILE_Y 16           // DGR/cuda_rasterizer/config.h:17
#define GCR_CHANNELS 3          // DGR/cuda_rasterizer/config.h:15

// ---------------------------------------------------------------------------------------
// Per-Gaussian / per-instance record: 48 B = 3 x float4, 16 B aligned so one tile's slice
// of the instance array can be moved into shared memory with a single TMA bulk copy.
//   q0 = { mean2D.x, mean2D.y, conic.x (A), conic.y (B) }
//   q1 = { conic.z (C), opacity, 2*ln(255*opacity) (cull threshold), unused }     <- projection
//   q2 = { r, g, b, unused }                                                      <- colour stage
// ---------------------------------------------------------------------------------------
struct __align__(16) GcrRecord {
  float4 q0, q1, q2;
};
static_assert(sizeof(GcrRecord) == 48, "record must be 48 bytes");

// Per-Gaussian gradient accumulator written by the backward blend with vector reductions
// (red.global.add.v4.f32): 48 B = 3 x float4.
//   g0 = { dL/dmean2D.x, dL/dmean2D.y, dL/dconic.x, dL/dconic.y }
//   g1 = { dL/dconic.w, dL/dopacity, dL/dcolor.r, dL/dcolor.g }
//   g2 = { dL/dcolor.b, 0, 0, 0 }
struct __align__(16) GcrGradAcc {
  float4 g0, g1, g2;
};
static_assert(sizeof(GcrGradAcc) == 48, "grad accumulator must be 48 bytes");

// Spherical harmonics constants (DGR/cuda_rasterizer/auxiliary.h:21-30)
__device__ static const float GCR_SH_C0 = 0.28209479177387814f;
__device__ static const float GCR_SH_C1 = 0.4886025119029199f;
__device__ static const float GCR_SH_C2[] = {1.0925484305920792f, -1.0925484305920792f,
                                             0.31539156525252005f, -1.0925484305920792f,
                                             0.5462742152960396f};
__device__ static const float GCR_SH_C3[] = {-0.5900435899266435f, 2.890611442640554f,
                                             -0.4570457994644658f, 0.3731763325901154f,
                                             -0.4570457994644658f, 1.445305721320277f,
                                             -0.5900435899266435f};

// Pixel mapping, evaluated in double exactly like DGR auxiliary.h:32-34 (literals are double).
__forceinline__ __device__ float gcr_ndc2pix(float v, int S) {
  return ((v + 1.0) * S - 1.0) * 0.5;
}

// Tile rectangle of a splat; float arithmetic + C truncation as DGR auxiliary.h:36-46.
__forceinline__ __device__ void gcr_get_rect(float px, float py, int max_radius, int gx, int gy,
                                             uint2& rmin, uint2& rmax) {
  rmin.x = min(gx, max(0, (int)((px - max_radius) / GCR_TILE_X)));
  rmin.y = min(gy, max(0, (int)((py - max_radius) / GCR_TILE_Y)));
  rmax.x = min(gx, max(0, (int)((px + max_radius + GCR_TILE_X - 1) / GCR_TILE_X)));
  rmax.y = min(gy, max(0, (int)((py + max_radius + GCR_TILE_Y - 1) / GCR_TILE_Y)));
}

// Pinned blend arithmetic (bit-exact with the SASS nvcc 12.9 emits for DGR forward.cu:307-330:
//   t1 = (C*dy)*dy ; s = fma(dx, A*dx, t1) ; t4 = (B*dx)*dy ; power = fma(s, -0.5, -t4)).
// Explicit _rn intrinsics are never re-contracted, so the result does not depend on how the
// surrounding code is scheduled.
__forceinline__ __device__ float gcr_power(float dx, float dy, float A, float B, float C) {
  float t1 = __fmul_rn(dy, C);
  t1 = __fmul_rn(dy, t1);
  float t2 = __fmul_rn(dx, A);
  float s = __fmaf_rn(dx, t2, t1);
  float t3 = __fmul_rn(dx, B);
  float t4 = __fmul_rn(dy, t3);
  return __fmaf_rn(s, -0.5f, -t4);
}

// ---------------------------------------------------------------------------------------
// mbarrier + TMA bulk-copy wrappers (PTX ISA: mbarrier.*, cp.async.bulk)
// ---------------------------------------------------------------------------------------
__forceinline__ __device__ uint32_t gcr_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__forceinline__ __device__ void gcr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gcr_smem_u32(bar)), "r"(count));
}
__forceinline__ __device__ void gcr_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__forceinline__ __device__ void gcr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gcr_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__forceinline__ __device__ void gcr_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = gcr_smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// One contiguous global->shared bulk copy (TMA engine; SASS UBLKCP.S.G). src, dst 16 B aligned,
// bytes a multiple of 16. Completion is signalled on `bar` via complete_tx.
__forceinline__ __device__ void gcr_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          gcr_smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(gcr_smem_u32(bar))
      : "memory");
}

// Asynchronous gather of ONE 48-byte record into shared memory: three 16-byte cp.async (SASS
// LDGSTS.E.BYPASS.128: global -> shared without a register round trip, L2 only), then an
// arrive-on-completion on `bar` (cp.async.mbarrier.arrive.noinc: the arrival counts against the
// barrier's expected count, which is the number of participating threads).  id < 0: nothing to
// fetch, the thread only arrives.  Every participating thread calls this once per phase.
__forceinline__ __device__ void gcr_gather_record(GcrRecord* dst_smem, const GcrRecord* records, int id,
                                                  uint64_t* bar) {
  if (id >= 0) {
    const uint32_t d = gcr_smem_u32(dst_smem);
    const char* src = reinterpret_cast<const char*>(records + id);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(src + 16) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 32), "l"(src + 32) : "memory");
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(gcr_smem_u32(bar)) : "memory");
}
__forceinline__ __device__ void gcr_cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// Vector float reduction to global memory (SASS REDG.E.ADD.F32x4); addr 16 B aligned.
__forceinline__ __device__ void gcr_red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__forceinline__ __device__ float4 gcr_ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// 256-bit global accesses (sm_100+: LDG.E.ENL2.256 / STG.E.ENL2.256): one full 32-byte sector
// per thread per instruction; addr must be 32 B aligned.
__forceinline__ __device__ void gcr_ldg_nc_v8(const float* p, float (&v)[8]) {
  // volatile: keeps the loads where they are written (up front); without it the compiler sinks
  // them next to their uses and the streaming kernels lose their memory-level parallelism
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
                 "=f"(v[7])
               : "l"(p));
}
__forceinline__ __device__ void gcr_stg_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Programmatic dependent launch (PDL).  A frame of GaussianCity's own size (<= 16 384 points) is a chain
// of ~13 dependent kernels of 4-10 us each, so launch latency between them is a fifth of the frame.
// Kernels of the chain start with gcr_pdl_wait() -- everything the predecessor wrote is visible after it;
// a no-op when the kernel was launched the ordinary way -- followed by gcr_pdl_trigger(), which lets the
// NEXT kernel of the stream be scheduled (its CTAs then sit in their own gcr_pdl_wait()) as soon as every
// CTA of this one has started.  Since every kernel of a chain waits before touching memory, completion
// stays transitive: a kernel that has finished implies all its predecessors have.
__forceinline__ __device__ void gcr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__forceinline__ __device__ void gcr_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline size_t gcr_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Opt-in dynamic shared-memory size: a per-device function attribute that costs microseconds to
// set, so it is set once per (kernel, device).  `mask` is one std::atomic<uint64_t> per kernel
// (bit = device ordinal): safe from any host thread, no lock, and the return code is propagated.
#ifdef __cplusplus
#include <atomic>
template <typename Kernel>
static inline cudaError_t gcr_set_dynamic_smem_once(Kernel kernel, int bytes, std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
  if (bit != 0ull && (mask.load(std::memory_order_acquire) & bit) != 0ull) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && bit != 0ull) mask.fetch_or(bit, std::memory_order_release);
  return e;
}

// <<<>>> with the programmatic-stream-serialization attribute when PDL is on (gcr_pdl_edges(), api.cu).
// Edge classes: the forward chain, the backward blend (its predecessor is a memset), the geometry backward
// (its predecessor is the blend); all three are on when PDL is on (api.cu).
enum { GCR_EDGE_FWD = 1, GCR_EDGE_BLEND_BWD = 2, GCR_EDGE_GEOM_BWD = 4 };
int gcr_pdl_edges();   // bit mask of the edge classes launched programmatically (0 = PDL off)
template <int kEdge = GCR_EDGE_FWD, typename... KArgs, typename... Args>
static inline cudaError_t gcr_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (gcr_pdl_edges() & kEdge) != 0 ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
