// binning.cu -- cub-free stable radix sort, fused scan + tile-instance emission, and tile ranges
// for sm_100a.
//
// Behavioural spec (reference): cub::DeviceScan::InclusiveSum + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs(64-bit tile|depth keys) + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:66-124, 228-270).
//
// B200-first redesign: the reference sorts R tile-instances on 32+log2(T) key bits (6 passes of
// 24 B/pair at 1080p).  Here the visible Gaussians are depth-sorted ONCE (4 passes over 8-byte
// pairs; culled Gaussians are dropped by the first pass), instances are emitted in that order,
// and a stable split on the tile id (2 passes over R 8-byte pairs at 1080p) finishes the job.
// A stable sort on (tile, depth) of index-ordered input equals a stable depth sort followed by a
// stable tile sort, so point_list and ranges are bit-identical to the reference's, ties
// included, at ~1/3 of the HBM traffic.
//
// Every element count after the preprocess lives on the DEVICE (visible Gaussians, R): grids are
// sized for the host-known upper bound and surplus CTAs leave at once, so the host pipeline
// never has to wait for a count (api.cu reads R once, only to size the binning buffer in its
// "exact" mode).
#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

constexpr int kSortThreads = 256;
constexpr int kBins = 256;
constexpr int kMaxPasses = 4;
constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;   // depth key of a culled Gaussian (preprocess.cu)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Lanes of `vmask` holding the same `nbits`-bit digit as the caller.  Built from one ballot per
// digit bit: on sm_100 MATCH.ANY resolves one distinct value at a time (hundreds of cycles for a
// high-entropy digit, measured: it bounded the whole sort), ballots pipeline.
__device__ __forceinline__ unsigned digit_peers(uint32_t d, int nbits, unsigned vmask) {
  unsigned peers = vmask;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    if (b < nbits) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(vmask, bit);
      peers &= bit ? bal : ~bal;
    }
  }
  return peers;
}

// exclusive scan of one value per DIGIT (threads 0..255 carry values, any others pass 0);
// *total (optional) receives the sum.  smem: 9 words.
__device__ __forceinline__ uint32_t digit_excl_scan(uint32_t v, uint32_t* smem, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31 && warp < 8) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = (lane < 8) ? smem[lane] : 0;
    const uint32_t winc = warp_incl_scan(w, lane);
    if (lane < 8) smem[lane] = winc - w;
    if (lane == 7) smem[8] = winc;
  }
  __syncthreads();
  const uint32_t res = inc - v + ((warp < 8) ? smem[warp] : 0u);
  if (total != nullptr) *total = smem[8];
  __syncthreads();
  return res;
}

__device__ __forceinline__ uint32_t load_count(const uint32_t* n_ptr, uint32_t n_max) {
  return n_ptr != nullptr ? min(*n_ptr, n_max) : n_max;
}

// ------------------------------------------------------------------------------------------
// Onesweep-style radix sort (Adinets & Merrill): ONE histogram kernel computes the global digit
// histograms of every pass (they are permutation invariant), then each pass is a single kernel
// whose tiles obtain their output offsets with a decoupled look-back over per-tile status words.
// 16 B/pair/pass of HBM traffic and 1 + passes launches per sort.  Tiles are handed out by an
// atomic ticket so a tile's predecessors have always started: the look-back cannot deadlock.
//
// Order inside a tile ("counts first", measured 10-15 % faster than rank-then-publish on B200,
// profiles/r02_variants_ab.md): the digit counts come first -- one shared-memory atomic per key,
// no order needed -- and the tile's aggregate is published before any ranking, so ranking is off
// the inter-tile dependency chain; the stable ranking (ballot-built peer masks + per-warp
// counters) then starts from the scanned per-warp bases and drops keys straight into the
// digit-ordered staging buffer, the values follow, and the look-back runs last, when the
// predecessors have long published.  Global stores are coalesced per digit run.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kFlagAgg = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1u;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_all_kernel(const uint32_t* __restrict__ keys, uint32_t n_max,
                      const uint32_t* __restrict__ n_ptr, int end_bit, bool skip_invalid,
                      uint32_t* __restrict__ ghist /* [kMaxPasses][kBins] */) {
  constexpr int kItems = 16;
  constexpr int kChunk = kSortThreads * kItems;
  __shared__ uint32_t hist[kMaxPasses][kBins];
  gcr_pdl_wait();
  gcr_pdl_trigger();
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = load_count(n_ptr, n_max);
  const int npass = (end_bit + 7) / 8;
  const uint32_t nchunks = (n + kChunk - 1) / kChunk;
  for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const uint32_t base = c * kChunk + threadIdx.x;
    uint32_t key[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const uint32_t i = base + (uint32_t)r * kSortThreads;
      key[r] = i < n ? keys[i] : kInvalidKey;
    }
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const uint32_t i = base + (uint32_t)r * kSortThreads;
      if (i < n && !(skip_invalid && key[r] == kInvalidKey)) {
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p) {
          if (p < npass) {
            const int bits = min(8, end_bit - 8 * p);
            atomicAdd(&hist[p][(key[r] >> (8 * p)) & ((1u << bits) - 1u)], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p)
    if (p < npass && hist[p][threadIdx.x] != 0) atomicAdd(&ghist[p * kBins + threadIdx.x], hist[p][threadIdx.x]);
}

// kThreads x kItems pairs per tile (256 x 16 = 4096 by default; 256 x 4 = 1024 for small inputs,
// where the 16 serial ranking rounds of a few lonely CTAs are the whole latency); all shared
// memory dynamic.  kFirstDepth: first pass of the depth sort -- values are generated as the
// element index and culled Gaussians (key 0xFFFFFFFF) are dropped, so every later stage works on
// the compacted, still index-ordered (stable) list.  `n_out` (tile 0 writes it) receives the
// number of elements this pass outputs.
template <bool kFirstDepth, int kThreads, int kItems>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 4 : 2)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     uint32_t n_max, const uint32_t* __restrict__ n_ptr, int shift,
                     uint32_t digit_mask, int nbits,
                     const uint32_t* __restrict__ ghist_pass /* [kBins] */,
                     volatile uint32_t* status /* [tiles][kBins], zeroed */, uint32_t* ticket,
                     uint32_t* __restrict__ n_out) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kTile = kThreads * kItems;
  extern __shared__ __align__(16) uint32_t os_smem[];
  gcr_pdl_wait();
  gcr_pdl_trigger();
  uint32_t (*warp_hist)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(os_smem);
  uint32_t* gbase = os_smem + kWarps * kBins;
  uint32_t* bstart = gbase + kBins;
  uint32_t* sm = bstart + kBins;          // 16 words
  uint32_t* st_keys = sm + 16;
  uint32_t* st_vals = st_keys + kTile;
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t n = load_count(n_ptr, n_max);
  if ((uint64_t)tile * kTile >= n) {
    // surplus CTA (the grid is sized for n_max; on a striped frame most CTAs of the later passes
    // end here, so nothing is done before this test).  An empty input still has to report its count.
    if (n_out != nullptr && tile == 0 && tid == 0) *n_out = 0;
    return;
  }
  for (int i = tid; i < kWarps * kBins; i += kThreads) (&warp_hist[0][0])[i] = 0;
  __syncthreads();

  const uint32_t chunk_base = tile * (uint32_t)kTile;
  const uint32_t base = chunk_base + (uint32_t)warp * (32 * kItems);
  uint32_t key[kItems];
  uint16_t pos16[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const uint32_t i = base + (uint32_t)r * 32 + lane;
    key[r] = i < n ? keys_in[i] : kInvalidKey;
  }
  // 1. counts of this warp's keys per digit (order-free)
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const uint32_t i = base + (uint32_t)r * 32 + lane;
    const bool valid = i < n && !(kFirstDepth && key[r] == kInvalidKey);
    if (valid) atomicAdd(&warp_hist[warp][(key[r] >> shift) & digit_mask], 1u);
  }
  __syncthreads();

  // 2. tile counts -> publish; per-warp starting positions inside the tile
  uint32_t acc = 0, dstart = 0, tile_total = 0;
  volatile uint32_t* mine = status + (size_t)tile * kBins + tid;
  {
    if (tid < kBins) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t t = warp_hist[w][tid];
        warp_hist[w][tid] = acc;
        acc += t;
      }
      *mine = (tile == 0 ? kFlagIncl : kFlagAgg) | acc;
    }
    const uint32_t bs = digit_excl_scan(acc, sm, &tile_total);
    uint32_t n_total;
    dstart = digit_excl_scan(tid < kBins ? ghist_pass[tid] : 0u, sm, &n_total);
    if (n_out != nullptr && tile == 0 && tid == 0) *n_out = n_total;
    if (tid < kBins) {
      bstart[tid] = bs;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) warp_hist[w][tid] += bs;
    }
  }
  __syncthreads();

  // 3. stable ranking from the scanned bases: positions in the tile's digit-ordered staging buffer
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const uint32_t i = base + (uint32_t)r * 32 + lane;
    const bool valid = i < n && !(kFirstDepth && key[r] == kInvalidKey);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    pos16[r] = 0xFFFFu;
    uint32_t d = 0, before = 0;
    unsigned m = 0;
    if (valid) {
      d = (key[r] >> shift) & digit_mask;
      m = digit_peers(d, nbits, vmask);
      before = warp_hist[warp][d];
    }
    __syncwarp();   // every lane: the counters read above are rewritten below, and read again next round
    if (valid && (m & ((1u << lane) - 1)) == 0) warp_hist[warp][d] = before + __popc(m);
    __syncwarp();
    if (valid) {
      const uint32_t pos = before + __popc(m & ((1u << lane) - 1));
      st_keys[pos] = key[r];
      pos16[r] = (uint16_t)pos;
    }
  }
  // 4. values
  {
    uint32_t val[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const uint32_t i = base + (uint32_t)r * 32 + lane;
      val[r] = kFirstDepth ? i : (i < n ? vals_in[i] : 0u);
    }
#pragma unroll
    for (int r = 0; r < kItems; ++r)
      if (pos16[r] != 0xFFFFu) st_vals[pos16[r]] = val[r];
  }
  // 5. decoupled look-back, last
  if (tid < kBins) {
    uint32_t excl = 0;
    if (tile != 0) {
      long long t = (long long)tile - 1;
      while (true) {
        const uint32_t v = status[(size_t)t * kBins + tid];
        const uint32_t f = v & ~kValMask;
        if (f == 0u) continue;              // predecessor has not published yet: spin
        excl += v & kValMask;
        if (f == kFlagIncl) break;
        --t;
      }
      *mine = kFlagIncl | (excl + acc);
    }
    gbase[tid] = dstart + excl;
  }
  __syncthreads();

  for (uint32_t i = tid; i < tile_total; i += kThreads) {
    const uint32_t k = st_keys[i];
    const uint32_t d = (k >> shift) & digit_mask;
    const size_t dst = (size_t)gbase[d] + (i - bstart[d]);
    keys_out[dst] = k;
    vals_out[dst] = st_vals[i];
  }
}

template <bool kFirstDepth, int kThreads, int kItems>
cudaError_t launch_onesweep_pass(unsigned tiles, cudaStream_t stream, const uint32_t* kin,
                                 const uint32_t* vin, uint32_t* kout, uint32_t* vout, uint32_t n_max,
                                 const uint32_t* n_ptr, int shift, uint32_t mask, int bits,
                                 const uint32_t* ghist, uint32_t* st, uint32_t* ticket,
                                 uint32_t* n_out) {
  constexpr int smem = ((kThreads / 32) * kBins + 2 * kBins + 16 + 2 * kThreads * kItems) * 4;
  static std::atomic<unsigned long long> configured{0ull};   // one per template instance
  cudaError_t e = gcr_set_dynamic_smem_once(onesweep_pass_kernel<kFirstDepth, kThreads, kItems>, smem, configured);
  if (e != cudaSuccess) return e;
  return gcr_launch_chain(onesweep_pass_kernel<kFirstDepth, kThreads, kItems>, dim3(tiles), dim3(kThreads), smem, stream,
                          kin, vin, kout, vout, n_max, n_ptr, shift, mask, bits, ghist, st, ticket, n_out);
}

constexpr size_t kSmallTileMaxN = 262144;   // up to here the 1024-pair tiles are used
inline bool small_tiles_for(size_t n_max) { return n_max <= kSmallTileMaxN; }
inline size_t sort_tiles_for(size_t n_max) {
  return small_tiles_for(n_max) ? (n_max + 1023) / 1024 : (n_max + 4095) / 4096;
}

// ------------------------------------------------------------------------------------------
// Fused scan + emission (duplicateWithKeys, rasterizer_impl.cu:66-99, with the inclusive scan of
// :228-231 folded in).  A CTA takes chunks of 256 depth-ordered Gaussians from an atomic ticket,
// scans their tile counts, obtains the chunk's global offset by a decoupled look-back over 64-bit
// status words, and expands the chunk: each warp's 32 Gaussians occupy ONE contiguous output
// span, so lanes stride over the span's elements, find the owning Gaussian by a shuffle binary
// search over the lanes' offsets, and write fully coalesced (the per-thread loop of the
// reference writes 32 scattered runs per instruction).  Row-major over the tile rect, clipped to
// the tile-row stripe [row0,row1) this rank owns.  The digit histograms of the tile sort are
// accumulated on the way (shared memory, flushed once per CTA), the last chunk publishes R.
// ------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;
// (macros: 64-bit namespace-scope constexprs are not usable as device-side lvalues)
#define kEmitAgg (1ull << 62)
#define kEmitIncl (2ull << 62)
#define kEmitVal ((1ull << 62) - 1ull)

struct EmitArgs {
  uint32_t n_max;                 // upper bound of visible Gaussians (P)
  const uint32_t* n_vis;          // device count
  const uint32_t* sorted_gauss;
  const uint2* rects;             // x0 | y0 << 16, width | rows << 16 (stripe-clipped)
  int grid_x;
  uint32_t* tile_keys;
  uint32_t* gauss_vals;
  uint32_t cap;                   // capacity of tile_keys / gauss_vals
  volatile uint64_t* status;      // [ceil(n_max/256)], zeroed
  uint32_t* ticket;               // zeroed
  uint32_t* ghist_tile;           // [kMaxPasses][kBins], zeroed
  int tile_end_bit;
  uint32_t* counters;             // GcrCounters
  uint32_t* offsets_out;          // optional (debug): inclusive offsets in depth order
};

__global__ void __launch_bounds__(kEmitThreads)
emit_scan_kernel(EmitArgs a) {
  __shared__ uint32_t hist[kMaxPasses][kBins];
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t s_chunk;
  __shared__ uint64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npass = (a.tile_end_bit + 7) / 8;
  gcr_pdl_wait();
  gcr_pdl_trigger();
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p) hist[p][tid] = 0;
  const uint32_t n_vis = load_count(a.n_vis, a.n_max);
  const uint32_t nchunks = (n_vis + kEmitThreads - 1) / kEmitThreads;
  __syncthreads();

  while (true) {
    if (tid == 0) s_chunk = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    if (chunk >= nchunks) {
      if (chunk == 0 && tid == 0) {   // nothing visible at all
        a.counters[GCR_CNT_R] = 0;
        a.counters[GCR_CNT_OVERFLOW] = 0;
      }
      break;
    }
    const uint32_t i = chunk * kEmitThreads + tid;
    uint32_t g = 0, n = 0, x0 = 0, w = 1, y0 = 0;
    if (i < n_vis) {
      g = a.sorted_gauss[i];
      const uint2 rc = a.rects[g];   // one 8-byte gather per Gaussian (every listed one has tiles)
      x0 = rc.x & 0xFFFFu;
      y0 = rc.x >> 16;
      w = rc.y & 0xFFFFu;
      n = w * (rc.y >> 16);
    }
    // chunk-local inclusive scan of the tile counts
    const uint32_t winc = warp_incl_scan(n, lane);
    if (lane == 31) wsum[warp] = winc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t t = wsum[k];
      if (k < warp) wbase += t;
      total += t;
    }
    // publish the aggregate, then look back (warp 0, 32 predecessors at a time)
    if (warp == 0) {
      if (lane == 0) a.status[chunk] = (chunk == 0 ? kEmitIncl : kEmitAgg) | (uint64_t)total;
      uint64_t excl = 0;
      if (chunk != 0) {
        long long t = (long long)chunk - 1;
        while (true) {
          const long long idx = t - lane;
          uint64_t v;
          do {
            v = idx >= 0 ? a.status[idx] : kEmitIncl;   // before the first chunk: inclusive 0
          } while (__any_sync(0xffffffffu, (v & ~kEmitVal) == 0ull));
          const unsigned incl = __ballot_sync(0xffffffffu, (v & ~kEmitVal) == kEmitIncl);
          // lanes 0..first carry the contributions (nearest predecessors first)
          const int first = __ffs(incl) - 1;   // incl != 0 once idx < 0 is reached at the latest
          uint64_t c = (incl == 0u || lane <= first) ? (v & kEmitVal) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
          excl += c;
          if (incl != 0u) break;
          t -= 32;
        }
        if (lane == 0) a.status[chunk] = kEmitIncl | (excl + total);
      }
      if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const uint64_t excl64 = s_excl;
    if (chunk == nchunks - 1 && tid == 0) {
      const uint64_t R = excl64 + total;
      a.counters[GCR_CNT_R] = (uint32_t)min(R, (uint64_t)0xFFFFFFFFu);
      a.counters[GCR_CNT_OVERFLOW] = R > (uint64_t)a.cap ? 1u : 0u;
    }
    // saturate at 2^32-1: anything above the capacity is dropped below and reported as overflow
    const uint64_t end64 = excl64 + wbase + winc;
    uint32_t end = (uint32_t)min(end64, (uint64_t)0xFFFFFFFFu);
    if (a.offsets_out != nullptr && i < n_vis) a.offsets_out[i] = end;

    // ---- expansion: lanes past n_vis repeat the last end so offsets stay non-decreasing ----
    const uint32_t warp_first = chunk * kEmitThreads + warp * 32;
    const int last_lane = (int)min(31u, n_vis > warp_first ? n_vis - 1 - warp_first : 0u);
    const uint32_t last_end = __shfl_sync(0xffffffffu, end, last_lane);
    if (i >= n_vis) end = last_end;
    const uint32_t start = end - n;
    // instances at or beyond the binning capacity are dropped (reported through GCR_CNT_OVERFLOW)
    const uint32_t span0 = min(__shfl_sync(0xffffffffu, start, 0), a.cap);
    const uint32_t span1 = min(__shfl_sync(0xffffffffu, end, 31), a.cap);
    for (uint32_t eb = span0; eb < span1; eb += 32) {   // warp-uniform trip count
      const bool act = eb + lane < span1;
      const uint32_t e = act ? eb + lane : span1 - 1;
      // owner = first lane whose inclusive end exceeds e  (lanes with n == 0 have end == start)
      int lo = 0, hi = 31;
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int mid = (lo + hi) >> 1;
        const uint32_t em = __shfl_sync(0xffffffffu, end, mid);
        if (em > e) hi = mid; else lo = mid + 1;
      }
      const int own = lo;
      const uint32_t os = __shfl_sync(0xffffffffu, start, own);
      const uint32_t og = __shfl_sync(0xffffffffu, g, own);
      const uint32_t ox0 = __shfl_sync(0xffffffffu, x0, own);
      const uint32_t ow = __shfl_sync(0xffffffffu, w, own);
      const uint32_t oy0 = __shfl_sync(0xffffffffu, y0, own);
      const uint32_t k = e - os;
      const uint32_t ry = k / ow, rx = k - ry * ow;
      const uint32_t key = (oy0 + ry) * (uint32_t)a.grid_x + ox0 + rx;
      const bool wr = act;
      if (wr) {
        a.tile_keys[e] = key;
        a.gauss_vals[e] = og;
        atomicAdd(&hist[0][key & (a.tile_end_bit >= 8 ? 0xFFu : ((1u << a.tile_end_bit) - 1u))], 1u);
      }
      // higher digits: a warp's 32 consecutive instances share one or two values -> one
      // shared-memory atomic per distinct value (MATCH.ANY is fast at this entropy)
      const unsigned wmask = __ballot_sync(0xffffffffu, wr);
      for (int p = 1; p < npass; ++p) {
        if (wr) {
          const int bits = min(8, a.tile_end_bit - 8 * p);
          const uint32_t d = (key >> (8 * p)) & ((1u << bits) - 1u);
          const unsigned m = __match_any_sync(wmask, d);
          if ((m & ((1u << lane) - 1)) == 0) atomicAdd(&hist[p][d], (uint32_t)__popc(m));
        }
      }
    }
    __syncthreads();   // s_chunk / s_excl / wsum are rewritten by the next chunk
  }
  __syncthreads();
  for (int p = 0; p < npass; ++p)
    if (hist[p][tid] != 0) atomicAdd(&a.ghist_tile[p * kBins + tid], hist[p][tid]);
}

// identifyTileRanges (rasterizer_impl.cu:104-124) on the sorted tile ids; `ranges` was zeroed.
__global__ void __launch_bounds__(256)
tile_ranges_kernel(uint32_t n_max, const uint32_t* __restrict__ n_ptr,
                   const uint32_t* __restrict__ keys, uint2* __restrict__ ranges) {
  gcr_pdl_wait();
  gcr_pdl_trigger();
  const uint32_t R = load_count(n_ptr, n_max);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += stride) {
    const uint32_t cur = keys[i];
    if (i == 0) {
      ranges[cur].x = 0;
    } else {
      const uint32_t prev = keys[i - 1];
      if (cur != prev) {
        ranges[prev].y = i;
        ranges[cur].x = i;
      }
    }
    if (i == R - 1) ranges[cur].y = R;
  }
}

}  // namespace

// ---- workspace layouts (uint32 words) ---------------------------------------------------------
// sort:  ghist [4][256] | tickets [4] (+pad to 64) | status [passes][tiles][256]
size_t gcr_sort_workspace_bytes(size_t n_max) {
  const size_t tiles = sort_tiles_for(n_max);
  return gcr_align_up((kMaxPasses * kBins + 64 + kMaxPasses * tiles * kBins) * sizeof(uint32_t), 256);
}
// emit:  ticket (+pad to 64 words) | status64 [chunks]
size_t gcr_emit_workspace_bytes(size_t n_max) {
  const size_t chunks = (n_max + kEmitThreads - 1) / kEmitThreads;
  return gcr_align_up(64 * sizeof(uint32_t) + (chunks + 1) * sizeof(uint64_t), 256);
}
uint32_t* gcr_sort_ghist(void* sort_workspace) { return static_cast<uint32_t*>(sort_workspace); }

cudaError_t gcr_launch_radix_sort(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                                  uint32_t* vals_out, size_t n_max, const uint32_t* n_ptr,
                                  int end_bit, bool depth_mode, bool hist_done, uint32_t* n_out,
                                  void* workspace, cudaStream_t stream) {
  if (n_max == 0) return cudaSuccess;
  if (n_max >= (size_t)kValMask) return cudaErrorInvalidValue;   // 30-bit look-back counters
  if (end_bit <= 0) end_bit = 1;  // at least one pass so values are materialised
  if (end_bit > 8 * kMaxPasses) return cudaErrorInvalidValue;
  uint32_t* ws = static_cast<uint32_t*>(workspace);
  uint32_t* ghist = ws;
  uint32_t* tickets = ws + kMaxPasses * kBins;
  uint32_t* status = ws + kMaxPasses * kBins + 64;
  const int npass = (end_bit + 7) / 8;
  const bool small_tiles = small_tiles_for(n_max);
  const size_t ntiles = sort_tiles_for(n_max);
  const uint32_t n32 = (uint32_t)n_max;
  if (!hist_done) {
    const size_t chunks = (n_max + 4095) / 4096;
    const unsigned hgrid = (unsigned)(chunks < (size_t)148 * 8 ? chunks : (size_t)148 * 8);
    cudaError_t e = gcr_launch_chain(radix_hist_all_kernel, dim3(hgrid), dim3(kSortThreads), 0, stream, keys_in, n32,
                                     n_ptr, end_bit, depth_mode, ghist);
    if (e != cudaSuccess) return e;
  }
  uint32_t* kin = keys_in; uint32_t* vin = vals_in; uint32_t* kout = keys_out; uint32_t* vout = vals_out;
  for (int p = 0; p < npass; ++p) {
    const int shift = 8 * p;
    const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    const uint32_t mask = (1u << bits) - 1u;
    uint32_t* st = status + (size_t)p * ntiles * kBins;
    const bool first_depth = depth_mode && p == 0;
    // after the compacting first pass of the depth sort the element count is the device-side n_out
    const uint32_t* np = (depth_mode && p > 0) ? n_out : n_ptr;
    uint32_t* no = (p == 0) ? n_out : nullptr;
    cudaError_t e;
    const unsigned tiles = (unsigned)ntiles;
    if (small_tiles) {
      e = first_depth ? launch_onesweep_pass<true, 256, 4>(tiles, stream, kin, vin, kout, vout, n32, np, shift, mask, bits, ghist + p * kBins, st, tickets + p, no)
                      : launch_onesweep_pass<false, 256, 4>(tiles, stream, kin, vin, kout, vout, n32, np, shift, mask, bits, ghist + p * kBins, st, tickets + p, no);
    } else {
      e = first_depth ? launch_onesweep_pass<true, 256, 16>(tiles, stream, kin, vin, kout, vout, n32, np, shift, mask, bits, ghist + p * kBins, st, tickets + p, no)
                      : launch_onesweep_pass<false, 256, 16>(tiles, stream, kin, vin, kout, vout, n32, np, shift, mask, bits, ghist + p * kBins, st, tickets + p, no);
    }
    if (e != cudaSuccess) return e;
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  return cudaSuccess;
}

cudaError_t gcr_launch_emit_scan(const GcrEmitLaunch& l, cudaStream_t stream) {
  if (l.n_max == 0) return cudaSuccess;
  EmitArgs a;
  a.n_max = l.n_max; a.n_vis = l.n_vis; a.sorted_gauss = l.sorted_gauss;
  a.rects = l.rects; a.grid_x = l.grid_x;
  a.tile_keys = l.tile_keys; a.gauss_vals = l.gauss_vals; a.cap = l.cap;
  uint32_t* ws = static_cast<uint32_t*>(l.workspace);
  a.ticket = ws;
  a.status = reinterpret_cast<volatile uint64_t*>(ws + 64);
  a.ghist_tile = l.ghist_tile; a.tile_end_bit = l.tile_end_bit <= 0 ? 1 : l.tile_end_bit;
  a.counters = l.counters; a.offsets_out = l.offsets_out;
  const unsigned chunks = (l.n_max + kEmitThreads - 1) / kEmitThreads;
  const unsigned grid = chunks < 148u * 8u ? chunks : 148u * 8u;   // persistent: 8 CTAs per SM
  return gcr_launch_chain(emit_scan_kernel, dim3(grid), dim3(kEmitThreads), 0, stream, a);
}

cudaError_t gcr_launch_tile_ranges(uint32_t n_max, const uint32_t* n_ptr, const uint32_t* sorted_keys,
                                   uint2* ranges, cudaStream_t stream) {
  if (n_max == 0) return cudaSuccess;
  const unsigned blocks = (n_max + 255) / 256;
  return gcr_launch_chain(tile_ranges_kernel, dim3(blocks < 148u * 16u ? blocks : 148u * 16u), dim3(256), 0, stream,
                          n_max, n_ptr, sorted_keys, ranges);
}
