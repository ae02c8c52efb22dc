// binning.cu -- scan, cub-free stable radix sort, tile-instance emission, tile ranges and the
// per-instance record gather, for sm_100a.
//
// Behavioural spec (reference): cub::DeviceScan::InclusiveSum + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs(64-bit tile|depth keys) + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:66-124, 228-270).
//
// B200-first redesign: the reference sorts R tile-instances on 32+log2(T) key bits (6 passes of
// 24 B/pair at 1080p).  Here the P Gaussians are depth-sorted ONCE (4 passes over P 8-byte
// pairs), instances are emitted in that order, and a stable split on the tile id (2 passes over
// R 8-byte pairs at 1080p) finishes the job.  A stable sort on (tile, depth) of index-ordered
// input equals a stable depth sort followed by a stable tile sort, so point_list and ranges are
// bit-identical to the reference's, ties included, at ~1/3 of the HBM traffic.
#include <cstdlib>

#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix,
// *total receives the block sum. smem: 8 words + 1.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* smem, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < 8) ? smem[lane] : 0;
    uint32_t winc = warp_incl_scan(w, lane);
    if (lane < 8) smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 7) smem[8] = winc;
  }
  __syncthreads();
  uint32_t res = inc - v + smem[warp];
  *total = smem[8];
  __syncthreads();
  return res;
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

__device__ __forceinline__ uint32_t scan_load(const uint32_t* __restrict__ in,
                                              const uint32_t* __restrict__ gather, size_t i,
                                              size_t n) {
  if (i >= n) return 0;
  return gather ? in[gather[i]] : in[i];
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather, size_t n,
                   uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sm[16];
  const size_t base = (size_t)blockIdx.x * kScanChunk;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    s += scan_load(in, gather, base + (size_t)k * kScanThreads + threadIdx.x, n);
  uint32_t total;
  block_excl_scan_256(s, sm, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block sums, in place
__global__ void __launch_bounds__(kScanThreads)
scan_spine_kernel(uint32_t* __restrict__ block_sums, size_t nb) {
  __shared__ uint32_t sm[16];
  uint32_t carry = 0;
  for (size_t base = 0; base < nb; base += kScanThreads) {
    const size_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? block_sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_excl_scan_256(v, sm, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads)
scan_final_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather, size_t n,
                  const uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ out) {
  __shared__ uint32_t sm[16];
  // blocked arrangement: thread t owns items [t*8, t*8+8) of the chunk
  const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = scan_load(in, gather, base + k, n);
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan_256(s, sm, &total) + block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    run += v[k];
    if (base + k < n) out[base + k] = run;
  }
}

// ------------------------------------------------------------------------------------------
// Radix sort: 8-bit digits, per pass  histogram -> per-digit scan over blocks -> stable scatter.
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortChunk = kSortThreads * kSortItems;  // 4096
constexpr int kBins = 256;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t digit_mask,
                  int nbits, uint32_t* __restrict__ table, uint32_t nblk) {
  __shared__ uint32_t hist[kBins];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * kSortChunk + (size_t)warp * (32 * kSortItems);
  // issue all loads first: ballots are convergence points the compiler does not hoist loads over
  uint32_t key[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    key[r] = i < n ? keys[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const unsigned m = digit_peers(d, nbits, vmask);
      if ((m & ((1u << lane) - 1)) == 0) atomicAdd(&hist[d], (uint32_t)__popc(m));
    }
  }
  __syncthreads();
  table[(size_t)threadIdx.x * nblk + blockIdx.x] = hist[threadIdx.x];
}

// one block per digit: exclusive scan of its row over blocks, in place; totals[d] = row sum
__global__ void __launch_bounds__(kSortThreads)
radix_scan_rows_kernel(uint32_t* __restrict__ table, uint32_t nblk, uint32_t* __restrict__ totals) {
  __shared__ uint32_t sm[16];
  uint32_t* row = table + (size_t)blockIdx.x * nblk;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblk; base += kSortThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nblk ? row[i] : 0;
    uint32_t total;
    const uint32_t ex = block_excl_scan_256(v, sm, &total);
    if (i < nblk) row[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

template <bool kIota>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, uint32_t digit_mask, int nbits, const uint32_t* __restrict__ table,
                     uint32_t nblk, const uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_hist[8][kBins];   // per-warp digit counts -> per-warp prefixes
  __shared__ uint32_t gbase[kBins];          // global output base of this block's digit run
  __shared__ uint32_t bstart[kBins];         // start of the digit run inside the block staging
  __shared__ uint32_t sm[16];
  __shared__ uint32_t st_keys[kSortChunk];
  __shared__ uint32_t st_vals[kSortChunk];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int w = 0; w < 8; ++w) warp_hist[w][tid] = 0;
  __syncthreads();

  const size_t chunk_base = (size_t)blockIdx.x * kSortChunk;
  const size_t base = chunk_base + (size_t)warp * (32 * kSortItems);
  uint32_t key[kSortItems], val[kSortItems];
  uint16_t rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {   // all loads in flight before the first ballot
    const size_t i = base + (size_t)r * 32 + lane;
    key[r] = i < n ? keys_in[i] : 0u;
    val[r] = kIota ? (uint32_t)i : (i < n ? vals_in[i] : 0u);
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const unsigned m = digit_peers(d, nbits, vmask);
      const uint32_t before = warp_hist[warp][d];
      rank[r] = (uint16_t)(before + __popc(m & ((1u << lane) - 1)));
      __syncwarp(vmask);
      if ((m & ((1u << lane) - 1)) == 0) warp_hist[warp][d] = before + __popc(m);
      __syncwarp(vmask);
    }
  }
  __syncthreads();

  // digit `tid`: prefix over warps, block total, global base
  {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t t = warp_hist[w][tid];
      warp_hist[w][tid] = acc;
      acc += t;
    }
    uint32_t dummy;
    const uint32_t bs = block_excl_scan_256(acc, sm, &dummy);
    bstart[tid] = bs;
    const uint32_t dstart = block_excl_scan_256(totals[tid], sm, &dummy);
    gbase[tid] = dstart + table[(size_t)tid * nblk + blockIdx.x];
  }
  __syncthreads();

#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const uint32_t pos = bstart[d] + warp_hist[warp][d] + rank[r];
      st_keys[pos] = key[r];
      st_vals[pos] = val[r];
    }
  }
  __syncthreads();

  const uint32_t count = (uint32_t)min((size_t)kSortChunk, n - chunk_base);
  for (uint32_t i = tid; i < count; i += kSortThreads) {
    const uint32_t k = st_keys[i];
    const uint32_t d = (k >> shift) & digit_mask;
    const size_t dst = (size_t)gbase[d] + (i - bstart[d]);
    keys_out[dst] = k;
    vals_out[dst] = st_vals[i];
  }
}

// ------------------------------------------------------------------------------------------
// Onesweep-style passes (Adinets & Merrill): ONE histogram kernel computes the global digit
// histograms of every pass (they are permutation invariant), then each pass is a single kernel
// whose blocks obtain their output offsets with a chained scan / decoupled look-back over a
// per-tile status word instead of a separate histogram + scan launch.  16 B/pair/pass of HBM
// traffic and 1 + passes launches per sort (was 20 B and 3 x passes).  Tiles are handed out by
// an atomic ticket so a tile's predecessors have always started: the look-back cannot deadlock.
// ------------------------------------------------------------------------------------------
constexpr size_t kSmallTileMaxN = 262144;   // up to here the 1024-pair tiles are used
constexpr uint32_t kFlagAgg = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1u;
constexpr int kMaxPasses = 4;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_all_kernel(const uint32_t* __restrict__ keys, size_t n, int end_bit,
                      uint32_t* __restrict__ ghist /* [kMaxPasses][kBins] */) {
  __shared__ uint32_t hist[kMaxPasses][kBins];
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
  const int npass = (end_bit + 7) / 8;
  const size_t nchunks = (n + kSortChunk - 1) / kSortChunk;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const size_t base = c * kSortChunk + threadIdx.x;
    uint32_t key[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const size_t i = base + (size_t)r * kSortThreads;
      key[r] = i < n ? keys[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const size_t i = base + (size_t)r * kSortThreads;
      if (i < n) {
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

// exclusive scan of one value per DIGIT (threads 0..255 carry values, any others pass 0)
__device__ __forceinline__ uint32_t digit_excl_scan(uint32_t v, uint32_t* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31 && warp < 8) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = (lane < 8) ? smem[lane] : 0;
    const uint32_t winc = warp_incl_scan(w, lane);
    if (lane < 8) smem[lane] = winc - w;
  }
  __syncthreads();
  const uint32_t res = inc - v + ((warp < 8) ? smem[warp] : 0u);
  __syncthreads();
  return res;
}

// kThreads x kItems pairs per tile (256 x 16 = 4096 by default; 256 x 4 = 1024 for small inputs,
// where the 16 serial ranking rounds of a few lonely CTAs are the whole latency); all shared
// memory dynamic.
template <bool kIota, int kThreads, int kItems>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 4 : 2)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, uint32_t digit_mask, int nbits,
                     const uint32_t* __restrict__ ghist_pass /* [kBins] */,
                     volatile uint32_t* status /* [tiles][kBins], zeroed */, uint32_t* ticket) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kTile = kThreads * kItems;
  extern __shared__ __align__(16) uint32_t os_smem[];
  uint32_t (*warp_hist)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(os_smem);
  uint32_t* gbase = os_smem + kWarps * kBins;
  uint32_t* bstart = gbase + kBins;
  uint32_t* sm = bstart + kBins;          // 16 words
  uint32_t* st_keys = sm + 16;
  uint32_t* st_vals = st_keys + kTile;
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kWarps * kBins; i += kThreads) (&warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;

  const size_t chunk_base = (size_t)tile * kTile;
  const size_t base = chunk_base + (size_t)warp * (32 * kItems);
  // keys stay in registers across the ranking; values are fetched only when they are staged
  // (keeps the kernel at <= 64 registers; it is bandwidth/latency bound)
  uint32_t key[kItems];
  uint16_t rank[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    key[r] = i < n ? keys_in[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const unsigned m = digit_peers(d, nbits, vmask);
      const uint32_t before = warp_hist[warp][d];
      rank[r] = (uint16_t)(before + __popc(m & ((1u << lane) - 1)));
      __syncwarp(vmask);
      if ((m & ((1u << lane) - 1)) == 0) warp_hist[warp][d] = before + __popc(m);
      __syncwarp(vmask);
    }
  }
  __syncthreads();

  {
    uint32_t acc = 0;   // this tile's count of digit `tid` (threads >= kBins idle here)
    if (tid < kBins) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t t = warp_hist[w][tid];
        warp_hist[w][tid] = acc;
        acc += t;
      }
    }
    const uint32_t bs = digit_excl_scan(acc, sm);
    const uint32_t dstart = digit_excl_scan(tid < kBins ? ghist_pass[tid] : 0u, sm);
    if (tid < kBins) {
      bstart[tid] = bs;
      // decoupled look-back over the predecessors' status words for this digit
      volatile uint32_t* mine = status + (size_t)tile * kBins + tid;
      uint32_t excl = 0;
      if (tile == 0) {
        *mine = kFlagIncl | acc;
      } else {
        *mine = kFlagAgg | acc;
        // (an 8-wide batched walk was measured: no faster -- the wait is for predecessors to
        // publish, not for the L2 round trips of the walk itself)
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
  }
  __syncthreads();

  {
    uint32_t val[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const size_t i = base + (size_t)r * 32 + lane;
      val[r] = kIota ? (uint32_t)i : (i < n ? vals_in[i] : 0u);
    }
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const size_t i = base + (size_t)r * 32 + lane;
      if (i < n) {
        const uint32_t d = (key[r] >> shift) & digit_mask;
        const uint32_t pos = bstart[d] + warp_hist[warp][d] + rank[r];
        st_keys[pos] = key[r];
        st_vals[pos] = val[r];
      }
    }
  }
  __syncthreads();

  const uint32_t count = (uint32_t)min((size_t)kTile, n - chunk_base);
  for (uint32_t i = tid; i < count; i += kThreads) {
    const uint32_t k = st_keys[i];
    const uint32_t d = (k >> shift) & digit_mask;
    const size_t dst = (size_t)gbase[d] + (i - bstart[d]);
    keys_out[dst] = k;
    vals_out[dst] = st_vals[i];
  }
}

// Early-counts order (experimental, env GCR_SORT_ORDER=counts; not yet measured on hardware).
// The ncu source view of the default order (profiles/r01_ncu_source_hotspots.md) puts ~30 % of the
// samples in the look-back spin and ~12 % on the value loads issued behind it: a tile only publishes its digit counts after the 16 ballot-ranking rounds,
// so ranking sits on the inter-tile dependency chain (the look-back of tile t waits for tile
// t-1's ranking).  Here the counts come first -- one shared-memory atomic per key, no order
// needed -- and the aggregate is published before any ranking; the ranking then starts from the
// scanned per-warp bases and yields tile positions directly (keys go to the staging buffer inside
// the ranking loop), the values follow, and the look-back runs last, when the predecessors have
// long published.  Same outputs as onesweep_pass_kernel (stable within tile and across tiles).
// `spin_ns` > 0 (env GCR_SORT_SPIN_NS) backs the remaining spin off with __nanosleep.
template <bool kIota, int kThreads, int kItems>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 4 : 2)
onesweep_pass_counts_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                            int shift, uint32_t digit_mask, int nbits,
                            const uint32_t* __restrict__ ghist_pass /* [kBins] */,
                            volatile uint32_t* status /* [tiles][kBins], zeroed */, uint32_t* ticket,
                            unsigned spin_ns) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kTile = kThreads * kItems;
  extern __shared__ __align__(16) uint32_t os_smem[];
  uint32_t (*warp_hist)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(os_smem);
  uint32_t* gbase = os_smem + kWarps * kBins;
  uint32_t* bstart = gbase + kBins;
  uint32_t* sm = bstart + kBins;          // 16 words
  uint32_t* st_keys = sm + 16;
  uint32_t* st_vals = st_keys + kTile;
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kWarps * kBins; i += kThreads) (&warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;

  const size_t chunk_base = (size_t)tile * kTile;
  const size_t base = chunk_base + (size_t)warp * (32 * kItems);
  uint32_t key[kItems];
  uint16_t pos16[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    key[r] = i < n ? keys_in[i] : 0u;
  }
  // 1. counts of this warp's keys per digit (order-free)
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    if (i < n) atomicAdd(&warp_hist[warp][(key[r] >> shift) & digit_mask], 1u);
  }
  __syncthreads();

  // 2. tile counts -> publish; per-warp starting positions inside the tile
  uint32_t acc = 0, dstart = 0;
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
    const uint32_t bs = digit_excl_scan(acc, sm);
    dstart = digit_excl_scan(tid < kBins ? ghist_pass[tid] : 0u, sm);
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
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    pos16[r] = 0;
    if (valid) {
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const unsigned m = digit_peers(d, nbits, vmask);
      const uint32_t before = warp_hist[warp][d];
      const uint32_t pos = before + __popc(m & ((1u << lane) - 1));
      __syncwarp(vmask);
      if ((m & ((1u << lane) - 1)) == 0) warp_hist[warp][d] = before + __popc(m);
      __syncwarp(vmask);
      st_keys[pos] = key[r];
      pos16[r] = (uint16_t)pos;
    }
  }
  // 4. values
  {
    uint32_t val[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const size_t i = base + (size_t)r * 32 + lane;
      val[r] = kIota ? (uint32_t)i : (i < n ? vals_in[i] : 0u);
    }
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      const size_t i = base + (size_t)r * 32 + lane;
      if (i < n) st_vals[pos16[r]] = val[r];
    }
  }
  // 5. decoupled look-back, last
  if (tid < kBins) {
    uint32_t excl = 0;
    if (tile != 0) {
      long long t = (long long)tile - 1;
      while (true) {
        const uint32_t v = status[(size_t)t * kBins + tid];
        const uint32_t f = v & ~kValMask;
        if (f == 0u) {
          if (spin_ns) __nanosleep(spin_ns);
          continue;
        }
        excl += v & kValMask;
        if (f == kFlagIncl) break;
        --t;
      }
      *mine = kFlagIncl | (excl + acc);
    }
    gbase[tid] = dstart + excl;
  }
  __syncthreads();

  const uint32_t count = (uint32_t)min((size_t)kTile, n - chunk_base);
  for (uint32_t i = tid; i < count; i += kThreads) {
    const uint32_t k = st_keys[i];
    const uint32_t d = (k >> shift) & digit_mask;
    const size_t dst = (size_t)gbase[d] + (i - bstart[d]);
    keys_out[dst] = k;
    vals_out[dst] = st_vals[i];
  }
}

struct SortTuning { bool counts; unsigned spin_ns; };
static SortTuning sort_tuning() {
  static const SortTuning t = [] {
    SortTuning r{false, 0u};
    const char* o = getenv("GCR_SORT_ORDER");
    r.counts = o != nullptr && o[0] == 'c';
    const char* s = getenv("GCR_SORT_SPIN_NS");
    if (s != nullptr) r.spin_ns = (unsigned)max(0, atoi(s));
    return r;
  }();
  return t;
}

template <bool kIota, int kThreads, int kItems>
void launch_onesweep_pass(unsigned tiles, cudaStream_t stream, const uint32_t* kin, const uint32_t* vin,
                          uint32_t* kout, uint32_t* vout, size_t n, int shift, uint32_t mask, int bits,
                          const uint32_t* ghist, uint32_t* st, uint32_t* ticket) {
  constexpr int smem = ((kThreads / 32) * kBins + 2 * kBins + 16 + 2 * kThreads * kItems) * 4;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(onesweep_pass_kernel<kIota, kThreads, kItems>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(onesweep_pass_counts_kernel<kIota, kThreads, kItems>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const SortTuning t = sort_tuning();
  if (t.counts)
    onesweep_pass_counts_kernel<kIota, kThreads, kItems><<<tiles, kThreads, smem, stream>>>(
        kin, vin, kout, vout, n, shift, mask, bits, ghist, st, ticket, t.spin_ns);
  else
    onesweep_pass_kernel<kIota, kThreads, kItems><<<tiles, kThreads, smem, stream>>>(kin, vin, kout, vout, n, shift, mask, bits, ghist, st, ticket);
}

// ------------------------------------------------------------------------------------------
// Small inputs (GaussianCity's own regime: P <= 16 384 points per frame, R of a few 10^4):
// the multi-kernel sort above is launch-latency bound (3 launches per digit pass), so all digit
// passes run inside ONE single-CTA kernel: 32 warps, chunks of 16 384 items, the same stable
// peer-mask ranking, ping-pong through global memory (L2 resident at this size).
// MEASURED SLOWER than the multi-block path on B200 -> disabled by default (see below).
// ------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 1024;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallRounds = 16;
constexpr int kSmallChunk = kSmallThreads * kSmallRounds;  // 16384
constexpr int kSmallMaxChunks = 4;
constexpr size_t kSmallSortMaxN = (size_t)kSmallChunk * kSmallMaxChunks;  // 65536
// Measured on B200 (profiles/r01_small_scenes.md): a single SM is slower than the multi-block
// path even at 16 k items (one CTA cannot hide its own L2 round trips), so the single-CTA kernels
// are kept for reference but disabled; env GCR_SMALL_SORT=1 re-enables them for experiments.
static bool small_paths_enabled() {
  static const bool on = getenv("GCR_SMALL_SORT") != nullptr;
  return on;
}
constexpr int kSmallSortSmem = (kSmallWarps * kBins + kSmallMaxChunks * kBins + kBins) * 4 + kSmallChunk * 2;

template <bool kIotaFirst>
__global__ void __launch_bounds__(kSmallThreads)
small_sort_kernel(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n,
                  int end_bit) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  uint32_t (*warp_hist)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(sm_raw);
  uint32_t (*chunk_base)[kBins] = reinterpret_cast<uint32_t (*)[kBins]>(sm_raw + kSmallWarps * kBins * 4);
  uint32_t* digit_tot = reinterpret_cast<uint32_t*>(sm_raw + (kSmallWarps + kSmallMaxChunks) * kBins * 4);
  uint16_t* ranks = reinterpret_cast<uint16_t*>(sm_raw + (kSmallWarps + kSmallMaxChunks + 1) * kBins * 4);
  __shared__ uint32_t scan_tmp[40];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nchunks = (n + kSmallChunk - 1) / kSmallChunk;
  uint32_t* kin = keys_a; uint32_t* vin = vals_a; uint32_t* kout = keys_b; uint32_t* vout = vals_b;

  for (int shift = 0; shift < end_bit; shift += 8) {
    const int nbits = min(8, end_bit - shift);
    const uint32_t dmask = (1u << nbits) - 1u;
    // ---- sweep 1: per-chunk digit counts (skipped for a single chunk) ----
    if (nchunks > 1) {
      for (int i = tid; i < kSmallMaxChunks * kBins; i += kSmallThreads) (&chunk_base[0][0])[i] = 0;
      __syncthreads();
      for (uint32_t c = 0; c < nchunks; ++c) {
        uint32_t k1[kSmallRounds];
#pragma unroll
        for (int r = 0; r < kSmallRounds; ++r) {
          const uint32_t i = c * kSmallChunk + warp * (32 * kSmallRounds) + r * 32 + lane;
          k1[r] = i < n ? kin[i] : 0u;
        }
#pragma unroll
        for (int r = 0; r < kSmallRounds; ++r) {
          const uint32_t i = c * kSmallChunk + warp * (32 * kSmallRounds) + r * 32 + lane;
          const bool valid = i < n;
          const unsigned vmask = __ballot_sync(0xffffffffu, valid);
          if (valid) {
            const uint32_t d = (k1[r] >> shift) & dmask;
            const unsigned m = digit_peers(d, nbits, vmask);
            if ((m & ((1u << lane) - 1)) == 0) atomicAdd(&chunk_base[c][d], (uint32_t)__popc(m));
          }
        }
      }
      __syncthreads();
    }
    // ---- sweep 2: rank + scatter, chunk by chunk ----
    for (uint32_t c = 0; c < nchunks; ++c) {
      for (int i = tid; i < kSmallWarps * kBins; i += kSmallThreads) (&warp_hist[0][0])[i] = 0;
      // all loads of the chunk are issued up front (16 independent requests per thread): a
      // single CTA has no other warps to hide L2 latency behind
      uint32_t key[kSmallRounds], val[kSmallRounds];
      const uint32_t loc0 = warp * (32 * kSmallRounds) + lane;
#pragma unroll
      for (int r = 0; r < kSmallRounds; ++r) {
        const uint32_t i = c * kSmallChunk + loc0 + r * 32;
        key[r] = i < n ? kin[i] : 0u;
        val[r] = (kIotaFirst && shift == 0) ? i : (i < n ? vin[i] : 0u);
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kSmallRounds; ++r) {
        const uint32_t loc = loc0 + r * 32;
        const bool valid = c * kSmallChunk + loc < n;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          const uint32_t d = (key[r] >> shift) & dmask;
          const unsigned m = digit_peers(d, nbits, vmask);
          const uint32_t before = warp_hist[warp][d];
          ranks[loc] = (uint16_t)(before + __popc(m & ((1u << lane) - 1)));
          __syncwarp(vmask);
          if ((m & ((1u << lane) - 1)) == 0) warp_hist[warp][d] = before + __popc(m);
          __syncwarp(vmask);
        }
      }
      __syncthreads();
      // digit tid: exclusive prefix over warps; chunk total
      uint32_t tot = 0;
      if (tid < kBins) {
#pragma unroll 4
        for (int w = 0; w < kSmallWarps; ++w) {
          const uint32_t t = warp_hist[w][tid];
          warp_hist[w][tid] = tot;
          tot += t;
        }
      }
      __syncthreads();
      if (nchunks == 1 || c == 0) {
        // global digit starts: exclusive scan over digits of the per-digit totals of ALL chunks
        uint32_t all = tot;
        if (nchunks > 1 && tid < kBins) {
          all = 0;
          for (uint32_t cc = 0; cc < nchunks; ++cc) all += chunk_base[cc][tid];
        }
        uint32_t v = (tid < kBins) ? all : 0;
        uint32_t inc = warp_incl_scan(v, lane);
        if (tid < kBins && lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
          uint32_t wv = (lane < 8) ? scan_tmp[lane] : 0;
          uint32_t winc = warp_incl_scan(wv, lane);
          if (lane < 8) scan_tmp[lane] = winc - wv;
        }
        __syncthreads();
        if (tid < kBins) digit_tot[tid] = inc - v + scan_tmp[warp];   // digit start
        __syncthreads();
        if (nchunks > 1 && tid < kBins) {
          uint32_t run = digit_tot[tid];
          for (uint32_t cc = 0; cc < nchunks; ++cc) {
            const uint32_t t = chunk_base[cc][tid];
            chunk_base[cc][tid] = run;
            run += t;
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < kSmallRounds; ++r) {
        const uint32_t loc = loc0 + r * 32;
        if (c * kSmallChunk + loc < n) {
          const uint32_t d = (key[r] >> shift) & dmask;
          const uint32_t base = (nchunks > 1) ? chunk_base[c][d] : digit_tot[d];
          const uint32_t dst = base + warp_hist[warp][d] + ranks[loc];
          kout[dst] = key[r];
          vout[dst] = val[r];
        }
      }
      __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
}

// single-CTA inclusive scan (with optional gather) for n <= 65536: 16 consecutive items per
// thread, all (gathered) loads issued before the first dependent instruction
__global__ void __launch_bounds__(kSmallThreads)
small_scan_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather, uint32_t n,
                  uint32_t* __restrict__ out) {
  __shared__ uint32_t wsum[kSmallWarps];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += kSmallChunk) {
    const uint32_t i0 = base + tid * kSmallRounds;
    uint32_t v[kSmallRounds];
#pragma unroll
    for (int k = 0; k < kSmallRounds; ++k) {
      const uint32_t i = i0 + k;
      v[k] = i < n ? (gather ? in[gather[i]] : in[i]) : 0u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kSmallRounds; ++k) s += v[k];
    const uint32_t inc = warp_incl_scan(s, lane);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = wsum[lane];
      const uint32_t winc = warp_incl_scan(w, lane);
      wsum[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = inc - s + wsum[warp] + carry_s;
#pragma unroll
    for (int k = 0; k < kSmallRounds; ++k) {
      run += v[k];
      if (i0 + k < n) out[i0 + k] = run;
    }
    __syncthreads();
    if (tid == kSmallThreads - 1) carry_s = run;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// One warp expands 32 consecutive depth-ordered Gaussians: their instances occupy ONE contiguous
// output span (offsets are a scan in this order), so lanes stride over the span's elements, find
// the owning Gaussian by a shuffle binary search over the lanes' start offsets, and write fully
// coalesced (the per-thread loop of the reference, rasterizer_impl.cu:85-98, writes 32 scattered
// runs per instruction).  Row-major over the tile rect, owned tile rows only.
__global__ void __launch_bounds__(256)
emit_pairs_kernel(int P, const uint32_t* __restrict__ sorted_gauss,
                  const uint32_t* __restrict__ offsets_incl,
                  const uint32_t* __restrict__ tiles_touched,
                  const GcrRecord* __restrict__ records, const int* __restrict__ radii, int grid_x,
                  int grid_y, int shard_rank, int shard_count, uint32_t* __restrict__ tile_keys,
                  uint32_t* __restrict__ gauss_vals) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // warp covers i0 .. i0+31
  uint32_t g = 0, n = 0, end = 0, x0 = 0, w = 1, y0 = 0;
  if (i < P) {
    g = sorted_gauss[i];
    n = tiles_touched[g];
    end = offsets_incl[i];
    if (n != 0) {
      const float4 q0 = records[g].q0;
      uint2 rmin, rmax;
      gcr_get_rect(q0.x, q0.y, radii[g], grid_x, grid_y, rmin, rmax);
      x0 = rmin.x;
      w = rmax.x - rmin.x;
      // first owned tile row >= rmin.y (all rows are owned when shard_count == 1)
      y0 = rmin.y;
      if (shard_count > 1) {
        const uint32_t m = rmin.y % (uint32_t)shard_count;
        y0 = rmin.y + (((uint32_t)shard_rank + (uint32_t)shard_count - m) % (uint32_t)shard_count);
      }
    }
  }
  // inclusive end offsets are non-decreasing across lanes; lanes past P repeat the last end
  const uint32_t last_end = __shfl_sync(0xffffffffu, end, min(31, max(0, P - 1 - (i - lane))));
  if (i >= P) end = last_end;
  const uint32_t start = end - n;
  const uint32_t span0 = __shfl_sync(0xffffffffu, start, 0);
  const uint32_t span1 = __shfl_sync(0xffffffffu, end, 31);
  for (uint32_t eb = span0; eb < span1; eb += 32) {   // warp-uniform trip count (full-mask shuffles)
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
    if (act) {
      tile_keys[e] = (oy0 + ry * (uint32_t)shard_count) * (uint32_t)grid_x + ox0 + rx;
      gauss_vals[e] = og;
    }
  }
}

__global__ void __launch_bounds__(256)
ranges_gather_kernel(size_t R, const uint32_t* __restrict__ keys,
                     const uint32_t* __restrict__ point_list,
                     const GcrRecord* __restrict__ records, uint2* __restrict__ ranges,
                     GcrRecord* __restrict__ inst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t cur = keys[i];
  if (i == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = keys[i - 1];
    if (cur != prev) {
      ranges[prev].y = (uint32_t)i;
      ranges[cur].x = (uint32_t)i;
    }
  }
  if (i == R - 1) ranges[cur].y = (uint32_t)R;
  const GcrRecord* src = records + point_list[i];
  const float4 a = __ldg(&src->q0), b = __ldg(&src->q1), c = __ldg(&src->q2);
  inst[i].q0 = a;
  inst[i].q1 = b;
  inst[i].q2 = c;
}

}  // namespace

size_t gcr_scan_workspace_bytes(size_t n) {
  const size_t nb = (n + kScanChunk - 1) / kScanChunk;
  return gcr_align_up((nb + 1) * sizeof(uint32_t), 256);
}

void gcr_launch_inclusive_scan(const uint32_t* in, const uint32_t* gather, uint32_t* out, size_t n,
                               void* workspace, cudaStream_t stream) {
  if (n == 0) return;
  if (n <= kSmallSortMaxN && small_paths_enabled()) {
    small_scan_kernel<<<1, kSmallThreads, 0, stream>>>(in, gather, (uint32_t)n, out);
    return;
  }
  const size_t nb = (n + kScanChunk - 1) / kScanChunk;
  uint32_t* sums = static_cast<uint32_t*>(workspace);
  scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, gather, n, sums);
  scan_spine_kernel<<<1, kScanThreads, 0, stream>>>(sums, nb);
  scan_final_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, gather, n, sums, out);
}

size_t gcr_sort_workspace_bytes(size_t n) {
  size_t nblk = (n + kSortChunk - 1) / kSortChunk;
  const size_t small = (n + 1023) / 1024;   // 1024-pair tiles may be used (always when n is small, or by env)
  if (small > nblk) nblk = small;
  // onesweep: global histograms [4][256] + tickets [4] (+pad) + status [4][nblk][256]
  return gcr_align_up((kMaxPasses * kBins + 64 + kMaxPasses * nblk * kBins) * sizeof(uint32_t), 256);
}

int gcr_launch_radix_sort(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                          size_t n, int end_bit, bool vals_iota, void* workspace,
                          cudaStream_t stream) {
  if (n == 0) return 0;
  if (end_bit <= 0) end_bit = 1;  // at least one pass so values are materialised
  if (n <= kSmallSortMaxN && small_paths_enabled()) {
    static const bool configured = [] {
      cudaFuncSetAttribute(small_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSortSmem);
      cudaFuncSetAttribute(small_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSortSmem);
      return true;
    }();
    (void)configured;
    if (vals_iota)
      small_sort_kernel<true><<<1, kSmallThreads, kSmallSortSmem, stream>>>(keys_a, vals_a, keys_b, vals_b, (uint32_t)n, end_bit);
    else
      small_sort_kernel<false><<<1, kSmallThreads, kSmallSortSmem, stream>>>(keys_a, vals_a, keys_b, vals_b, (uint32_t)n, end_bit);
    return ((end_bit + 7) / 8) & 1;
  }
  const uint32_t nblk = (uint32_t)((n + kSortChunk - 1) / kSortChunk);
  static const bool classic = getenv("GCR_SORT_CLASSIC") != nullptr;
  if (!classic && n < (size_t)kValMask && end_bit <= 8 * kMaxPasses) {
    uint32_t* ws = static_cast<uint32_t*>(workspace);
    uint32_t* ghist = ws;
    uint32_t* tickets = ws + kMaxPasses * kBins;
    uint32_t* status = ws + kMaxPasses * kBins + 64;
    const int npass = (end_bit + 7) / 8;
    static const int tile_env = [] { const char* e = getenv("GCR_SORT_TILE"); return e ? atoi(e) : 0; }();
    const bool big_tiles = tile_env >= 8192;
    // small inputs: 1024-pair tiles (4 ranking rounds instead of 16) -- latency, not bandwidth
    const bool small_tiles = (tile_env == 0 && n <= kSmallTileMaxN) || tile_env == 1024;
    const size_t ntiles = small_tiles ? (n + 1023) / 1024 : nblk;   // status stride per pass
    cudaMemsetAsync(ws, 0, (kMaxPasses * kBins + 64 + (size_t)npass * ntiles * kBins) * sizeof(uint32_t), stream);
    const unsigned hgrid = (unsigned)min((size_t)nblk, (size_t)148 * 8);
    radix_hist_all_kernel<<<hgrid, kSortThreads, 0, stream>>>(keys_a, n, end_bit, ghist);
    uint32_t* kin = keys_a; uint32_t* vin = vals_a; uint32_t* kout = keys_b; uint32_t* vout = vals_b;
    for (int p = 0; p < npass; ++p) {
      const int shift = 8 * p;
      const int bits = min(8, end_bit - shift);
      const uint32_t mask = (1u << bits) - 1u;
      uint32_t* st = status + (size_t)p * ntiles * kBins;
      const bool iota = vals_iota && p == 0;
      if (small_tiles) {
        const unsigned tiles = (unsigned)((n + 1023) / 1024);
        if (iota) launch_onesweep_pass<true, 256, 4>(tiles, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
        else launch_onesweep_pass<false, 256, 4>(tiles, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
      } else if (big_tiles) {
        const unsigned tiles = (unsigned)((n + 8191) / 8192);
        if (iota) launch_onesweep_pass<true, 512, 16>(tiles, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
        else launch_onesweep_pass<false, 512, 16>(tiles, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
      } else {
        if (iota) launch_onesweep_pass<true, 256, 16>(nblk, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
        else launch_onesweep_pass<false, 256, 16>(nblk, stream, kin, vin, kout, vout, n, shift, mask, bits, ghist + p * kBins, st, tickets + p);
      }
      uint32_t* t = kin; kin = kout; kout = t;
      t = vin; vin = vout; vout = t;
    }
    return npass & 1;
  }
  uint32_t* table = static_cast<uint32_t*>(workspace);
  uint32_t* totals = table + (size_t)nblk * kBins;
  uint32_t* kin = keys_a;
  uint32_t* vin = vals_a;
  uint32_t* kout = keys_b;
  uint32_t* vout = vals_b;
  int where = 0;
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int bits = min(8, end_bit - shift);
    const uint32_t mask = (1u << bits) - 1u;
    radix_hist_kernel<<<nblk, kSortThreads, 0, stream>>>(kin, n, shift, mask, bits, table, nblk);
    radix_scan_rows_kernel<<<kBins, kSortThreads, 0, stream>>>(table, nblk, totals);
    if (vals_iota && shift == 0)
      radix_scatter_kernel<true><<<nblk, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift,
                                                                    mask, bits, table, nblk, totals);
    else
      radix_scatter_kernel<false><<<nblk, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift,
                                                                     mask, bits, table, nblk, totals);
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
    where ^= 1;
  }
  return where;
}

void gcr_launch_emit_pairs(int P, const uint32_t* sorted_gauss, const uint32_t* offsets_incl,
                           const uint32_t* tiles_touched, const GcrRecord* records,
                           const int* radii, int grid_x, int grid_y, int shard_rank,
                           int shard_count, uint32_t* tile_keys, uint32_t* gauss_vals,
                           cudaStream_t stream) {
  if (P <= 0) return;
  emit_pairs_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, sorted_gauss, offsets_incl,
                                                         tiles_touched, records, radii, grid_x,
                                                         grid_y, shard_rank, shard_count,
                                                         tile_keys, gauss_vals);
}

void gcr_launch_ranges_and_gather(size_t R, const uint32_t* sorted_tile_keys,
                                  const uint32_t* point_list, const GcrRecord* records,
                                  uint2* ranges, GcrRecord* inst, cudaStream_t stream) {
  if (R == 0) return;
  ranges_gather_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(
      R, sorted_tile_keys, point_list, records, ranges, inst);
}
