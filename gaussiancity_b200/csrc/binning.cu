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
                  uint32_t* __restrict__ table, uint32_t nblk) {
  __shared__ uint32_t hist[kBins];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * kSortChunk + (size_t)warp * (32 * kSortItems);
#pragma unroll 4
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = (keys[i] >> shift) & digit_mask;
      const unsigned m = __match_any_sync(vmask, d);
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
                     int shift, uint32_t digit_mask, const uint32_t* __restrict__ table,
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
  for (int r = 0; r < kSortItems; ++r) {
    const size_t i = base + (size_t)r * 32 + lane;
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    key[r] = 0;
    val[r] = 0;
    rank[r] = 0;
    if (valid) {
      key[r] = keys_in[i];
      val[r] = kIota ? (uint32_t)i : vals_in[i];
      const uint32_t d = (key[r] >> shift) & digit_mask;
      const unsigned m = __match_any_sync(vmask, d);
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
__global__ void __launch_bounds__(256)
emit_pairs_kernel(int P, const uint32_t* __restrict__ sorted_gauss,
                  const uint32_t* __restrict__ offsets_incl,
                  const uint32_t* __restrict__ tiles_touched,
                  const GcrRecord* __restrict__ records, const int* __restrict__ radii, int grid_x,
                  int grid_y, int shard_rank, int shard_count, uint32_t* __restrict__ tile_keys,
                  uint32_t* __restrict__ gauss_vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t g = sorted_gauss[i];
  const uint32_t n = tiles_touched[g];
  if (n == 0) return;
  uint32_t off = offsets_incl[i] - n;
  const float4 q0 = records[g].q0;
  uint2 rmin, rmax;
  gcr_get_rect(q0.x, q0.y, radii[g], grid_x, grid_y, rmin, rmax);
  for (uint32_t y = rmin.y; y < rmax.y; ++y) {
    if (shard_count > 1 && (int)(y % (uint32_t)shard_count) != shard_rank) continue;
    for (uint32_t x = rmin.x; x < rmax.x; ++x) {
      tile_keys[off] = y * grid_x + x;
      gauss_vals[off] = g;
      ++off;
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
  const size_t nb = (n + kScanChunk - 1) / kScanChunk;
  uint32_t* sums = static_cast<uint32_t*>(workspace);
  scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, gather, n, sums);
  scan_spine_kernel<<<1, kScanThreads, 0, stream>>>(sums, nb);
  scan_final_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, gather, n, sums, out);
}

size_t gcr_sort_workspace_bytes(size_t n) {
  const size_t nblk = (n + kSortChunk - 1) / kSortChunk;
  return gcr_align_up((nblk * kBins + kBins) * sizeof(uint32_t), 256);
}

int gcr_launch_radix_sort(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                          size_t n, int end_bit, bool vals_iota, void* workspace,
                          cudaStream_t stream) {
  if (n == 0) return 0;
  const uint32_t nblk = (uint32_t)((n + kSortChunk - 1) / kSortChunk);
  uint32_t* table = static_cast<uint32_t*>(workspace);
  uint32_t* totals = table + (size_t)nblk * kBins;
  uint32_t* kin = keys_a;
  uint32_t* vin = vals_a;
  uint32_t* kout = keys_b;
  uint32_t* vout = vals_b;
  int where = 0;
  if (end_bit <= 0) end_bit = 1;  // at least one pass so values are materialised
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int bits = min(8, end_bit - shift);
    const uint32_t mask = (1u << bits) - 1u;
    radix_hist_kernel<<<nblk, kSortThreads, 0, stream>>>(kin, n, shift, mask, table, nblk);
    radix_scan_rows_kernel<<<kBins, kSortThreads, 0, stream>>>(table, nblk, totals);
    if (vals_iota && shift == 0)
      radix_scatter_kernel<true><<<nblk, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift,
                                                                    mask, table, nblk, totals);
    else
      radix_scatter_kernel<false><<<nblk, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, shift,
                                                                     mask, table, nblk, totals);
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
