// peer.cu -- cross-GPU barrier over peer-mapped memory (NVLink 5 / NVSwitch) for tile-sharded
// frames.  The backward blend of every rank adds straight into the accumulator of the rank that
// owns each Gaussian (blend_bwd.cu); before an owner may read its accumulator, every rank's blend
// must have finished and its remote reductions must have landed.  One tiny kernel per rank:
//   fence.sys (orders the preceding kernel's peer reductions before the flag)
//   -> st.release.sys epoch into slot [rank] of every peer's flag array
//   -> ld.acquire.sys spin on the own array until every slot has reached the epoch.
// Epochs only grow, so a slot never has to be reset and a fast rank cannot be overtaken.
#include "gcr_common.cuh"
#include "gcr_kernels.h"

namespace {

__global__ void peer_barrier_kernel(GcrPeerFlags flags, int rank, int world, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  uint32_t* remote = flags.p[t] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t* mine = flags.p[rank] + t;
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
  } while ((int32_t)(v - epoch) < 0);
  __threadfence_system();
}

}  // namespace

cudaError_t gcr_launch_peer_barrier(const GcrPeerFlags& flags, int rank, int world, uint32_t epoch,
                                    cudaStream_t stream) {
  if (world <= 1) return cudaSuccess;
  peer_barrier_kernel<<<1, 32, 0, stream>>>(flags, rank, world, epoch);
  return cudaGetLastError();
}
