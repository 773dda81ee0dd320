// One-shot all-reduce of the packed sync-BN statistics over NVLink peer memory (NVSwitch): every rank's
// [G][C][2] fp64 partial sums live in a symmetric buffer that all peers map; after a flag exchange each rank
// reads all peers' slots directly (peer loads) and adds them in rank order (bit-identical result on every
// rank).  Replaces the NCCL all-reduce that SyncBatchNorm issues per layer and pass (train_adamml.py:125-127;
// SURVEY.md §2.4 C2): one ~5 us kernel instead of a collective + two cross-stream synchronisations, and it is
// captured into the step's CUDA graph like any other kernel.
//
// Synchronisation: `flags` is a symmetric array [lanes][world] of 32-bit epochs per rank.  Lane = the stream
// (backbone) the caller runs on: every rank issues the same sequence of reductions per lane.  Rank r publishes
// epoch e by writing flags_of_peer[p][lane][r] = e on every peer p (release, system scope) and then waits until
// its own flags[lane][q] >= e for every q (acquire).  The epoch counter lives in device memory and is advanced
// by the kernel itself, so graph replays keep counting.  Every BN layer owns its own slots, so a slot is only
// rewritten a whole step (hundreds of flag exchanges) after its last remote read.
#include "common.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024)
p2p_allreduce_f64_kernel(const unsigned long long* __restrict__ peer_bufs, const unsigned long long* __restrict__ peer_flags,
                         long long slot_off, double* __restrict__ out, int n, int world, int rank, int lane,
                         unsigned* __restrict__ epoch_ctr, unsigned* __restrict__ err_flag, long long timeout_cycles) {
  __shared__ unsigned s_epoch;
  __shared__ int s_fail;
  if (threadIdx.x == 0) {
    s_epoch = ++epoch_ctr[lane];
    s_fail = 0;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  __threadfence_system();  // this rank's partial sums (written by the preceding kernels) are visible to the peers
  if ((int)threadIdx.x < world) {
    unsigned* pf = reinterpret_cast<unsigned*>(peer_flags[threadIdx.x]);
    st_release_sys(pf + (long long)lane * world + rank, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(peer_flags[rank]) + (long long)lane * world + threadIdx.x;
    const long long t0 = clock64();
    // epochs only grow; (int) difference tolerates wrap-around
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        s_fail = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (s_fail) {
    // a peer never arrived: make the failure impossible to miss -- the reduced statistics become NaN (so do the
    // activations, the loss and every gradient of this step) and err_flag is raised for P2PStats.check()
    if (threadIdx.x == 0) atomicExch(err_flag, 1u);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = nan;
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) {
      const double* pb = reinterpret_cast<const double*>(peer_bufs[r]) + slot_off;
      acc += __ldcv(pb + i);
    }
    out[i] = acc;
  }
}

}  // namespace

extern "C" {

// peer_bufs / peer_flags: device arrays [world] of the peers' symmetric base addresses (as mapped in this process).
int adamml_p2p_allreduce_f64(const unsigned long long* peer_bufs, const unsigned long long* peer_flags,
                             long long slot_off, double* out, int n, int world, int rank, int lane,
                             unsigned* epoch_ctr, unsigned* err_flag, cudaStream_t stream) {
  ADAMML_REQUIRE(n > 0 && world > 1 && world <= 64 && rank >= 0 && rank < world && lane >= 0,
                 "p2p_allreduce: bad arguments");
  // how long a rank waits for its peers before it poisons the step (see the kernel): default 120 s at ~2 GHz, i.e.
  // longer than a checkpoint save or a data-loader stall; ADAMML_B200_P2P_TIMEOUT_S overrides (0 = wait forever and
  // leave stragglers to the process-group watchdog, like the NCCL path)
  static const long long timeout_cycles = []() {
    const char* e = getenv("ADAMML_B200_P2P_TIMEOUT_S");
    const double sec = e ? atof(e) : 120.0;
    return sec <= 0.0 ? (long long)0x7fffffffffffffffLL : (long long)(sec * 2.0e9);
  }();
  p2p_allreduce_f64_kernel<<<1, 1024, 0, stream>>>(peer_bufs, peer_flags, slot_off, out, n, world, rank, lane,
                                                  epoch_ctr, err_flag, timeout_cycles);
  return adamml_check_launch("p2p_allreduce_f64");
}

}  // extern "C"
