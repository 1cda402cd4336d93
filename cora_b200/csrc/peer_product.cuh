// peer_product.cuh -- the row-partitioned data-matrix product (cora_b200/rowpart.py) with its two
// exchanges done by THE LIBRARY'S OWN KERNELS over peer-mapped memory (NVLink / NVSwitch loads), not by collectives:
//
//   k_peer_halo     cross-GPU barrier (every rank has its operand ready), then the ghost pose rows of the operand are
//                   PULLED from their owners' buffers (plain loads through the peer mapping)
//   local product   the persistent SpMM kernel on the rank's rows (asynchronous, same stream)
//   k_peer_reduce   the rank's partial landmark rows go to an exported staging buffer, cross-GPU barrier, then every
//                   rank sums the partials of ALL ranks in rank order (deterministic, identical on every rank)
//
// Each rank exports three allocations with cudaIpcGetMemHandle (operand, result staging, flags) and opens the peers'
// with cudaIpcOpenMemHandle; the handles travel over any out-of-band channel (rowpart.py: torch.distributed).  The
// barrier is a flag exchange: rank a writes the product's epoch into flags[a] of every peer (system-scope release),
// and spins (system-scope acquire) until its own flags hold the epoch from everybody -- with a wall-clock timeout so
// that a missing peer turns into an error code instead of a hung GPU.  Measured against the NCCL formulation
// (all_to_all_single + all_reduce, ~35 us each): see profiles/README.md, r02.
#pragma once
#include "solver.cuh"

namespace cora_b200 {

constexpr int kPeerMaxWorld = 16;

struct PeerCtx {
  H *h = nullptr;
  int world = 0, rank = 0, r = 0;
  const double *peer_x[kPeerMaxWorld] = {};      // operand buffers (internal layout) of every rank, peer mapped
  const double *peer_stage[kPeerMaxWorld] = {};  // landmark staging buffers
  unsigned long long *peer_flags[kPeerMaxWorld] = {};  // [2][kPeerMaxWorld] epochs per rank
  void *opened[3 * kPeerMaxWorld] = {};
  DevBuf<double> stage;
  DevBuf<unsigned long long> flags;
  DevBuf<int> ghost_peer, ghost_src, ghost_dst, lm_rows;
  DevBuf<int> err;
  int n_ghost = 0, n_lm = 0;
  unsigned long long epoch = 0;
};

struct PeerPtrs {
  const double *x[kPeerMaxWorld];
  const double *stage[kPeerMaxWorld];
  unsigned long long *flags[kPeerMaxWorld];
};

// flag barrier `which` (0: operands ready, 1: partial landmark rows staged) at `epoch`; one warp
__device__ __forceinline__ void peer_barrier(const PeerPtrs &P, int world, int rank, int which, unsigned long long epoch,
                                             int *err) {
  __threadfence_system();
  const int lane = threadIdx.x & 31;
  if (lane < world) {
    unsigned long long *dst = P.flags[lane] + (size_t)which * kPeerMaxWorld + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
    const unsigned long long *src = P.flags[rank] + (size_t)which * kPeerMaxWorld + lane;
    unsigned long long t0, t, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 5000000000ull) { *err = 1; break; }  // 5 s: a peer never arrived
    } while (v < epoch);
  }
  __syncwarp();
  __threadfence_system();
}

static __global__ void __launch_bounds__(256) k_peer_halo(PeerPtrs P, int world, int rank, unsigned long long epoch, int r,
                                                          int n_ghost, const int *__restrict__ peer,
                                                          const int *__restrict__ src_row, const int *__restrict__ dst_row,
                                                          double *x, int *err) {
  if (threadIdx.x < 32) peer_barrier(P, world, rank, 0, epoch, err);
  __syncthreads();
  for (int e = threadIdx.x; e < n_ghost * r; e += blockDim.x) {
    const int g = e / r, c = e - g * r;
    x[(size_t)dst_row[g] * r + c] = P.x[peer[g]][(size_t)src_row[g] * r + c];
  }
}

static __global__ void __launch_bounds__(256) k_peer_reduce(PeerPtrs P, int world, int rank, unsigned long long epoch, int r,
                                                            int n_lm, const int *__restrict__ lm_rows, double *stage,
                                                            double *y, int *err) {
  for (int e = threadIdx.x; e < n_lm * r; e += blockDim.x) {
    const int j = e / r, c = e - j * r;
    stage[e] = y[(size_t)lm_rows[j] * r + c];
  }
  __syncthreads();
  if (threadIdx.x < 32) peer_barrier(P, world, rank, 1, epoch, err);
  __syncthreads();
  for (int e = threadIdx.x; e < n_lm * r; e += blockDim.x) {
    const int j = e / r, c = e - j * r;
    double s = 0.0;
    for (int q = 0; q < world; ++q) s += P.stage[q][e];  // rank order on every rank: identical sums
    y[(size_t)lm_rows[j] * r + c] = s;
  }
}

}  // namespace cora_b200
