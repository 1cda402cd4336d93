// gen_chol_dev.cuh -- device side of the general sparse block Cholesky (gen_chol.hpp): the level-scheduled
// triangular solves of the pose system for graphs with loop closures / several robots (SURVEY 8f-2;
// src/CORA_preconditioners.cpp:46-83 blockCholeskySolve, src/CORA_utils.cpp:33-57).
//
// ONE cooperative launch per solve.  A warp owns a cluster (a few consecutive poses of the elimination order: a
// whole small subtree or a piece of a tree path); the clusters of one dependency level run concurrently, levels are
// separated by a grid barrier (one L2 counter, release-add / acquire-spin), forward sweep up the levels, backward
// sweep down.  Lane (a, c) owns row a of the pose block and right-hand-side column c (B x CW lanes, CW = 32 / B
// columns per pass).
//
// A cluster K is eliminated by two products, never by a serial chain over its poses:
//     forward   r = b_K - sum over the blocks L_vu with u OUTSIDE the cluster of L_vu y_u ;   y_K = Linv_K r
//     backward  r = y_K - sum over the blocks L_wv with w outside of L_wv^T x_w ;              x_K = Linv_K^T r
// with Linv_K = L_KK^-1 (dense, poses x poses blocks, gen_cluster_inverses); the inverse arrives in shared memory
// as 16-byte coalesced loads.  The forward sum runs over the ROWS of L, which are long near the root (the few fat
// separators read ~1000 blocks each while a thousand warps idle), so it is split PUSH-style: the cluster that
// produces y_u also writes the products c = L_vu y_u for its (short) columns into a buffer laid out in row order;
// the consumer only sums a contiguous, index-free stream with many loads in flight.  The backward sum runs over
// columns (short) and gathers.  Measured history on TIERS (9768 poses, 14 levels,
// 3 columns): one launch per level + pose-by-pose elimination from global memory 2.2 ms per solve; cooperative
// kernel, shared-memory staging per cluster 1.8 ms -- a pose step is ~5600 cycles of dependent latency (one global
// round trip is ~500 cycles), 12 poses per cluster, 28 level sweeps; the two-product form removes the chain.
// Sums run in a fixed order: results are bit-reproducible.
#pragma once
#include "gen_chol.hpp"
#include "ops.cuh"

namespace cora_b200 {

struct GenSymDev {  // device copy of the GenSym arrays the kernels read
  DevBuf<int> perm, colptr, rowidx, rowptr, colidx, lvl_cl, lvl_ptr, desc;
  DevBuf<unsigned char> erow_f, erow_b;
  DevBuf<int> col2row;
  DevBuf<unsigned long long> bar, prof;
  std::vector<int> lvl_sizes;
  int nlevels = 0, max_level_clusters = 0;
  int grid = 0;  // cooperative grid of k_gen_solve (occupancy query cached per handle)
};

struct GenFactorDev {
  DevBuf<double> Lval, Dinv, Linv;        // nnzL blocks in column order, n inverse diagonal blocks, the per-cluster
                                          // inverses
  DevBuf<double> yw;                      // n * B * cols work vector in elimination order
  DevBuf<double> cbuf;                    // nnzL * B * cols products L_vu y_u of the forward sweep, row order
  int yw_cols = 0;
};

struct GenSolveArgs {
  const int *perm, *rowidx, *colidx, *lvl_cl, *lvl_ptr, *desc;
  const unsigned char *erow_f, *erow_b;
  const int *col2row;
  const double *Lval, *Linv;
  double *X, *yw, *cbuf;
  unsigned long long *bar;
  unsigned long long *prof;  // optional: CTA 0's clock after every grid barrier (CORA_B200_GEN_PROFILE)
  const CgCtrl *ctrl;
  int nlevels, ld, ncols;
};

constexpr int kGenWarps = 8;  // warps (clusters in flight) per CTA
constexpr int kGenBatch = 8;   // off-diagonal blocks in flight per lane (gathers: index, then block + operand)
constexpr int kGenStream = 32; // products in flight per lane when summing the contiguous forward stream

template <int B>
__host__ __device__ constexpr int gen_smem_doubles_per_warp() {
  return kGenClusterMax * 32 + kGenClusterMax * kGenClusterMax * B * B;  // r / results, the cluster inverse
}
template <int B>
constexpr size_t gen_smem_bytes() {
  return (size_t)kGenWarps * gen_smem_doubles_per_warp<B>() * sizeof(double);
}

__device__ __forceinline__ void gen_grid_sync(unsigned long long *bar, unsigned long long &target, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += G;
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// one cluster, one direction, one pass of CW columns.
// part / nparts: in the thinner levels of the tree (fewer clusters than warps in the grid) groups of 2, 4 or 8
// warps of a CTA share ONE cluster: each sums a slice of the off-diagonal products into its own copy of r (a single
// warp keeps only ~8 loads in flight: measured 4400 cycles per batch of 64), the group leader adds the copies in a
// fixed order, multiplies by the inverse, and all warps of the group push a slice of the column products.
template <int B, bool FWD>
__device__ __forceinline__ void gen_cluster(const GenSolveArgs &A, int K, int c0, double *own, double *linv, int lane,
                                            int part, int nparts, int bar_id) {
  // barrier of the nparts warps that share this cluster (a named barrier: the groups of a CTA run independently)
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nparts * 32) : "memory"); };
  constexpr int BB = B * B, CW = 32 / B;
  const int dv = lane < 8 ? A.desc[(size_t)K * 8 + lane] : 0;
  const int p0 = __shfl_sync(0xffffffffu, dv, 0), nodes = __shfl_sync(0xffffffffu, dv, 1);
  const int e0 = __shfl_sync(0xffffffffu, dv, FWD ? 2 : 4), nent = __shfl_sync(0xffffffffu, dv, FWD ? 3 : 5);
  const int loff = __shfl_sync(0xffffffffu, dv, 6);
  const int p1 = p0 + nodes;
  const int a = lane / CW, cc = lane - a * CW, c = c0 + cc;
  const int ncols = A.ncols;
  const bool row = a < B, active = row && c < ncols;
  const int al = row ? a : 0;
  const int *nb = (FWD ? A.colidx : A.rowidx) + e0;
  const unsigned char *er = (FWD ? A.erow_f : A.erow_b) + e0;
  const double *Ls = A.Lval + (size_t)e0 * BB;  // (backward: the cluster's columns)
  // Every loop below is "issue a batch of independent loads, THEN consume them".  The loads are unconditional
  // (clamped indices) and a compiler fence separates the two halves: with guarded loads the compiler merges the
  // equal guards and sinks each load next to its use -- LDG, STS, LDG, STS ... in the SASS, one ~500-cycle round
  // trip per element (measured: 255 000 cycles for the 518 products of one root cluster).
  // (a compiler fence alone is not enough: ptxas re-pairs the loads with their stores.  GEN_JOIN makes a branch
  //  depend on every loaded value -- a never-taken nanosleep -- so all loads of a batch are issued before it.)
  unsigned long long chk = 0;
#define GEN_KEEP(x) chk ^= (unsigned long long)(x)
#define GEN_KEEPD(x) chk ^= (unsigned long long)__double_as_longlong(x)
#define GEN_FENCE() do { if (chk == 0x9E3779B97F4A7C15ull) asm volatile("nanosleep.u32 1;"); asm volatile("" ::: "memory"); } while (0)
  const int cl = min(c, ncols - 1);  // clamped column: lanes beyond the last column load valid memory, store nothing
  // ---- the cluster inverse -> shared memory (16-byte loads, all in flight) ----
  if (part == 0) {
    const double2 *src = reinterpret_cast<const double2 *>(A.Linv + (size_t)loff * BB);  // loff even: 16-byte aligned
    double2 *dst = reinterpret_cast<double2 *>(linv);
    const int n2 = (nodes * nodes * BB + 1) / 2;
    for (int base = 0; base < n2; base += 32 * 12) {
      double2 v[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) { v[k] = src[min(base + k * 32 + lane, n2 - 1)]; GEN_KEEPD(v[k].x); GEN_KEEPD(v[k].y); }
      GEN_FENCE();
#pragma unroll
      for (int k = 0; k < 12; ++k) { const int i = base + k * 32 + lane; if (i < n2) dst[i] = v[k]; }
    }
  }
  // ---- right-hand sides of the cluster's poses ----
  if (part != 0) {
#pragma unroll
    for (int j = 0; j < kGenClusterMax; ++j)
      if (j < nodes) own[j * 32 + lane] = 0.0;
  } else {
    double v[kGenClusterMax];
    if (FWD) {
      int pp[kGenClusterMax];
#pragma unroll
      for (int j = 0; j < kGenClusterMax; ++j) { pp[j] = A.perm[p0 + min(j, nodes - 1)]; GEN_KEEP(pp[j]); }
      GEN_FENCE();
#pragma unroll
      for (int j = 0; j < kGenClusterMax; ++j) { v[j] = A.X[((size_t)pp[j] * B + al) * A.ld + cl]; GEN_KEEPD(v[j]); }
    } else {
#pragma unroll
      for (int j = 0; j < kGenClusterMax; ++j) { v[j] = A.yw[((size_t)(p0 + min(j, nodes - 1)) * B + al) * ncols + cl]; GEN_KEEPD(v[j]); }
    }
    GEN_FENCE();
#pragma unroll
    for (int j = 0; j < kGenClusterMax; ++j)
      if (j < nodes) own[j * 32 + lane] = active ? v[j] : 0.0;
  }
  // ---- minus the products with the blocks that couple the cluster to poses outside it ----
  // (this warp's slice: whole batches, so the order of the sums does not depend on the slicing of other warps)
  constexpr int kStep = FWD ? kGenStream : kGenBatch;
  const int nbatch = (nent + kStep - 1) / kStep;
  const int tbeg = (int)((long long)nbatch * part / nparts) * kStep;
  const int tend = min(nent, (int)((long long)nbatch * (part + 1) / nparts) * kStep);
  if (FWD) {  // produced by the clusters below (push): a contiguous stream in row order
    const double *cb = A.cbuf + ((size_t)e0 * B + al) * ncols + cl;
    for (int t0 = tbeg; t0 < tend; t0 += kGenStream) {
      int fl[kGenStream];
      double v[kGenStream];
#pragma unroll
      for (int k = 0; k < kGenStream; ++k) { fl[k] = (int)er[min(t0 + k, nent - 1)]; GEN_KEEP(fl[k]); }
#pragma unroll
      for (int k = 0; k < kGenStream; ++k) { v[k] = cb[(size_t)min(t0 + k, nent - 1) * B * ncols]; GEN_KEEPD(v[k]); }
      GEN_FENCE();
      // entries are ordered by row: sum runs of equal rows in registers, one shared-memory update per run (a
      // read-modify-write per entry is a ~100-cycle dependent chain through shared memory)
      double run = 0.0;
      int rrow = fl[0] & 0x7f;
#pragma unroll
      for (int k = 0; k < kGenStream; ++k) {
        const int rk = fl[k] & 0x7f;
        if (rk != rrow) {
          if (active) own[rrow * 32 + lane] -= run;
          run = 0.0;
          rrow = rk;
        }
        if (t0 + k < nent && !(fl[k] & 0x80)) run += v[k];
      }
      if (active) own[rrow * 32 + lane] -= run;
    }
  } else {  // gather over the (short) columns
    for (int t0 = tbeg; t0 < tend; t0 += kGenBatch) {
      int u[kGenBatch], fl[kGenBatch];
#pragma unroll
      for (int k = 0; k < kGenBatch; ++k) {
        const int t = min(t0 + k, nent - 1);
        fl[k] = (int)er[t];
        u[k] = nb[t];
        GEN_KEEP(fl[k]); GEN_KEEP(u[k]);
      }
      GEN_FENCE();
      double Lv[kGenBatch][B], yv[kGenBatch][B];
#pragma unroll
      for (int k = 0; k < kGenBatch; ++k) {
        const int t = min(t0 + k, nent - 1);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          Lv[k][b] = Ls[(size_t)t * BB + b * B + al];
          yv[k][b] = A.yw[((size_t)u[k] * B + b) * ncols + cl];
          GEN_KEEPD(Lv[k][b]); GEN_KEEPD(yv[k][b]);
        }
      }
      GEN_FENCE();
      double run = 0.0;  // (runs of equal columns, as above; this lane is the only writer of its element)
      int rrow = fl[0] & 0x7f;
#pragma unroll
      for (int k = 0; k < kGenBatch; ++k) {
        const int rk = fl[k] & 0x7f;
        if (rk != rrow) {
          if (active) own[rrow * 32 + lane] -= run;
          run = 0.0;
          rrow = rk;
        }
        if (t0 + k < nent && !(fl[k] & 0x80)) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < B; ++b) s += Lv[k][b] * yv[k][b];
          run += s;
        }
      }
      if (active) own[rrow * 32 + lane] -= run;
    }
  }
  // push (forward sweep): c = L_wu y_u for the blocks of the cluster's columns whose row w lies outside (above) the
  // cluster, written where the owner of row w will stream them; `res_tab`: the cluster's results in shared memory.
  // In CTA mode every warp pushes a slice of the column blocks.
  auto push = [&](const double *res_tab) {
    const int e0c = __shfl_sync(0xffffffffu, dv, 4), nentc = __shfl_sync(0xffffffffu, dv, 5);
    const unsigned char *ec = A.erow_b + e0c;
    const int *c2r = A.col2row + e0c;
    const double *Lc = A.Lval + (size_t)e0c * BB;
    const int nb_c = (nentc + kGenBatch - 1) / kGenBatch;
    const int cbeg = (int)((long long)nb_c * part / nparts) * kGenBatch;
    const int cend = min(nentc, (int)((long long)nb_c * (part + 1) / nparts) * kGenBatch);
    for (int t0 = cbeg; t0 < cend; t0 += kGenBatch) {
      int fl[kGenBatch], dst[kGenBatch];
      double Lv[kGenBatch][B];
#pragma unroll
      for (int k = 0; k < kGenBatch; ++k) {
        const int t = min(t0 + k, nentc - 1);
        fl[k] = (int)ec[t];
        dst[k] = c2r[t];
        GEN_KEEP(fl[k]); GEN_KEEP(dst[k]);
#pragma unroll
        for (int b = 0; b < B; ++b) { Lv[k][b] = Lc[(size_t)t * BB + al * B + b]; GEN_KEEPD(Lv[k][b]); }
      }
      GEN_FENCE();
#pragma unroll
      for (int k = 0; k < kGenBatch; ++k)
        if (t0 + k < nentc && !(fl[k] & 0x80) && active) {
          const double *y = res_tab + (fl[k] & 0x7f) * 32 + cc;
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < B; ++b) s += Lv[k][b] * y[b * CW];
          A.cbuf[((size_t)dst[k] * B + a) * ncols + c] = s;
        }
    }
    __syncwarp();
  };
  __syncwarp();
  if (nparts > 1) {  // (uniform over the group)
    group_sync();
    const int stride = gen_smem_doubles_per_warp<B>();
    if (part != 0) {
      if (FWD) {
        group_sync();                  // the leader has combined, multiplied by the inverse and stored the results
        push(own - (size_t)part * stride);
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < kGenClusterMax; ++j)
      if (j < nodes) {
        double r = own[j * 32 + lane];
        for (int w = 1; w < nparts; ++w) r += own[(size_t)w * stride + j * 32 + lane];
        own[j * 32 + lane] = r;
      }
    __syncwarp();
  }
  // ---- times the cluster inverse (transposed on the way down) ----
  double res[kGenClusterMax];
#pragma unroll
  for (int j = 0; j < kGenClusterMax; ++j) {
    res[j] = 0.0;
    if (j < nodes) {
      double s = 0.0;
      const int j0 = FWD ? 0 : j, j1 = FWD ? j : nodes - 1;
      for (int jj = j0; jj <= j1; ++jj) {
        const double *Lb = FWD ? linv + ((size_t)j * nodes + jj) * BB + al * B : linv + ((size_t)jj * nodes + j) * BB + al;
        const double *r = own + jj * 32 + cc;
#pragma unroll
        for (int b = 0; b < B; ++b) s += (FWD ? Lb[b] : Lb[b * B]) * r[b * CW];
      }
      res[j] = s;
    }
  }
  __syncwarp();
  int ppw[kGenClusterMax];
  if (!FWD) {
#pragma unroll
    for (int j = 0; j < kGenClusterMax; ++j) { ppw[j] = A.perm[p0 + min(j, nodes - 1)]; GEN_KEEP(ppw[j]); }
    GEN_FENCE();
  }
#pragma unroll
  for (int j = 0; j < kGenClusterMax; ++j)
    if (j < nodes) {
      if (active) {
        A.yw[((size_t)(p0 + j) * B + a) * ncols + c] = res[j];
        if (!FWD) A.X[((size_t)ppw[j] * B + a) * A.ld + c] = res[j];
      }
      if (FWD && row) own[j * 32 + lane] = res[j];
    }
  __syncwarp();
  if (FWD) {
    if (nparts > 1) group_sync();  // the results are in this warp's `own`: the other warps of the group push too
    push(own);
  }
#undef GEN_FENCE
#undef GEN_KEEP
#undef GEN_KEEPD
}

template <int B>
__global__ void __launch_bounds__(kGenWarps * 32) k_gen_solve(const GenSolveArgs A) {
  if (A.ctrl != nullptr && *((volatile const int *)&A.ctrl->state) != 0) return;
  extern __shared__ double gen_smem[];
  constexpr int CW = 32 / B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *own = gen_smem + (size_t)warp * gen_smem_doubles_per_warp<B>();
  double *linv = own + kGenClusterMax * 32;
  const int gwarp = blockIdx.x * kGenWarps + warp, nwarps = gridDim.x * kGenWarps;
  unsigned long long target = 0;
  // the largest group (8, 4, 2 warps of a CTA per cluster) that still handles the level in one round
  auto group_size = [&](int nclusters) {
    int g = kGenWarps;
    while (g > 1 && (long long)gridDim.x * (kGenWarps / g) < nclusters) g >>= 1;
    return g;
  };
  int np = 0;
  auto stamp = [&]() {
    if (A.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      A.prof[np] = t;
    }
    ++np;
  };
  stamp();
  for (int c0 = 0; c0 < A.ncols; c0 += CW) {
    for (int t = 0; t < A.nlevels; ++t) {
      const int l0 = A.lvl_ptr[t], l1 = A.lvl_ptr[t + 1];
      {
        const int g = group_size(l1 - l0);  // warps per cluster at this level: 1, 2, 4 or 8
        if (g == 1) {
          for (int k = l0 + gwarp; k < l1; k += nwarps) gen_cluster<B, true>(A, A.lvl_cl[k], c0, own, linv, lane, 0, 1, 0);
        } else {
          const int k = l0 + (int)blockIdx.x * (kGenWarps / g) + warp / g;
          if (k < l1) gen_cluster<B, true>(A, A.lvl_cl[k], c0, own, linv, lane, warp % g, g, 1 + warp / g);
        }
      }
      gen_grid_sync(A.bar, target, gridDim.x);
      stamp();
    }
    for (int t = A.nlevels - 1; t >= 0; --t) {
      const int l0 = A.lvl_ptr[t], l1 = A.lvl_ptr[t + 1];
      {
        const int g = group_size(l1 - l0);
        if (g == 1) {
          for (int k = l0 + gwarp; k < l1; k += nwarps) gen_cluster<B, false>(A, A.lvl_cl[k], c0, own, linv, lane, 0, 1, 0);
        } else {
          const int k = l0 + (int)blockIdx.x * (kGenWarps / g) + warp / g;
          if (k < l1) gen_cluster<B, false>(A, A.lvl_cl[k], c0, own, linv, lane, warp % g, g, 1 + warp / g);
        }
      }
      if (t > 0 || c0 + CW < A.ncols) gen_grid_sync(A.bar, target, gridDim.x);
      stamp();
    }
  }
}

inline void gen_sym_upload(H *h, const GenSym &S, GenSymDev &D) {
  cudaStream_t s = h->stream;
  auto up = [&](DevBuf<int> &b, const std::vector<int32_t> &v) {
    std::vector<int> t(v.begin(), v.end());
    b.upload(t, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  };
  up(D.perm, S.perm); up(D.colptr, S.colptr); up(D.rowidx, S.rowidx); up(D.rowptr, S.rowptr);
  up(D.colidx, S.colidx); up(D.lvl_cl, S.lvl_cl); up(D.lvl_ptr, S.lvl_ptr); up(D.desc, S.desc);
  D.erow_f.upload(S.erow_f, s); D.erow_b.upload(S.erow_b, s); up(D.col2row, S.col2row);
  CUDA_CHECK(cudaStreamSynchronize(s));
  D.bar.alloc(1);
  D.nlevels = S.levels();
  D.max_level_clusters = 0;
  D.lvl_sizes.clear();
  for (int t = 0; t < D.nlevels; ++t) {
    D.lvl_sizes.push_back(S.lvl_ptr[t + 1] - S.lvl_ptr[t]);
    D.max_level_clusters = std::max(D.max_level_clusters, S.lvl_ptr[t + 1] - S.lvl_ptr[t]);
  }
}

// X ([n][B][ld] in pose order, device) <- T^-1 X for `ncols` columns
template <int B>
inline void gen_solve_device(H *h, GenSymDev &D, GenFactorDev &F, int n, double *X, int ld, int ncols, const CgCtrl *ctrl) {
  if (n <= 0 || D.nlevels <= 0) return;
  if (F.yw_cols < ncols) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    F.yw.alloc((size_t)n * B * ncols);
    F.cbuf.alloc(std::max<size_t>(D.rowidx.n, 1) * B * ncols);
    F.yw_cols = ncols;
  }
  cudaStream_t s = h->stream;
  GenSolveArgs A{};
  A.perm = D.perm.p; A.rowidx = D.rowidx.p; A.colidx = D.colidx.p;
  A.lvl_cl = D.lvl_cl.p; A.lvl_ptr = D.lvl_ptr.p; A.desc = D.desc.p;
  A.erow_f = D.erow_f.p; A.erow_b = D.erow_b.p; A.col2row = D.col2row.p;
  A.Lval = F.Lval.p; A.Linv = F.Linv.p;
  A.X = X; A.yw = F.yw.p; A.cbuf = F.cbuf.p; A.bar = D.bar.p; A.ctrl = ctrl;
  A.nlevels = D.nlevels; A.ld = ld; A.ncols = ncols;
  const size_t smem = gen_smem_bytes<B>();
  void *kfn = (void *)k_gen_solve<B>;
  if (D.grid == 0) {
    CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kGenWarps * 32, smem));
    if (per_sm < 1) throw Error(CORA_B200_ERUNTIME, "general Cholesky solve kernel does not fit on an SM");
    D.grid = std::max(1, std::min(h->sm_count * per_sm, (D.max_level_clusters + kGenWarps - 1) / kGenWarps));
  }
  const int G = D.grid;
  CUDA_CHECK(cudaMemsetAsync(D.bar.p, 0, sizeof(unsigned long long), s));
  static const bool profile = getenv("CORA_B200_GEN_PROFILE") != nullptr;
  const int nstamp = 1 + 2 * D.nlevels * ((ncols + 32 / B - 1) / (32 / B));
  if (profile) { if (D.prof.n < (size_t)nstamp) D.prof.alloc(nstamp); A.prof = D.prof.p; }
  void *args[] = {(void *)&A};
  CUDA_CHECK(cudaLaunchCooperativeKernel(kfn, dim3(G), dim3(kGenWarps * 32), args, smem, s));
  check_launch(h);
  if (profile && ctrl == nullptr) {
    std::vector<unsigned long long> t(nstamp);
    CUDA_CHECK(cudaMemcpyAsync(t.data(), D.prof.p, nstamp * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::printf("[gen solve] grid %d x %d warps, %d columns, %d levels, total %.1f us; per level (clusters: fwd us / bwd us):", G, kGenWarps,
                ncols, D.nlevels, (t[nstamp - 1] - t[0]) * 1e-3);
    for (int lv = 0; lv < D.nlevels; ++lv)
      std::printf(" [%d: %.1f / %.1f]", D.lvl_sizes[lv], (t[1 + lv] - t[lv]) * 1e-3,
                  (t[1 + D.nlevels + (D.nlevels - 1 - lv)] - t[D.nlevels + (D.nlevels - 1 - lv)]) * 1e-3);
    std::printf("\n");
  }
}

}  // namespace cora_b200
