// gen_chol_dev.cuh -- device side of the general sparse block Cholesky (gen_chol.hpp): the level-scheduled
// triangular solves of the pose system for graphs with loop closures / several robots (SURVEY 8f-2;
// src/CORA_preconditioners.cpp:46-83 blockCholeskySolve, src/CORA_utils.cpp:33-57).
//
// One WARP per cluster (a few consecutive poses of the elimination order: a whole small subtree or a piece of a
// tree path), clusters of one dependency level per launch.  Lane (a, c) owns row a of the pose block and
// right-hand-side column c (B x CW lanes, CW = 32 / B columns per pass); the poses of a cluster are eliminated
// serially by the warp, results of earlier poses of the same cluster are read back after __syncwarp().  Sums run
// in pattern order: results are bit-reproducible and equal to gen_solve_host's.
#pragma once
#include "gen_chol.hpp"
#include "ops.cuh"

namespace cora_b200 {

struct GenSymDev {  // device copy of the GenSym arrays the kernels read
  DevBuf<int> perm, colptr, rowidx, rowptr, colidx, rowslot, cl_ptr, lvl_cl;
  std::vector<int32_t> lvl_ptr;  // host
};

struct GenFactorDev {
  DevBuf<double> Lval, Dinv;  // nnzL / n blocks
  DevBuf<double> yw;          // n * B * cols work vector in elimination order
  int yw_cols = 0;
};

template <int B>
__global__ void __launch_bounds__(128) k_gen_forward(int ncl, const int *__restrict__ lvl_cl, const int *__restrict__ cl_ptr,
                                                     const int *__restrict__ perm, const int *__restrict__ rowptr,
                                                     const int *__restrict__ colidx, const int *__restrict__ rowslot,
                                                     const double *__restrict__ Lval, const double *__restrict__ Dinv,
                                                     const double *X, double *yw, int ld, int ncols, const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  constexpr int BB = B * B, CW = 32 / B;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ncl) return;
  const int K = lvl_cl[warp];
  const int p0 = cl_ptr[K], p1 = cl_ptr[K + 1];
  const int a = lane / CW, cc = lane - a * CW;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    const int c = c0 + cc;
    const bool active = a < B && c < ncols;
    for (int p = p0; p < p1; ++p) {
      double acc = 0.0;
      if (active) {
        acc = X[((size_t)perm[p] * B + a) * ld + c];
        for (int q = rowptr[p]; q < rowptr[p + 1]; ++q) {
          const double *Lb = Lval + (size_t)rowslot[q] * BB + a * B;
          const double *yu = yw + (size_t)colidx[q] * B * ncols + c;
#pragma unroll
          for (int b = 0; b < B; ++b) acc -= Lb[b] * yu[(size_t)b * ncols];
        }
      }
      double s = 0.0;
      const double *Li = Dinv + (size_t)p * BB + (a < B ? a : 0) * B;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const double ab = __shfl_sync(0xffffffffu, acc, b * CW + cc);
        if (b <= a) s += Li[b] * ab;
      }
      if (active) yw[((size_t)p * B + a) * ncols + c] = s;
      __syncwarp();
    }
  }
}

template <int B>
__global__ void __launch_bounds__(128) k_gen_backward(int ncl, const int *__restrict__ lvl_cl, const int *__restrict__ cl_ptr,
                                                      const int *__restrict__ perm, const int *__restrict__ colptr,
                                                      const int *__restrict__ rowidx, const double *__restrict__ Lval,
                                                      const double *__restrict__ Dinv, double *X, double *yw, int ld,
                                                      int ncols, const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  constexpr int BB = B * B, CW = 32 / B;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ncl) return;
  const int K = lvl_cl[warp];
  const int p0 = cl_ptr[K], p1 = cl_ptr[K + 1];
  const int a = lane / CW, cc = lane - a * CW;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    const int c = c0 + cc;
    const bool active = a < B && c < ncols;
    for (int p = p1 - 1; p >= p0; --p) {
      double acc = 0.0;
      if (active) {
        acc = yw[((size_t)p * B + a) * ncols + c];
        for (int q = colptr[p]; q < colptr[p + 1]; ++q) {
          const double *Lb = Lval + (size_t)q * BB + a;
          const double *xw = yw + (size_t)rowidx[q] * B * ncols + c;
#pragma unroll
          for (int b = 0; b < B; ++b) acc -= Lb[b * B] * xw[(size_t)b * ncols];
        }
      }
      double s = 0.0;
      const double *Li = Dinv + (size_t)p * BB + (a < B ? a : 0);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const double ab = __shfl_sync(0xffffffffu, acc, b * CW + cc);
        if (b >= a) s += Li[b * B] * ab;
      }
      __syncwarp();  // every lane has read y_p through the shuffles before it is overwritten
      if (active) {
        yw[((size_t)p * B + a) * ncols + c] = s;
        X[((size_t)perm[p] * B + a) * ld + c] = s;
      }
      __syncwarp();
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
  up(D.colidx, S.colidx); up(D.rowslot, S.rowslot); up(D.cl_ptr, S.cl_ptr); up(D.lvl_cl, S.lvl_cl);
  D.lvl_ptr = S.lvl_ptr;
}

// X ([n][B][ld] in pose order, device) <- T^-1 X for `ncols` columns
template <int B>
inline void gen_solve_device(H *h, const GenSymDev &D, GenFactorDev &F, int n, double *X, int ld, int ncols,
                             const CgCtrl *ctrl) {
  if (n <= 0) return;
  if (F.yw_cols < ncols) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    F.yw.alloc((size_t)n * B * ncols);
    F.yw_cols = ncols;
  }
  cudaStream_t s = h->stream;
  const int nl = (int)D.lvl_ptr.size() - 1;
  for (int t = 0; t < nl; ++t) {
    const int ncl = D.lvl_ptr[t + 1] - D.lvl_ptr[t];
    k_gen_forward<B><<<(ncl + 3) / 4, 128, 0, s>>>(ncl, D.lvl_cl.p + D.lvl_ptr[t], D.cl_ptr.p, D.perm.p, D.rowptr.p,
                                                   D.colidx.p, D.rowslot.p, F.Lval.p, F.Dinv.p, X, F.yw.p, ld, ncols, ctrl);
    check_launch(h);
  }
  for (int t = nl - 1; t >= 0; --t) {
    const int ncl = D.lvl_ptr[t + 1] - D.lvl_ptr[t];
    k_gen_backward<B><<<(ncl + 3) / 4, 128, 0, s>>>(ncl, D.lvl_cl.p + D.lvl_ptr[t], D.cl_ptr.p, D.perm.p, D.colptr.p,
                                                    D.rowidx.p, F.Lval.p, F.Dinv.p, X, F.yw.p, ld, ncols, ctrl);
    check_launch(h);
  }
}

}  // namespace cora_b200
