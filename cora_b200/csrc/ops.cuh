// ops.cuh -- host-side launch helpers over the kernels (one stream per handle).
#pragma once
#include "handle.cuh"
#include "kernels.cuh"

namespace cora_b200 {

typedef cora_b200_handle H;
constexpr int kMaxGeomRank = 24;  // shared-memory bound of the staged-tile epilogues

inline void check_geom_rank(int r) {
  if (r < 1 || r > kMaxGeomRank)
    throw Error(CORA_B200_EINVAL, "relaxation rank must be in [1, 24], got " + std::to_string(r));
}

#define DISPATCH_D(h, ...)                                      \
  do {                                                          \
    if ((h)->DL.d == 2) { constexpr int DD = 2; __VA_ARGS__; }  \
    else { constexpr int DD = 3; __VA_ARGS__; }                 \
  } while (0)

inline void check_launch(H *h) {
  ++h->launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(CORA_B200_ECUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e));
}

inline int flat_grid(H *h, long long nE) {
  long long b = (nE + kThreads * 4 - 1) / (kThreads * 4);
  const long long cap = (long long)h->sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <int D>
inline size_t smem_q(const H *h, int r, int mode) {
  const int nvec = mode == QM_SPMM ? 1 : (mode == QM_GRAD ? 2 : 4);
  return qsmem_bytes<D>(h->DL.maxSlots, h->DL.TP, h->DL.TR, r, nvec);
}
inline size_t smem_vec(const H *h, int r, int nvec) {
  return ((size_t)nvec * ((size_t)h->DL.TR * (r | 1) + h->DL.TP) + 64) * sizeof(double);
}

template <typename K>
inline void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

inline void ensure_workspace(H *h, int r) {
  if (r <= 0) throw Error(CORA_B200_EINVAL, "rank must be positive");
  if (r > 64) throw Error(CORA_B200_EINVAL, "rank above 64 is not supported");
  if (r <= h->ws_r) return;
  const int cap = std::max(r + 2, 8);
  const size_t rows = (size_t)h->DL.numTiles * h->DL.TR;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int v = 0; v < V_COUNT; ++v) {
    if (v == V_X && h->resident_r > 0 && h->ws[v].p) {
      double *np_ = nullptr;
      CUDA_CHECK(cudaMalloc((void **)&np_, rows * cap * sizeof(double)));
      CUDA_CHECK(cudaMemcpy(np_, h->ws[v].p, (size_t)h->DL.N * h->resident_r * sizeof(double), cudaMemcpyDeviceToDevice));
      cudaFree(h->ws[v].p);
      h->ws[v].p = np_;
      h->ws[v].n = rows * cap;
    } else {
      h->ws[v].alloc(rows * cap);
    }
  }
  h->d_stage.alloc((size_t)h->DL.N * cap);
  h->d_longbuf.alloc((size_t)std::max(1, h->DL.numLong) * h->DL.D1 * cap);
  h->ws_r = cap;
  // opt in to large dynamic shared memory for this rank range
  const int gcap = std::min(cap, kMaxGeomRank);
  const size_t lim = 227 * 1024;
  DISPATCH_D(h, {
    allow_smem(k_qprod<DD>, std::min(lim, std::max(smem_q<DD>(h, gcap, QM_HESS), smem_q<DD>(h, cap, QM_SPMM))));
    allow_smem(k_cg_update<DD>, std::min(lim, smem_vec(h, gcap, 3)));
    allow_smem(k_retract<DD>, std::min(lim, smem_vec(h, gcap, 1)));
    allow_smem(k_tangent<DD>, std::min(lim, smem_vec(h, gcap, 2)));
  });
}

// Q X (+ epilogue) with the explicit data matrix.  Lalt: alternative value set on the same structure (S + eta I).
inline void launch_qprod_raw(H *h, int mode, const double *X, const double *Y, const double *G, double *out,
                             double *out2, int r, int post, int slot, CgCtrl *ctrl,
                             const DevLayout *Lalt = nullptr) {
  const DevLayout &L = Lalt ? *Lalt : h->DL;
  if (L.numLong > 0) {
    DISPATCH_D(h, k_long_groups<DD><<<L.numLong, kThreads, (size_t)L.D1 * kThreads * sizeof(double), h->stream>>>(
                      L, X, h->d_longbuf.p, r, ctrl));
    check_launch(h);
  }
  QArgs A{};
  A.X = X; A.Y = Y; A.G = G; A.out = out; A.out2 = out2;
  A.longbuf = h->d_longbuf.p;
  A.partials = h->d_partials.p; A.counter = h->d_counter.p; A.scal = h->d_scal.p;
  A.ctrl = ctrl; A.r = r; A.mode = mode; A.post = post; A.slot = slot;
  const bool prof = h->prof_on && mode == QM_HESS && ctrl != nullptr && h->prof_n + 2 <= h->prof_ev.size();
  if (prof) CUDA_CHECK(cudaEventRecord(h->prof_ev[h->prof_n++], h->stream));
  DISPATCH_D(h, k_qprod<DD><<<L.numTiles, kThreads, smem_q<DD>(h, r, mode), h->stream>>>(L, A));
  check_launch(h);
  if (prof) CUDA_CHECK(cudaEventRecord(h->prof_ev[h->prof_n++], h->stream));
}

// Formulation::Implicit (implicit.cuh): the translation-completed copy of X, and zeroing of translation rows
inline const double *implicit_complete(H *h, const double *X, int r, CgCtrl *ctrl);
inline void zero_translation_rows(H *h, double *V, int r, const CgCtrl *ctrl);

// Problem::dataMatrixProduct (src/CORA_problem.cpp:742-757) + epilogue in the handle's formulation.  Implicit:
// Qmain Y - T L^-1 T^T Y = rows of Q [Y; t*(Y)] -- the same fused kernel on the completed copy; the translation
// rows of the results (zero up to rounding) are cleared.  Products with an alternative value set (the certificate
// matrix, always the translation-explicit one: src/CORA_problem.cpp:1055-1059) are never completed.
inline void launch_qprod(H *h, int mode, const double *X, const double *Y, const double *G, double *out,
                         double *out2, int r, int post, int slot, CgCtrl *ctrl,
                         const DevLayout *Lalt = nullptr) {
  if (h->formulation != CORA_B200_FORMULATION_IMPLICIT || Lalt != nullptr) {
    launch_qprod_raw(h, mode, X, Y, G, out, out2, r, post, slot, ctrl, Lalt);
    return;
  }
  const double *Xc = implicit_complete(h, X, r, ctrl);
  launch_qprod_raw(h, mode, Xc, Y == X ? Xc : Y, G, out, out2, r, post, slot, ctrl, nullptr);
  if (out) zero_translation_rows(h, out, r, ctrl);
  if (out2) zero_translation_rows(h, out2, r, ctrl);
}

inline void launch_update(H *h, const UArgs &A0) {
  UArgs A = A0;
  A.partials = h->d_partials.p; A.counter = h->d_counter.p; A.scal = h->d_scal.p;
  DISPATCH_D(h, k_cg_update<DD><<<h->DL.numTiles, kThreads, smem_vec(h, A.r, 3), h->stream>>>(h->DL, A));
  check_launch(h);
}

// out = project(Y + alpha V) ; stores <V,V>, <Gr,V> at scal[slot..slot+1] when slot >= 0
inline void launch_retract(H *h, const double *Y, const double *V, double alpha, const double *Gr,
                           double *out, int r, int slot) {
  DISPATCH_D(h, k_retract<DD><<<h->DL.numTiles, kThreads, smem_vec(h, r, 1), h->stream>>>(
                    h->DL, Y, V, alpha, Gr, out, r, slot >= 0 ? h->d_partials.p : nullptr, h->d_counter.p,
                    h->d_scal.p, slot));
  check_launch(h);
}

inline void launch_tangent(H *h, const double *Y, const double *V, double *out, int r) {
  DISPATCH_D(h, k_tangent<DD><<<h->DL.numTiles, kThreads, smem_vec(h, r, 2), h->stream>>>(h->DL, Y, V, out, r));
  check_launch(h);
}

inline void launch_dot2(H *h, const double *a0, const double *b0, const double *a1, const double *b1,
                        long long nE, int slot) {
  int grid = flat_grid(h, nE);
  if (grid > h->DL.numTiles) grid = std::max(1, h->DL.numTiles);
  k_dot2<<<grid, kThreads, 0, h->stream>>>(a0, b0, a1, b1, nE, h->d_partials.p, h->d_counter.p, h->d_scal.p, slot);
  check_launch(h);
}

inline void launch_axpby(H *h, double a, const double *x, double b, const double *y, double *out, long long nE) {
  k_axpby<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(a, x, b, y, out, nE);
  check_launch(h);
}

// Host matrices are column-major with getExpectedVariableSize() rows (src/CORA_problem.cpp:944-954): N in the
// explicit formulation, d n + m (rotations and ranges) in the implicit one.  rows < 0: the handle's formulation.
inline void import_matrix(H *h, const double *host, int src_cols, double *dst, int r, int rows = -1) {
  const int io = rows < 0 ? h->io_rows() : rows;
  const size_t bytes = (size_t)io * src_cols * sizeof(double);
  CUDA_CHECK(cudaMemcpyAsync(h->d_stage.p, host, bytes, cudaMemcpyHostToDevice, h->stream));
  const long long nE = (long long)h->DL.N * r;
  k_import<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(h->DL.int2ref, h->d_stage.p, dst, h->DL.N, r, src_cols, io);
  check_launch(h);
}
inline void export_matrix(H *h, const double *src, int r, double *host, int rows = -1) {
  const int io = rows < 0 ? h->io_rows() : rows;
  const long long nE = (long long)h->DL.N * r;
  k_export<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(h->DL.int2ref, src, h->d_stage.p, h->DL.N, r, io);
  check_launch(h);
  CUDA_CHECK(cudaMemcpyAsync(host, h->d_stage.p, (size_t)io * r * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

inline void read_scal(H *h) {
  CUDA_CHECK(cudaMemcpyAsync(h->h_scal, h->d_scal.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

}  // namespace cora_b200
