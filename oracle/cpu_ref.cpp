// cpu_ref.cpp -- TEST INFRASTRUCTURE: a dependency-free C++17 restatement of the reference's CPU
// path (MarineRoboticsGroup/cora @ 015dc43) for the staircase inner loop.  It exists to (a) time
// "the reference's algorithm on the host cores" beside the CUDA path (bench.py cpu_baseline and
// `--impl reference`) and (b) cross-check oracle/cora_oracle.py.  Nothing under cora_b200/ links,
// loads or calls it.  The reference itself cannot be compiled here (Eigen3 and SuiteSparse are not
// installed), so this port keeps the reference's data layouts and its operation counts:
//   * data matrix: CSR int32 / f64, both triangles (Eigen::SparseMatrix<double,RowMajor>,
//     include/CORA/CORA_types.h:70); dense iterates: column-major N x r (Eigen::MatrixXd, :47-48);
//   * Q*Y one pass over Q per dense column, as Eigen's row-major-sparse x col-major-dense product
//     does (src/CORA_problem.cpp:746);
//   * per-pose loops for SymBlockDiagProduct / tangent projection / polar retraction
//     (src/StiefelProduct.cpp:8-55, src/ObliqueManifold.cpp:6-27, src/CORA_problem.cpp:782-938);
//   * STPCG (libs/Optimization/.../LinearAlgebra/IterativeSolvers.h:166-426) and TNT
//     (.../Riemannian/TNT.h:242-689) with the closures of src/CORA.cpp:52-122, including the
//     reference's redundant work: f and QM each do their own Q*Y (TNT.h:508,573), the model
//     decrease costs one more Hessian-vector product (:511-512), <r,v> is recomputed three times
//     per CG iteration (IterativeSolvers.h:290,341,408) and the metric forms the r x r product
//     V1^T V2 before taking its trace (src/CORA.cpp:119-122).
// Preconditioner: Jacobi (src/CORA_problem.cpp:616-618,888-889), none, or RegularizedCholesky
// (src/CORA_problem.cpp:544-614, src/CORA_preconditioners.cpp:46-83).  The reference hands
// (Q + lambda I)[0:N-1, 0:N-1] to CHOLMOD (absent here); for the graphs of the BASELINE configurations -- one
// odometry chain plus landmark / range factors -- a fill-free elimination order exists and the factor is
// restated directly: range rows first (each is a diagonal entry with two couplings), the poses as a
// block-tridiagonal Cholesky in trajectory order, the landmarks last through their dense Schur complement.
// Same matrix, same solution, no fill beyond the landmark border (what AMD + CHOLMOD produce on a chain);
// sequential like CHOLMOD's solve, one dense column per worker thread at most.
// Threads: the reference is single threaded.  cpu_ref_set_threads(T > 1) runs the row
// loops of the products and the per-pose loops on a small std::thread pool (this image has no
// libgomp) -- a stronger comparator than the reference; `cpu_ref_threads()` reports what is in use.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../include/cora_b200.h"  // cora_b200_tnt_params / cora_b200_tnt_result (plain C structs)

namespace {

// Cholesky factor of an [odometry chain + landmark border + range rows] matrix, the last row pinned to zero
// (src/CORA_preconditioners.cpp:77-80) when `pinned`.
struct ChainFactor {
  bool built = false, pos_def = false, pinned = false;
  int B = 0, n = 0, l = 0, m = 0;
  std::vector<double> rdinv;          // m: 1 / diagonal of the range rows
  std::vector<int32_t> rx;            // 2m: translation index of the two couplings (pose i -> i, landmark j -> n + j)
  std::vector<double> re;             // 2m: coupling values
  std::vector<double> Ld, G;          // n x B x B: Cholesky factors of the pivots, sub-diagonal blocks G_i = U_i^T L_i^-T
  std::vector<double> Y;              // (n B) x l column-major: L^-1 (border)
  std::vector<double> LC;             // l x l lower Cholesky factor of the landmark Schur complement
};

struct Ref {
  int d = 0, n = 0, m = 0, nt = 0;
  int64_t N = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<double> val, jac;  // jac = 1 / diag(Q)
  int precond = CORA_B200_PRECON_JACOBI;
  int64_t spmm_count = 0;
  ChainFactor chol;              // RegularizedCholesky: factor of (Q + lambda I) with the last row pinned
  double lambda_reg = 0.0;
};

typedef std::vector<double> Mat;  // column-major N x r

// ------------------------------------------------------------- tiny thread pool ----
class Pool {
 public:
  ~Pool() { resize(1); }
  int size() const { return nthreads_; }
  void resize(int n) {
    n = std::max(1, n);
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
    workers_.clear();
    stop_ = false;
    nthreads_ = n;
    for (int w = 1; w < n; ++w) workers_.emplace_back([this, w] { loop(w); });
  }
  // fn(begin, end, worker) over [0, n) split in contiguous chunks
  void run(int64_t n, const std::function<void(int64_t, int64_t, int)> &fn) {
    if (nthreads_ == 1 || n < 4096) { fn(0, n, 0); return; }
    {
      std::unique_lock<std::mutex> lk(mu_);
      fn_ = &fn; n_ = n; pending_ = nthreads_ - 1; ++epoch_;
    }
    cv_.notify_all();
    chunk(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void chunk(int w) {
    const int64_t b = n_ * w / nthreads_, e = n_ * (w + 1) / nthreads_;
    if (b < e) (*fn_)(b, e, w);
  }
  void loop(int w) {
    uint64_t seen = 0;
    {
      std::unique_lock<std::mutex> lk(mu_);
      seen = epoch_;
    }
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      chunk(w);
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t, int64_t, int)> *fn_ = nullptr;
  int64_t n_ = 0;
  int pending_ = 0, nthreads_ = 1;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};
Pool g_pool;
constexpr int kMaxWorkers = 256;

template <typename F>
inline void pfor(int64_t n, F &&body) {  // body(i)
  g_pool.run(n, [&](int64_t b, int64_t e, int) {
    for (int64_t i = b; i < e; ++i) body(i);
  });
}

// src/CORA_problem.cpp:742-757 (Explicit): out = Q * Y, one pass over Q per column
void data_matrix_product(Ref &P, const Mat &Y, int r, Mat &out) {
  const int64_t N = P.N;
  out.resize((size_t)N * r);
  for (int c = 0; c < r; ++c) {
    const double *y = Y.data() + (size_t)c * N;
    double *o = out.data() + (size_t)c * N;
    pfor(N, [&](int64_t i_) {
      const int64_t i = (int64_t)i_;
      double s = 0.0;
      for (int32_t k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k) s += P.val[k] * y[P.col[k]];
      o[i] = s;
    });
  }
  ++P.spmm_count;
}

// src/CORA.cpp:119-122 + MatrixManifold.h:55-59: trace(V1^T V2), forming the r x r product
double metric(const Ref &P, const Mat &A, const Mat &B, int r) {
  const int64_t N = P.N;
  double tr = 0.0;
  for (int a = 0; a < r; ++a)
    for (int b = 0; b < r; ++b) {
      const double *x = A.data() + (size_t)a * N, *y = B.data() + (size_t)b * N;
      double part[kMaxWorkers] = {0.0};
      g_pool.run(N, [&](int64_t b0, int64_t e0, int w) {
        double t = 0.0;
        for (int64_t i = b0; i < e0; ++i) t += x[i] * y[i];
        part[w] = t;
      });
      double s = 0.0;
      for (int w = 0; w < g_pool.size(); ++w) s += part[w];
      if (a == b) tr += s;
    }
  return tr;
}

inline double &at(Mat &M, int64_t N, int64_t i, int c) { return M[(size_t)c * N + i]; }
inline double at(const Mat &M, int64_t N, int64_t i, int c) { return M[(size_t)c * N + i]; }

// StiefelProduct::SymBlockDiagProduct (src/StiefelProduct.cpp:38-55) in row form:
// R_i -= sym(B_i C_i^T) A_i for the d x r blocks of pose i (A, B, C, R: N x r)
void sym_block_diag_sub(const Ref &P, const Mat &A, const Mat &B, const Mat &C, Mat &R, int r) {
  const int d = P.d;
  const int64_t N = P.N;
  pfor(P.n, [&](int64_t i_) {
    const int i = (int)i_;
    double S[9];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        double s = 0.0;
        for (int c = 0; c < r; ++c) s += at(B, N, (int64_t)d * i + a, c) * at(C, N, (int64_t)d * i + b, c);
        S[a * 3 + b] = s;
      }
    for (int a = 0; a < d; ++a)
      for (int b = a + 1; b < d; ++b) {
        const double s = 0.5 * (S[a * 3 + b] + S[b * 3 + a]);
        S[a * 3 + b] = S[b * 3 + a] = s;
      }
    for (int c = 0; c < r; ++c)
      for (int a = 0; a < d; ++a) {
        double s = 0.0;
        for (int b = 0; b < d; ++b) s += S[a * 3 + b] * at(A, N, (int64_t)d * i + b, c);
        at(R, N, (int64_t)d * i + a, c) -= s;
      }
  });
}

// Problem::tangent_space_projection (src/CORA_problem.cpp:782-820): out = proj_Y(V)
void tangent_projection(const Ref &P, const Mat &Y, const Mat &V, int r, Mat &out) {
  const int64_t N = P.N, dn = (int64_t)P.d * P.n;
  out = V;
  sym_block_diag_sub(P, Y, Y, V, out, r);  // V_i - sym(Y_i V_i^T) Y_i
  pfor(P.m, [&](int64_t k_) {
    const int k = (int)k_;
    double s = 0.0;
    for (int c = 0; c < r; ++c) s += at(Y, N, dn + k, c) * at(V, N, dn + k, c);
    for (int c = 0; c < r; ++c) at(out, N, dn + k, c) -= s * at(Y, N, dn + k, c);
  });
}

// Problem::Riemannian_Hessian_vector_product (src/CORA_problem.cpp:822-867)
void hessvec(Ref &P, const Mat &Y, const Mat &G, const Mat &Yd, int r, Mat &out) {
  const int64_t N = P.N, dn = (int64_t)P.d * P.n;
  Mat W;
  data_matrix_product(P, Yd, r, W);
  sym_block_diag_sub(P, Yd, Y, G, W, r);  // W_i -= sym(Y_i G_i^T) Yd_i
  pfor(P.m, [&](int64_t k_) {
    const int k = (int)k_;
    double s = 0.0;
    for (int c = 0; c < r; ++c) s += at(G, N, dn + k, c) * at(Y, N, dn + k, c);
    for (int c = 0; c < r; ++c) at(W, N, dn + k, c) -= s * at(Yd, N, dn + k, c);
  });
  tangent_projection(P, Y, W, r, out);
}

// ------------------------------------------------- chain Cholesky (RegularizedCholesky) ----
// Dense helpers on small row-major B x B blocks.
static bool chol_lower(double *A, int B) {  // in place, lower triangle; false: not positive definite
  for (int j = 0; j < B; ++j) {
    double s = A[j * B + j];
    for (int k = 0; k < j; ++k) s -= A[j * B + k] * A[j * B + k];
    if (!(s > 0.0)) return false;
    const double dj = std::sqrt(s);
    A[j * B + j] = dj;
    for (int i = j + 1; i < B; ++i) {
      double t = A[i * B + j];
      for (int k = 0; k < j; ++k) t -= A[i * B + k] * A[j * B + k];
      A[i * B + j] = t / dj;
    }
    for (int k = j + 1; k < B; ++k) A[j * B + k] = 0.0;
  }
  return true;
}
static void fwd_lower(const double *L, int B, double *x) {  // x <- L^-1 x
  for (int i = 0; i < B; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= L[i * B + k] * x[k];
    x[i] = s / L[i * B + i];
  }
}
static void bwd_lower(const double *L, int B, double *x) {  // x <- L^-T x
  for (int i = B - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < B; ++k) s -= L[k * B + i] * x[k];
    x[i] = s / L[i * B + i];
  }
}

// Factor M = A + shift I restricted to rows/columns 0..N-1 (pin_last: the last row and column replaced by the
// identity), A given as CSR in the reference row order [d n rotations | m ranges | n translations | l landmarks].
// Returns false (with a message) when the graph is not a chain + landmark border.
static bool chain_factor(ChainFactor &F, int d, int n, int m, int nt, const int32_t *rowptr, const int32_t *col,
                         const double *val, double shift, bool pin_last, std::string &err) {
  const int B = d + 1, l = nt - n, BB = B * B;
  const int64_t dn = (int64_t)d * n, t0 = dn + m, N = dn + m + nt;
  F = ChainFactor();
  F.B = B; F.n = n; F.l = l; F.m = m; F.pinned = pin_last;
  // unknowns of pose i: rotation rows d i .. d i + d - 1, translation t0 + i  -> local index 0..d-1, d
  auto pose_of = [&](int64_t row, int &a) -> int {
    if (row < dn) { a = (int)(row % d); return (int)(row / d); }
    if (row >= t0 && row < t0 + n) { a = d; return (int)(row - t0); }
    return -1;
  };
  std::vector<double> A((size_t)std::max(n, 1) * BB, 0.0), U((size_t)std::max(n, 1) * BB, 0.0);
  std::vector<double> C((size_t)std::max(l, 1) * std::max(l, 1), 0.0), Bd((size_t)std::max(n, 1) * B * std::max(l, 1), 0.0);
  F.rdinv.assign(std::max(m, 1), 0.0); F.rx.assign((size_t)std::max(m, 1) * 2, -1); F.re.assign((size_t)std::max(m, 1) * 2, 0.0);
  bool pd = true;
  for (int64_t row = 0; row < N; ++row) {
    int a = 0;
    const int i = pose_of(row, a);
    for (int32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const int64_t c = col[k];
      const double v = val[k];
      int b = 0;
      const int j = pose_of(c, b);
      if (i >= 0) {                                   // pose row
        if (j >= 0) {
          if (j == i) A[(size_t)i * BB + a * B + b] += v;
          else if (j == i + 1) U[(size_t)i * BB + a * B + b] += v;
          else if (j == i - 1) { /* transpose of U[i-1] */ }
          else if (v != 0.0) { err = "not an odometry chain: a pose is coupled to a non-adjacent pose"; return false; }
        } else if (c >= t0 + n) {
          Bd[((size_t)(c - t0 - n)) * n * B + (size_t)i * B + a] += v;   // column-major by landmark
        } else { /* range column: taken from the range row */ }
      } else if (row >= t0 + n) {                     // landmark row
        if (c >= t0 + n) C[(size_t)(row - t0 - n) * l + (c - t0 - n)] += v;
      } else {                                        // range row k
        const int kk = (int)(row - dn);
        if (c == row) { F.rdinv[kk] += v; continue; }
        int x = -1;
        if (c >= t0 && c < t0 + n) x = (int)(c - t0);
        else if (c >= t0 + n) x = n + (int)(c - t0 - n);
        if (x < 0) { err = "a range row is coupled to a non-translation variable"; return false; }
        int slot = F.rx[(size_t)kk * 2] < 0 ? 0 : (F.rx[(size_t)kk * 2 + 1] < 0 ? 1 : 2);
        if (slot == 2) { err = "a range row has more than two couplings"; return false; }
        F.rx[(size_t)kk * 2 + slot] = x;
        F.re[(size_t)kk * 2 + slot] = v;
      }
    }
  }
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < B; ++a) A[(size_t)i * BB + a * B + a] += shift;
  for (int j = 0; j < l; ++j) C[(size_t)j * l + j] += shift;
  // eliminate the range rows: Schur complement onto their translations
  for (int k = 0; k < m; ++k) {
    const double delta = F.rdinv[k] + shift;
    if (!(delta > 0.0)) { pd = false; F.rdinv[k] = 1.0; continue; }
    F.rdinv[k] = 1.0 / delta;
    for (int p = 0; p < 2; ++p)
      for (int q = 0; q < 2; ++q) {
        const int x = F.rx[(size_t)k * 2 + p], y = F.rx[(size_t)k * 2 + q];
        if (x < 0 || y < 0) continue;
        const double v = -F.re[(size_t)k * 2 + p] * F.re[(size_t)k * 2 + q] * F.rdinv[k];
        if (x < n && y < n) {
          if (x == y) A[(size_t)x * BB + d * B + d] += v;
          else if (y == x + 1) U[(size_t)x * BB + d * B + d] += v;
          else if (y == x - 1) { }
          else { err = "a range factor joins two non-adjacent poses"; return false; }
        } else if (x < n) {
          Bd[(size_t)(y - n) * n * B + (size_t)x * B + d] += v;
        } else if (y >= n) {
          C[(size_t)(x - n) * l + (y - n)] += v;
        }
      }
  }
  // pin the last unknown
  if (pin_last) {
    if (l > 0) {
      const int j = l - 1;
      for (int q = 0; q < l; ++q) { C[(size_t)j * l + q] = 0.0; C[(size_t)q * l + j] = 0.0; }
      C[(size_t)j * l + j] = 1.0;
      for (size_t q = 0; q < (size_t)n * B; ++q) Bd[(size_t)j * n * B + q] = 0.0;
    } else if (n > 0) {
      const int i = n - 1;
      for (int a = 0; a < B; ++a) { A[(size_t)i * BB + d * B + a] = 0.0; A[(size_t)i * BB + a * B + d] = 0.0; }
      A[(size_t)i * BB + d * B + d] = 1.0;
      if (i > 0) for (int a = 0; a < B; ++a) U[(size_t)(i - 1) * BB + a * B + d] = 0.0;
    }
  }
  // block-tridiagonal Cholesky: S_i = A_i - G_{i-1} G_{i-1}^T, L_i = chol(S_i), G_i = U_i^T L_i^-T
  F.Ld.assign((size_t)std::max(n, 1) * BB, 0.0); F.G.assign((size_t)std::max(n, 1) * BB, 0.0);
  for (int i = 0; i < n; ++i) {
    double *S = &F.Ld[(size_t)i * BB];
    for (int e = 0; e < BB; ++e) S[e] = A[(size_t)i * BB + e];
    if (i > 0) {
      const double *Gp = &F.G[(size_t)(i - 1) * BB];
      for (int a = 0; a < B; ++a)
        for (int b = 0; b < B; ++b) {
          double t = 0.0;
          for (int k = 0; k < B; ++k) t += Gp[a * B + k] * Gp[b * B + k];
          S[a * B + b] -= t;
        }
    }
    if (!chol_lower(S, B)) {
      pd = false;
      for (int e = 0; e < BB; ++e) S[e] = 0.0;
      for (int a = 0; a < B; ++a) S[a * B + a] = 1.0;
    }
    if (i + 1 < n) {  // G_i row a = L_i^-1 (column a of U_i):  G_i = U_i^T L_i^-T  <=>  G_i^T = L_i^-1 U_i
      double *Gi = &F.G[(size_t)i * BB];
      for (int a = 0; a < B; ++a) {
        double x[4];
        for (int k = 0; k < B; ++k) x[k] = U[(size_t)i * BB + k * B + a];
        fwd_lower(S, B, x);
        for (int k = 0; k < B; ++k) Gi[a * B + k] = x[k];
      }
    }
  }
  // border: Y = L^-1 Bd, landmark Schur complement C - Y^T Y
  if (l > 0) {
    F.Y = Bd;
    for (int j = 0; j < l; ++j) {
      double *y = &F.Y[(size_t)j * n * B];
      for (int i = 0; i < n; ++i) {
        double *yi = y + (size_t)i * B;
        if (i > 0) {
          const double *Gp = &F.G[(size_t)(i - 1) * BB], *yp = y + (size_t)(i - 1) * B;
          for (int a = 0; a < B; ++a) {
            double t = 0.0;
            for (int k = 0; k < B; ++k) t += Gp[a * B + k] * yp[k];
            yi[a] -= t;
          }
        }
        fwd_lower(&F.Ld[(size_t)i * BB], B, yi);
      }
    }
    F.LC = C;
    for (int a = 0; a < l; ++a)
      for (int b = 0; b <= a; ++b) {
        const double *ya = &F.Y[(size_t)a * n * B], *yb = &F.Y[(size_t)b * n * B];
        double t = 0.0;
        for (size_t q = 0; q < (size_t)n * B; ++q) t += ya[q] * yb[q];
        F.LC[(size_t)a * l + b] -= t;
        F.LC[(size_t)b * l + a] = F.LC[(size_t)a * l + b];
      }
    if (!chol_lower(F.LC.data(), l)) {
      pd = false;
      std::fill(F.LC.begin(), F.LC.end(), 0.0);
      for (int a = 0; a < l; ++a) F.LC[(size_t)a * l + a] = 1.0;
    }
  }
  F.pos_def = pd;
  F.built = true;
  return true;
}

// out = M^-1 V with the last row of the result zero (blockCholeskySolve, src/CORA_preconditioners.cpp:46-83)
static void chain_solve(const Ref &P, const Mat &V, int r, Mat &out) {
  const ChainFactor &F = P.chol;
  const int B = F.B, n = F.n, l = F.l, m = F.m, d = P.d, BB = B * B;
  const int64_t N = P.N, dn = (int64_t)d * n, t0 = dn + m;
  out.assign(V.size(), 0.0);
  auto trow = [&](int x) -> int64_t { return x < n ? t0 + x : t0 + n + (x - n); };
  pfor(r, [&](int64_t c_) {
    const int c = (int)c_;
    const double *v = &V[(size_t)c * N];
    double *z = &out[(size_t)c * N];
    std::vector<double> u((size_t)std::max(n, 1) * B), w((size_t)std::max(l, 1));
    // right-hand side on [poses | landmarks] after eliminating the ranges
    for (int i = 0; i < n; ++i) {
      for (int a = 0; a < d; ++a) u[(size_t)i * B + a] = v[(size_t)d * i + a];
      u[(size_t)i * B + d] = v[t0 + i];
    }
    for (int j = 0; j < l; ++j) w[j] = v[t0 + n + j];
    for (int k = 0; k < m; ++k)
      for (int p = 0; p < 2; ++p) {
        const int x = F.rx[(size_t)k * 2 + p];
        if (x < 0) continue;
        const double t = F.re[(size_t)k * 2 + p] * F.rdinv[k] * v[dn + k];
        if (x < n) u[(size_t)x * B + d] -= t; else w[x - n] -= t;
      }
    if (F.pinned) { if (l > 0) w[l - 1] = 0.0; else if (n > 0) u[(size_t)(n - 1) * B + d] = 0.0; }
    // forward: u <- L^-1 u
    for (int i = 0; i < n; ++i) {
      double *ui = &u[(size_t)i * B];
      if (i > 0) {
        const double *Gp = &F.G[(size_t)(i - 1) * BB], *up = &u[(size_t)(i - 1) * B];
        for (int a = 0; a < B; ++a) {
          double t = 0.0;
          for (int k = 0; k < B; ++k) t += Gp[a * B + k] * up[k];
          ui[a] -= t;
        }
      }
      fwd_lower(&F.Ld[(size_t)i * BB], B, ui);
    }
    // landmarks: w <- LC^-T LC^-1 (w - Y^T u)
    if (l > 0) {
      for (int j = 0; j < l; ++j) {
        const double *y = &F.Y[(size_t)j * n * B];
        double t = 0.0;
        for (size_t q = 0; q < (size_t)n * B; ++q) t += y[q] * u[q];
        w[j] -= t;
      }
      fwd_lower(F.LC.data(), l, w.data());
      bwd_lower(F.LC.data(), l, w.data());
      for (int j = 0; j < l; ++j) {
        const double *y = &F.Y[(size_t)j * n * B];
        const double wj = w[j];
        for (size_t q = 0; q < (size_t)n * B; ++q) u[q] -= y[q] * wj;
      }
    }
    // backward: u <- L^-T u
    for (int i = n - 1; i >= 0; --i) {
      double *ui = &u[(size_t)i * B];
      if (i + 1 < n) {
        const double *Gi = &F.G[(size_t)i * BB], *un = &u[(size_t)(i + 1) * B];
        for (int a = 0; a < B; ++a) {
          double t = 0.0;
          for (int k = 0; k < B; ++k) t += Gi[k * B + a] * un[k];
          ui[a] -= t;
        }
      }
      bwd_lower(&F.Ld[(size_t)i * BB], B, ui);
    }
    for (int i = 0; i < n; ++i) {
      for (int a = 0; a < d; ++a) z[(size_t)d * i + a] = u[(size_t)i * B + a];
      z[t0 + i] = u[(size_t)i * B + d];
    }
    for (int j = 0; j < l; ++j) z[t0 + n + j] = w[j];
    // back-substitute the ranges
    for (int k = 0; k < m; ++k) {
      double s = v[dn + k];
      for (int p = 0; p < 2; ++p) {
        const int x = F.rx[(size_t)k * 2 + p];
        if (x >= 0) s -= F.re[(size_t)k * 2 + p] * z[trow(x)];
      }
      z[dn + k] = s * F.rdinv[k];
    }
    if (F.pinned) z[N - 1] = 0.0;
  });
}

// Problem::precondition (src/CORA_problem.cpp:869-903)
void precondition(const Ref &P, const Mat &V, int r, Mat &out) {
  const int64_t N = P.N;
  out.resize(V.size());
  if (P.precond == CORA_B200_PRECON_JACOBI) {
    for (int c = 0; c < r; ++c)
      pfor(N, [&](int64_t i_) { const int64_t i = (int64_t)i_; out[(size_t)c * N + i] = P.jac[i] * V[(size_t)c * N + i]; });
  } else if (P.precond == CORA_B200_PRECON_REG_CHOLESKY) {
    chain_solve(P, V, r, out);
  } else {
    out = V;
  }
}

// Polar factor of the d x r block W (rows) by one-sided Jacobi: the row form of the thin SVD
// U V^T of StiefelProduct.cpp:26-34.
void polar_rows(double *W, int d, int r) {  // W[a*r + c]
  double U[9];
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) U[a * 3 + b] = a == b ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < d; ++p)
      for (int q = p + 1; q < d; ++q) {
        double app = 0, aqq = 0, apq = 0;
        for (int c = 0; c < r; ++c) {
          app += W[p * r + c] * W[p * r + c];
          aqq += W[q * r + c] * W[q * r + c];
          apq += W[p * r + c] * W[q * r + c];
        }
        if (apq == 0.0 || std::fabs(apq) <= 1e-16 * std::sqrt(app * aqq)) continue;
        rotated = true;
        const double th = (aqq - app) / (2.0 * apq);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int c = 0; c < r; ++c) {
          const double x = W[p * r + c], y = W[q * r + c];
          W[p * r + c] = cs * x - sn * y;
          W[q * r + c] = sn * x + cs * y;
        }
        for (int k = 0; k < d; ++k) {
          const double x = U[k * 3 + p], y = U[k * 3 + q];
          U[k * 3 + p] = cs * x - sn * y;
          U[k * 3 + q] = sn * x + cs * y;
        }
      }
    if (!rotated) break;
  }
  double T[3 * 64];
  for (int a = 0; a < d; ++a) {
    double s = 0.0;
    for (int c = 0; c < r; ++c) s += W[a * r + c] * W[a * r + c];
    const double inv = 1.0 / std::sqrt(std::max(s, 1e-300));
    for (int c = 0; c < r; ++c) T[a * r + c] = W[a * r + c] * inv;
  }
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < r; ++c) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += U[a * 3 + k] * T[k * r + c];
      W[a * r + c] = s;
    }
}

// Problem::projectToManifold (src/CORA_problem.cpp:905-934)
void project_to_manifold(const Ref &P, const Mat &A, int r, Mat &out) {
  const int d = P.d;
  const int64_t N = P.N, dn = (int64_t)d * P.n;
  out = A;
  pfor(P.n, [&](int64_t i_) {
    const int i = (int)i_;
    double W[3 * 64];
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < r; ++c) W[a * r + c] = at(A, N, (int64_t)d * i + a, c);
    polar_rows(W, d, r);
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < r; ++c) at(out, N, (int64_t)d * i + a, c) = W[a * r + c];
  });
  pfor(P.m, [&](int64_t k_) {
    const int k = (int)k_;
    double s = 0.0;
    for (int c = 0; c < r; ++c) s += at(A, N, dn + k, c) * at(A, N, dn + k, c);
    const double inv = 1.0 / std::sqrt(s);
    for (int c = 0; c < r; ++c) at(out, N, dn + k, c) *= inv;
  });
}

void axpy(Mat &y, double a, const Mat &x) {
  const int64_t n = (int64_t)y.size();
  pfor(n, [&](int64_t i_) { const int64_t i = (int64_t)i_; y[i] += a * x[i]; });
}

// STPCG, IterativeSolvers.h:166-426.  Returns ||s||_M; *iters = CG iterations.
double stpcg(Ref &P, const Mat &Y, const Mat &G, const Mat &g, int r, double Delta, int max_it, double kappa_fgr,
             double theta, Mat &s, int *iters) {
  const double eps = 1e-8;
  auto H = [&](const Mat &v, Mat &out) { hessvec(P, Y, G, v, r, out); };
  auto Pop = [&](const Mat &v, Mat &out) {  // src/CORA.cpp:89-92
    Mat t;
    precondition(P, v, r, t);
    tangent_projection(P, Y, t, r, out);
  };
  s.assign(g.size(), 0.0);
  Mat rr = g, v, p, Hp;
  Pop(rr, v);
  p = v;
  for (double &x : p) x = -x;
  double sMp = 0.0, sM2 = 0.0, pM2 = metric(P, rr, v, r);
  const double Delta2 = Delta * Delta;
  const double r0 = std::sqrt(metric(P, rr, v, r));
  const double target = r0 * std::min(kappa_fgr, std::pow(r0, theta));  // :278-279
  int it = 0;
  while (it < max_it) {
    if (std::sqrt(metric(P, rr, v, r)) <= target) break;  // :290
    H(p, Hp);                                              // :294
    const double kappa = metric(P, p, Hp, r);
    if (std::sqrt(metric(P, Hp, Hp, r)) / std::sqrt(metric(P, p, p, r)) < eps) {  // :305-338
      double sg = 1.0;
      if (metric(P, p, rr, r) < 0) { sg = -1.0; sMp = -sMp; }
      const double sigma = (-sMp + std::sqrt(sMp * sMp + pM2 * (Delta2 - sM2))) / pM2;
      axpy(s, sg * sigma, p);
      *iters = it;
      return Delta;
    }
    const double alpha = metric(P, rr, v, r) / kappa;  // :341
    const double sM2n = sM2 + 2 * alpha * sMp + alpha * alpha * pM2;
    if (kappa <= 0 || sM2n > Delta2) {  // :347-362
      const double sigma = (-sMp + std::sqrt(sMp * sMp + pM2 * (Delta2 - sM2))) / pM2;
      axpy(s, sigma, p);
      *iters = it;
      return Delta;
    }
    axpy(s, alpha, p);    // :374
    axpy(rr, alpha, Hp);  // :377
    Pop(rr, v);           // :386
    const double rv = metric(P, rr, v, r);
    const double beta = rv / (alpha * kappa);  // :412
    sM2 = sM2n;
    sMp = beta * (sMp + alpha * pM2);
    pM2 = rv + beta * beta * pM2;
    const int64_t n = (int64_t)p.size();
    pfor(n, [&](int64_t i) { p[i] = -v[i] + beta * p[i]; });  // :420
    ++it;
  }
  *iters = it;
  return std::sqrt(sM2);
}

struct Trace {
  cora_b200_tnt_result *res;
  int ns = 0, ni = 0;
  void state(double t, double f, double g, double pg, double D) {
    if (ns < res->trace_capacity) {
      if (res->time) res->time[ns] = t;
      if (res->objective_values) res->objective_values[ns] = f;
      if (res->gradient_norms) res->gradient_norms[ns] = g;
      if (res->preconditioned_gradient_norms) res->preconditioned_gradient_norms[ns] = pg;
      if (res->trust_region_radius) res->trust_region_radius[ns] = D;
    }
    ++ns;
  }
  void iter(int inner, double hn, double hM, double rho) {
    if (ni < res->trace_capacity) {
      if (res->inner_iterations) res->inner_iterations[ni] = inner;
      if (res->update_step_norms) res->update_step_norms[ni] = hn;
      if (res->update_step_M_norms) res->update_step_M_norms[ni] = hM;
      if (res->gain_ratios) res->gain_ratios[ni] = rho;
    }
    ++ni;
  }
};

// TNT.h:242-689 with the closures of src/CORA.cpp:52-122
void tnt(Ref &P, int r, const double *X0, const cora_b200_tnt_params &prm, double *X_out, cora_b200_tnt_result *res) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto elapsed = [&]() { return std::chrono::duration<double>(clk::now() - t0).count(); };
  const size_t NE = (size_t)P.N * r;
  Mat x(X0, X0 + NE), G, grad, pg, tmp, h, xp, Hh;
  auto f = [&](const Mat &y) {  // src/CORA.cpp:52-55 -> evaluateObjective :759-762 (its own Q*Y)
    Mat qy;
    data_matrix_product(P, y, r, qy);
    return 0.5 * metric(P, y, qy, r);
  };
  auto QM = [&](const Mat &y) {  // :58-75
    data_matrix_product(P, y, r, G);
    tangent_projection(P, y, G, r, grad);
  };
  auto precon = [&](const Mat &y, const Mat &v, Mat &out) {
    precondition(P, v, r, tmp);
    tangent_projection(P, y, tmp, r, out);
  };
  const double sqrt_eps = std::sqrt(2.220446049250313e-16);
  Trace tw{res};
  double fx = f(x);
  QM(x);
  double gnorm = std::sqrt(metric(P, grad, grad, r));
  precon(x, grad, pg);
  double pgnorm = std::sqrt(metric(P, pg, pg, r));
  double Delta = prm.Delta0;
  int status = CORA_B200_TNT_ITERATION_LIMIT;
  int64_t total_inner = 0;
  for (int iteration = 0; iteration < prm.max_iterations; ++iteration) {
    const double el = elapsed();
    if (prm.max_computation_time > 0 && el > prm.max_computation_time) { status = CORA_B200_TNT_ELAPSED_TIME; break; }
    tw.state(el, fx, gnorm, pgnorm, Delta);
    if (gnorm < prm.gradient_tolerance) { status = CORA_B200_TNT_GRADIENT; break; }
    if (pgnorm < prm.preconditioned_gradient_tolerance) { status = CORA_B200_TNT_PRECONDITIONED_GRADIENT; break; }
    int inner = 0;
    const double hM = stpcg(P, x, G, grad, r, Delta, prm.max_TPCG_iterations, prm.kappa_fgr, prm.theta, h, &inner);
    total_inner += inner;
    const double hnorm = std::sqrt(metric(P, h, h, r));
    tmp = x;
    axpy(tmp, 1.0, h);
    project_to_manifold(P, tmp, r, xp);  // retract :505
    const double fp = f(xp);             // :508
    hessvec(P, x, G, h, r, Hh);          // :511-512
    const double dm = -metric(P, grad, h, r) - 0.5 * metric(P, h, Hh, r);
    const double df = fx - fp;
    const double rel = df / (sqrt_eps + std::fabs(fx));
    const double rho = df / dm;
    const bool accepted = !std::isnan(rho) && rho > prm.eta1;  // :532
    tw.iter(inner, hnorm, hM, rho);
    if (accepted) {
      x.swap(xp);
      fx = fp;
      if (rel < prm.relative_decrease_tolerance) { status = CORA_B200_TNT_RELATIVE_DECREASE; break; }
      if (hnorm < prm.stepsize_tolerance) { status = CORA_B200_TNT_STEPSIZE; break; }
      QM(x);  // :573
      gnorm = std::sqrt(metric(P, grad, grad, r));
      precon(x, grad, pg);
      pgnorm = std::sqrt(metric(P, pg, pg, r));
    }
    if (!std::isnan(rho) && rho >= prm.eta2) {
      Delta = std::max(prm.alpha2 * hM, Delta);
    } else if (std::isnan(rho) || rho < prm.eta1) {
      Delta = prm.alpha1 * hM;
      if (Delta < prm.Delta_tolerance) { status = CORA_B200_TNT_TRUST_REGION; break; }
    }
  }
  const double el = elapsed();
  tw.state(el, fx, gnorm, pgnorm, Delta);
  std::memcpy(X_out, x.data(), NE * sizeof(double));
  res->f = fx;
  res->gradfx_norm = gnorm;
  res->preconditioned_gradfx_norm = pgnorm;
  res->elapsed_time = el;
  res->device_time = 0.0;
  res->status = status;
  res->num_outer = tw.ni;
  res->total_inner = total_inner;
  res->kernel_launches = 0;
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char *cpu_ref_last_error(void) { return g_err.c_str(); }

int cpu_ref_threads(void) { return g_pool.size(); }

void cpu_ref_set_threads(int n) { g_pool.resize(std::min(std::max(n, 1), kMaxWorkers)); }

int cpu_ref_create(void **out, int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                   const int32_t *col, const double *val, int64_t nnz, int preconditioner) {
  if (!out || !rowptr || (d != 2 && d != 3)) { g_err = "bad argument"; return 1; }
  Ref *P = new Ref();
  P->d = d; P->n = n_poses; P->m = n_ranges; P->nt = n_trans;
  P->N = (int64_t)d * n_poses + n_ranges + n_trans;
  P->rowptr.assign(rowptr, rowptr + P->N + 1);
  P->col.assign(col, col + nnz);
  P->val.assign(val, val + nnz);
  P->precond = preconditioner;
  P->jac.assign((size_t)P->N, 0.0);
  for (int64_t i = 0; i < P->N; ++i) {
    double dg = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) dg += val[k];
    P->jac[i] = 1.0 / dg;  // src/CORA_problem.cpp:616-618
  }
  *out = P;
  return 0;
}

void cpu_ref_destroy(void *h) { delete (Ref *)h; }

int cpu_ref_data_matrix_product(void *h, int r, const double *Y, double *out) {
  Ref &P = *(Ref *)h;
  Mat y(Y, Y + (size_t)P.N * r), o;
  data_matrix_product(P, y, r, o);
  std::memcpy(out, o.data(), o.size() * sizeof(double));
  return 0;
}

int cpu_ref_hessvec(void *h, int r, const double *Y, const double *G, const double *Yd, double *out) {
  Ref &P = *(Ref *)h;
  const size_t NE = (size_t)P.N * r;
  Mat y(Y, Y + NE), g(G, G + NE), yd(Yd, Yd + NE), o;
  hessvec(P, y, g, yd, r, o);
  std::memcpy(out, o.data(), NE * sizeof(double));
  return 0;
}

int cpu_ref_tangent_proj(void *h, int r, const double *Y, const double *V, double *out) {
  Ref &P = *(Ref *)h;
  const size_t NE = (size_t)P.N * r;
  Mat y(Y, Y + NE), v(V, V + NE), o;
  tangent_projection(P, y, v, r, o);
  std::memcpy(out, o.data(), NE * sizeof(double));
  return 0;
}

int cpu_ref_project(void *h, int r, const double *A, double *out) {
  Ref &P = *(Ref *)h;
  const size_t NE = (size_t)P.N * r;
  if (r > 64) { g_err = "rank above 64 not supported"; return 1; }
  Mat a(A, A + NE), o;
  project_to_manifold(P, a, r, o);
  std::memcpy(out, o.data(), NE * sizeof(double));
  return 0;
}

int cpu_ref_tnt(void *h, int r, const double *X0, const cora_b200_tnt_params *p, double *X_out,
                cora_b200_tnt_result *res) {
  if (!h || !X0 || !p || !X_out || !res || r < 1 || r > 64) { g_err = "bad argument"; return 1; }
  Ref &P = *(Ref *)h;
  if (P.precond == CORA_B200_PRECON_REG_CHOLESKY && !P.chol.built) {
    g_err = "RegularizedCholesky: call cpu_ref_set_reg_cholesky(lambda) first";
    return 4;
  }
  if (P.precond != CORA_B200_PRECON_JACOBI && P.precond != CORA_B200_PRECON_NONE && P.precond != CORA_B200_PRECON_REG_CHOLESKY) {
    g_err = "cpu_ref restates Jacobi, RegularizedCholesky and no preconditioner";
    return 4;
  }
  tnt(P, r, X0, *p, X_out, res);
  return 0;
}

int64_t cpu_ref_spmm_count(void *h) { return ((Ref *)h)->spmm_count; }

// RegularizedCholesky with regularisation lambda (the caller passes the reference's lambda = ||Q||_2 / (c - 1),
// src/CORA_problem.cpp:556-591): factor (Q + lambda I) with the last row pinned and select the preconditioner.
int cpu_ref_set_reg_cholesky(void *h, double lambda) {
  Ref &P = *(Ref *)h;
  std::string err;
  if (!chain_factor(P.chol, P.d, P.n, P.m, P.nt, P.rowptr.data(), P.col.data(), P.val.data(), lambda, true, err)) {
    g_err = err;
    return 2;
  }
  if (!P.chol.pos_def) { g_err = "Q + lambda I is not positive definite"; return 3; }
  P.lambda_reg = lambda;
  P.precond = CORA_B200_PRECON_REG_CHOLESKY;
  return 0;
}

int cpu_ref_precondition(void *h, int r, const double *V, double *out) {
  Ref &P = *(Ref *)h;
  Mat v(V, V + (size_t)P.N * r), o;
  precondition(P, v, r, o);
  std::memcpy(out, o.data(), o.size() * sizeof(double));
  return 0;
}

// Positive-definiteness test of S + shift I for a symmetric matrix S on the same kind of graph, by Cholesky
// (the PSD half of fast_verification, src/CORA_utils.cpp:33-57).  *pos_def = 1 / 0.
int cpu_ref_chain_posdef(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr, const int32_t *col,
                         const double *val, double shift, int *pos_def) {
  ChainFactor F;
  std::string err;
  if (!chain_factor(F, d, n_poses, n_ranges, n_trans, rowptr, col, val, shift, false, err)) { g_err = err; return 2; }
  *pos_def = F.pos_def ? 1 : 0;
  return 0;
}

}  // extern "C"
