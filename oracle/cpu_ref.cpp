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
// Preconditioner: Jacobi (src/CORA_problem.cpp:616-618,888-889) or none; the RegularizedCholesky
// path needs CHOLMOD and is restated only in the NumPy/SciPy oracle.
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

struct Ref {
  int d = 0, n = 0, m = 0, nt = 0;
  int64_t N = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<double> val, jac;  // jac = 1 / diag(Q)
  int precond = CORA_B200_PRECON_JACOBI;
  int64_t spmm_count = 0;
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

// Problem::precondition (src/CORA_problem.cpp:869-903), Jacobi / none
void precondition(const Ref &P, const Mat &V, int r, Mat &out) {
  const int64_t N = P.N;
  out.resize(V.size());
  if (P.precond == CORA_B200_PRECON_JACOBI) {
    for (int c = 0; c < r; ++c)
      pfor(N, [&](int64_t i_) { const int64_t i = (int64_t)i_; out[(size_t)c * N + i] = P.jac[i] * V[(size_t)c * N + i]; });
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
  if (P.precond != CORA_B200_PRECON_JACOBI && P.precond != CORA_B200_PRECON_NONE) {
    g_err = "cpu_ref restates the Jacobi preconditioner only";
    return 4;
  }
  tnt(P, r, X0, *p, X_out, res);
  return 0;
}

int64_t cpu_ref_spmm_count(void *h) { return ((Ref *)h)->spmm_count; }

}  // extern "C"
