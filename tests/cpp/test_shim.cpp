// Exercises include/cora_b200.hpp (the C++ mirror of CORA::Problem / solveCORA) against a tiny
// SE(2) chain + ranges problem assembled through the C-ABI.  Without a CUDA device it checks the
// error path (std::runtime_error, no CPU fallback) and exits 0; with one it checks f, grad and a
// full staircase solve against values computed on the host in this file.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/cora_b200.hpp"

using namespace cora_b200;

static int fails = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } \
  } while (0)

int main() {
  const int d = 2, n = 40, l = 2;
  // ground truth: unit steps with a slow turn; odometry measured exactly (cost 0 at ground truth)
  std::vector<double> th(n), tx(n), ty(n);
  th[0] = tx[0] = ty[0] = 0;
  for (int i = 1; i < n; ++i) {
    th[i] = th[i - 1] + 0.15;
    tx[i] = tx[i - 1] + std::cos(th[i - 1]);
    ty[i] = ty[i - 1] + std::sin(th[i - 1]);
  }
  const double L[2][2] = {{3.0, 4.0}, {-2.0, 6.0}};
  std::vector<int64_t> rp_i, rp_j, rg_a, rg_b;
  std::vector<double> rp_t, rp_tau, rot_R, rot_kappa, rg_r, rg_w;
  for (int i = 0; i + 1 < n; ++i) {
    rp_i.push_back(i); rp_j.push_back(i + 1);
    rp_t.push_back(1.0); rp_t.push_back(0.0);  // body-frame step
    rp_tau.push_back(100.0);
    const double c = std::cos(0.15), s = std::sin(0.15);
    rot_R.insert(rot_R.end(), {c, -s, s, c});
    rot_kappa.push_back(50.0);
  }
  for (int i = 0; i < n; i += 3) {
    const int j = (i / 3) % l;
    rg_a.push_back(i); rg_b.push_back(n + j);
    rg_r.push_back(std::hypot(tx[i] - L[j][0], ty[i] - L[j][1]));
    rg_w.push_back(10.0);
  }
  const int64_t E = (int64_t)rp_tau.size(), m = (int64_t)rg_w.size();
  int64_t nnz = 0;
  check(cora_b200_assemble(d, n, l, E, rp_i.data(), rp_j.data(), rp_t.data(), rp_tau.data(), E, rp_i.data(), rp_j.data(),
                           rot_R.data(), rot_kappa.data(), m, rg_a.data(), rg_b.data(), rg_r.data(), rg_w.data(), &nnz,
                           nullptr, nullptr, nullptr));
  const int64_t N = (int64_t)d * n + m + n + l;
  std::vector<int32_t> rowptr((size_t)N + 1), col((size_t)nnz);
  std::vector<double> val((size_t)nnz);
  check(cora_b200_assemble(d, n, l, E, rp_i.data(), rp_j.data(), rp_t.data(), rp_tau.data(), E, rp_i.data(), rp_j.data(),
                           rot_R.data(), rot_kappa.data(), m, rg_a.data(), rg_b.data(), rg_r.data(), rg_w.data(), &nnz,
                           rowptr.data(), col.data(), val.data()));
  int ndev = 0;
  check(cora_b200_device_count(&ndev));
  if (ndev == 0) {
    bool threw = false;
    try {
      Problem p(d, n, (int)m, n + l, rowptr.data(), col.data(), val.data(), nnz, 3, Preconditioner::Jacobi);
    } catch (const std::runtime_error &e) {
      threw = true;
      std::printf("no CUDA device: Problem() threw std::runtime_error: %s\n", e.what());
    }
    CHECK(threw);
    std::printf(fails ? "FAILED\n" : "OK (cpu)\n");
    return fails ? 1 : 0;
  }
  Problem p(d, n, (int)m, n + l, rowptr.data(), col.data(), val.data(), nnz, 3, Preconditioner::Jacobi);
  CHECK(p.getDataMatrixSize() == N);
  // ground truth embedded in rank 3: rows of pose i are R_i^T, range rows unit bearings, translations
  Matrix X(N, 3);
  for (int i = 0; i < n; ++i) {
    const double c = std::cos(th[i]), s = std::sin(th[i]);
    X(2 * i, 0) = c; X(2 * i, 1) = s;       // first row of R^T
    X(2 * i + 1, 0) = -s; X(2 * i + 1, 1) = c;
    X(d * n + m + i, 0) = tx[i]; X(d * n + m + i, 1) = ty[i];
  }
  for (int j = 0; j < l; ++j) { X(d * n + m + n + j, 0) = L[j][0]; X(d * n + m + n + j, 1) = L[j][1]; }
  for (int k = 0; k < m; ++k) {
    const int i = (int)rg_a[k], j = (int)rg_b[k] - n;
    const double dx = tx[i] - L[j][0], dy = ty[i] - L[j][1], nr = std::hypot(dx, dy);
    X(d * n + k, 0) = dx / nr; X(d * n + k, 1) = dy / nr;
  }
  // host evaluation of f = 1/2 tr(X^T Q X) from the CSR
  double f_host = 0;
  for (int c = 0; c < 3; ++c)
    for (int64_t i = 0; i < N; ++i) {
      double s = 0;
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * X(col[k], c);
      f_host += 0.5 * X(i, c) * s;
    }
  const double f_dev = p.evaluateObjective(X);
  std::printf("f(ground truth): host %.3e device %.3e\n", f_host, f_dev);
  CHECK(std::fabs(f_host) < 1e-9 && std::fabs(f_dev) < 1e-9);  // noise-free: X_gt is in the null space of Q
  Matrix G = p.Riemannian_gradient(X);
  double gn = 0;
  for (int64_t i = 0; i < N; ++i) for (int c = 0; c < 3; ++c) gn += G(i, c) * G(i, c);
  CHECK(std::sqrt(gn) < 1e-8);
  // shape errors are std::invalid_argument as in the reference (checkMatrixShape)
  bool threw = false;
  try { p.evaluateObjective(Matrix(N - 1, 3)); } catch (const std::invalid_argument &) { threw = true; }
  CHECK(threw);
  // a perturbed start must come back to cost ~0, certified
  Matrix X0 = X;
  for (int64_t i = 0; i < N; ++i) for (int c = 0; c < 3; ++c) X0(i, c) += 0.05 * std::sin(1.0 + 0.37 * i + 1.3 * c);
  CoraResult res = solveCORA(p, p.projectToManifold(X0), 6);
  std::printf("solveCORA: f %.3e, status %d, x is %ld x %ld\n", res.first.f, (int)res.first.status, (long)res.first.x.rows(),
              (long)res.first.x.cols());
  CHECK(res.first.f < 1e-6);
  CHECK(res.first.x.cols() == d);
  // Formulation::Implicit (CORA_problem.h:338, src/CORA_problem.cpp:714-757): rotations + ranges only; the product of
  // the marginalised problem is the top of Q [Y; t*(Y)], and the ground truth stays in the kernel
  p.setRank(3);
  p.setFormulation(Formulation::Implicit);
  const int64_t K = p.getExpectedVariableSize();
  CHECK(K == (int64_t)d * n + m);
  Matrix Yi(K, 3);
  for (int64_t i = 0; i < K; ++i) for (int c = 0; c < 3; ++c) Yi(i, c) = X(i, c);
  CHECK(std::fabs(p.evaluateObjective(Yi)) < 1e-9);
  Matrix Xf = p.getTranslationExplicitSolution(Yi);
  CHECK(Xf.rows() == N && Xf.cols() == 3);
  double dt = 0;  // translations are recovered up to the pinned last one (a common shift)
  for (int64_t i = K; i < N; ++i) for (int c = 0; c < 3; ++c) dt = std::max(dt, std::fabs((Xf(i, c) - X(i, c)) - (Xf(K, c) - X(K, c))));
  CHECK(dt < 1e-8);
  threw = false;
  try { p.evaluateObjective(X); } catch (const std::invalid_argument &) { threw = true; }  // N rows are rejected now
  CHECK(threw);
  p.setFormulation(Formulation::Explicit);
  std::printf(fails ? "FAILED\n" : "OK (gpu)\n");
  return fails ? 1 : 0;
}
