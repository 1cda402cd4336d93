// lanczos.cuh -- small host-side dense helpers (symmetric tridiagonal eigen-solver) and the
// device Lanczos iterations built on the data-matrix product.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "solver.cuh"

namespace cora_b200 {

// Eigen-decomposition of the symmetric tridiagonal matrix (diag d[0..n), off-diagonal e[0..n-1))
// by the implicit QL algorithm (EISPACK tql2).  On return d holds the eigenvalues in ascending
// order and, when Z != nullptr, Z (n x n, row-major) the eigenvectors in its columns.
inline void tridiag_eig(int n, std::vector<double> &d, std::vector<double> e, std::vector<double> *Z) {
  if (Z) {
    Z->assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) (*Z)[(size_t)i * n + i] = 1.0;
  }
  e.resize(n, 0.0);
  if (n > 0) e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n - 1 && std::fabs(e[m]) > eps * tst1) ++m;
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 200) break;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          if (Z)
            for (int k = 0; k < n; ++k) {
              double *row = Z->data() + (size_t)k * n;
              h = row[i + 1];
              row[i + 1] = s * row[i] + c * h;
              row[i] = c * row[i] - s * h;
            }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1);
    }
    d[l] += f;
    e[l] = 0.0;
  }
  // ascending order
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  std::vector<double> d2(n);
  for (int i = 0; i < n; ++i) d2[i] = d[idx[i]];
  if (Z) {
    std::vector<double> Z2((size_t)n * n);
    for (int k = 0; k < n; ++k)
      for (int i = 0; i < n; ++i) Z2[(size_t)k * n + i] = (*Z)[(size_t)k * n + idx[i]];
    Z->swap(Z2);
  }
  d.swap(d2);
}

// ||Q||_2 = lambda_max(Q) (Q is PSD) by Lanczos on the device data-matrix product; the
// reference asks LOBPCG for a relative accuracy of 1e-2 (src/CORA_problem.cpp:556-578).
double estimate_spectral_norm(H *h) {
  ensure_workspace(h, 1);
  const long long nE = h->DL.N;
  double *q = h->ws[V_T0].p, *qp = h->ws[V_T1].p, *w = h->ws[V_Z].p;
  // deterministic start vector: x_i = 1 + (i mod 7)/7
  {
    std::vector<double> x((size_t)nE);
    double nrm = 0.0;
    for (long long i = 0; i < nE; ++i) { x[i] = 1.0 + (double)(i % 7) / 7.0 - 0.3 * (double)(i % 3); nrm += x[i] * x[i]; }
    nrm = std::sqrt(nrm);
    for (auto &v : x) v /= nrm;
    CUDA_CHECK(cudaMemcpyAsync(q, x.data(), (size_t)nE * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  CUDA_CHECK(cudaMemsetAsync(qp, 0, (size_t)nE * sizeof(double), h->stream));
  std::vector<double> al, be;
  double beta = 0.0, prev = 0.0;
  const int kmax = (int)std::min<long long>(60, nE);
  for (int k = 0; k < kmax; ++k) {
    launch_qprod_raw(h, QM_SPMM, q, nullptr, nullptr, w, nullptr, 1, POST_STORE, SC_TMP, nullptr);
    launch_dot2(h, q, w, nullptr, nullptr, nE, SC_TMP);
    read_scal(h);
    const double alpha = h->h_scal[SC_TMP];
    al.push_back(alpha);
    launch_axpby(h, 1.0, w, -alpha, q, w, nE);
    launch_axpby(h, 1.0, w, -beta, qp, w, nE);
    launch_dot2(h, w, w, nullptr, nullptr, nE, SC_TMP);
    read_scal(h);
    beta = std::sqrt(std::max(0.0, h->h_scal[SC_TMP]));
    std::vector<double> dd = al, ee = be;
    tridiag_eig((int)dd.size(), dd, ee, nullptr);
    const double top = dd.back();
    if (k >= 8 && std::fabs(top - prev) <= 1e-4 * std::fabs(top)) { prev = top; break; }
    prev = top;
    if (beta <= 1e-14 * std::fabs(top)) break;
    be.push_back(beta);
    launch_axpby(h, 1.0, q, 0.0, nullptr, qp, nE);
    launch_axpby(h, 1.0 / beta, w, 0.0, nullptr, q, nE);
  }
  return prev;
}

}  // namespace cora_b200
