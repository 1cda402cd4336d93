// assemble.hpp -- O(#factors log #factors) host assembly of the data matrix Q from the
// flattened measurement stacks.  Follows Problem::fillDataMatrix and friends
// (src/CORA_problem.cpp:115-377, 625-712; block diagram include/CORA/CORA_problem.h:147-184)
// written as direct triplets instead of the reference's five sparse triple products, and
// without its O(M^2) duplicate scans.
//
//   Q11 = L_rho + T' Omega_t T      pose-pose factor (i,j): +kappa I on (i,i),(j,j), -kappa R on
//                                   (i,j), -kappa R' on (j,i) (:314-342); + tau t t' on (i,i) (:208-213)
//   Q13 = T' Omega_t A_t            rows of pose i: +tau t at column t_i, -tau t at column t_j
//   Q22 = diag(w rho^2)   Q23 = D Omega_r A_r : row k: -w rho at first id, +w rho at second id
//   Q33 = A_t' Omega_t A_t + A_r' Omega_r A_r
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace cora_b200 {

struct Triplet {
  int32_t r, c;
  double v;
};

struct Measurements {
  int d, n, l;
  int64_t E;  // translation-carrying factors (pose-pose, pose priors, pose-landmark, landmark priors)
  const int64_t *rp_i, *rp_j;  // translation indices (0..n+l)
  const double *rp_t, *rp_tau;
  int64_t Ep;  // rotation-carrying factors
  const int64_t *rot_i, *rot_j;
  const double *rot_R, *rot_kappa;  // R row-major d x d
  int64_t m;
  const int64_t *rg_a, *rg_b;
  const double *rg_r, *rg_w;
};

inline void assemble_data_matrix(const Measurements &M, std::vector<int32_t> &rowptr,
                                 std::vector<int32_t> &col, std::vector<double> &val) {
  const int d = M.d;
  const int64_t n = M.n, nt = (int64_t)M.n + M.l, m = M.m;
  if (d != 2 && d != 3) throw std::invalid_argument("dimension must be 2 or 3");
  const int64_t dn = d * n, T0 = dn + m, N = dn + m + nt;
  if (N >= (int64_t)1 << 30) throw std::invalid_argument("problem too large");
  std::vector<Triplet> tr;
  tr.reserve((size_t)(M.Ep * (2 * d + 2 * d * d) + M.E * (d * d + 4 * d + 4) + m * 9));
  auto add = [&](int64_t r, int64_t c, double v) { tr.push_back({(int32_t)r, (int32_t)c, v}); };
  for (int64_t e = 0; e < M.Ep; ++e) {  // rotation connection Laplacian, :297-377
    const int64_t i = M.rot_i[e], j = M.rot_j[e];
    if (i < 0 || i >= n || j < 0 || j >= n) throw std::invalid_argument("rotation index out of range");
    const double k = M.rot_kappa[e];
    const double *R = M.rot_R + e * d * d;
    for (int a = 0; a < d; ++a) {
      add(i * d + a, i * d + a, k);
      add(j * d + a, j * d + a, k);
      for (int b = 0; b < d; ++b) {
        add(i * d + a, j * d + b, -k * R[a * d + b]);
        add(j * d + b, i * d + a, -k * R[a * d + b]);
      }
    }
  }
  for (int64_t e = 0; e < M.E; ++e) {
    const int64_t i = M.rp_i[e], j = M.rp_j[e];
    if (i < 0 || i >= n || j < 0 || j >= nt) throw std::invalid_argument("translation index out of range");
    const double tau = M.rp_tau[e];
    const double *t = M.rp_t + e * d;
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b < d; ++b) add(i * d + a, i * d + b, tau * t[a] * t[b]);
      const double v = tau * t[a];
      add(i * d + a, T0 + i, v);
      add(T0 + i, i * d + a, v);
      add(i * d + a, T0 + j, -v);
      add(T0 + j, i * d + a, -v);
    }
    add(T0 + i, T0 + i, tau);
    add(T0 + j, T0 + j, tau);
    add(T0 + i, T0 + j, -tau);
    add(T0 + j, T0 + i, -tau);
  }
  for (int64_t k = 0; k < m; ++k) {
    const int64_t a = M.rg_a[k], b = M.rg_b[k];
    if (a < 0 || a >= nt || b < 0 || b >= nt) throw std::invalid_argument("range index out of range");
    const double w = M.rg_w[k], rho = M.rg_r[k];
    add(dn + k, dn + k, w * rho * rho);
    add(dn + k, T0 + a, -w * rho);
    add(T0 + a, dn + k, -w * rho);
    add(dn + k, T0 + b, w * rho);
    add(T0 + b, dn + k, w * rho);
    add(T0 + a, T0 + a, w);
    add(T0 + b, T0 + b, w);
    add(T0 + a, T0 + b, -w);
    add(T0 + b, T0 + a, -w);
  }
  // counting sort by row, then sort each (short) row by column; stable so that duplicate
  // entries are summed in insertion order on every run
  std::vector<int64_t> cnt((size_t)N + 1, 0);
  for (const Triplet &t : tr) ++cnt[t.r + 1];
  for (int64_t i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
  std::vector<Triplet> byrow(tr.size());
  {
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (const Triplet &t : tr) byrow[pos[t.r]++] = t;
  }
  tr.clear();
  tr.shrink_to_fit();
  rowptr.assign((size_t)N + 1, 0);
  col.clear();
  val.clear();
  col.reserve(byrow.size());
  val.reserve(byrow.size());
  for (int64_t i = 0; i < N; ++i) {
    Triplet *b = byrow.data() + cnt[i], *e = byrow.data() + cnt[i + 1];
    std::stable_sort(b, e, [](const Triplet &x, const Triplet &y) { return x.c < y.c; });
    for (Triplet *p = b; p < e;) {
      double s = 0.0;
      Triplet *q = p;
      for (; q < e && q->c == p->c; ++q) s += q->v;
      if (s != 0.0) {  // exact zeros dropped (the MatrixMarket goldens drop them too)
        col.push_back(p->c);
        val.push_back(s);
      }
      p = q;
    }
    rowptr[i + 1] = (int32_t)col.size();
  }
}

}  // namespace cora_b200
