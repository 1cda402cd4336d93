// chain_factor_dev.cuh -- the chain Cholesky factorisation of chain_chol.cuh computed ON THE DEVICE.
//
// Round 1 copied the block values back to the host (38 MB at 100k poses, blocking) and factored on one CPU thread,
// once per preconditioner, once per PSD test and up to 60 more times in the shift search of the certification
// (src/CORA_problem.cpp:544-614 -> getBlockCholeskyFactorization, src/CORA_preconditioners.cpp:16-44; the PSD half
// of fast_verification, src/CORA_utils.cpp:33-57).  Here the STRUCTURE of the factorisation (which range touches
// which translation, the border entries, the chunk geometry of every level) is extracted once per handle from the
// layout -- it depends on Q's pattern only -- and every factorisation of M = values + shift I is a short sequence
// of kernels on the values resident in HBM (Q, or S = Q - Lambda patched on the same structure):
//   ranges:    rdinv_k = 1 / (sdiag_k + shift)
//   poses:     A_i, U_i from the block-ELL slots + the Schur terms of the eliminated ranges
//   border:    value = static coupling - sum over the ranges joining that pose and landmark
//   landmarks: C = static + shift + sdiag - range terms (one CTA per landmark)
//   levels:    factor_chunk<B> per chunk (the very routine the host test hook runs), separators -> next level
//   W = T^-1 B with the solve kernels, S_L = C - B^T W reduced per landmark; only the l x l inverse of S_L and the
//   positive-definiteness flag cross PCIe.
#pragma once
#include "chain_chol.cuh"

namespace cora_b200 {

struct ChainSymLevel {
  ChunkGeo G;
  DevBuf<double> A, U, SL, SR, CP;  // scratch of one factorisation (A, U: this level's chain; Schur pieces per chunk)
};

// Structure of the chain factorisation of this handle's graph (values of Q for the static couplings).
struct ChainSym {
  bool built = false;
  int B = 0, n = 0, l = 0, m = 0;
  std::vector<int32_t> rend_x;
  std::vector<double> rend_e;
  std::vector<int32_t> tinc_ptr, tinc_k;   // per translation (n poses, then l landmarks): incident ranges
  std::vector<double> tinc_e;
  std::vector<int32_t> bl_ptr, bl_row;     // border entries by landmark, sorted by pose-section row
  std::vector<double> bl_static;
  std::vector<int32_t> blc_ptr, blc_k;     // per border entry: ranges joining that translation and landmark
  std::vector<double> blc_coef;
  std::vector<int32_t> uc_ptr, uc_k;       // per pose i: ranges joining t_i and t_{i+1}
  std::vector<double> uc_coef;
  std::vector<int32_t> cc_j, cc_j2, cc_k;  // landmark-landmark ranges (both orders)
  std::vector<double> cc_coef;
  std::vector<double> Cstat;               // l x l static landmark-landmark couplings (off the diagonal array)
  std::vector<long long> Aoff, Uoff;       // per pose: offset of the diagonal / (i, i+1) slot in the block values (-1: none)
  // device copies
  DevBuf<int> d_rend_x, d_tinc_ptr, d_tinc_k, d_bl_ptr, d_bl_row, d_blc_ptr, d_blc_k, d_uc_ptr, d_uc_k, d_cc_j, d_cc_j2, d_cc_k;
  DevBuf<double> d_rend_e, d_tinc_e, d_bl_static, d_blc_coef, d_uc_coef, d_cc_coef, d_Cstat;
  DevBuf<long long> d_Aoff, d_Uoff;
  std::vector<ChainSymLevel *> lv;
  std::vector<DevBuf<double>> wsol, wrhs, wcL, wcR;  // work vectors of the W = T^-1 B solve (l columns), per level
  DevBuf<double> d_C, d_SL;                // l x l
  DevBuf<int> d_flag;
  // range incidence by translation for the solve, with the pinned translation left out: [0] no pin, [1] pinned
  std::vector<int32_t> rinc_ptr[2], rinc_k[2];
  std::vector<double> rinc_e[2];
  // ---- pose graph that is not a chain (loop closures, several robots): general sparse block Cholesky ----
  bool general = false;
  std::vector<int32_t> e_i, e_j;           // pose-pose couplings, i < j
  std::vector<long long> e_off;            // offset of block (i, j) in the block values (-1: only in the spill)
  std::vector<double> e_stat;              // static part of the block: spill entries of pose i (blocks beyond the ELL slots)
  std::vector<int32_t> ec_ptr, ec_k;       // per coupling: ranges joining t_i and t_j
  std::vector<double> ec_coef;
  GenSym gs;
  GenSymDev gsd;
  DevBuf<int> d_e_j, d_ec_ptr, d_ec_k;
  DevBuf<long long> d_e_off;
  DevBuf<double> d_e_stat, d_ec_coef, d_A, d_E;
  double *h_AE = nullptr;                  // pinned staging of the assembled pose blocks (A then E)
  std::vector<double> h_L, h_Dinv, h_Linv;
  ~ChainSym() {
    for (auto *p : lv) delete p;
    if (h_AE) cudaFreeHost(h_AE);
  }
};

// ---------------------------------------------------------------- symbolic ----
// general == false: the chain structure (throws ENOTIMPL as soon as a pose couples to a non-adjacent pose);
// general == true: every pose-pose coupling becomes an edge of the pose graph for gen_chol.hpp.
template <int B>
inline void chain_symbolic_build(ChainSym &S, const HostLayout &L, bool general) {
  constexpr int BB = B * B;
  const int n = L.n, l = L.l, m = L.m, D1 = L.D1, d = L.d;
  S.B = B; S.n = n; S.l = l; S.m = m;
  S.general = general;
  auto not_chain = [](const char *why) {
    throw Error(CORA_B200_ENOTIMPL,
                std::string("RegularizedCholesky / Cholesky certificate: unsupported coupling (") + why + ")");
  };
  std::map<std::pair<int32_t, int32_t>, int32_t> emap;
  auto edge_id = [&](int i, int j) -> int32_t {
    auto it = emap.find({(int32_t)i, (int32_t)j});
    if (it != emap.end()) return it->second;
    const int32_t id = (int32_t)S.e_i.size();
    emap[{(int32_t)i, (int32_t)j}] = id;
    S.e_i.push_back(i); S.e_j.push_back(j); S.e_off.push_back(-1);
    S.e_stat.insert(S.e_stat.end(), BB, 0.0);
    return id;
  };
  struct ECoup { int32_t id, k; double coef; };
  std::vector<ECoup> ecoup;
  S.Aoff.assign((size_t)std::max(n, 1), -1); S.Uoff.assign((size_t)std::max(n, 1), -1);
  for (int i = 0; i < n; ++i) {
    const int t = i / L.TP, p = i % L.TP;
    const int Sl = L.tile_slots[t];
    for (int s = 0; s < Sl; ++s) {
      const int j = L.bcol[L.tile_coff[t] + (int64_t)s * L.TP + p] / D1;
      const long long off = L.tile_boff[t] + (long long)s * BB * L.TP + p;
      bool nz = false;
      for (int e = 0; e < BB; ++e) nz = nz || L.bval[off + (long long)e * L.TP] != 0.0;
      if (s > 0 && j == i) continue;  // padding slot
      if (j == i) S.Aoff[i] = off;
      else if (general) { if (j > i && nz) S.e_off[edge_id(i, j)] = off; }  // (row j holds the transposed copy)
      else if (j == i + 1) S.Uoff[i] = off;
      else if (j == i - 1) { }
      else if (nz) not_chain("a pose is coupled to a non-adjacent pose");
    }
  }
  const int64_t rg0 = L.nPoseRows + l;
  auto trans_index = [&](int64_t ci) -> int {
    if (ci < L.nPoseRows) return (ci % D1 != d) ? -1 : (int)(ci / D1);
    if (ci < rg0) return n + (int)(ci - L.nPoseRows);
    return -1;
  };
  auto for_group = [&](int64_t g, auto &&fn) {
    for (int32_t k = L.grp_ptr[g]; k < L.grp_ptr[g + 1]; ++k) fn(L.rem_pk[k], L.rem_val[k]);
    auto it = std::lower_bound(L.long_grp.begin(), L.long_grp.end(), (int32_t)g);
    if (it != L.long_grp.end() && *it == g) {
      const size_t q = it - L.long_grp.begin();
      for (int32_t k = L.long_ptr[q]; k < L.long_ptr[q + 1]; ++k) fn(L.long_pk[k], L.long_val[k]);
    }
  };
  S.rend_x.assign((size_t)std::max(m, 1) * 2, -1);
  S.rend_e.assign((size_t)std::max(m, 1) * 2, 0.0);
  for (int k = 0; k < m; ++k) {
    int cnt = 0;
    for_group((int64_t)n + l + k, [&](uint32_t pk, double v) {
      const int x = trans_index(pk & kColMask);
      if (x < 0) not_chain("a range row is coupled to a non-translation variable");
      if (cnt >= 2) not_chain("a range row has more than two couplings");
      if (cnt == 1 && S.rend_x[(size_t)k * 2] == x) {  // both couplings on the same translation: one coupling
        S.rend_e[(size_t)k * 2] += v;
        return;
      }
      S.rend_x[(size_t)k * 2 + cnt] = x;
      S.rend_e[(size_t)k * 2 + cnt] = v;
      ++cnt;
    });
  }
  // incident ranges of every translation; adjacent-pose, pose-landmark and landmark-landmark pair terms
  const int nt = n + l;
  S.tinc_ptr.assign((size_t)nt + 1, 0);
  for (int k = 0; k < m; ++k)
    for (int p = 0; p < 2; ++p) if (S.rend_x[(size_t)k * 2 + p] >= 0) ++S.tinc_ptr[S.rend_x[(size_t)k * 2 + p] + 1];
  for (int x = 0; x < nt; ++x) S.tinc_ptr[x + 1] += S.tinc_ptr[x];
  S.tinc_k.assign((size_t)S.tinc_ptr[nt], 0); S.tinc_e.assign((size_t)S.tinc_ptr[nt], 0.0);
  {
    std::vector<int32_t> fill(S.tinc_ptr.begin(), S.tinc_ptr.end() - 1);
    for (int k = 0; k < m; ++k)
      for (int p = 0; p < 2; ++p) {
        const int x = S.rend_x[(size_t)k * 2 + p];
        if (x < 0) continue;
        S.tinc_k[fill[x]] = k; S.tinc_e[fill[x]] = S.rend_e[(size_t)k * 2 + p]; ++fill[x];
      }
  }
  struct BEnt { int32_t row, j, k; double v; };  // k >= 0: a range contribution with coefficient v (times rdinv_k)
  std::vector<BEnt> bents;
  std::vector<std::vector<std::pair<int32_t, double>>> uc((size_t)std::max(n, 1));
  S.Cstat.assign((size_t)std::max(l, 1) * std::max(l, 1), 0.0);
  for (int k = 0; k < m; ++k) {
    const int x0 = S.rend_x[(size_t)k * 2], x1 = S.rend_x[(size_t)k * 2 + 1];
    if (x0 < 0 || x1 < 0 || x0 == x1) continue;  // (x, x) terms come from the incidence lists
    const double coef = S.rend_e[(size_t)k * 2] * S.rend_e[(size_t)k * 2 + 1];
    const int a = std::min(x0, x1), b = std::max(x0, x1);
    if (b < n) {
      if (general) ecoup.push_back({edge_id(a, b), (int32_t)k, coef});
      else if (b == a + 1) uc[a].push_back({k, coef});
      else not_chain("a range factor joins two non-adjacent poses");
    } else if (a < n) {
      bents.push_back({(int32_t)(a * D1 + d), (int32_t)(b - n), k, coef});
    } else {
      S.cc_j.push_back(a - n); S.cc_j2.push_back(b - n); S.cc_k.push_back(k); S.cc_coef.push_back(coef);
      S.cc_j.push_back(b - n); S.cc_j2.push_back(a - n); S.cc_k.push_back(k); S.cc_coef.push_back(coef);
    }
  }
  S.uc_ptr.assign((size_t)std::max(n, 1) + 1, 0);
  for (int i = 0; i < n; ++i) {
    S.uc_ptr[i + 1] = S.uc_ptr[i] + (int32_t)uc[i].size();
    for (auto &e : uc[i]) { S.uc_k.push_back(e.first); S.uc_coef.push_back(e.second); }
  }
  for (int i = 0; i < n; ++i)
    for_group(i, [&](uint32_t pk, double v) {
      const int64_t ci = pk & kColMask;
      const int a = (int)(pk >> 30);
      if (ci < L.nPoseRows) {
        if (!general) not_chain("pose-pose coupling outside the block-ELL");
        const int j = (int)(ci / D1), b = (int)(ci % D1);
        if (j > i) S.e_stat[(size_t)edge_id(i, j) * BB + a * B + b] += v;
        return;
      }
      if (ci < rg0) bents.push_back({(int32_t)(i * D1 + a), (int32_t)(ci - L.nPoseRows), -1, v});
    });
  if (general) {
    const int ne = (int)S.e_i.size();
    S.ec_ptr.assign((size_t)ne + 1, 0);
    for (auto &c : ecoup) ++S.ec_ptr[c.id + 1];
    for (int e = 0; e < ne; ++e) S.ec_ptr[e + 1] += S.ec_ptr[e];
    S.ec_k.assign(ecoup.size(), 0); S.ec_coef.assign(ecoup.size(), 0.0);
    std::vector<int32_t> fill(S.ec_ptr.begin(), S.ec_ptr.end() - 1);
    for (auto &c : ecoup) { S.ec_k[fill[c.id]] = c.k; S.ec_coef[fill[c.id]] = c.coef; ++fill[c.id]; }
    gen_symbolic(S.gs, n, S.e_i, S.e_j);
  }
  for (int j = 0; j < l; ++j)
    for_group((int64_t)n + j, [&](uint32_t pk, double v) {
      const int64_t ci = pk & kColMask;
      if (ci >= L.nPoseRows && ci < rg0) S.Cstat[(size_t)j * l + (ci - L.nPoseRows)] += v;
    });
  std::stable_sort(bents.begin(), bents.end(), [](const BEnt &x, const BEnt &y) {
    return x.j != y.j ? x.j < y.j : x.row < y.row;
  });
  S.bl_ptr.assign((size_t)l + 1, 0);
  S.blc_ptr.assign(1, 0);
  for (size_t q = 0; q < bents.size();) {
    size_t e = q;
    double v = 0.0;
    while (e < bents.size() && bents[e].j == bents[q].j && bents[e].row == bents[q].row) {
      if (bents[e].k < 0) v += bents[e].v;
      else { S.blc_k.push_back(bents[e].k); S.blc_coef.push_back(bents[e].v); }
      ++e;
    }
    S.bl_row.push_back(bents[q].row);
    S.bl_static.push_back(v);
    S.blc_ptr.push_back((int32_t)S.blc_k.size());
    ++S.bl_ptr[bents[q].j + 1];
    q = e;
  }
  for (int j = 0; j < l; ++j) S.bl_ptr[j + 1] += S.bl_ptr[j];
  // range incidence for the solve kernels, with / without the pinned translation (src/CORA_preconditioners.cpp:77-80)
  for (int pin = 0; pin < 2; ++pin) {
    const int pinned = pin ? (l > 0 ? n + l - 1 : n - 1) : -1;
    S.rinc_ptr[pin].assign((size_t)nt + 1, 0);
    for (int x = 0; x < nt; ++x)
      S.rinc_ptr[pin][x + 1] = S.rinc_ptr[pin][x] + (x == pinned ? 0 : S.tinc_ptr[x + 1] - S.tinc_ptr[x]);
    S.rinc_k[pin].clear(); S.rinc_e[pin].clear();
    for (int x = 0; x < nt; ++x) {
      if (x == pinned) continue;
      for (int32_t q = S.tinc_ptr[x]; q < S.tinc_ptr[x + 1]; ++q) { S.rinc_k[pin].push_back(S.tinc_k[q]); S.rinc_e[pin].push_back(S.tinc_e[q]); }
    }
  }
}


// ----------------------------------------------------------------- kernels ----
static __global__ void __launch_bounds__(kThreads) k_cf_ranges(int m, int l, const double *__restrict__ sdiag, double shift,
                                                               double *rdinv, int *flag, int trans_only) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  if (trans_only) { rdinv[k] = 0.0; return; }  // the range rows are not part of the translation Laplacian
  const double delta = sdiag[l + k] + shift;
  if (!(delta > 0.0)) { rdinv[k] = 1.0; atomicAnd(flag, 0); }
  else rdinv[k] = 1.0 / delta;
}

template <int B>
__global__ void __launch_bounds__(kThreads) k_cf_pose_blocks(int n, int TP, const double *__restrict__ bval,
                                                             const long long *__restrict__ Aoff,
                                                             const long long *__restrict__ Uoff, double shift,
                                                             const int *__restrict__ tinc_ptr, const int *__restrict__ tinc_k,
                                                             const double *__restrict__ tinc_e,
                                                             const int *__restrict__ uc_ptr, const int *__restrict__ uc_k,
                                                             const double *__restrict__ uc_coef,
                                                             const double *__restrict__ rdinv, int pinned_pose_row,
                                                             double *A, double *U, int trans_only) {
  constexpr int BB = B * B, d = B - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[BB], u[BB];
  for (int e = 0; e < BB; ++e) {
    a[e] = Aoff[i] >= 0 ? bval[Aoff[i] + (long long)e * TP] : 0.0;
    u[e] = Uoff[i] >= 0 ? bval[Uoff[i] + (long long)e * TP] : 0.0;
  }
  for (int q = 0; q < B; ++q) a[q * B + q] += shift;
  double s = 0.0;  // Schur complement of the ranges incident to t_i (in list order: deterministic)
  for (int q = tinc_ptr[i]; q < tinc_ptr[i + 1]; ++q) s += tinc_e[q] * tinc_e[q] * rdinv[tinc_k[q]];
  a[d * B + d] -= s;
  double su = 0.0;
  for (int q = uc_ptr[i]; q < uc_ptr[i + 1]; ++q) su += uc_coef[q] * rdinv[uc_k[q]];
  u[d * B + d] -= su;
  if (pinned_pose_row >= 0) {
    if (i == n - 1) {
      for (int q = 0; q < B; ++q) { a[d * B + q] = 0.0; a[q * B + d] = 0.0; }
      a[d * B + d] = 1.0;
    }
    if (i == n - 2)
      for (int q = 0; q < B; ++q) u[q * B + d] = 0.0;
  }
  if (trans_only) {  // translation Laplacian only: identity on the rotation rows, no coupling to them
    for (int p = 0; p < B; ++p)
      for (int q = 0; q < B; ++q)
        if (p != d || q != d) { a[p * B + q] = (p == q) ? 1.0 : 0.0; u[p * B + q] = 0.0; }
  }
  for (int e = 0; e < BB; ++e) { A[(size_t)i * BB + e] = a[e]; U[(size_t)i * BB + e] = u[e]; }
}

// general pose graph: thread t < n assembles the diagonal block of pose t, thread n + e the coupling block of edge e
// (M_ij, rows of pose i, columns of pose j, i < j) = block-ELL slot + spill part - Schur terms of the ranges
template <int B>
__global__ void __launch_bounds__(kThreads) k_gen_pose_blocks(int n, int ne, int TP, const double *__restrict__ bval,
                                                              const long long *__restrict__ Aoff, double shift,
                                                              const int *__restrict__ tinc_ptr, const int *__restrict__ tinc_k,
                                                              const double *__restrict__ tinc_e,
                                                              const double *__restrict__ rdinv, int pinned_pose_row,
                                                              const int *__restrict__ e_j, const long long *__restrict__ e_off,
                                                              const double *__restrict__ e_stat,
                                                              const int *__restrict__ ec_ptr, const int *__restrict__ ec_k,
                                                              const double *__restrict__ ec_coef, double *A, double *E,
                                                              int trans_only) {
  constexpr int BB = B * B, d = B - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const int i = t;
    double a[BB];
    for (int e = 0; e < BB; ++e) a[e] = Aoff[i] >= 0 ? bval[Aoff[i] + (long long)e * TP] : 0.0;
    for (int q = 0; q < B; ++q) a[q * B + q] += shift;
    double s = 0.0;
    for (int q = tinc_ptr[i]; q < tinc_ptr[i + 1]; ++q) s += tinc_e[q] * tinc_e[q] * rdinv[tinc_k[q]];
    a[d * B + d] -= s;
    if (pinned_pose_row >= 0 && i == n - 1) {
      for (int q = 0; q < B; ++q) { a[d * B + q] = 0.0; a[q * B + d] = 0.0; }
      a[d * B + d] = 1.0;
    }
    if (trans_only)
      for (int p = 0; p < B; ++p)
        for (int q = 0; q < B; ++q)
          if (p != d || q != d) a[p * B + q] = (p == q) ? 1.0 : 0.0;
    for (int e = 0; e < BB; ++e) A[(size_t)i * BB + e] = a[e];
  } else if (t - n < ne) {
    const int e = t - n;
    double u[BB];
    for (int q = 0; q < BB; ++q) u[q] = (e_off[e] >= 0 ? bval[e_off[e] + (long long)q * TP] : 0.0) + e_stat[(size_t)e * BB + q];
    double su = 0.0;
    for (int q = ec_ptr[e]; q < ec_ptr[e + 1]; ++q) su += ec_coef[q] * rdinv[ec_k[q]];
    u[d * B + d] -= su;
    if (pinned_pose_row >= 0 && e_j[e] == n - 1)
      for (int q = 0; q < B; ++q) u[q * B + d] = 0.0;
    if (trans_only)
      for (int q = 0; q < BB; ++q) if (q != d * B + d) u[q] = 0.0;
    for (int q = 0; q < BB; ++q) E[(size_t)e * BB + q] = u[q];
  }
}

static __global__ void __launch_bounds__(kThreads) k_cf_border(int nb, const int *__restrict__ bl_ptr, int l,
                                                               const double *__restrict__ bl_static,
                                                               const int *__restrict__ blc_ptr, const int *__restrict__ blc_k,
                                                               const double *__restrict__ blc_coef,
                                                               const double *__restrict__ rdinv, const int *__restrict__ bl_row,
                                                               int pinned_landmark, int pinned_pose_row, double *bl_val,
                                                               int trans_only, int D1) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb) return;
  if (trans_only && bl_row[q] % D1 != D1 - 1) { bl_val[q] = 0.0; return; }  // rotation row of a pose
  double v = bl_static[q];
  for (int e = blc_ptr[q]; e < blc_ptr[q + 1]; ++e) v -= blc_coef[e] * rdinv[blc_k[e]];
  if (bl_row[q] == pinned_pose_row) v = 0.0;
  if (pinned_landmark >= 0 && q >= bl_ptr[pinned_landmark] && q < bl_ptr[pinned_landmark + 1]) v = 0.0;
  bl_val[q] = v;
}

// one CTA per landmark j: row j of C = static + (sdiag_j + shift) on the diagonal - range terms
static __global__ void __launch_bounds__(kThreads) k_cf_landmarks(int n, int l, const double *__restrict__ sdiag, double shift,
                                                                  const double *__restrict__ Cstat,
                                                                  const int *__restrict__ tinc_ptr, const int *__restrict__ tinc_k,
                                                                  const double *__restrict__ tinc_e,
                                                                  const double *__restrict__ rdinv, int ncc,
                                                                  const int *__restrict__ cc_j, const int *__restrict__ cc_j2,
                                                                  const int *__restrict__ cc_k, const double *__restrict__ cc_coef,
                                                                  int pinned_landmark, double *C) {
  __shared__ double sred[kThreads];
  const int j = blockIdx.x, tid = threadIdx.x;
  double s = 0.0;
  for (int q = tinc_ptr[n + j] + tid; q < tinc_ptr[n + j + 1]; q += kThreads) s += tinc_e[q] * tinc_e[q] * rdinv[tinc_k[q]];
  sred[tid] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (tid < o) sred[tid] += sred[tid + o];
    __syncthreads();
  }
  for (int j2 = tid; j2 < l; j2 += kThreads) {
    double v = Cstat[(size_t)j * l + j2];
    if (j2 == j) v += sdiag[j] + shift - sred[0];
    for (int e = 0; e < ncc; ++e)
      if (cc_j[e] == j && cc_j2[e] == j2) v -= cc_coef[e] * rdinv[cc_k[e]];
    if (pinned_landmark >= 0 && (j == pinned_landmark || j2 == pinned_landmark)) v = (j == j2) ? 1.0 : 0.0;
    C[(size_t)j * l + j2] = v;
  }
}

template <int B>
__global__ void __launch_bounds__(128) k_cf_factor_level(const ChunkGeo G, const double *A, const double *U, double *fwd,
                                                         double *bwd, double *UR, double *SL, double *SR, double *CP,
                                                         int *flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= G.K) return;
  if (!factor_chunk<B>(G, k, A, U, fwd, bwd, UR, SL, SR, CP)) atomicAnd(flag, 0);
}

template <int B>
__global__ void __launch_bounds__(kThreads) k_cf_next_level(const ChunkGeo G, int nn, const double *__restrict__ A,
                                                            const double *__restrict__ SL, const double *__restrict__ SR,
                                                            const double *__restrict__ CP, double *A2, double *U2) {
  constexpr int BB = B * B;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nn * BB) return;
  const int q = t / BB, e = t - q * BB;
  const size_t s = (size_t)((q + 1) * G.c - 1);
  A2[t] = A[s * BB + e] - SR[(size_t)q * BB + e] - SL[(size_t)(q + 1) * BB + e];
  U2[t] = (q < nn - 1) ? CP[(size_t)(q + 1) * BB + e] : 0.0;
}

static __global__ void __launch_bounds__(kThreads) k_cf_scatter_border(int nb, int l, const int *__restrict__ bl_ptr,
                                                                       const int *__restrict__ bl_row,
                                                                       const double *__restrict__ bl_val, double *W) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb) return;
  int j = 0;  // landmark of entry q (binary search in bl_ptr)
  int lo = 0, hi = l;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (bl_ptr[mid] <= q) lo = mid; else hi = mid; }
  j = lo;
  W[(size_t)bl_row[q] * l + j] = bl_val[q];
}

// one CTA per landmark j: row j of S_L = C - B^T W (fixed summation order per thread, tree over the CTA)
static __global__ void __launch_bounds__(kThreads) k_cf_schur(int l, const int *__restrict__ bl_ptr, const int *__restrict__ bl_row,
                                                              const double *__restrict__ bl_val, const double *__restrict__ W,
                                                              const double *__restrict__ C, double *SLm) {
  __shared__ double sred[kThreads];
  const int j = blockIdx.x, tid = threadIdx.x;
  for (int j2 = 0; j2 < l; ++j2) {
    double s = 0.0;
    for (int q = bl_ptr[j] + tid; q < bl_ptr[j + 1]; q += kThreads) s += bl_val[q] * W[(size_t)bl_row[q] * l + j2];
    sred[tid] = s;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
      if (tid < o) sred[tid] += sred[tid + o];
      __syncthreads();
    }
    if (tid == 0) SLm[(size_t)j * l + j2] = C[(size_t)j * l + j2] - sred[0];
    __syncthreads();
  }
}

// ------------------------------------------------------------------- driver ----
template <typename T, typename Tsrc>
inline void upload_as(DevBuf<T> &b, const std::vector<Tsrc> &v, cudaStream_t s) {
  std::vector<T> t(v.begin(), v.end());
  b.upload(t, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
}

inline void chain_symbolic_upload(H *h, ChainSym &S) {
  cudaStream_t s = h->stream;
  upload_as(S.d_rend_x, S.rend_x, s); upload_as(S.d_rend_e, S.rend_e, s);
  upload_as(S.d_tinc_ptr, S.tinc_ptr, s); upload_as(S.d_tinc_k, S.tinc_k, s); upload_as(S.d_tinc_e, S.tinc_e, s);
  upload_as(S.d_bl_ptr, S.bl_ptr, s); upload_as(S.d_bl_row, S.bl_row, s); upload_as(S.d_bl_static, S.bl_static, s);
  upload_as(S.d_blc_ptr, S.blc_ptr, s); upload_as(S.d_blc_k, S.blc_k, s); upload_as(S.d_blc_coef, S.blc_coef, s);
  upload_as(S.d_uc_ptr, S.uc_ptr, s); upload_as(S.d_uc_k, S.uc_k, s); upload_as(S.d_uc_coef, S.uc_coef, s);
  upload_as(S.d_cc_j, S.cc_j, s); upload_as(S.d_cc_j2, S.cc_j2, s); upload_as(S.d_cc_k, S.cc_k, s);
  upload_as(S.d_cc_coef, S.cc_coef, s);
  upload_as(S.d_Cstat, S.Cstat, s); upload_as(S.d_Aoff, S.Aoff, s); upload_as(S.d_Uoff, S.Uoff, s);
  const int BB = S.B * S.B;
  if (S.general) {
    const size_t ne = S.e_i.size();
    upload_as(S.d_e_j, S.e_j, s); upload_as(S.d_e_off, S.e_off, s); upload_as(S.d_e_stat, S.e_stat, s);
    upload_as(S.d_ec_ptr, S.ec_ptr, s); upload_as(S.d_ec_k, S.ec_k, s); upload_as(S.d_ec_coef, S.ec_coef, s);
    S.d_A.alloc((size_t)std::max(S.n, 1) * BB); S.d_E.alloc(std::max<size_t>(ne, 1) * BB);
    CUDA_CHECK(cudaMallocHost((void **)&S.h_AE, ((size_t)std::max(S.n, 1) + std::max<size_t>(ne, 1)) * BB * sizeof(double)));
    gen_sym_upload(h, S.gs, S.gsd);
  }
  int cur = S.general ? 0 : S.n;
  while (!S.general) {
    ChainSymLevel *Lv = new ChainSymLevel();
    Lv->G = make_geo(cur);
    Lv->A.alloc((size_t)std::max(cur, 1) * BB); Lv->U.alloc((size_t)std::max(cur, 1) * BB);
    Lv->SL.alloc((size_t)(Lv->G.K + 1) * BB); Lv->SR.alloc((size_t)(Lv->G.K + 1) * BB); Lv->CP.alloc((size_t)(Lv->G.K + 1) * BB);
    S.lv.push_back(Lv);
    const int nn = Lv->G.top ? 0 : Lv->G.K - 1;
    if (nn == 0) break;
    cur = nn;
  }
  S.d_C.alloc((size_t)std::max(S.l, 1) * std::max(S.l, 1));
  S.d_SL.alloc((size_t)std::max(S.l, 1) * std::max(S.l, 1));
  S.d_flag.alloc(1);
}

// Factor M = values + shift I on the device.  d_bval / d_sdiag: block-ELL values and scalar-row diagonal on the
// handle's structure (Q, or S = Q - Lambda).  Throws ENOTIMPL when the graph is not a chain + landmark border.
template <int B>
inline void chain_factor_device(H *h, ChainSym &S, ChainChol *C, const double *d_bval, const double *d_sdiag, double shift,
                                bool pin_last, bool want_solve, bool trans_only) {
  const int to = trans_only ? 1 : 0;  // factor the translation Laplacian Q33 only (Formulation::Implicit)
  constexpr int BB = B * B;
  const int n = S.n, l = S.l, m = S.m, d = B - 1;
  cudaStream_t s = h->stream;
  ChainFactorHost &F = C->host;
  F.B = B; F.n = n; F.l = l; F.m = m; F.pos_def = true;
  F.pinned_landmark = -1; F.pinned_pose_row = -1;
  if (pin_last) {
    if (l > 0) F.pinned_landmark = l - 1;
    else if (n > 0) F.pinned_pose_row = B * (n - 1) + d;
  }
  const int pin = pin_last ? 1 : 0;
  F.rend_x = S.rend_x; F.bl_ptr = S.bl_ptr;
  F.rinc_ptr = S.rinc_ptr[pin];  // (host copies of the structure: the persistent kernel's chunking reads them)
  const int one = 1;
  CUDA_CHECK(cudaMemcpyAsync(S.d_flag.p, &one, sizeof(int), cudaMemcpyHostToDevice, s));
  C->rdinv.reserve((size_t)std::max(m, 1));
  if (m > 0) {
    k_cf_ranges<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(m, l, d_sdiag, shift, C->rdinv.p, S.d_flag.p, to);
    check_launch(h);
  }
  C->general = S.general;
  C->gsd = &S.gsd;
  if (n > 0 && S.general) {
    const int ne = (int)S.e_i.size();
    k_gen_pose_blocks<B><<<(n + ne + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        n, ne, h->HL.TP, d_bval, S.d_Aoff.p, shift, S.d_tinc_ptr.p, S.d_tinc_k.p, S.d_tinc_e.p, C->rdinv.p,
        F.pinned_pose_row, S.d_e_j.p, S.d_e_off.p, S.d_e_stat.p, S.d_ec_ptr.p, S.d_ec_k.p, S.d_ec_coef.p, S.d_A.p, S.d_E.p, to);
    check_launch(h);
  } else if (n > 0) {
    k_cf_pose_blocks<B><<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        n, h->HL.TP, d_bval, S.d_Aoff.p, S.d_Uoff.p, shift, S.d_tinc_ptr.p, S.d_tinc_k.p, S.d_tinc_e.p, S.d_uc_ptr.p,
        S.d_uc_k.p, S.d_uc_coef.p, C->rdinv.p, F.pinned_pose_row, S.lv[0]->A.p, S.lv[0]->U.p, to);
    check_launch(h);
  }
  const int nb = (int)S.bl_row.size();
  C->bl_val.reserve((size_t)std::max(nb, 1));
  if (nb > 0) {
    k_cf_border<<<(nb + kThreads - 1) / kThreads, kThreads, 0, s>>>(nb, S.d_bl_ptr.p, l, S.d_bl_static.p, S.d_blc_ptr.p,
                                                                   S.d_blc_k.p, S.d_blc_coef.p, C->rdinv.p, S.d_bl_row.p,
                                                                   F.pinned_landmark, F.pinned_pose_row, C->bl_val.p, to, B);
    check_launch(h);
  }
  if (l > 0) {
    k_cf_landmarks<<<l, kThreads, 0, s>>>(n, l, d_sdiag, shift, S.d_Cstat.p, S.d_tinc_ptr.p, S.d_tinc_k.p, S.d_tinc_e.p,
                                         C->rdinv.p, (int)S.cc_j.size(), S.d_cc_j.p, S.d_cc_j2.p, S.d_cc_k.p, S.d_cc_coef.p,
                                         F.pinned_landmark, S.d_C.p);
    check_launch(h);
  }
  // general pose graph: the assembled blocks go to the host, the left-looking block Cholesky of gen_chol.hpp runs
  // there on the structure built once per handle, and the factor comes back (a few MB at TIERS / MR.CLAM sizes)
  bool gen_pd = true;
  if (S.general && n > 0) {
    const size_t ne = S.e_i.size();
    CUDA_CHECK(cudaMemcpyAsync(S.h_AE, S.d_A.p, (size_t)n * BB * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ne) CUDA_CHECK(cudaMemcpyAsync(S.h_AE + (size_t)n * BB, S.d_E.p, ne * BB * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    gen_pd = gen_numeric<B>(S.gs, S.h_AE, S.h_AE + (size_t)n * BB, S.h_L, S.h_Dinv);
    gen_cluster_inverses<B>(S.gs, S.h_L, S.h_Dinv, S.h_Linv);
    C->gen.Lval.reserve(std::max<size_t>(S.h_L.size(), 1));
    C->gen.Dinv.reserve(S.h_Dinv.size()); C->gen.Linv.reserve(S.h_Linv.size());
    CUDA_CHECK(cudaMemcpyAsync(C->gen.Linv.p, S.h_Linv.data(), S.h_Linv.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    if (!S.h_L.empty())
      CUDA_CHECK(cudaMemcpyAsync(C->gen.Lval.p, S.h_L.data(), S.h_L.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(C->gen.Dinv.p, S.h_Dinv.data(), S.h_Dinv.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  // levels
  F.levels.clear();
  if (C->levels.size() != S.lv.size()) {  // (a recycled factor of the same handle keeps its level buffers)
    for (auto *p : C->levels) delete p;
    C->levels.clear();
    for (size_t lvi = 0; lvi < S.lv.size(); ++lvi) C->levels.push_back(new ChainLevelDev());
  }
  for (size_t lvi = 0; lvi < S.lv.size(); ++lvi) {
    ChainSymLevel *Sv = S.lv[lvi];
    const ChunkGeo G = Sv->G;
    ChainLevelDev *D = C->levels[lvi];
    D->G = G;
    const size_t nf = (size_t)G.c * 3 * BB * G.K, nbw = (size_t)G.c * 2 * BB * G.K, nu = (size_t)G.K * BB;
    D->fwd.reserve(nf); D->bwd.reserve(nbw); D->UR.reserve(nu);
    CUDA_CHECK(cudaMemsetAsync(D->fwd.p, 0, nf * sizeof(double), s));
    CUDA_CHECK(cudaMemsetAsync(D->bwd.p, 0, nbw * sizeof(double), s));
    CUDA_CHECK(cudaMemsetAsync(D->UR.p, 0, nu * sizeof(double), s));
    ChainLevelHost Hl;
    Hl.G = G;
    F.levels.push_back(Hl);
    if (n > 0) {
      k_cf_factor_level<B><<<(G.K + 127) / 128, 128, 0, s>>>(G, Sv->A.p, Sv->U.p, D->fwd.p, D->bwd.p, D->UR.p, Sv->SL.p,
                                                            Sv->SR.p, Sv->CP.p, S.d_flag.p);
      check_launch(h);
    }
    const int nn = G.top ? 0 : G.K - 1;
    if (nn == 0) break;
    k_cf_next_level<B><<<(nn * BB + kThreads - 1) / kThreads, kThreads, 0, s>>>(G, nn, Sv->A.p, Sv->SL.p, Sv->SR.p, Sv->CP.p,
                                                                              S.lv[lvi + 1]->A.p, S.lv[lvi + 1]->U.p);
    check_launch(h);
  }
  // landmark Schur complement S_L = C - B^T T^-1 B
  if (l > 0) {
    C->W.reserve((size_t)std::max(n, 1) * B * l);
    CUDA_CHECK(cudaMemsetAsync(C->W.p, 0, (size_t)std::max(n, 1) * B * l * sizeof(double), s));
    if (nb > 0) {
      k_cf_scatter_border<<<(nb + kThreads - 1) / kThreads, kThreads, 0, s>>>(nb, l, S.d_bl_ptr.p, S.d_bl_row.p, C->bl_val.p,
                                                                             C->W.p);
      check_launch(h);
    }
    if (n > 0 && S.general) {
      gen_solve_device<B>(h, S.gsd, C->gen, n, C->W.p, l, l, nullptr);
    } else if (n > 0) {  // W <- T^-1 W with the solve kernels (l columns, leading dimension l)
      const int nl = (int)C->levels.size();
      std::vector<DevBuf<double>> &sol = S.wsol, &rhs = S.wrhs, &cL = S.wcL, &cR = S.wcR;  // kept across factorisations
      if ((int)sol.size() != nl) { sol = std::vector<DevBuf<double>>(nl); rhs = std::vector<DevBuf<double>>(nl);
                                   cL = std::vector<DevBuf<double>>(nl); cR = std::vector<DevBuf<double>>(nl); }
      for (int lv = 0; lv < nl; ++lv) {
        const ChunkGeo G = C->levels[lv]->G;
        if (lv > 0) { sol[lv].reserve((size_t)std::max(G.n, 1) * B * l); rhs[lv].reserve((size_t)std::max(G.n, 1) * B * l); }
        cL[lv].reserve((size_t)(G.K + 1) * B * l); cR[lv].reserve((size_t)(G.K + 1) * B * l);
        CUDA_CHECK(cudaMemsetAsync(cL[lv].p, 0, (size_t)(G.K + 1) * B * l * sizeof(double), s));
        CUDA_CHECK(cudaMemsetAsync(cR[lv].p, 0, (size_t)(G.K + 1) * B * l * sizeof(double), s));
      }
      auto solp = [&](int lv) { return lv == 0 ? C->W.p : sol[lv].p; };
      auto rhsp = [&](int lv) { return lv == 0 ? C->W.p : rhs[lv].p; };
      for (int lv = 0; lv < nl; ++lv) {
        ChainLevelDev *D = C->levels[lv];
        const int grid = (D->G.K * l + 127) / 128;
        const bool top = (lv == nl - 1);
        k_chain_forward<B><<<grid, 128, 0, s>>>(D->G, l, l, D->fwd.p, D->UR.p, solp(lv), rhsp(lv),
                                                lv > 0 ? rhsp(lv - 1) : nullptr, lv > 0 ? C->levels[lv - 1]->G.c : 0,
                                                lv > 0 ? cL[lv - 1].p : nullptr, lv > 0 ? cR[lv - 1].p : nullptr, cL[lv].p,
                                                cR[lv].p, top ? D->bwd.p : nullptr, nullptr);
        check_launch(h);
      }
      for (int lv = nl - 2; lv >= 0; --lv) {
        ChainLevelDev *D = C->levels[lv];
        const int grid = (D->G.K * l + 127) / 128;
        k_chain_backward<B><<<grid, 128, 0, s>>>(D->G, l, l, D->bwd.p, solp(lv), solp(lv + 1), nullptr);
        check_launch(h);
      }
    }
    k_cf_schur<<<l, kThreads, 0, s>>>(l, S.d_bl_ptr.p, S.d_bl_row.p, C->bl_val.p, C->W.p, S.d_C.p, S.d_SL.p);
    check_launch(h);
  }
  // what crosses PCIe: the flag and the l x l Schur complement
  int flag = 1;
  CUDA_CHECK(cudaMemcpyAsync(&flag, S.d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  std::vector<double> SLm((size_t)std::max(l, 1) * std::max(l, 1), 0.0);
  if (l > 0) CUDA_CHECK(cudaMemcpyAsync(SLm.data(), S.d_SL.p, (size_t)l * l * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  F.pos_def = flag != 0 && gen_pd;
  if (l > 0) {
    for (int a = 0; a < l; ++a)
      for (int b = a + 1; b < l; ++b) {
        const double v = 0.5 * (SLm[(size_t)a * l + b] + SLm[(size_t)b * l + a]);
        SLm[(size_t)a * l + b] = SLm[(size_t)b * l + a] = v;
      }
    F.pos_def = dense_spd_inverse(SLm, l) && F.pos_def;
    if (want_solve && F.pos_def) C->SLinv.upload(SLm, s);
  }
  if (want_solve && F.pos_def) {
    C->rend_e.upload(S.rend_e, s);
    C->rinc_e.upload(S.rinc_e[pin], s);
    { std::vector<int> t(S.rinc_ptr[pin].begin(), S.rinc_ptr[pin].end()); C->rinc_ptr.upload(t, s);
      std::vector<int> k(S.rinc_k[pin].begin(), S.rinc_k[pin].end()); C->rinc_k.upload(k, s);
      std::vector<int> x(S.rend_x.begin(), S.rend_x.end()); C->rend_x.upload(x, s);
      std::vector<int> p(S.bl_ptr.begin(), S.bl_ptr.end()); C->bl_ptr.upload(p, s);
      std::vector<int> rw(S.bl_row.begin(), S.bl_row.end()); C->bl_row.upload(rw, s); }
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
}


// The structure of this handle's chain factorisation, built on first use (throws ENOTIMPL for non-chain graphs,
// every time: the verdict is cached).
inline ChainSym &chain_symbolic(H *h) {
  if (h->chain_sym_state == 2) throw Error(CORA_B200_ENOTIMPL, h->chain_sym_error);
  if (h->chain_sym_state == 0) {
    ChainSym *S = new ChainSym();
    try {
      try {
        if (h->HL.D1 == 3) chain_symbolic_build<3>(*S, h->HL, false);
        else chain_symbolic_build<4>(*S, h->HL, false);
      } catch (const Error &e) {
        if (e.code != CORA_B200_ENOTIMPL) throw;
        // not a chain (loop closures, several robots): the general sparse block Cholesky
        // (CORA_B200_GENERAL_CHOLESKY=0 keeps the "no factorisation" behaviour reachable for its tests)
        if (const char *g = getenv("CORA_B200_GENERAL_CHOLESKY")) if (atoi(g) == 0) throw;
        delete S;
        S = new ChainSym();
        if (h->HL.D1 == 3) chain_symbolic_build<3>(*S, h->HL, true);
        else chain_symbolic_build<4>(*S, h->HL, true);
      }
    } catch (const Error &e) {
      delete S;
      if (e.code == CORA_B200_ENOTIMPL) { h->chain_sym_state = 2; h->chain_sym_error = e.what(); }
      throw;
    }
    chain_symbolic_upload(h, *S);
    S->built = true;
    h->chain_sym = S;
    h->chain_sym_state = 1;
  }
  return *h->chain_sym;
}

inline ChainChol *build_chain_chol(H *h, const double *d_bval, const double *d_sdiag, double shift, bool pin_last,
                                   bool *pos_def, bool want_solve, bool trans_only) {
  ChainSym &S = chain_symbolic(h);
  // recycle the buffers of the last released factor (not for the long-lived translation factor)
  ChainChol *C = (h->chol_spare && !trans_only) ? h->chol_spare : new ChainChol();
  if (C == h->chol_spare) h->chol_spare = nullptr;
  C->ws_cols = C->ws_cols;  // (work vectors of the apply are sized by columns only: still valid)
  try {
    C->B = h->HL.D1;
    if (h->HL.D1 == 3) chain_factor_device<3>(h, S, C, d_bval, d_sdiag, shift, pin_last, want_solve, trans_only);
    else chain_factor_device<4>(h, S, C, d_bval, d_sdiag, shift, pin_last, want_solve, trans_only);
    *pos_def = C->host.pos_def;
  } catch (...) {
    delete C;
    throw;
  }
  return C;
}

inline void destroy_chain_sym(ChainSym *s) { delete s; }

// Release a factor: its device buffers are kept for the next factorisation of the same handle (the PSD tests and the
// shift search of the certification factor the same structure again and again).
inline void release_chain_chol(H *h, ChainChol *C) {
  if (!C) return;
  if (h->chol_spare == nullptr) h->chol_spare = C;
  else delete C;
}

}  // namespace cora_b200
