// chain_chol.cuh -- exact Cholesky-type factorisation and solve of M = (values + shift I)
// on the structure of Q for ODOMETRY-CHAIN + LANDMARK graphs (every BASELINE configuration;
// SURVEY F12), used for
//   * Preconditioner::RegularizedCholesky: M = (Q + lambda I) with the last row/column removed
//     and the last row of the result pinned to zero (src/CORA_problem.cpp:544-614,
//     src/CORA_preconditioners.cpp:16-83), and
//   * the positive-definiteness test of S + eta I in fast_verification
//     (src/CORA_utils.cpp:33-57), where the reference calls CholmodSupernodalLLT.
// The reference hands both to CHOLMOD (a general sparse Cholesky, not in /root/reference);
// the same matrix is factored here in a GPU-friendly elimination order:
//   1. range rows (diagonal, each coupled to two translations) are eliminated first,
//   2. the pose chain -- block tridiagonal, blocks of d+1 -- is factored by recursive
//      chunking: chunks of kChunk consecutive blocks are eliminated independently (one
//      thread per chunk and right-hand-side column), their last block is a separator; the
//      separators form a kChunk-times shorter chain that is treated the same way,
//   3. the l landmark columns are a dense border closed with an l x l Schur complement.
// Positive definiteness <=> every pivot block of this block elimination is positive definite.
// Graphs whose pose coupling is not block tridiagonal (loop closures, several robots: TIERS, MR.CLAM) keep
// steps 1 and 3 and replace step 2 by the general sparse block Cholesky of gen_chol.hpp (nested-dissection
// order, level-scheduled solves).
#pragma once
#include <cmath>
#include <map>

#include "gen_chol_dev.cuh"
#include "ops.cuh"

#ifdef __CUDACC__
#define CB_HD __host__ __device__ __forceinline__
#define CB_UNROLL _Pragma("unroll")
#else
#define CB_HD inline
#define CB_UNROLL
#endif

namespace cora_b200 {

// prefetch one cache line into L2 (device only): the per-chunk substitutions are serial chains of
// small block steps whose coefficients stream from HBM; requesting step j+1 while step j computes
// turns a DRAM round trip per step into an L2 hit
CB_HD void chain_prefetch(const double *p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

constexpr int kChunk = 8;    // chain blocks per chunk (7 interior + 1 separator): the substitutions are serial per chunk, so the chunk length times the number of levels is the latency of one apply
constexpr int kPrefetchAllK = 4096;  // levels with at most this many chunks request a whole chunk's coefficients at once
constexpr int kTopMax = 8;   // a level with at most this many blocks is solved by one thread per column

// ----------------------------------------------------------- B x B block helpers ---
template <int B>
CB_HD void blk_zero(double *X) {
  for (int i = 0; i < B * B; ++i) X[i] = 0.0;
}
template <int B>
CB_HD void blk_copy(const double *X, double *Y) {
  for (int i = 0; i < B * B; ++i) Y[i] = X[i];
}
// Z = X * Y
template <int B>
CB_HD void blk_mul(const double *X, const double *Y, double *Z) {
  for (int a = 0; a < B; ++a)
    for (int b = 0; b < B; ++b) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += X[a * B + k] * Y[k * B + b];
      Z[a * B + b] = s;
    }
}
// Z = X^T * Y
template <int B>
CB_HD void blk_mul_tn(const double *X, const double *Y, double *Z) {
  for (int a = 0; a < B; ++a)
    for (int b = 0; b < B; ++b) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += X[k * B + a] * Y[k * B + b];
      Z[a * B + b] = s;
    }
}
// Z = X * Y^T
template <int B>
CB_HD void blk_mul_nt(const double *X, const double *Y, double *Z) {
  for (int a = 0; a < B; ++a)
    for (int b = 0; b < B; ++b) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += X[a * B + k] * Y[b * B + k];
      Z[a * B + b] = s;
    }
}
// Cholesky-based inverse of a symmetric B x B block; false if a pivot is not positive.
template <int B>
CB_HD bool blk_spd_inverse(const double *P, double *Pinv) {
  double L[B * B];
  bool ok = true;
  for (int j = 0; j < B; ++j) {
    double s = P[j * B + j];
    for (int k = 0; k < j; ++k) s -= L[j * B + k] * L[j * B + k];
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    const double dj = sqrt(s);
    L[j * B + j] = dj;
    for (int i = j + 1; i < B; ++i) {
      double t = 0.5 * (P[i * B + j] + P[j * B + i]);
      for (int k = 0; k < j; ++k) t -= L[i * B + k] * L[j * B + k];
      L[i * B + j] = t / dj;
    }
  }
  // Linv (lower triangular), then Pinv = Linv^T Linv
  double Li[B * B];
  for (int i = 0; i < B * B; ++i) Li[i] = 0.0;
  for (int j = 0; j < B; ++j) {
    Li[j * B + j] = 1.0 / L[j * B + j];
    for (int i = j + 1; i < B; ++i) {
      double t = 0.0;
      for (int k = j; k < i; ++k) t -= L[i * B + k] * Li[k * B + j];
      Li[i * B + j] = t / L[i * B + i];
    }
  }
  for (int a = 0; a < B; ++a)
    for (int b = 0; b < B; ++b) {
      double s = 0.0;
      for (int k = (a > b ? a : b); k < B; ++k) s += Li[k * B + a] * Li[k * B + b];
      Pinv[a * B + b] = s;
    }
  return ok;
}

// ------------------------------------------------------------ chunk geometry ------
struct ChunkGeo {
  int n, c, K;
  bool top;
  CB_HD int first(int k) const { return k * c; }
  CB_HD bool has_left(int k) const { return k > 0; }
  CB_HD bool has_right(int k) const { return !top && k < K - 1; }
  CB_HD int interior(int k) const {
    if (top) return n;
    return k < K - 1 ? c - 1 : n - k * c;
  }
};
inline ChunkGeo make_geo(int n) {
  ChunkGeo g;
  g.n = n;
  if (n <= kTopMax) { g.top = true; g.c = n > 0 ? n : 1; g.K = 1; }
  else { g.top = false; g.c = kChunk; g.K = (n + kChunk - 1) / kChunk; }
  return g;
}

// Factor one chunk.  A, U: level matrices ([node][B*B] row-major; U[g] = block (g, g+1)).
// fwd: [j][3*BB][K] = (U_{g-1}^T, P^-1, F); bwd: [j][2*BB][K] = (P^-1 U_g, P^-1 F^T);
// UR[k] = U of the last interior block (coupling to the right separator);
// SL/SR[k]: Schur updates of the left/right separator diagonal; CP[k]: new coupling (sL, sR).
template <int B>
CB_HD bool factor_chunk(const ChunkGeo G, int k, const double *A, const double *U, double *fwd, double *bwd,
                        double *UR, double *SL, double *SR, double *CP) {
  constexpr int BB = B * B;
  const int g0 = G.first(k), L = G.interior(k), K = G.K;
  const bool hasL = G.has_left(k), hasR = G.has_right(k);
  double F[BB], Pinv[BB], Pprev[BB], acc[BB], T1[BB], T2[BB], Ug[BB], Ut[BB];
  bool ok = true;
  if (hasL) blk_copy<B>(U + (size_t)(g0 - 1) * BB, F); else blk_zero<B>(F);
  blk_zero<B>(acc);
  blk_zero<B>(Pprev);
  for (int j = 0; j < L; ++j) {
    const int g = g0 + j;
    double P[BB];
    blk_copy<B>(A + (size_t)g * BB, P);
    if (j > 0) {
      const double *Up = U + (size_t)(g - 1) * BB;
      blk_mul<B>(Pprev, Up, T1);     // P_{g-1}^-1 U_{g-1}
      blk_mul_tn<B>(Up, T1, T2);     // U^T P^-1 U
      for (int i = 0; i < BB; ++i) P[i] -= T2[i];
      for (int a = 0; a < B; ++a)
        for (int b = 0; b < B; ++b) Ut[a * B + b] = Up[b * B + a];
    } else {
      blk_zero<B>(Ut);
    }
    ok = blk_spd_inverse<B>(P, Pinv) && ok;
    if (j < L - 1 || hasR) blk_copy<B>(U + (size_t)g * BB, Ug); else blk_zero<B>(Ug);
    double Ub[BB], Gm[BB];
    blk_mul<B>(Pinv, Ug, Ub);
    blk_mul_nt<B>(Pinv, F, Gm);
    for (int e = 0; e < BB; ++e) {
      fwd[((size_t)j * 3 * BB + e) * K + k] = Ut[e];
      fwd[((size_t)j * 3 * BB + BB + e) * K + k] = Pinv[e];
      fwd[((size_t)j * 3 * BB + 2 * BB + e) * K + k] = F[e];
      bwd[((size_t)j * 2 * BB + e) * K + k] = Ub[e];
      bwd[((size_t)j * 2 * BB + BB + e) * K + k] = Gm[e];
    }
    blk_mul<B>(F, Gm, T1);  // F P^-1 F^T
    for (int i = 0; i < BB; ++i) acc[i] += T1[i];
    blk_mul<B>(F, Ub, T1);  // F P^-1 U_g  -> fill (sL, g+1) = -that
    for (int i = 0; i < BB; ++i) F[i] = -T1[i];
    blk_copy<B>(Pinv, Pprev);
    if (j == L - 1) {
      blk_copy<B>(Ug, UR + (size_t)k * BB);
      blk_mul_tn<B>(Ug, Ub, T2);  // U^T P^-1 U
      blk_copy<B>(T2, SR + (size_t)k * BB);
    }
  }
  blk_copy<B>(acc, SL + (size_t)k * BB);
  blk_copy<B>(F, CP + (size_t)k * BB);
  return ok;
}

// Forward elimination of one chunk for one right-hand-side column.  Vectors are
// [node][B][ld] with the column at offset col.  rhs_prev == nullptr: level 0 (b is already in
// sol, in place); otherwise b_g = rhs_prev[sep_prev(g)] - cR_prev[g] - cL_prev[g+1] and the
// separator's b is materialised in rhs_cur for the next level.
template <int B>
CB_HD void forward_chunk(const ChunkGeo G, int k, int col, int ld, const double *fwd, const double *UR,
                         double *sol, double *rhs_cur, const double *rhs_prev, int c_prev,
                         const double *cL_prev, const double *cR_prev, double *cL, double *cR) {
  constexpr int BB = B * B;
  const int g0 = G.first(k), L = G.interior(k), K = G.K;
  const bool hasR = G.has_right(k);
  double w[B], acc[B];
  CB_UNROLL
  for (int a = 0; a < B; ++a) { w[a] = 0.0; acc[a] = 0.0; }
  const int nodes = L + (hasR ? 1 : 0);
  for (int j = 0; j < nodes; ++j) {
    const int g = g0 + j;
    double b[B];
    if (rhs_prev != nullptr) {
      const size_t s = (size_t)((g + 1) * c_prev - 1);
      CB_UNROLL
      for (int a = 0; a < B; ++a)
        b[a] = rhs_prev[(s * B + a) * ld + col] - cR_prev[((size_t)g * B + a) * ld + col] -
               cL_prev[((size_t)(g + 1) * B + a) * ld + col];
      if (j == L)  // the separator: keep its right-hand side for the next level
        CB_UNROLL
        for (int a = 0; a < B; ++a) rhs_cur[((size_t)g * B + a) * ld + col] = b[a];
    } else {
      if (j == L) break;  // level 0: the separator's b stays where it is (sol is in place)
      CB_UNROLL
      for (int a = 0; a < B; ++a) b[a] = sol[((size_t)g * B + a) * ld + col];
    }
    if (j == L) break;
    const double *f = fwd + (size_t)j * 3 * BB * K + k;
    if (K <= kPrefetchAllK) {
      // a small level: its few threads walk the chunk at one memory round trip per step, and the coefficients
      // (streamed out of L2 by the rest of the CG iteration) come from HBM -- request the whole chunk up front,
      // once per chunk (the columns of a chunk share the coefficients)
      if (j == 0 && col == 0)
        for (int jj = 1; jj < L; ++jj)
          for (int e = 0; e < 3 * BB; ++e) chain_prefetch(fwd + ((size_t)jj * 3 * BB + e) * K + k);
    } else if (j + 1 < L) {
      for (int e = 0; e < 3 * BB; ++e) chain_prefetch(f + (size_t)(3 * BB + e) * K);
    }
    double y[B];
    CB_UNROLL
    for (int a = 0; a < B; ++a) {
      double s = b[a];
      CB_UNROLL
      for (int q = 0; q < B; ++q) s -= f[(size_t)(a * B + q) * K] * w[q];
      y[a] = s;
    }
    CB_UNROLL
    for (int a = 0; a < B; ++a) {
      double s = 0.0;
      CB_UNROLL
      for (int q = 0; q < B; ++q) s += f[(size_t)(BB + a * B + q) * K] * y[q];
      w[a] = s;
    }
    CB_UNROLL
    for (int a = 0; a < B; ++a) {
      double s = acc[a];
      CB_UNROLL
      for (int q = 0; q < B; ++q) s += f[(size_t)(2 * BB + a * B + q) * K] * w[q];
      acc[a] = s;
    }
    CB_UNROLL
    for (int a = 0; a < B; ++a) sol[((size_t)g * B + a) * ld + col] = w[a];
  }
  CB_UNROLL
  for (int a = 0; a < B; ++a) {
    cL[((size_t)k * B + a) * ld + col] = acc[a];
    double s = 0.0;
    if (hasR)
      CB_UNROLL
      for (int q = 0; q < B; ++q) s += UR[(size_t)k * BB + q * B + a] * w[q];  // U^T w_last
    cR[((size_t)k * B + a) * ld + col] = s;
  }
}

// Back substitution of one chunk for one column; xsep = solution of the next level
// (separator k is its node k), nullptr at the top level.
template <int B>
CB_HD void backward_chunk(const ChunkGeo G, int k, int col, int ld, const double *bwd, double *sol,
                          const double *xsep) {
  constexpr int BB = B * B;
  const int g0 = G.first(k), L = G.interior(k), K = G.K;
  double xL[B], xn[B];
  CB_UNROLL
  for (int a = 0; a < B; ++a) {
    xL[a] = (xsep != nullptr && G.has_left(k)) ? xsep[((size_t)(k - 1) * B + a) * ld + col] : 0.0;
    xn[a] = (xsep != nullptr && G.has_right(k)) ? xsep[((size_t)k * B + a) * ld + col] : 0.0;
  }
  if (G.has_right(k))
    CB_UNROLL
    for (int a = 0; a < B; ++a) sol[((size_t)(g0 + L) * B + a) * ld + col] = xn[a];
  for (int j = L - 1; j >= 0; --j) {
    const int g = g0 + j;
    const double *f = bwd + (size_t)j * 2 * BB * K + k;
    if (K <= kPrefetchAllK) {
      if (j == L - 1 && col == 0)
        for (int jj = 0; jj < L - 1; ++jj)
          for (int e = 0; e < 2 * BB; ++e) chain_prefetch(bwd + ((size_t)jj * 2 * BB + e) * K + k);
    } else if (j > 0) {
      for (int e = 0; e < 2 * BB; ++e) chain_prefetch(f - (size_t)(2 * BB - e) * K);
    }
    double x[B];
    CB_UNROLL
    for (int a = 0; a < B; ++a) {
      double s = sol[((size_t)g * B + a) * ld + col];
      CB_UNROLL
      for (int q = 0; q < B; ++q)
        s -= f[(size_t)(a * B + q) * K] * xn[q] + f[(size_t)(BB + a * B + q) * K] * xL[q];
      x[a] = s;
    }
    CB_UNROLL
    for (int a = 0; a < B; ++a) {
      sol[((size_t)g * B + a) * ld + col] = x[a];
      xn[a] = x[a];
    }
  }
}

// ------------------------------------------------------------------ host factor ---
struct ChainLevelHost {
  ChunkGeo G;
  std::vector<double> A, U, fwd, bwd, UR;
};

struct ChainFactorHost {
  int B = 0, n = 0, l = 0, m = 0;
  bool pos_def = true;
  std::vector<ChainLevelHost> levels;
  // ranges
  std::vector<double> rdinv;                 // 1/delta_k
  std::vector<int32_t> rinc_ptr, rinc_k;     // per translation (n poses, then l landmarks): incident ranges
  std::vector<double> rinc_e;
  std::vector<int32_t> rend_x;               // per range: 2 endpoints (translation index, -1 = none)
  std::vector<double> rend_e;
  // border
  std::vector<int32_t> bl_ptr, bl_row;       // per landmark: (pose-section row, value)
  std::vector<double> bl_val;
  std::vector<double> W;                     // (B*n) x l row-major: T^-1 Bdense
  std::vector<double> SLinv;                 // l x l
  int pinned_landmark = -1;                  // landmark index pinned to zero (or -1)
  int pinned_pose_row = -1;                  // pose-section row pinned to zero (or -1)
  // pose graph with loop closures / several robots: general sparse block Cholesky (gen_chol.hpp) instead of levels
  bool general = false;
  GenSym gsym;
  std::vector<double> gL, gDinv;
};

template <int B>
inline void chain_factor_levels(ChainFactorHost &F, std::vector<double> &A0, std::vector<double> &U0, int n) {
  constexpr int BB = B * B;
  F.levels.clear();
  std::vector<double> A = std::move(A0), U = std::move(U0);
  int cur = n;
  while (true) {
    ChainLevelHost Lv;
    Lv.G = make_geo(cur);
    const ChunkGeo G = Lv.G;
    Lv.fwd.assign((size_t)G.c * 3 * BB * G.K, 0.0);
    Lv.bwd.assign((size_t)G.c * 2 * BB * G.K, 0.0);
    Lv.UR.assign((size_t)G.K * BB, 0.0);
    std::vector<double> SL((size_t)G.K * BB, 0.0), SR((size_t)G.K * BB, 0.0), CP((size_t)G.K * BB, 0.0);
    bool ok = true;
    for (int k = 0; k < G.K; ++k)
      ok = factor_chunk<B>(G, k, A.data(), U.data(), Lv.fwd.data(), Lv.bwd.data(), Lv.UR.data(), SL.data(),
                           SR.data(), CP.data()) && ok;
    F.pos_def = F.pos_def && ok;
    const int nn = G.top ? 0 : G.K - 1;
    std::vector<double> A2((size_t)std::max(nn, 1) * BB, 0.0), U2((size_t)std::max(nn, 1) * BB, 0.0);
    for (int q = 0; q < nn; ++q) {
      const size_t s = (size_t)((q + 1) * G.c - 1);
      for (int e = 0; e < BB; ++e) {
        A2[(size_t)q * BB + e] = A[s * BB + e] - SR[(size_t)q * BB + e] - SL[(size_t)(q + 1) * BB + e];
        U2[(size_t)q * BB + e] = (q < nn - 1) ? CP[(size_t)(q + 1) * BB + e] : 0.0;
      }
    }
    Lv.A.swap(A);
    Lv.U.swap(U);
    Lv.A.clear(); Lv.A.shrink_to_fit();
    Lv.U.clear(); Lv.U.shrink_to_fit();
    F.levels.push_back(std::move(Lv));
    if (nn == 0) break;
    A.swap(A2);
    U.swap(U2);
    cur = nn;
  }
}

// Host execution of the chain solve (same per-chunk functions the kernels call); used at
// set-up time for W = T^-1 B and by the CPU test hook.  X: [n][B][ld], in place.
template <int B>
inline void chain_solve_host(const ChainFactorHost &F, double *X, int ld, int ncols) {
  if (F.general) {
    gen_solve_host<B>(F.gsym, F.gL.data(), F.gDinv.data(), X, ld, ncols);
    return;
  }
  const int nl = (int)F.levels.size();
  std::vector<std::vector<double>> sol(nl), rhs(nl), cL(nl), cR(nl);
  for (int lv = 0; lv < nl; ++lv) {
    const ChunkGeo G = F.levels[lv].G;
    if (lv > 0) { sol[lv].assign((size_t)G.n * B * ld, 0.0); rhs[lv].assign((size_t)G.n * B * ld, 0.0); }
    cL[lv].assign((size_t)(G.K + 1) * B * ld, 0.0);
    cR[lv].assign((size_t)(G.K + 1) * B * ld, 0.0);
  }
  auto solp = [&](int lv) { return lv == 0 ? X : sol[lv].data(); };
  auto rhsp = [&](int lv) { return lv == 0 ? X : rhs[lv].data(); };
  for (int lv = 0; lv < nl; ++lv) {
    const ChainLevelHost &Lv = F.levels[lv];
    for (int k = 0; k < Lv.G.K; ++k)
      for (int c = 0; c < ncols; ++c)
        forward_chunk<B>(Lv.G, k, c, ld, Lv.fwd.data(), Lv.UR.data(), solp(lv), rhsp(lv),
                         lv > 0 ? rhsp(lv - 1) : nullptr, lv > 0 ? F.levels[lv - 1].G.c : 0,
                         lv > 0 ? cL[lv - 1].data() : nullptr, lv > 0 ? cR[lv - 1].data() : nullptr,
                         cL[lv].data(), cR[lv].data());
  }
  for (int lv = nl - 1; lv >= 0; --lv) {
    const ChainLevelHost &Lv = F.levels[lv];
    for (int k = 0; k < Lv.G.K; ++k)
      for (int c = 0; c < ncols; ++c)
        backward_chunk<B>(Lv.G, k, c, ld, Lv.bwd.data(), solp(lv), lv + 1 < nl ? solp(lv + 1) : nullptr);
  }
}

// Dense symmetric positive definite inverse (l x l, landmark Schur complement), host.
inline bool dense_spd_inverse(std::vector<double> &S, int l) {
  std::vector<double> Lm((size_t)l * l, 0.0);
  bool ok = true;
  for (int j = 0; j < l; ++j) {
    double s = S[(size_t)j * l + j];
    for (int k = 0; k < j; ++k) s -= Lm[(size_t)j * l + k] * Lm[(size_t)j * l + k];
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    const double dj = std::sqrt(s);
    Lm[(size_t)j * l + j] = dj;
    for (int i = j + 1; i < l; ++i) {
      double t = S[(size_t)i * l + j];
      for (int k = 0; k < j; ++k) t -= Lm[(size_t)i * l + k] * Lm[(size_t)j * l + k];
      Lm[(size_t)i * l + j] = t / dj;
    }
  }
  std::vector<double> Li((size_t)l * l, 0.0);
  for (int j = 0; j < l; ++j) {
    Li[(size_t)j * l + j] = 1.0 / Lm[(size_t)j * l + j];
    for (int i = j + 1; i < l; ++i) {
      double t = 0.0;
      for (int k = j; k < i; ++k) t -= Lm[(size_t)i * l + k] * Li[(size_t)k * l + j];
      Li[(size_t)i * l + j] = t / Lm[(size_t)i * l + i];
    }
  }
  for (int a = 0; a < l; ++a)
    for (int b = 0; b < l; ++b) {
      double s = 0.0;
      for (int k = std::max(a, b); k < l; ++k) s += Li[(size_t)k * l + a] * Li[(size_t)k * l + b];
      S[(size_t)a * l + b] = s;
    }
  return ok;
}

// Build the factor of M = values + shift*I from the host layout (structure + spill values) and
// the given block-ELL / scalar-diagonal values.  Throws ENOTIMPL when the graph is not a chain.
template <int B>
inline void chain_factor_host(ChainFactorHost &F, const HostLayout &L, const double *bval,
                              const double *sdiag, double shift, bool pin_last, bool want_solve) {
  constexpr int BB = B * B;
  const int n = L.n, l = L.l, m = L.m, D1 = L.D1, d = L.d;
  F.B = B; F.n = n; F.l = l; F.m = m; F.pos_def = true;
  F.pinned_landmark = -1; F.pinned_pose_row = -1;
  if (pin_last) {
    if (l > 0) F.pinned_landmark = l - 1;
    else if (n > 0) F.pinned_pose_row = D1 * (n - 1) + d;
  }
  auto not_chain = [](const char *why) {
    throw Error(CORA_B200_ENOTIMPL,
                std::string("RegularizedCholesky / Cholesky certificate: unsupported coupling (") + why + ")");
  };
  // couplings between non-adjacent poses (loop closures, other robots): block M_ij (rows of i, columns of j), i < j
  std::map<std::pair<int32_t, int32_t>, std::vector<double>> extra;
  auto extra_block = [&](int i, int j) -> double * {
    auto &b = extra[{(int32_t)i, (int32_t)j}];
    if (b.empty()) b.assign(BB, 0.0);
    return b.data();
  };
  // ---- chain blocks from the block-ELL ----
  std::vector<double> A((size_t)std::max(n, 1) * BB, 0.0), U((size_t)std::max(n, 1) * BB, 0.0);
  for (int i = 0; i < n; ++i) {
    const int t = i / L.TP, p = i % L.TP;
    const int S = L.tile_slots[t];
    for (int s = 0; s < S; ++s) {
      const int j = L.bcol[L.tile_coff[t] + (int64_t)s * L.TP + p] / D1;
      const double *bv = bval + L.tile_boff[t] + (int64_t)s * BB * L.TP + p;
      bool nz = false;
      for (int e = 0; e < BB; ++e) nz = nz || bv[(int64_t)e * L.TP] != 0.0;
      if (s > 0 && j == i) continue;  // padding slot
      if (j == i) {
        for (int e = 0; e < BB; ++e) A[(size_t)i * BB + e] = bv[(int64_t)e * L.TP];
      } else if (j == i + 1) {
        for (int e = 0; e < BB; ++e) U[(size_t)i * BB + e] = bv[(int64_t)e * L.TP];
      } else if (j == i - 1) {
        // lower block = transpose of U[i-1] (Q is symmetric); nothing to store
      } else if (nz && j > i) {  // (the transposed copy in row j is the same coupling)
        double *eb = extra_block(i, j);
        for (int e = 0; e < BB; ++e) eb[e] += bv[(int64_t)e * L.TP];
      }
    }
    for (int a = 0; a < B; ++a) A[(size_t)i * BB + a * B + a] += shift;
  }
  // ---- landmark block C (dense), border B (sparse, by landmark) ----
  std::vector<double> C((size_t)std::max(l, 1) * std::max(l, 1), 0.0);
  for (int j = 0; j < l; ++j) C[(size_t)j * l + j] = sdiag[j] + shift;
  struct BEnt { int32_t row, j; double v; };
  std::vector<BEnt> bents;
  // ---- range rows ----
  F.rdinv.assign(std::max(m, 1), 0.0);
  F.rend_x.assign((size_t)std::max(m, 1) * 2, -1);
  F.rend_e.assign((size_t)std::max(m, 1) * 2, 0.0);
  const int64_t rg0 = L.nPoseRows + l;
  auto trans_index = [&](int64_t ci) -> int {  // internal row -> translation index (poses, then landmarks)
    if (ci < L.nPoseRows) {
      if (ci % D1 != d) return -1;
      return (int)(ci / D1);
    }
    if (ci < rg0) return n + (int)(ci - L.nPoseRows);
    return -1;
  };
  auto for_group = [&](int64_t g, auto &&fn) {
    for (int32_t k = L.grp_ptr[g]; k < L.grp_ptr[g + 1]; ++k) fn(L.rem_pk[k], L.rem_val[k]);
    // long groups
    auto it = std::lower_bound(L.long_grp.begin(), L.long_grp.end(), (int32_t)g);
    if (it != L.long_grp.end() && *it == g) {
      const size_t q = it - L.long_grp.begin();
      for (int32_t k = L.long_ptr[q]; k < L.long_ptr[q + 1]; ++k) fn(L.long_pk[k], L.long_val[k]);
    }
  };
  for (int k = 0; k < m; ++k) {
    const double delta = sdiag[l + k] + shift;
    if (!(delta > 0.0)) { F.pos_def = false; F.rdinv[k] = 1.0; }
    else F.rdinv[k] = 1.0 / delta;
    int cnt = 0;
    for_group((int64_t)n + l + k, [&](uint32_t pk, double v) {
      const int64_t ci = pk & kColMask;
      const int x = trans_index(ci);
      if (x < 0) not_chain("a range row is coupled to a non-translation variable");
      if (cnt >= 2) not_chain("a range row has more than two couplings");
      F.rend_x[(size_t)k * 2 + cnt] = x;
      F.rend_e[(size_t)k * 2 + cnt] = v;
      ++cnt;
    });
  }
  auto pinned_trans = [&](int x) {
    if (x < n) return F.pinned_pose_row >= 0 && x == n - 1;
    return F.pinned_landmark >= 0 && (x - n) == F.pinned_landmark;
  };
  // Schur complement of the range rows onto the translations
  auto add_tt = [&](int x, int y, double v) {  // M[t_x, t_y] += v   (called for both orders)
    if (x < n && y < n) {
      if (x == y) A[(size_t)x * BB + d * B + d] += v;
      else if (y == x + 1) U[(size_t)x * BB + d * B + d] += v;
      else if (y == x - 1) { /* transpose of the above */ }
      else if (y > x) extra_block(x, y)[d * B + d] += v;
    } else if (x < n && y >= n) {
      bents.push_back({(int32_t)(x * D1 + d), (int32_t)(y - n), v});
    } else if (x >= n && y >= n) {
      C[(size_t)(x - n) * l + (y - n)] += v;
    }
  };
  for (int k = 0; k < m; ++k) {
    const double di = F.rdinv[k];
    for (int p = 0; p < 2; ++p)
      for (int q = 0; q < 2; ++q) {
        const int x = F.rend_x[(size_t)k * 2 + p], y = F.rend_x[(size_t)k * 2 + q];
        if (x < 0 || y < 0) continue;
        add_tt(x, y, -F.rend_e[(size_t)k * 2 + p] * F.rend_e[(size_t)k * 2 + q] * di);
      }
  }
  // pose rows: couplings to landmarks (border) from the spill
  for (int i = 0; i < n; ++i)
    for_group(i, [&](uint32_t pk, double v) {
      const int64_t ci = pk & kColMask;
      const int a = (int)(pk >> 30);
      if (ci < L.nPoseRows) {  // pose-pose coupling beyond the block-ELL slots
        const int j = (int)(ci / D1), b = (int)(ci % D1);
        if (j == i + 1) U[(size_t)i * BB + a * B + b] += v;
        else if (j > i) extra_block(i, j)[a * B + b] += v;
        return;
      }
      if (ci < rg0) bents.push_back({(int32_t)(i * D1 + a), (int32_t)(ci - L.nPoseRows), v});
    });
  // landmark rows: landmark-landmark couplings
  for (int j = 0; j < l; ++j)
    for_group((int64_t)n + j, [&](uint32_t pk, double v) {
      const int64_t ci = pk & kColMask;
      if (ci >= L.nPoseRows && ci < rg0) C[(size_t)j * l + (ci - L.nPoseRows)] += v;
    });
  // ---- pinning (src/CORA_preconditioners.cpp:77-80: last row of the solution is zero) ----
  if (F.pinned_pose_row >= 0) {
    const int i = n - 1;
    for (int a = 0; a < B; ++a) { A[(size_t)i * BB + d * B + a] = 0.0; A[(size_t)i * BB + a * B + d] = 0.0; }
    A[(size_t)i * BB + d * B + d] = 1.0;
    if (i > 0)
      for (int a = 0; a < B; ++a) U[(size_t)(i - 1) * BB + a * B + d] = 0.0;
    for (auto &kv : extra)
      if (kv.first.second == i)
        for (int a = 0; a < B; ++a) kv.second[a * B + d] = 0.0;
  }
  if (F.pinned_landmark >= 0) {
    const int j = F.pinned_landmark;
    for (int q = 0; q < l; ++q) { C[(size_t)j * l + q] = 0.0; C[(size_t)q * l + j] = 0.0; }
    C[(size_t)j * l + j] = 1.0;
  }
  // border by landmark (sorted by row so that the reduction order is fixed)
  std::stable_sort(bents.begin(), bents.end(), [](const BEnt &x, const BEnt &y) {
    return x.j != y.j ? x.j < y.j : x.row < y.row;
  });
  F.bl_ptr.assign((size_t)l + 1, 0);
  F.bl_row.clear(); F.bl_val.clear();
  for (size_t q = 0; q < bents.size();) {
    size_t e = q;
    double v = 0.0;
    while (e < bents.size() && bents[e].j == bents[q].j && bents[e].row == bents[q].row) v += bents[e++].v;
    const bool pinned = (bents[q].j == F.pinned_landmark) || (bents[q].row == F.pinned_pose_row);
    if (!pinned && v != 0.0) {
      F.bl_row.push_back(bents[q].row);
      F.bl_val.push_back(v);
      ++F.bl_ptr[bents[q].j + 1];
    }
    q = e;
  }
  for (int j = 0; j < l; ++j) F.bl_ptr[j + 1] += F.bl_ptr[j];
  // ---- factor the pose system: chain levels, or the general sparse block Cholesky ----
  F.general = !extra.empty();
  if (F.general) {
    std::vector<int32_t> ei, ej;
    std::vector<double> E;
    for (int i = 0; i + 1 < n; ++i) {
      ei.push_back(i); ej.push_back(i + 1);
      E.insert(E.end(), U.begin() + (size_t)i * BB, U.begin() + (size_t)(i + 1) * BB);
    }
    for (auto &kv : extra) {
      ei.push_back(kv.first.first); ej.push_back(kv.first.second);
      E.insert(E.end(), kv.second.begin(), kv.second.end());
    }
    gen_symbolic(F.gsym, n, ei, ej);
    F.pos_def = gen_numeric<B>(F.gsym, A.data(), E.data(), F.gL, F.gDinv) && F.pos_def;
  } else {
    chain_factor_levels<B>(F, A, U, n);
  }
  // ---- landmark Schur complement S_L = C - B^T T^-1 B ----
  if (l > 0) {
    std::vector<double> W((size_t)std::max(n, 1) * B * l, 0.0);
    for (int j = 0; j < l; ++j)
      for (int32_t q = F.bl_ptr[j]; q < F.bl_ptr[j + 1]; ++q) W[(size_t)F.bl_row[q] * l + j] += F.bl_val[q];
    if (n > 0) chain_solve_host<B>(F, W.data(), l, l);
    std::vector<double> SLm = C;
    for (int j = 0; j < l; ++j)
      for (int32_t q = F.bl_ptr[j]; q < F.bl_ptr[j + 1]; ++q) {
        const double bv = F.bl_val[q];
        const double *wr = W.data() + (size_t)F.bl_row[q] * l;
        for (int j2 = 0; j2 < l; ++j2) SLm[(size_t)j * l + j2] -= bv * wr[j2];
      }
    for (int a = 0; a < l; ++a)
      for (int b = a + 1; b < l; ++b) {
        const double s = 0.5 * (SLm[(size_t)a * l + b] + SLm[(size_t)b * l + a]);
        SLm[(size_t)a * l + b] = SLm[(size_t)b * l + a] = s;
      }
    F.pos_def = dense_spd_inverse(SLm, l) && F.pos_def;
    F.SLinv.swap(SLm);
    if (want_solve) F.W.swap(W);
  }
  // ---- range incidence by translation (for the forward elimination of the ranges) ----
  if (want_solve) {
    const int nt = n + l;
    F.rinc_ptr.assign((size_t)nt + 1, 0);
    for (int k = 0; k < m; ++k)
      for (int p = 0; p < 2; ++p) {
        const int x = F.rend_x[(size_t)k * 2 + p];
        if (x >= 0 && !pinned_trans(x)) ++F.rinc_ptr[x + 1];
      }
    for (int x = 0; x < nt; ++x) F.rinc_ptr[x + 1] += F.rinc_ptr[x];
    F.rinc_k.assign((size_t)F.rinc_ptr[nt], 0);
    F.rinc_e.assign((size_t)F.rinc_ptr[nt], 0.0);
    std::vector<int32_t> fill(F.rinc_ptr.begin(), F.rinc_ptr.end() - 1);
    for (int k = 0; k < m; ++k)
      for (int p = 0; p < 2; ++p) {
        const int x = F.rend_x[(size_t)k * 2 + p];
        if (x >= 0 && !pinned_trans(x)) {
          F.rinc_k[fill[x]] = k;
          F.rinc_e[fill[x]] = F.rend_e[(size_t)k * 2 + p];
          ++fill[x];
        }
      }
  }
}

// Host execution of the whole M^-1 apply (test hook + reference for the device kernels).
// V, Z: N x r row-major in the INTERNAL row order.
template <int B>
inline void chain_apply_host(const ChainFactorHost &F, const HostLayout &L, const double *V, double *Z, int r) {
  const int n = F.n, l = F.l, m = F.m, D1 = L.D1, d = L.d;
  const size_t np = (size_t)L.nPoseRows;
  std::vector<double> Y((size_t)L.N * r);
  for (size_t i = 0; i < (size_t)L.N * r; ++i) Y[i] = V[i];
  auto trow = [&](int x) -> size_t { return x < n ? (size_t)x * D1 + d : np + (x - n); };
  // 1. eliminate ranges
  for (int x = 0; x < n + l; ++x)
    for (int32_t q = F.rinc_ptr[x]; q < F.rinc_ptr[x + 1]; ++q) {
      const int k = F.rinc_k[q];
      for (int c = 0; c < r; ++c) Y[trow(x) * r + c] -= F.rinc_e[q] * F.rdinv[k] * V[(np + l + k) * r + c];
    }
  if (F.pinned_pose_row >= 0) for (int c = 0; c < r; ++c) Y[(size_t)F.pinned_pose_row * r + c] = 0.0;
  if (F.pinned_landmark >= 0) for (int c = 0; c < r; ++c) Y[(np + F.pinned_landmark) * r + c] = 0.0;
  // 2. chain
  if (n > 0) chain_solve_host<B>(F, Y.data(), r, r);
  // 3. landmarks
  std::vector<double> u((size_t)std::max(l, 1) * r, 0.0), zL((size_t)std::max(l, 1) * r, 0.0);
  for (int j = 0; j < l; ++j)
    for (int c = 0; c < r; ++c) {
      double s = Y[(np + j) * r + c];
      for (int32_t q = F.bl_ptr[j]; q < F.bl_ptr[j + 1]; ++q) s -= F.bl_val[q] * Y[(size_t)F.bl_row[q] * r + c];
      u[(size_t)j * r + c] = s;
    }
  for (int j = 0; j < l; ++j)
    for (int c = 0; c < r; ++c) {
      double s = 0.0;
      for (int j2 = 0; j2 < l; ++j2) s += F.SLinv[(size_t)j * l + j2] * u[(size_t)j2 * r + c];
      zL[(size_t)j * r + c] = (j == F.pinned_landmark) ? 0.0 : s;
    }
  // 4. correct poses, back-substitute ranges
  for (size_t row = 0; row < np; ++row)
    for (int c = 0; c < r; ++c) {
      double s = Y[row * r + c];
      for (int j = 0; j < l; ++j) s -= F.W[row * l + j] * zL[(size_t)j * r + c];
      Z[row * r + c] = ((int)row == F.pinned_pose_row) ? 0.0 : s;
    }
  for (int j = 0; j < l; ++j)
    for (int c = 0; c < r; ++c) Z[(np + j) * r + c] = zL[(size_t)j * r + c];
  for (int k = 0; k < m; ++k)
    for (int c = 0; c < r; ++c) {
      double s = V[(np + l + k) * r + c];
      for (int p = 0; p < 2; ++p) {
        const int x = F.rend_x[(size_t)k * 2 + p];
        if (x >= 0) s -= F.rend_e[(size_t)k * 2 + p] * Z[trow(x) * r + c];
      }
      Z[(np + l + k) * r + c] = s * F.rdinv[k];
    }
}

// ================================================================ device side ====
struct ChainLevelDev {
  ChunkGeo G;
  DevBuf<double> fwd, bwd, UR;
  DevBuf<double> sol, rhs, cL, cR;  // work vectors (level >= 1 / per level), capacity ws_r columns
};

struct ChainChol {
  int B = 0;
  ChainFactorHost host;  // small parts stay on the host too (geometry)
  std::vector<ChainLevelDev *> levels;
  DevBuf<double> rdinv, rinc_e, rend_e, bl_val, W, SLinv, u, zL, Y;
  DevBuf<int> rinc_ptr, rinc_k, rend_x, bl_ptr, bl_row;
  int ws_cols = 0;
  // pose graph with loop closures / several robots: general sparse block factor (gen_chol_dev.cuh) instead of levels
  bool general = false;
  GenFactorDev gen;
  GenSymDev *gsd = nullptr;  // owned by the handle's ChainSym
  ~ChainChol() {
    for (auto *p : levels) delete p;
  }
};

inline void destroy_chain_chol(ChainChol *c) { delete c; }

template <int B>
__global__ void __launch_bounds__(128) k_chain_forward(const ChunkGeo G, int ld, int ncols, const double *fwd,
                                                       const double *UR, double *sol, double *rhs_cur,
                                                       const double *rhs_prev, int c_prev,
                                                       const double *cL_prev, const double *cR_prev,
                                                       double *cL, double *cR, const double *bwd_top,
                                                       const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = t / ncols, col = t - k * ncols;
  if (k >= G.K) return;
  forward_chunk<B>(G, k, col, ld, fwd, UR, sol, rhs_cur, rhs_prev, c_prev, cL_prev, cR_prev, cL, cR);
  if (bwd_top != nullptr) backward_chunk<B>(G, k, col, ld, bwd_top, sol, nullptr);
}

template <int B>
__global__ void __launch_bounds__(128) k_chain_backward(const ChunkGeo G, int ld, int ncols, const double *bwd,
                                                        double *sol, const double *xsep, const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = t / ncols, col = t - k * ncols;
  if (k >= G.K) return;
  backward_chunk<B>(G, k, col, ld, bwd, sol, xsep);
}

// Step 1 of the apply: Y = V on pose and landmark rows, minus the range elimination on the
// translation rows; pinned rows zeroed.  Blocks [0, l) reduce one landmark each; the rest
// cover the pose-section elements (flat).
static __global__ void __launch_bounds__(kThreads) k_chain_pre(int n, int l, int D1, int r, const double *__restrict__ V,
                                                        double *__restrict__ Y, const int *rinc_ptr,
                                                        const int *rinc_k, const double *rinc_e,
                                                        const double *rdinv, int pinned_pose_row,
                                                        int pinned_landmark, const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  const size_t np = (size_t)n * D1;
  const size_t rg0 = np + l;
  if ((int)blockIdx.x < l) {
    __shared__ double sacc[kThreads];
    const int j = blockIdx.x;
    const int per = blockDim.x / r, e = threadIdx.x / r, c = threadIdx.x - e * r;
    double acc = 0.0;
    if (e < per)
      for (int q = rinc_ptr[n + j] + e; q < rinc_ptr[n + j + 1]; q += per)
        acc += rinc_e[q] * rdinv[rinc_k[q]] * V[(rg0 + rinc_k[q]) * r + c];
    sacc[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < r) {
      double s = 0.0;
      for (int i = 0; i < per; ++i) s += sacc[i * r + threadIdx.x];
      Y[(np + j) * r + threadIdx.x] = (j == pinned_landmark) ? 0.0 : V[(np + j) * r + threadIdx.x] - s;
    }
    return;
  }
  const size_t nE = np * r;
  for (size_t e = (size_t)(blockIdx.x - l) * blockDim.x + threadIdx.x; e < nE;
       e += (size_t)(gridDim.x - l) * blockDim.x) {
    const size_t row = e / r;
    const int c = (int)(e - row * r);
    double v = V[e];
    const int a = (int)(row % D1);
    if (a == D1 - 1) {
      const int x = (int)(row / D1);
      for (int q = rinc_ptr[x]; q < rinc_ptr[x + 1]; ++q)
        v -= rinc_e[q] * rdinv[rinc_k[q]] * V[(rg0 + rinc_k[q]) * r + c];
      if ((int)row == pinned_pose_row) v = 0.0;
    }
    Y[e] = v;
  }
}

// Step 3: u_j = y_L[j] - sum_B B[row, j] y[row]  (one CTA per landmark), then the last CTA
// applies S_L^-1 (l x l) and writes z_L.
static __global__ void __launch_bounds__(kThreads) k_chain_border(int n, int l, int D1, int r, const double *__restrict__ Y,
                                                           const int *bl_ptr, const int *bl_row,
                                                           const double *bl_val, const double *SLinv, double *u,
                                                           double *zL, unsigned *counter, int pinned_landmark,
                                                           const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  __shared__ double sacc[kThreads];
  __shared__ int s_last;
  const size_t np = (size_t)n * D1;
  const int j = blockIdx.x;
  const int per = blockDim.x / r, e = threadIdx.x / r, c = threadIdx.x - e * r;
  double acc = 0.0;
  if (e < per)
    for (int q = bl_ptr[j] + e; q < bl_ptr[j + 1]; q += per) acc += bl_val[q] * Y[(size_t)bl_row[q] * r + c];
  sacc[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < r) {
    double s = 0.0;
    for (int i = 0; i < per; ++i) s += sacc[i * r + threadIdx.x];
    u[(size_t)j * r + threadIdx.x] = Y[(np + j) * r + threadIdx.x] - s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicInc(counter, gridDim.x - 1);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < l * r; i += blockDim.x) {
    const int jj = i / r, cc = i - jj * r;
    double s = 0.0;
    for (int j2 = 0; j2 < l; ++j2) s += SLinv[(size_t)jj * l + j2] * __ldcg(u + (size_t)j2 * r + cc);
    zL[i] = (jj == pinned_landmark) ? 0.0 : s;
  }
}

// Step 4: z_P = y_P - W z_L ; z_L ; ranges back-substituted.
static __global__ void __launch_bounds__(kThreads) k_chain_post(int n, int l, int m, int D1, int r,
                                                         const double *__restrict__ V, const double *__restrict__ Y,
                                                         const double *__restrict__ W, const double *__restrict__ zL,
                                                         const int *rend_x, const double *rend_e,
                                                         const double *rdinv, int pinned_pose_row, double *Z,
                                                         const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  const size_t np = (size_t)n * D1, rg0 = np + l;
  const size_t nE = (rg0 + m) * r;
  auto zpose = [&](size_t row, int c) -> double {
    if ((int)row == pinned_pose_row) return 0.0;
    double s = Y[row * r + c];
    for (int j = 0; j < l; ++j) s = fma(-W[row * l + j], zL[(size_t)j * r + c], s);
    return s;
  };
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / r;
    const int c = (int)(e - row * r);
    double out;
    if (row < np) {
      out = zpose(row, c);
    } else if (row < rg0) {
      out = zL[(row - np) * r + c];
    } else {
      const size_t k = row - rg0;
      double s = V[e];
      for (int p = 0; p < 2; ++p) {
        const int x = rend_x[k * 2 + p];
        if (x < 0) continue;
        const double zx = x < n ? zpose((size_t)x * D1 + (D1 - 1), c) : zL[(size_t)(x - n) * r + c];
        s = fma(-rend_e[k * 2 + p], zx, s);
      }
      out = s * rdinv[k];
    }
    Z[e] = out;
  }
}

template <typename T>
inline void upload_vec(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t s) { b.upload(v, s); }

// build_chain_chol: chain_factor_dev.cuh (the factorisation runs on the device)
// trans_only: factor the translation Laplacian Q33 alone (rotation and range rows decoupled: identity / dropped),
// the LtransCholRed_ of the implicit formulation (src/CORA_problem.cpp:714-741) when pin_last is set
ChainChol *build_chain_chol(H *h, const double *d_bval, const double *d_sdiag, double shift, bool pin_last,
                            bool *pos_def, bool want_solve = true, bool trans_only = false);

inline void chain_ensure_ws(H *h, ChainChol *C, int r) {
  if (r <= C->ws_cols) return;
  const int cap = std::max(r, h->ws_r);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (size_t lv = 0; lv < C->levels.size(); ++lv) {
    ChainLevelDev *D = C->levels[lv];
    const size_t nb = (size_t)std::max(D->G.n, 1) * C->B * cap;
    if (lv > 0) { D->sol.alloc(nb); D->rhs.alloc(nb); }
    D->cL.alloc((size_t)(D->G.K + 1) * C->B * cap);
    D->cR.alloc((size_t)(D->G.K + 1) * C->B * cap);
    CUDA_CHECK(cudaMemset(D->cL.p, 0, D->cL.n * sizeof(double)));
    CUDA_CHECK(cudaMemset(D->cR.p, 0, D->cR.n * sizeof(double)));
  }
  C->u.alloc((size_t)std::max(C->host.l, 1) * cap);
  C->zL.alloc((size_t)std::max(C->host.l, 1) * cap);
  C->Y.alloc((size_t)h->DL.N * cap);
  C->ws_cols = cap;
}

// Z = M^-1 V on the device (V, Z: N x r internal row-major, V != Z).
inline void chain_solve(H *h, ChainChol *C, const double *V, double *Z, int r, const CgCtrl *ctrl) {
  if (r > kThreads) throw Error(CORA_B200_EINVAL, "too many columns for the chain solve");
  chain_ensure_ws(h, C, r);
  const ChainFactorHost &F = C->host;
  const int n = F.n, l = F.l, m = F.m, D1 = h->DL.D1;
  cudaStream_t s = h->stream;
  double *Y = C->Y.p;
  {
    const size_t nE = (size_t)n * D1 * r;
    const int gb = (int)std::min<size_t>((nE + kThreads - 1) / kThreads, (size_t)h->sm_count * 8);
    k_chain_pre<<<l + std::max(gb, 1), kThreads, 0, s>>>(n, l, D1, r, V, Y, C->rinc_ptr.p, C->rinc_k.p, C->rinc_e.p,
                                                       C->rdinv.p, F.pinned_pose_row, F.pinned_landmark, ctrl);
    check_launch(h);
  }
  const int nl = (int)C->levels.size();
  auto solp = [&](int lv) { return lv == 0 ? Y : C->levels[lv]->sol.p; };
  auto rhsp = [&](int lv) { return lv == 0 ? Y : C->levels[lv]->rhs.p; };
  if (n > 0 && C->general) {
    if (C->B == 3) gen_solve_device<3>(h, *C->gsd, C->gen, n, Y, r, r, ctrl);
    else gen_solve_device<4>(h, *C->gsd, C->gen, n, Y, r, r, ctrl);
  } else if (n > 0) {
    for (int lv = 0; lv < nl; ++lv) {
      ChainLevelDev *D = C->levels[lv];
      const int threads = D->G.K * r;
      const int grid = (threads + 127) / 128;
      const bool top = (lv == nl - 1);
      if (C->B == 3)
        k_chain_forward<3><<<grid, 128, 0, s>>>(D->G, r, r, D->fwd.p, D->UR.p, solp(lv), rhsp(lv),
                                                lv > 0 ? rhsp(lv - 1) : nullptr, lv > 0 ? C->levels[lv - 1]->G.c : 0,
                                                lv > 0 ? C->levels[lv - 1]->cL.p : nullptr,
                                                lv > 0 ? C->levels[lv - 1]->cR.p : nullptr, D->cL.p, D->cR.p,
                                                top ? D->bwd.p : nullptr, ctrl);
      else
        k_chain_forward<4><<<grid, 128, 0, s>>>(D->G, r, r, D->fwd.p, D->UR.p, solp(lv), rhsp(lv),
                                                lv > 0 ? rhsp(lv - 1) : nullptr, lv > 0 ? C->levels[lv - 1]->G.c : 0,
                                                lv > 0 ? C->levels[lv - 1]->cL.p : nullptr,
                                                lv > 0 ? C->levels[lv - 1]->cR.p : nullptr, D->cL.p, D->cR.p,
                                                top ? D->bwd.p : nullptr, ctrl);
      check_launch(h);
    }
    for (int lv = nl - 2; lv >= 0; --lv) {
      ChainLevelDev *D = C->levels[lv];
      const int threads = D->G.K * r;
      const int grid = (threads + 127) / 128;
      if (C->B == 3) k_chain_backward<3><<<grid, 128, 0, s>>>(D->G, r, r, D->bwd.p, solp(lv), solp(lv + 1), ctrl);
      else k_chain_backward<4><<<grid, 128, 0, s>>>(D->G, r, r, D->bwd.p, solp(lv), solp(lv + 1), ctrl);
      check_launch(h);
    }
  }
  if (l > 0) {
    k_chain_border<<<l, kThreads, 0, s>>>(n, l, D1, r, Y, C->bl_ptr.p, C->bl_row.p, C->bl_val.p, C->SLinv.p, C->u.p,
                                          C->zL.p, h->d_counter.p + 1, F.pinned_landmark, ctrl);
    check_launch(h);
  }
  {
    const size_t nE = (size_t)h->DL.N * r;
    const int gb = (int)std::min<size_t>((nE + kThreads - 1) / kThreads, (size_t)h->sm_count * 8);
    k_chain_post<<<std::max(gb, 1), kThreads, 0, s>>>(n, l, m, D1, r, V, Y, C->W.p, C->zL.p, C->rend_x.p, C->rend_e.p,
                                                     C->rdinv.p, F.pinned_pose_row, Z, ctrl);
    check_launch(h);
  }
}

// ||Q||_2 by Lanczos on the device product: lanczos.cuh
double estimate_spectral_norm(H *h);

}  // namespace cora_b200
