// kernels.cuh -- sm_100a kernels of the CORA staircase inner loop.
//
// All dense iterates are N x r ROW-major in the internal pose-major row order
// (layout.hpp), leading dimension exactly r: element (row, c) is at row*r + c, so
// every vector pass is a flat, fully coalesced stream.  One CTA processes one TILE
// of TR consecutive rows (TP = TR/(d+1) poses):
//   phase 0  stage the tile's block-ELL slice and the tile rows of Y/G/Ydot in
//            shared memory (coalesced),
//   phase 1  Q*X for the tile: one thread per (pose, column) accumulates the d+1
//            rows of the pose over its block slots (x gathered through L1/L2 with
//            ld.global.nc, block values broadcast from shared memory), plus the CSR
//            spill; scalar rows (landmarks, ranges) one thread per (row, column),
//   phase 2  hub rows (landmarks) are added from the buffer k_long_groups filled,
//   phase 3  Riemannian epilogue on the staged tile, one thread per pose / range row,
//   phase 4  coalesced store + the inner products of the CG recurrences, block
//            reduced, one partial per tile; the LAST CTA to finish sums the partials
//            in a fixed order (deterministic) and runs the scalar logic of STPCG on
//            the device, so no host round trip happens inside a CG solve.
//
// Reference semantics: src/CORA_problem.cpp:742-938 (operators),
// src/StiefelProduct.cpp:8-55, src/ObliqueManifold.cpp:6-27 (geometry),
// libs/Optimization/.../IterativeSolvers.h:207-426 (STPCG).
#pragma once
#include "internal.cuh"

namespace cora_b200 {

constexpr int kThreads = 256;
constexpr int kNPart = 4;  // partial sums per tile

// ------------------------------------------------------------ small helpers ----
template <int D>
struct Geo {
  static constexpr int D1 = D + 1;
  static constexpr int PADP = (D1 % 2 == 0) ? 1 : 0;
  int r, RS;
  __device__ __forceinline__ Geo(int r_) : r(r_), RS(r_ | 1) {}
  // shared-memory offset of (local row, column): odd row stride and odd pose
  // stride keep the one-thread-per-pose epilogue free of bank conflicts
  __device__ __forceinline__ int soff(int lrow, int c) const {
    return lrow * RS + (PADP ? lrow / D1 : 0) + c;
  }
  __device__ __forceinline__ int pose_base(int p) const { return p * (D1 * RS + PADP); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum K values over the block in a fixed order; result valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *sred) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) sred[w * K + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (int i = 0; i < nw; ++i) s += sred[i * K + k];
      v[k] = s;
    }
  }
}

// ---------------------------------------------------- STPCG scalar logic -------
// IterativeSolvers.h:294-362: after Hp = H(p).
__device__ inline void cg_post_hess(CgCtrl *c, double pHp, double HpHp, double pp) {
  c->kappa = pHp;
  c->pp = pp;
  c->HpHp = HpHp;
  if (sqrt(HpHp) / sqrt(pp) < c->eps) {  // :305-338 -- finished on the host (rare)
    c->exit_reason = CG_EXIT_KERNEL;
    c->hM = c->Delta;
    __threadfence();
    c->state = 1;
    return;
  }
  const double alpha = c->rv / pHp;  // :341
  const double sM2n = c->sM2 + 2.0 * alpha * c->sMp + alpha * alpha * c->pM2;
  if (pHp <= 0.0 || sM2n > c->Delta2) {  // :347-362
    c->sigma = (-c->sMp + sqrt(c->sMp * c->sMp + c->pM2 * (c->Delta2 - c->sM2))) / c->pM2;
    c->mode = CG_MODE_BOUNDARY;
    c->hM = c->Delta;
    c->exit_reason = CG_EXIT_BOUNDARY;
  } else {
    c->mode = CG_MODE_STEP;
    c->alpha = alpha;
    c->sM2_next = sM2n;
  }
}

// IterativeSolvers.h:408-417 + loop head :285-291.
__device__ inline void cg_post_update(CgCtrl *c, double rv_new) {
  const double beta = rv_new / (c->alpha * c->kappa);
  c->beta = beta;
  c->sM2 = c->sM2_next;
  c->sMp = beta * (c->sMp + c->alpha * c->pM2);
  c->pM2 = rv_new + beta * beta * c->pM2;
  c->rv = rv_new;
  c->it += 1;
  int done = 0;
  if (c->it >= c->max_it) {
    c->exit_reason = CG_EXIT_MAXIT;
    done = 1;
  } else if (sqrt(rv_new) <= c->target) {
    c->exit_reason = CG_EXIT_TARGET;
    done = 1;
  }
  if (done) {
    c->hM = sqrt(c->sM2);
    __threadfence();
    c->state = 1;
  }
}

// Last-CTA reduction of the per-tile partials + post operation.
template <int K>
__device__ __forceinline__ void finish_reduction(double (&v)[K], double *sred, double *partials,
                                                 unsigned *counter, double *scal, CgCtrl *ctrl,
                                                 int post, int slot) {
  __shared__ int s_last;
  block_sum<K>(v, sred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * kNPart + k] = v[k];
    __threadfence();
    const unsigned prev = atomicInc(counter, gridDim.x - 1);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < K; ++k) t[k] += __ldcg(partials + (size_t)i * kNPart + k);
  block_sum<K>(t, sred);
  if (threadIdx.x == 0) {
    if (post == POST_STORE) {
#pragma unroll
      for (int k = 0; k < K; ++k) scal[slot + k] = t[k];
    } else if (post == POST_CG_HESS) {
      cg_post_hess(ctrl, t[0], K > 1 ? t[K > 1 ? 1 : 0] : 0.0, K > 2 ? t[K > 2 ? 2 : 0] : 0.0);
    } else if (post == POST_CG_UPDATE) {
      cg_post_update(ctrl, t[0]);
    }
  }
}

// ------------------------------------------------- per-pose / per-row geometry --
// One thread owns one pose block (d x r, rows RS apart in shared memory).
// w -= sym(y w^T) y        (StiefelProduct.h:79-81 + StiefelProduct.cpp:38-55, row form)
template <int D>
__device__ __forceinline__ void stiefel_tangent(const double *y, double *w, int r, int RS) {
  double P[D][D];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) P[a][b] = 0.0;
  for (int c = 0; c < r; ++c) {
    double yc[D], wc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) { yc[a] = y[a * RS + c]; wc[a] = w[a * RS + c]; }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) P[a][b] = fma(yc[a], wc[b], P[a][b]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a + 1; b < D; ++b) { const double s = 0.5 * (P[a][b] + P[b][a]); P[a][b] = s; P[b][a] = s; }
  for (int c = 0; c < r; ++c) {
    double yc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) yc[a] = y[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double s = w[a * RS + c];
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(-P[a][b], yc[b], s);
      w[a * RS + c] = s;
    }
  }
}

// w -= sym(y g^T) dd       (src/CORA_problem.cpp:839-850)
template <int D>
__device__ __forceinline__ void stiefel_curvature(const double *y, const double *g, const double *dd,
                                                  double *w, int r, int RS) {
  double P[D][D];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) P[a][b] = 0.0;
  for (int c = 0; c < r; ++c) {
    double yc[D], gc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) { yc[a] = y[a * RS + c]; gc[a] = g[a * RS + c]; }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) P[a][b] = fma(yc[a], gc[b], P[a][b]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a + 1; b < D; ++b) { const double s = 0.5 * (P[a][b] + P[b][a]); P[a][b] = s; P[b][a] = s; }
  for (int c = 0; c < r; ++c) {
    double dc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) dc[a] = dd[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double s = w[a * RS + c];
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(-P[a][b], dc[b], s);
      w[a * RS + c] = s;
    }
  }
}

// ObliqueManifold.cpp:16-27: w -= (y.w) y
__device__ __forceinline__ void oblique_tangent(const double *y, double *w, int r) {
  double s = 0.0;
  for (int c = 0; c < r; ++c) s = fma(y[c], w[c], s);
  for (int c = 0; c < r; ++c) w[c] = fma(-s, y[c], w[c]);
}
// src/CORA_problem.cpp:853-864: w -= (g.y) dd
__device__ __forceinline__ void oblique_curvature(const double *y, const double *g, const double *dd,
                                                  double *w, int r) {
  double s = 0.0;
  for (int c = 0; c < r; ++c) s = fma(g[c], y[c], s);
  for (int c = 0; c < r; ++c) w[c] = fma(-s, dd[c], w[c]);
}

// Symmetric eigen-decomposition of a D x D matrix by cyclic Jacobi; on return A is
// (numerically) diagonal and V holds the eigenvectors in its columns.
template <int D>
__device__ __forceinline__ void jacobi_eig(double (&A)[D][D], double (&V)[D][D]) {
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, dg = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      dg += fabs(A[a][a]);
#pragma unroll
      for (int b = a + 1; b < D; ++b) off += fabs(A[a][b]);
    }
    if (off <= 1e-18 * dg) break;
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
      for (int q = p + 1; q < D; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double tt = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
#pragma unroll
        for (int k = 0; k < D; ++k) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {  // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
}

// Accurate path for ill-conditioned blocks: one-sided (Hestenes) Jacobi on the d rows of w.
// Rotations from the left make the rows mutually orthogonal, U^T w = Sigma V^T; the polar factor
// is U V^T = U * (rows of U^T w normalised).  Relative accuracy eps*cond(w) instead of the
// eps*cond(w)^2 of the Gram-matrix route.
template <int D>
__device__ __noinline__ void stiefel_polar_hestenes(double *w, int r, int RS) {
  double U[D][D];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) U[a][b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
      for (int q = p + 1; q < D; ++q) {
        double app = 0.0, aqq = 0.0, apq = 0.0;
        for (int c = 0; c < r; ++c) {
          const double x = w[p * RS + c], y = w[q * RS + c];
          app = fma(x, x, app); aqq = fma(y, y, aqq); apq = fma(x, y, apq);
        }
        if (fabs(apq) <= 1e-16 * sqrt(app * aqq) || apq == 0.0) continue;
        rotated = true;
        const double th = (aqq - app) / (2.0 * apq);
        const double tt = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
        for (int c = 0; c < r; ++c) {
          const double x = w[p * RS + c], y = w[q * RS + c];
          w[p * RS + c] = cs * x - sn * y;
          w[q * RS + c] = sn * x + cs * y;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {  // U <- U J  (so that w_orig = U * w)
          const double ukp = U[k][p], ukq = U[k][q];
          U[k][p] = cs * ukp - sn * ukq;
          U[k][q] = sn * ukp + cs * ukq;
        }
      }
    if (!rotated) break;
  }
  double inv[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = 0.0;
    for (int c = 0; c < r; ++c) s = fma(w[a * RS + c], w[a * RS + c], s);
    inv[a] = 1.0 / sqrt(fmax(s, 1e-300));
  }
  for (int c = 0; c < r; ++c) {
    double v[D];
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = w[a * RS + c] * inv[a];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(U[a][k], v[k], s);
      w[a * RS + c] = s;
    }
  }
}

// Polar factor of the d x r block w (rows RS apart), in place:
// w <- (w w^T)^{-1/2} w, the row form of StiefelProduct.cpp:26-34 (thin SVD -> U V^T).
template <int D>
__device__ __forceinline__ void stiefel_polar(double *w, int r, int RS) {
  double Gm[D][D], A[D][D], V[D][D], M[D][D];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) Gm[a][b] = 0.0;
  for (int c = 0; c < r; ++c) {
    double wc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) wc[a] = w[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) Gm[a][b] = fma(wc[a], wc[b], Gm[a][b]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) A[a][b] = Gm[a][b];
  jacobi_eig<D>(A, V);
  {
    double lmin = A[0][0], lmax = A[0][0];
#pragma unroll
    for (int a = 1; a < D; ++a) { lmin = fmin(lmin, A[a][a]); lmax = fmax(lmax, A[a][a]); }
    if (lmin < 1e-6 * lmax) {  // ill-conditioned block (random initial guesses only)
      stiefel_polar_hestenes<D>(w, r, RS);
      return;
    }
  }
  double is[D];
#pragma unroll
  for (int a = 0; a < D; ++a) is[a] = 1.0 / sqrt(fmax(A[a][a], 1e-300));
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(V[a][k] * is[k], V[b][k], s);
      M[a][b] = s;
    }
  // two Newton-Schulz polishing steps on M ~ G^{-1/2}:  M <- M (1.5 I - 0.5 G M M)
#pragma unroll
  for (int step = 0; step < 2; ++step) {
    double T[D][D], U[D][D], Mn[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(M[a][k], M[k][b], s);
        T[a][b] = s;
      }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(Gm[a][k], T[k][b], s);
        U[a][b] = ((a == b) ? 1.5 : 0.0) - 0.5 * s;
      }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(M[a][k], U[k][b], s);
        Mn[a][b] = s;
      }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) M[a][b] = 0.5 * (Mn[a][b] + Mn[b][a]);
  }
  for (int c = 0; c < r; ++c) {
    double wc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) wc[a] = w[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(M[a][b], wc[b], s);
      w[a * RS + c] = s;
    }
  }
  // final Newton-Schulz step on the block itself, w <- (1.5 I - 0.5 w w^T) w: the Gram-matrix
  // route above is accurate to eps*cond(G); this restores orthonormality to rounding level
  // for ill-conditioned blocks (random initial guesses)
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) Gm[a][b] = 0.0;
  for (int c = 0; c < r; ++c) {
    double wc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) wc[a] = w[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) Gm[a][b] = fma(wc[a], wc[b], Gm[a][b]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) M[a][b] = ((a == b) ? 1.5 : 0.0) - 0.25 * (Gm[a][b] + Gm[b][a]);
  for (int c = 0; c < r; ++c) {
    double wc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) wc[a] = w[a * RS + c];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(M[a][b], wc[b], s);
      w[a * RS + c] = s;
    }
  }
}

// ------------------------------------------------------------ tile geometry ----
struct TileInfo {
  int row0, nR, nP, nS, S;
  long long ebase;
};
template <int D>
__device__ __forceinline__ TileInfo tile_info(const DevLayout &L, int t, int r) {
  TileInfo T;
  T.row0 = t * L.TR;
  T.nR = min(L.TR, L.N - T.row0);
  T.nP = max(0, min(L.TP, L.n - t * L.TP));
  T.nS = T.nR - T.nP * (D + 1);
  T.S = L.tile_slots[t];
  T.ebase = (long long)T.row0 * r;
  return T;
}

// Riemannian epilogue over a staged tile: tangent projection (and optionally the
// curvature correction first) for every pose block and range row of the tile.
template <int D, bool CURV>
__device__ __forceinline__ void tile_epilogue(const DevLayout &L, const TileInfo &T, const Geo<D> &geo,
                                              const double *sY, const double *sG, const double *sD,
                                              double *sW) {
  const int r = geo.r, RS = geo.RS;
  for (int u = threadIdx.x; u < T.nP + T.nS; u += blockDim.x) {
    if (u < T.nP) {
      const int o = geo.pose_base(u);
      if (CURV) stiefel_curvature<D>(sY + o, sG + o, sD + o, sW + o, r, RS);
      stiefel_tangent<D>(sY + o, sW + o, r, RS);
    } else {
      const int lrow = T.nP * (D + 1) + (u - T.nP);
      if (T.row0 + lrow >= L.nPoseRows + L.l) {  // range row (landmark rows: Euclidean)
        const int o = geo.soff(lrow, 0);
        if (CURV) oblique_curvature(sY + o, sG + o, sD + o, sW + o, r);
        oblique_tangent(sY + o, sW + o, r);
      }
    }
  }
}

// ------------------------------------------------ shared-memory carve-up -------
struct QSmem {
  double *sval, *sW, *sY, *sG, *sD, *sred;
  int *scol;
};
template <int D>
__host__ __device__ inline size_t qsmem_bytes(int maxSlots, int TP, int TR, int r, int nvec) {
  const int D1 = D + 1;
  const size_t nbv = (size_t)maxSlots * D1 * D1 * TP;
  const size_t ncol = ((size_t)maxSlots * TP + 1) & ~(size_t)1;
  const size_t vstride = (size_t)TR * (r | 1) + TP;
  return (nbv + (size_t)nvec * vstride + 64) * sizeof(double) + ncol * sizeof(int);
}

// ===================================================================== k_qprod ==
// out = Q X with an optional Riemannian epilogue (QMode).
template <int D>
__global__ void __launch_bounds__(kThreads) k_qprod(const DevLayout L, const QArgs A) {
  constexpr int D1 = D + 1;
  if (A.ctrl != nullptr && *((volatile int *)&A.ctrl->state) != 0) return;
  extern __shared__ double smem[];
  const int t = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int r = A.r;
  const Geo<D> geo(r);
  const TileInfo T = tile_info<D>(L, t, r);
  const int TP = L.TP;
  const size_t nbv = (size_t)L.maxSlots * D1 * D1 * TP;
  const size_t ncol = ((size_t)L.maxSlots * TP + 1) & ~(size_t)1;
  const size_t vstride = (size_t)L.TR * geo.RS + TP;
  double *sred = smem;  // 64 doubles
  double *sval = smem + 64;
  int *scol = (int *)(sval + nbv);
  double *sW = (double *)(scol + ncol);
  double *sY = sW + vstride;
  double *sG = sY + vstride;
  double *sD = sG + vstride;
  const double *__restrict__ X = A.X;

  // ---- phase 0: stage block-ELL slice and the tile rows of the dense operands ----
  {
    const double *gb = L.bval + L.tile_boff[t];
    const int nb = T.S * D1 * D1 * TP;
    for (int i = tid; i < nb; i += nth) sval[i] = __ldg(gb + i);
    const int *gc = L.bcol + L.tile_coff[t];
    const int nc = T.S * TP;
    for (int i = tid; i < nc; i += nth) scol[i] = __ldg(gc + i);
  }
  const int nE = T.nR * r;
  if (A.mode != QM_SPMM) {
    for (int le = tid; le < nE; le += nth) {
      const int lrow = le / r, c = le - lrow * r;
      const int so = geo.soff(lrow, c);
      sY[so] = A.Y[T.ebase + le];
      if (A.mode == QM_HESS) {
        sG[so] = A.G[T.ebase + le];
        sD[so] = X[T.ebase + le];
      }
    }
  }
  __syncthreads();

  // ---- phase 1: Q X ----
  const int nPoseItems = T.nP * r;
  const int nItems = nPoseItems + T.nS * r;
  for (int it = tid; it < nItems; it += nth) {
    if (it < nPoseItems) {
      const int p = it / r, c = it - p * r;
      double acc[D1];
#pragma unroll
      for (int a = 0; a < D1; ++a) acc[a] = 0.0;
      for (int s = 0; s < T.S; ++s) {
        const int jb = scol[s * TP + p];
        const double *xp = X + (size_t)jb * r + c;
        double x[D1];
#pragma unroll
        for (int b = 0; b < D1; ++b) x[b] = __ldg(xp + b * r);
        const double *bv = sval + (size_t)s * D1 * D1 * TP + p;
#pragma unroll
        for (int a = 0; a < D1; ++a)
#pragma unroll
          for (int b = 0; b < D1; ++b) acc[a] = fma(bv[(a * D1 + b) * TP], x[b], acc[a]);
      }
      const int g = t * TP + p;
      const int k0 = __ldg(L.grp_ptr + g), k1 = __ldg(L.grp_ptr + g + 1);
      for (int k = k0; k < k1; ++k) {
        const unsigned pk = __ldg(L.rem_pk + k);
        const double v = __ldg(L.rem_val + k);
        const int lr = (int)(pk >> 30);
        const double xv = v * __ldg(X + (size_t)(pk & kColMask) * r + c);
#pragma unroll
        for (int a = 0; a < D1; ++a) acc[a] += (lr == a) ? xv : 0.0;
      }
      const int o = geo.pose_base(p) + c;
#pragma unroll
      for (int a = 0; a < D1; ++a) sW[o + a * geo.RS] = acc[a];
    } else {
      const int j = it - nPoseItems;
      const int sr = j / r, c = j - sr * r;
      const int lrow = T.nP * D1 + sr;
      const int grow = T.row0 + lrow;
      const int sidx = grow - L.nPoseRows;
      double acc = __ldg(L.sdiag + sidx) * __ldg(X + (size_t)grow * r + c);
      const int g = L.n + sidx;
      const int k0 = __ldg(L.grp_ptr + g), k1 = __ldg(L.grp_ptr + g + 1);
      for (int k = k0; k < k1; ++k) {
        const unsigned pk = __ldg(L.rem_pk + k);
        acc = fma(__ldg(L.rem_val + k), __ldg(X + (size_t)(pk & kColMask) * r + c), acc);
      }
      sW[geo.soff(lrow, c)] = acc;
    }
  }
  __syncthreads();

  // ---- phase 2: hub groups reduced by k_long_groups ----
  {
    const int q0 = L.tile_long_ptr[t], q1 = L.tile_long_ptr[t + 1];
    for (int q = q0; q < q1; ++q) {
      const int g = L.long_grp[q];
      const int lrow0 = (g < L.n ? g * D1 : L.nPoseRows + (g - L.n)) - T.row0;
      const int nrow = g < L.n ? D1 : 1;
      for (int i = tid; i < nrow * r; i += nth) {
        const int a = i / r, c = i - a * r;
        sW[geo.soff(lrow0 + a, c)] += A.longbuf[((size_t)q * D1 + a) * r + c];
      }
    }
    if (q1 > q0) __syncthreads();
  }

  if (A.mode == QM_SPMM) {
    for (int le = tid; le < nE; le += nth) {
      const int lrow = le / r, c = le - lrow * r;
      A.out[T.ebase + le] = sW[geo.soff(lrow, c)];
    }
    return;
  }

  double acc[3] = {0.0, 0.0, 0.0};
  if (A.mode == QM_GRAD) {
    // Euclidean gradient out, <Y, QY>; then grad = proj_Y(QY)
    for (int le = tid; le < nE; le += nth) {
      const int lrow = le / r, c = le - lrow * r;
      const int so = geo.soff(lrow, c);
      const double w = sW[so];
      A.out2[T.ebase + le] = w;
      acc[0] = fma(sY[so], w, acc[0]);
    }
    __syncthreads();
    tile_epilogue<D, false>(L, T, geo, sY, nullptr, nullptr, sW);
    __syncthreads();
    for (int le = tid; le < nE; le += nth) {
      const int lrow = le / r, c = le - lrow * r;
      const double w = sW[geo.soff(lrow, c)];
      A.out[T.ebase + le] = w;
      acc[1] = fma(w, w, acc[1]);
    }
  } else {  // QM_HESS
    tile_epilogue<D, true>(L, T, geo, sY, sG, sD, sW);
    __syncthreads();
    for (int le = tid; le < nE; le += nth) {
      const int lrow = le / r, c = le - lrow * r;
      const int so = geo.soff(lrow, c);
      const double w = sW[so], dd = sD[so];
      A.out[T.ebase + le] = w;
      acc[0] = fma(dd, w, acc[0]);
      acc[1] = fma(w, w, acc[1]);
      acc[2] = fma(dd, dd, acc[2]);
    }
  }
  finish_reduction<3>(acc, sred, A.partials, A.counter, A.scal, A.ctrl, A.post, A.slot);
}

// ================================================================ k_long_groups ==
// One CTA per hub group (landmark rows with thousands of entries): longbuf[q][a][c].
template <int D>
__global__ void __launch_bounds__(kThreads) k_long_groups(const DevLayout L, const double *__restrict__ X,
                                                          double *longbuf, int r, const CgCtrl *ctrl) {
  constexpr int D1 = D + 1;
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  extern __shared__ double smem[];  // D1 * nth
  const int q = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int per = nth / r;  // entries processed concurrently
  const int e = tid / r, c = tid - e * r;
  double acc[D1];
#pragma unroll
  for (int a = 0; a < D1; ++a) acc[a] = 0.0;
  if (e < per) {
    const int k0 = L.long_ptr[q], k1 = L.long_ptr[q + 1];
    for (int k = k0 + e; k < k1; k += per) {
      const unsigned pk = __ldg(L.long_pk + k);
      const int lr = (int)(pk >> 30);
      const double xv = __ldg(L.long_val + k) * __ldg(X + (size_t)(pk & kColMask) * r + c);
#pragma unroll
      for (int a = 0; a < D1; ++a) acc[a] += (lr == a) ? xv : 0.0;
    }
  }
#pragma unroll
  for (int a = 0; a < D1; ++a) smem[a * nth + tid] = acc[a];
  __syncthreads();
  if (tid < D1 * r) {
    const int a = tid / r, cc = tid - a * r;
    double s = 0.0;
    for (int i = 0; i < per; ++i) s += smem[a * nth + i * r + cc];
    longbuf[((size_t)q * D1 + a) * r + cc] = s;
  }
}

// ================================================================= k_cg_update ==
// STPCG update (IterativeSolvers.h:374-386) fused with the preconditioner closure
// of src/CORA.cpp:89-92:  s += alpha p ; r += alpha Hp ; v = proj_Y(M^-1 r) ; <r,v>.
template <int D>
__global__ void __launch_bounds__(kThreads) k_cg_update(const DevLayout L, const UArgs A) {
  CgCtrl *ctrl = A.ctrl;
  if (A.gated && *((volatile int *)&ctrl->state) != 0) return;
  extern __shared__ double smem[];
  const int t = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int r = A.r;
  const Geo<D> geo(r);
  const TileInfo T = tile_info<D>(L, t, r);
  const int nE = T.nR * r;
  double alpha = 0.0;
  if (A.do_axpy) {
    if (ctrl->mode == CG_MODE_BOUNDARY) {  // :355-361  s += sigma p, done
      const double sigma = ctrl->sigma;
      for (int le = tid; le < nE; le += nth) A.S[T.ebase + le] = fma(sigma, A.P[T.ebase + le], A.S[T.ebase + le]);
      __shared__ int s_last;
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicInc(A.counter, gridDim.x - 1);
        s_last = (prev == gridDim.x - 1);
        if (s_last) {
          __threadfence();
          ctrl->state = 1;
        }
      }
      return;
    }
    alpha = ctrl->alpha;
  }
  const size_t vstride = (size_t)L.TR * geo.RS + L.TP;
  double *sred = smem;
  double *sZ = smem + 64;
  double *sY = sZ + vstride;
  double *sR = sY + vstride;
  for (int le = tid; le < nE; le += nth) {
    const long long e = T.ebase + le;
    double rr = A.R[e];
    if (A.do_axpy) {
      A.S[e] = fma(alpha, A.P[e], A.S[e]);
      rr = fma(alpha, A.HP[e], rr);
      A.R[e] = rr;
    }
    if (A.do_proj) {
      const int lrow = le / r, c = le - lrow * r;
      const int so = geo.soff(lrow, c);
      double z;
      if (A.zsrc == 0) z = rr * __ldg(L.dinv + T.row0 + lrow);
      else if (A.zsrc == 1) z = rr;
      else z = A.Z[e];
      sZ[so] = z;
      sR[so] = rr;
      sY[so] = A.Y[e];
    }
  }
  if (!A.do_proj) return;
  __syncthreads();
  tile_epilogue<D, false>(L, T, geo, sY, nullptr, nullptr, sZ);
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  for (int le = tid; le < nE; le += nth) {
    const int lrow = le / r, c = le - lrow * r;
    const int so = geo.soff(lrow, c);
    const double v = sZ[so];
    A.V[T.ebase + le] = v;
    acc[0] = fma(sR[so], v, acc[0]);
    acc[1] = fma(v, v, acc[1]);
  }
  finish_reduction<2>(acc, sred, A.partials, A.counter, A.scal, ctrl, A.post, A.slot);
}

// p = -v + beta p  (IterativeSolvers.h:420)
static __global__ void __launch_bounds__(kThreads) k_cg_pupdate(const CgCtrl *ctrl, const double *__restrict__ V,
                                                         double *__restrict__ P, long long nE) {
  if (*((volatile const int *)&ctrl->state) != 0) return;
  const double beta = ctrl->beta;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x)
    P[e] = fma(beta, P[e], -V[e]);
}

// STPCG initialisation (IterativeSolvers.h:207-279): s = 0, r = g, p = -v and the
// scalar state; v = P(g) and <g, v> were produced by the caller (TNT.h:383-392
// computes the same preconditioned gradient for its stopping test).
static __global__ void __launch_bounds__(kThreads) k_cg_init(CgCtrl *ctrl, const double *scal, int rv_slot,
                                                      double Delta, int max_it, double kappa_fgr,
                                                      double theta, double eps,
                                                      const double *__restrict__ Gr,
                                                      const double *__restrict__ V, double *__restrict__ S,
                                                      double *__restrict__ R, double *__restrict__ P,
                                                      long long nE) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    S[e] = 0.0;
    R[e] = Gr[e];
    P[e] = -V[e];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double rv = scal[rv_slot];
    ctrl->mode = CG_MODE_STEP;
    ctrl->it = 0;
    ctrl->max_it = max_it;
    ctrl->exit_reason = CG_EXIT_NONE;
    ctrl->rv = rv;
    ctrl->Delta = Delta;
    ctrl->Delta2 = Delta * Delta;
    ctrl->sMp = 0.0;
    ctrl->sM2 = 0.0;
    ctrl->pM2 = rv;
    ctrl->sM2_next = 0.0;
    ctrl->alpha = ctrl->beta = ctrl->kappa = ctrl->sigma = 0.0;
    ctrl->hM = 0.0;
    ctrl->eps = eps;
    ctrl->kappa_fgr = kappa_fgr;
    ctrl->theta = theta;
    const double r0 = sqrt(rv);
    ctrl->target = r0 * fmin(kappa_fgr, pow(r0, theta));  // :278-279
    int done = 0;
    if (max_it <= 0) { ctrl->exit_reason = CG_EXIT_MAXIT; done = 1; }
    else if (r0 <= ctrl->target) { ctrl->exit_reason = CG_EXIT_TARGET; done = 1; }
    ctrl->state = done;
  }
}

// =================================================================== k_retract ==
// out = projectToManifold(Y + alpha V)  (src/CORA_problem.cpp:905-938); with V == nullptr it
// is projectToManifold(Y).  Also <V,V> and <Gr,V> (TNT.h:503,511) when Gr != nullptr.
template <int D>
__global__ void __launch_bounds__(kThreads) k_retract(const DevLayout L, const double *__restrict__ Y,
                                                      const double *__restrict__ V, double alpha,
                                                      const double *__restrict__ Gr, double *out, int r,
                                                      double *partials, unsigned *counter, double *scal,
                                                      int slot) {
  extern __shared__ double smem[];
  const int t = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const Geo<D> geo(r);
  const TileInfo T = tile_info<D>(L, t, r);
  const int nE = T.nR * r;
  double *sred = smem;
  double *sW = smem + 64;
  double acc[2] = {0.0, 0.0};
  for (int le = tid; le < nE; le += nth) {
    const int lrow = le / r, c = le - lrow * r;
    double w = Y[T.ebase + le];
    if (V != nullptr) {
      const double v = V[T.ebase + le];
      w = fma(alpha, v, w);
      acc[0] = fma(v, v, acc[0]);
      if (Gr != nullptr) acc[1] = fma(Gr[T.ebase + le], v, acc[1]);
    }
    sW[geo.soff(lrow, c)] = w;
  }
  __syncthreads();
  for (int u = tid; u < T.nP + T.nS; u += nth) {
    if (u < T.nP) {
      stiefel_polar<D>(sW + geo.pose_base(u), r, geo.RS);
    } else {
      const int lrow = T.nP * (D + 1) + (u - T.nP);
      if (T.row0 + lrow >= L.nPoseRows + L.l) {  // ObliqueManifold.cpp:6-14
        double *w = sW + geo.soff(lrow, 0);
        double s = 0.0;
        for (int c = 0; c < r; ++c) s = fma(w[c], w[c], s);
        const double inv = 1.0 / sqrt(s);
        for (int c = 0; c < r; ++c) w[c] *= inv;
      }
    }
  }
  __syncthreads();
  for (int le = tid; le < nE; le += nth) {
    const int lrow = le / r, c = le - lrow * r;
    out[T.ebase + le] = sW[geo.soff(lrow, c)];
  }
  if (partials != nullptr) finish_reduction<2>(acc, sred, partials, counter, scal, nullptr, POST_STORE, slot);
}

// ==================================================================== k_tangent ==
// out = proj_Y(V)   (src/CORA_problem.cpp:782-820), tier-1 entry.
template <int D>
__global__ void __launch_bounds__(kThreads) k_tangent(const DevLayout L, const double *__restrict__ Y,
                                                      const double *__restrict__ V, double *out, int r) {
  extern __shared__ double smem[];
  const int t = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const Geo<D> geo(r);
  const TileInfo T = tile_info<D>(L, t, r);
  const int nE = T.nR * r;
  const size_t vstride = (size_t)L.TR * geo.RS + L.TP;
  double *sW = smem, *sY = smem + vstride;
  for (int le = tid; le < nE; le += nth) {
    const int lrow = le / r, c = le - lrow * r;
    const int so = geo.soff(lrow, c);
    sW[so] = V[T.ebase + le];
    sY[so] = Y[T.ebase + le];
  }
  __syncthreads();
  tile_epilogue<D, false>(L, T, geo, sY, nullptr, nullptr, sW);
  __syncthreads();
  for (int le = tid; le < nE; le += nth) {
    const int lrow = le / r, c = le - lrow * r;
    out[T.ebase + le] = sW[geo.soff(lrow, c)];
  }
}

// ===================================================================== k_lambda ==
// Lambda blocks (src/CORA_problem.cpp:1105-1131): lam_st[i] = sym((QY)_i Y_i^T)
// (row-major d x d per pose), lam_ob[k] = (QY)_k . y_k.  One thread per pose / range.
template <int D>
__global__ void __launch_bounds__(kThreads) k_lambda(const DevLayout L, const double *__restrict__ Y,
                                                     const double *__restrict__ QY, double *lam_st,
                                                     double *lam_ob, int r) {
  constexpr int D1 = D + 1;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < L.n) {
    const double *y = Y + (size_t)u * D1 * r, *g = QY + (size_t)u * D1 * r;
    double P[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) P[a][b] = 0.0;
    for (int c = 0; c < r; ++c)
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) P[a][b] = fma(g[a * r + c], y[b * r + c], P[a][b]);
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) lam_st[((size_t)u * D + a) * D + b] = 0.5 * (P[a][b] + P[b][a]);
  } else if (u < L.n + L.m) {
    const int k = u - L.n;
    const size_t row = (size_t)L.nPoseRows + L.l + k;
    double s = 0.0;
    for (int c = 0; c < r; ++c) s = fma(QY[row * r + c], Y[row * r + c], s);
    lam_ob[k] = s;
  }
}

// Build the values of S + eta I = Q - Lambda + eta I on the layout of Q: only slot 0
// of the block-ELL and the scalar diagonal change (layout.hpp).
template <int D>
__global__ void __launch_bounds__(kThreads) k_patch_certificate(const DevLayout L, const double *lam_st,
                                                                const double *lam_ob, double eta,
                                                                double *bvalS, double *sdiagS) {
  constexpr int D1 = D + 1;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < L.n) {
    const int t = u / L.TP, p = u - t * L.TP;
    double *bv = bvalS + L.tile_boff[t] + p;  // slot 0
#pragma unroll
    for (int a = 0; a < D1; ++a)
#pragma unroll
      for (int b = 0; b < D1; ++b) {
        double v = bv[(size_t)(a * D1 + b) * L.TP];
        if (a < D && b < D) v -= lam_st[((size_t)u * D + a) * D + b];
        if (a == b) v += eta;
        bv[(size_t)(a * D1 + b) * L.TP] = v;
      }
  } else if (u < L.n + L.l + L.m) {
    const int k = u - L.n;  // scalar row index
    double v = L.sdiag[k] + eta;
    if (k >= L.l) v -= lam_ob[k - L.l];
    sdiagS[k] = v;
  }
}

// ------------------------------------------------------------- layout changes ---
// reference column-major N x r  ->  internal row-major (row permuted)
static __global__ void __launch_bounds__(kThreads) k_import(const int *__restrict__ int2ref, const double *__restrict__ src,
                                                     double *__restrict__ dst, int N, int r, int src_cols,
                                                     int io_rows) {
  // io_rows < N: the host matrix holds only the leading io_rows reference rows (Formulation::Implicit: rotations
  // and ranges); the remaining internal rows (translations) are zero
  const long long nE = (long long)N * r;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / r), c = (int)(e - (long long)row * r);
    const int ref = int2ref[row];
    dst[e] = (c < src_cols && ref < io_rows) ? src[(size_t)c * io_rows + ref] : 0.0;
  }
}
static __global__ void __launch_bounds__(kThreads) k_export(const int *__restrict__ int2ref, const double *__restrict__ src,
                                                     double *__restrict__ dst, int N, int r, int io_rows) {
  const long long nE = (long long)N * r;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / r), c = (int)(e - (long long)row * r);
    const int ref = int2ref[row];
    if (ref < io_rows) dst[(size_t)c * io_rows + ref] = src[e];
  }
}

// Formulation::Implicit helpers.  Translation rows in the internal order: row d of every pose block, then the l
// landmark rows.  mode 0: out = x with the translation rows zeroed; mode 1: out's translation rows = -z's
// (out otherwise untouched); mode 2: zero the translation rows of out in place.
static __global__ void __launch_bounds__(kThreads) k_translation_rows(int mode, int nPoseRows, int D1, int l, int r,
                                                               const double *__restrict__ x, const double *__restrict__ z,
                                                               double *out, long long nE, const CgCtrl *ctrl) {
  if (ctrl != nullptr && *((volatile const int *)&ctrl->state) != 0) return;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / r;
    const bool tr = row < nPoseRows ? (row % D1 == D1 - 1) : (row < nPoseRows + l);
    if (mode == 0) out[e] = tr ? 0.0 : x[e];
    else if (tr) out[e] = (mode == 1) ? -z[e] : 0.0;
  }
}

// out = a*x + b*y (flat)
static __global__ void __launch_bounds__(kThreads) k_axpby(double a, const double *__restrict__ x, double b,
                                                    const double *__restrict__ y, double *out, long long nE) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = a * x[e] + (y != nullptr ? b * y[e] : 0.0);
}
// z = V * dinv  (Jacobi, src/CORA_problem.cpp:888-889)
static __global__ void __launch_bounds__(kThreads) k_jacobi(const double *__restrict__ dinv, const double *__restrict__ V,
                                                     double *out, int r, long long nE) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = V[e] * dinv[e / r];
}

// Deterministic flat inner products: up to 2 pairs, one partial per CTA.
static __global__ void __launch_bounds__(kThreads) k_dot2(const double *__restrict__ a0, const double *__restrict__ b0,
                                                   const double *__restrict__ a1, const double *__restrict__ b1,
                                                   long long nE, double *partials, unsigned *counter,
                                                   double *scal, int slot) {
  __shared__ double sred[64];
  double acc[2] = {0.0, 0.0};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    acc[0] = fma(a0[e], b0[e], acc[0]);
    if (a1 != nullptr) acc[1] = fma(a1[e], b1[e], acc[1]);
  }
  finish_reduction<2>(acc, sred, partials, counter, scal, nullptr, POST_STORE, slot);
}

}  // namespace cora_b200
