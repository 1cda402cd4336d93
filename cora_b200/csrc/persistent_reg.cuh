// persistent_reg.cuh -- register / warp-shuffle formulation of the preconditioned update phase of a CG
// iteration for the persistent TNT kernel at ranks without a compiled streaming path (stream.cuh).
//
// The shared-memory tile pipeline (persistent.cuh) is latency bound at two CTAs per SM: four block
// barriers per tile and a thread-per-pose epilogue leave the LSU idle (ncu: issue active 24 %, 0.34
// eligible warps per cycle, DRAM 16 % of peak; profiles/README.md r01b).  The blocks of this problem are
// tiny ((d+1)x(d+1) against r columns) and carry no reuse beyond the r columns of a pose, so here
//   * a GROUP of GS = 2^ceil(log2 r) adjacent lanes owns one pose, lane c its column c; every operand is
//     loaded straight from global memory (block values [slot][a][b][pose] give one full 32-byte sector per
//     warp load at GS = 8; the r lanes of a pose read the same value -> one L1 request),
//   * all loads of a pose are independent, so each lane keeps ~20 of them in flight and the warps never
//     meet at a block barrier inside the phase,
//   * the d x d Gram-type products of the Riemannian epilogue (sym(Y W^T)) are reduced across the group
//     with __shfl_xor (fixed order: deterministic).
// Reference semantics as in persistent.cuh: src/CORA_problem.cpp:742-903, StiefelProduct.cpp:38-55,
// ObliqueManifold.cpp:16-27, IterativeSolvers.h:374-386.
#pragma once
#include "persistent.cuh"

#ifndef CORA_UPDATE_UNROLL
#define CORA_UPDATE_UNROLL 2  // pose groups per warp step of update_reg
#endif
#ifndef CORA_UPDATE_UNROLL_SCALAR
#define CORA_UPDATE_UNROLL_SCALAR 4  // scalar-row groups per warp step of update_reg
#endif

namespace cora_b200 {

__device__ __forceinline__ double group_sum(double v, int GS) {
  for (int o = GS >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int group_size(int r) {
  int g = 1;
  while (g < r) g <<= 1;
  return g;
}

// Tangent projection of the pose block held one column per lane: w <- w - sym(Y W^T) y  (rows a < D).
// Returns the symmetric matrix in Sm (every lane of the group holds all of it).
template <int D>
__device__ __forceinline__ void group_tangent(const double (&y)[D], double (&w)[D + 1], int GS, bool active,
                                              double (&Sm)[D * D]) {
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) Sm[a * D + b] = group_sum(active ? y[a] * w[b] : 0.0, GS);
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a + 1; b < D; ++b) {
      const double s = 0.5 * (Sm[a * D + b] + Sm[b * D + a]);
      Sm[a * D + b] = s;
      Sm[b * D + a] = s;
    }
  double t[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = w[a];
#pragma unroll
    for (int b = 0; b < D; ++b) s = fma(-Sm[a * D + b], y[b], s);
    t[a] = s;
  }
#pragma unroll
  for (int a = 0; a < D; ++a) w[a] = t[a];
}

// STPCG update + preconditioner closure: R += alpha HP (AXPY) ; V = proj_Y(z), z = R*dinv | R | Z
// acc[0] += <R,V>, acc[1] += <V,V>
template <int D, bool AXPY>
__device__ __forceinline__ void update_reg(const DevLayout &L, PCtx &c, const double *Y, const double *HP, double *R,
                                           const double *Z, double *V, double alpha, int zsrc, double *acc,
                                           double *S = nullptr, const double *P = nullptr) {  // S != nullptr: s += alpha p
  constexpr int D1 = D + 1;
  const int r = c.r, TP = L.TP;
  const int GS = group_size(r), PPW = 32 / GS;
  const int lane = c.tid & 31, warp = c.tid >> 5, nwarps = c.nth >> 5;
  const int sub = lane / GS, cc = lane - sub * GS;
  const bool col_ok = cc < r;
  ph_begin(c);
  const int P0 = min(c.t0 * TP, L.n), P1 = min(c.t1 * TP, L.n);
  // U pose groups per warp step: all loads of the U groups are issued before the first dependent shuffle / store, which
  // doubles the bytes in flight per warp (the phase is latency bound at 2 CTAs/SM).  The dot accumulation order per
  // thread is unchanged (group 0 rows, then group 1 rows = the order of the one-group loop).
  constexpr int U = CORA_UPDATE_UNROLL;
  const int step = nwarps * PPW;
  for (int pb = P0 + warp * PPW; pb < P1; pb += U * step) {
    double rr[U][D1], z[U][D1], y[U][D], hp[U][D1], dv[U][D1];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * step + sub;
      act[u] = col_ok && p < P1;
#pragma unroll
      for (int a = 0; a < D1; ++a) { rr[u][a] = 0.0; z[u][a] = 0.0; hp[u][a] = 0.0; dv[u][a] = 0.0; }
#pragma unroll
      for (int a = 0; a < D; ++a) y[u][a] = 0.0;
      if (act[u]) {
#pragma unroll
        for (int a = 0; a < D1; ++a) {
          const size_t e = ((size_t)p * D1 + a) * r + cc;
          rr[u][a] = R[e];
          if (AXPY) hp[u][a] = HP[e];
          if (zsrc == 0) dv[u][a] = __ldg(L.dinv + p * D1 + a);
          else if (zsrc == 2) z[u][a] = Z[e];
          if (a < D) y[u][a] = Y[e];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * step + sub;
      if (act[u]) {
#pragma unroll
        for (int a = 0; a < D1; ++a) {
          const size_t e = ((size_t)p * D1 + a) * r + cc;
          double v = rr[u][a];
          if (AXPY) {
            v = fma(alpha, hp[u][a], v);
            R[e] = v;
            if (S != nullptr) S[e] = fma(alpha, P[e], S[e]);
          }
          rr[u][a] = v;
          if (zsrc == 0) z[u][a] = v * dv[u][a];
          else if (zsrc == 1) z[u][a] = v;
        }
      }
      double Sm[D * D];
      group_tangent<D>(y[u], z[u], GS, act[u], Sm);
      if (act[u]) {
#pragma unroll
        for (int a = 0; a < D1; ++a) {
          V[((size_t)p * D1 + a) * r + cc] = z[u][a];
          acc[0] = fma(rr[u][a], z[u][a], acc[0]);
          acc[1] = fma(z[u][a], z[u][a], acc[1]);
        }
      }
    }
  }
  const int R0 = max(c.t0 * L.TR, L.nPoseRows), R1 = min(c.t1 * L.TR, L.N);
  // scalar rows (landmark + range rows): 4 loads per row only, so US row groups per warp step keep enough bytes in
  // flight; the CTAs that own these rows were the stragglers of the phase (40 us against a median of 25 us)
  constexpr int US = CORA_UPDATE_UNROLL_SCALAR;
  for (int rb = R0 + warp * PPW; rb < R1; rb += US * step) {
    double rr[US], z[US], yv[US], hp[US], dv[US];
    bool act[US];
#pragma unroll
    for (int u = 0; u < US; ++u) {
      const int row = rb + u * step + sub;
      act[u] = col_ok && row < R1;
      rr[u] = 0.0; z[u] = 0.0; yv[u] = 0.0; hp[u] = 0.0; dv[u] = 0.0;
      if (act[u]) {
        const size_t e = (size_t)row * r + cc;
        rr[u] = R[e];
        if (AXPY) hp[u] = HP[e];
        if (zsrc == 0) dv[u] = __ldg(L.dinv + row);
        else if (zsrc == 2) z[u] = Z[e];
        yv[u] = Y[e];
      }
    }
#pragma unroll
    for (int u = 0; u < US; ++u) {
      const int row = rb + u * step + sub;
      const bool is_range = row >= L.nPoseRows + L.l;
      const size_t e = (size_t)row * r + cc;
      if (act[u]) {
        if (AXPY) {
          rr[u] = fma(alpha, hp[u], rr[u]);
          R[e] = rr[u];
          if (S != nullptr) S[e] = fma(alpha, P[e], S[e]);
        }
        if (zsrc == 0) z[u] = rr[u] * dv[u];
        else if (zsrc == 1) z[u] = rr[u];
      }
      const double sd = group_sum((act[u] && is_range) ? yv[u] * z[u] : 0.0, GS);
      if (act[u]) {
        if (is_range) z[u] = fma(-sd, yv[u], z[u]);
        V[e] = z[u];
        acc[0] = fma(rr[u], z[u], acc[0]);
        acc[1] = fma(z[u], z[u], acc[1]);
      }
    }
  }
  ph_end(c, AXPY ? PH_UPDATE : PH_PRECOND);
}

}  // namespace cora_b200
