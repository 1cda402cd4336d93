// persistent_reg.cuh -- register / warp-shuffle formulation of the two hot phases of a CG iteration
// (the Hessian product and the preconditioned update) for the persistent TNT kernel.
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

// ------------------------------------------------------------------------------------------------
// Hybrid: the operands arrive through the double-buffered tile pipeline of persistent.cuh (TMA bulk
// copies of the data-matrix slice, cp.async of the dense tile rows + halo: many bytes in flight at no
// register cost), the compute is the group-per-pose register formulation above reading from the staged
// buffers.  Two block barriers per tile (acquire / release) instead of five, every warp busy through
// the epilogue, results stored straight from registers.
template <int D, int MODE>
__device__ __forceinline__ void qprod_hyb(const DevLayout &L, PCtx &c, const double *X, const double *Y, double *out,
                                          double *out2, const double *longpart, double *lam, double *lamS,
                                          double *acc) {
  constexpr int D1 = D + 1;
  constexpr int NV = (MODE == QM_HESS) ? 2 : 1;
  const int r = c.r, TP = L.TP;
  const PGeo<D> geo(r);
  const int GS = group_size(r), PPW = 32 / GS;
  const int lane = c.tid & 31, warp = c.tid >> 5, nwarps = c.nth >> 5;
  const int sub = lane / GS, cc = lane - sub * GS;
  const bool col_ok = cc < r;
  const int pstride = D1 * geo.RS + geo.PADP;
  ph_begin(c);
  int buf = 0;
  const double *bsrc = (MODE == QM_HESS) ? lam : nullptr;
  if (c.t0 < c.t1) tile_prefetch<D, true, NV>(L, c, c.t0, 0, X, Y, nullptr, bsrc);
  for (int t = c.t0; t < c.t1; ++t) {
    sub_begin(c);
    tile_acquire<D, true, NV>(L, c, t, buf, X, Y, nullptr, bsrc);
    sub_end(c, PH_Q_WAIT);
    const TileBuf B = c.pick(buf);
    const TileInfo T = tile_geom<D>(L, t, r);
    const TileMeta M = tile_meta(L, c, t);
    const double *sX = B.slot[0], *sY = B.slot[1];
    const int S = B.meta[0];
    const int winLo = max(T.row0 - D1, 0);
    const int winHi = min(min(T.row0 + T.nR + D1, L.N), L.nPoseRows);
    double *hub = c.sW;
    const int hs = hub_stride<D>(L, M, r);
    if (M.lq1 > M.lq0) {
      tile_hub_sums<D>(L, c, M, longpart, hub);
      __syncthreads();
    }
    // ---------------- pose blocks of the tile ----------------
    for (int pb = warp * PPW; pb < T.nP; pb += nwarps * PPW) {
      const int pl = pb + sub;
      const bool active = col_ok && pl < T.nP;
      const int p = t * TP + pl;
      double w[D1], xo[D1];
#pragma unroll
      for (int a = 0; a < D1; ++a) { w[a] = 0.0; xo[a] = 0.0; }
      if (active) {
        // spill columns first: their L2 round trip overlaps the block products
        const int k0 = B.gptr[pl], k1 = B.gptr[pl + 1];
        double xs0 = 0.0, xs1 = 0.0;
        if (k0 < k1) xs0 = X[(size_t)(B.spk[k0] & kColMask) * r + cc];
        if (k0 + 1 < k1) xs1 = X[(size_t)(B.spk[k0 + 1] & kColMask) * r + cc];
        const double *xop = sX + pl * pstride + cc;
#pragma unroll
        for (int q = 0; q < D1; ++q) xo[q] = xop[q * geo.RS];
        for (int s = 0; s < S; ++s) {
          const int jb = B.scol[s * TP + pl];
          double x[D1];
          if (jb >= winLo && jb + D1 <= winHi) {
            const int lp = (jb - T.row0 + D1) / D1 - 1;
            const double *xp = sX + lp * pstride + cc;
#pragma unroll
            for (int q = 0; q < D1; ++q) x[q] = xp[q * geo.RS];
          } else {
            const double *xp = X + (size_t)jb * r + cc;
#pragma unroll
            for (int q = 0; q < D1; ++q) x[q] = xp[(size_t)q * r];
          }
          const double *bv = B.sval + (size_t)s * D1 * D1 * TP + pl;
#pragma unroll
          for (int a = 0; a < D1; ++a)
#pragma unroll
            for (int q = 0; q < D1; ++q) w[a] = fma(bv[(a * D1 + q) * TP], x[q], w[a]);
        }
        for (int k = k0; k < k1; ++k) {
          const unsigned pk = B.spk[k];
          const int lr = (int)(pk >> 30);
          const double xg = k == k0 ? xs0 : (k == k0 + 1 ? xs1 : X[(size_t)(pk & kColMask) * r + cc]);
          const double xv = B.spv[k] * xg;
#pragma unroll
          for (int a = 0; a < D1; ++a) w[a] += (lr == a) ? xv : 0.0;
        }
        for (int q = M.lq0; q < M.lq1; ++q) {
          if (L.long_grp[q] != p) continue;
#pragma unroll
          for (int a = 0; a < D1; ++a) w[a] += hub[(q - M.lq0) * hs + a * r + cc];
        }
      }
      if (MODE == QM_SPMM) {
        if (active)
#pragma unroll
          for (int a = 0; a < D1; ++a) out[((size_t)p * D1 + a) * r + cc] = w[a];
        continue;
      }
      double y[D];
#pragma unroll
      for (int a = 0; a < D; ++a) y[a] = 0.0;
      if (active) {
        if (MODE == QM_GRAD) {
#pragma unroll
          for (int a = 0; a < D; ++a) y[a] = xo[a];
#pragma unroll
          for (int a = 0; a < D1; ++a) {
            out2[((size_t)p * D1 + a) * r + cc] = w[a];
            acc[0] = fma(xo[a], w[a], acc[0]);
          }
        } else {
          const double *yp = sY + pl * pstride + cc;
#pragma unroll
          for (int a = 0; a < D; ++a) y[a] = yp[a * geo.RS];
        }
      }
      double Sm[D * D];
      group_tangent<D>(y, w, GS, active, Sm);
      if (active) {
        if (MODE == QM_GRAD) {
          if (cc < D) {  // lane cc writes row cc of the diagonal block of Q - Lambda
            double *lg = lam + M.boff + pl;
#pragma unroll
            for (int b = 0; b < D; ++b) {
              double v = Sm[b];
#pragma unroll
              for (int a = 1; a < D; ++a) v = (cc == a) ? Sm[a * D + b] : v;
              lg[(cc * D1 + b) * TP] = B.sval[(cc * D1 + b) * TP + pl] - v;
            }
          }
#pragma unroll
          for (int a = 0; a < D1; ++a) {
            out[((size_t)p * D1 + a) * r + cc] = w[a];
            acc[1] = fma(w[a], w[a], acc[1]);
          }
        } else {
#pragma unroll
          for (int a = 0; a < D1; ++a) {
            out[((size_t)p * D1 + a) * r + cc] = w[a];
            acc[0] = fma(xo[a], w[a], acc[0]);
            acc[1] = fma(w[a], w[a], acc[1]);
            acc[2] = fma(xo[a], xo[a], acc[2]);
          }
        }
      }
    }
    // ---------------- scalar rows of the tile ----------------
    for (int rb = warp * PPW; rb < T.nS; rb += nwarps * PPW) {
      const int sr = rb + sub;
      const bool active = col_ok && sr < T.nS;
      const int lrow = T.nP * D1 + sr;
      const int row = T.row0 + lrow;
      const int sidx = row - L.nPoseRows;
      const bool is_range = row >= L.nPoseRows + L.l;
      double w = 0.0, xo = 0.0, yv = 0.0;
      if (active) {
        const int u = T.nP + sr;
        const int k0 = B.gptr[u], k1 = B.gptr[u + 1];
        double xs0 = 0.0, xs1 = 0.0;
        if (k0 < k1) xs0 = X[(size_t)(B.spk[k0] & kColMask) * r + cc];
        if (k0 + 1 < k1) xs1 = X[(size_t)(B.spk[k0 + 1] & kColMask) * r + cc];
        xo = sX[geo.soff(lrow, cc)];
        const double dg = (MODE == QM_HESS) ? lamS[sidx] : __ldg(L.sdiag + sidx);
        w = dg * xo;
        if (k0 < k1) w = fma(B.spv[k0], xs0, w);
        if (k0 + 1 < k1) w = fma(B.spv[k0 + 1], xs1, w);
        for (int k = k0 + 2; k < k1; ++k) w = fma(B.spv[k], X[(size_t)(B.spk[k] & kColMask) * r + cc], w);
        for (int q = M.lq0; q < M.lq1; ++q) {
          if (L.long_grp[q] != L.n + sidx) continue;
          w += hub[(q - M.lq0) * hs + cc];
        }
      }
      if (MODE == QM_SPMM) {
        if (active) out[(size_t)row * r + cc] = w;
        continue;
      }
      if (active) {
        if (MODE == QM_GRAD) {
          yv = xo;
          out2[(size_t)row * r + cc] = w;
          acc[0] = fma(xo, w, acc[0]);
        } else {
          yv = sY[geo.soff(lrow, cc)];
        }
      }
      const double s = group_sum((active && is_range) ? yv * w : 0.0, GS);
      if (active) {
        if (is_range) w = fma(-s, yv, w);
        out[(size_t)row * r + cc] = w;
        if (MODE == QM_GRAD) {
          if (cc == 0) lamS[sidx] = __ldg(L.sdiag + sidx) - (is_range ? s : 0.0);
          acc[1] = fma(w, w, acc[1]);
        } else {
          acc[0] = fma(xo, w, acc[0]);
          acc[1] = fma(w, w, acc[1]);
          acc[2] = fma(xo, xo, acc[2]);
        }
      }
    }
    sub_end(c, PH_Q_QX);
    if (MODE == QM_GRAD) asm volatile("fence.proxy.async.global;" ::: "memory");
    tile_release<D, true, NV>(L, c, t, buf, X, Y, nullptr, bsrc);
    sub_end(c, PH_Q_STORE);
  }
  ph_end(c, MODE == QM_HESS ? PH_HESS : PH_GRAD);
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
