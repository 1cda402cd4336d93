// persistent_chain.cuh -- Z = M^-1 V with the chain Cholesky factor (chain_chol.cuh), executed INSIDE
// the persistent TNT kernel: the launches of chain_solve() become phases separated by the grid
// barrier (Preconditioner::RegularizedCholesky, src/CORA_problem.cpp:869-903 ->
// blockCholeskySolve, src/CORA_preconditioners.cpp:46-83).
//   pre | forward level 0 .. top (+ top back substitution) | backward levels | border | post
// Work is spread over all threads of the grid by global thread id; data written in one phase is
// read in the next only after a grid barrier (whose fence invalidates L1), so plain loads are safe
// and the per-chunk routines are the very same host/device functions the host test hook runs.
#pragma once
#include "chain_chol.cuh"
#include "persistent.cuh"

namespace cora_b200 {

constexpr int kMaxChainLevels = 8;

struct ChainDev {
  int nl, n, l, m;
  int pinned_pose_row, pinned_landmark;
  ChunkGeo G[kMaxChainLevels];
  const double *fwd[kMaxChainLevels], *bwd[kMaxChainLevels], *UR[kMaxChainLevels];
  double *sol[kMaxChainLevels], *rhs[kMaxChainLevels], *cL[kMaxChainLevels], *cR[kMaxChainLevels];
  const int *rinc_ptr, *rinc_k, *rend_x, *bl_ptr, *bl_row;
  const double *rinc_e, *rdinv, *rend_e, *bl_val, *W, *SLinv;
  double *u, *Y;
  // landmark reductions split in chunks of kLmChunk entries over all CTAs: partials [l][max chunks][r]
  int max_rinc_chunks, max_bl_chunks;
  double *part_pre, *part_bl;
};
constexpr int kLmChunk = 256;

// CTA-wide reduction of one landmark's sparse column: out[c] = sum_q val[q] * X[row[q]*r + c]
__device__ __forceinline__ void cta_sparse_dot(PCtx &c, int q0, int q1, const int *idx, const double *val,
                                               const double *scale_by_idx, const double *X, size_t row_off, int r,
                                               double *out_smem) {
  const int per = c.nth / r, e = c.tid / r, cc = c.tid - e * r;
  double acc = 0.0;
  if (e < per)
    for (int q = q0 + e; q < q1; q += per) {
      const int k = idx[q];
      double v = val[q];
      if (scale_by_idx != nullptr) v *= scale_by_idx[k];
      acc = fma(v, X[(row_off + (size_t)k) * r + cc], acc);
    }
  __syncthreads();
  c.sW[c.tid] = acc;
  __syncthreads();
  if (c.tid < r) {
    double s = 0.0;
    for (int i = 0; i < per; ++i) s += c.sW[i * r + c.tid];
    out_smem[c.tid] = s;
  }
  __syncthreads();
}

template <int D>
__device__ __forceinline__ void chain_apply_persistent(const ChainDev &C, PCtx &c, const double *V, double *Z) {
  constexpr int B = D + 1;
  const int r = c.r, n = C.n, l = C.l, m = C.m;
  const size_t np = (size_t)n * B, rg0 = np + l;
  const long long gtid = (long long)c.b * c.nth + c.tid, gsize = (long long)c.G * c.nth;
  double *Y = C.Y;
  ph_begin(c);
  // ---- pre: Y = V on pose rows minus the range elimination (k_chain_pre); the landmark rows' reductions
  // over their incident ranges are split in chunks over all CTAs and combined after the border phase ----
  for (int item = c.b; item < l * C.max_rinc_chunks; item += c.G) {
    const int j = item / C.max_rinc_chunks, ch = item - j * C.max_rinc_chunks;
    const int q0 = C.rinc_ptr[n + j] + ch * kLmChunk, q1 = min(C.rinc_ptr[n + j + 1], q0 + kLmChunk);
    if (q0 >= q1) continue;  // uniform per CTA
    cta_sparse_dot(c, q0, q1, C.rinc_k, C.rinc_e, C.rdinv, V, rg0, r, c.sW + c.nth);
    if (c.tid < r) C.part_pre[((size_t)j * C.max_rinc_chunks + ch) * r + c.tid] = c.sW[c.nth + c.tid];
    __syncthreads();
  }
  {
    const unsigned nE = (unsigned)(np * r), ur = (unsigned)r;
    for (unsigned e = (unsigned)gtid; e < nE; e += (unsigned)gsize) {
      const unsigned row = e / ur;
      const int cc = (int)(e - row * ur);
      double v = V[e];
      if (row % B == B - 1) {
        const int x = (int)(row / B);
        for (int q = C.rinc_ptr[x]; q < C.rinc_ptr[x + 1]; ++q)
          v -= C.rinc_e[q] * C.rdinv[C.rinc_k[q]] * V[(rg0 + C.rinc_k[q]) * r + cc];
        if ((int)row == C.pinned_pose_row) v = 0.0;
      }
      Y[e] = v;
    }
  }
  ph_end(c, PH_CH_PRE);
  grid_sync(c);
  // ---- the pose chain: forward levels, top solve, backward levels ----
  if (n > 0) {
    for (int lv = 0; lv < C.nl; ++lv) {
      const ChunkGeo G = C.G[lv];
      const bool top = (lv == C.nl - 1);
      const long long items = (long long)G.K * r;
      for (long long t = gtid; t < items; t += gsize) {
        const int k = (int)(t / r), col = (int)(t - (long long)k * r);
        forward_chunk<B>(G, k, col, r, C.fwd[lv], C.UR[lv], lv == 0 ? Y : C.sol[lv], lv == 0 ? Y : C.rhs[lv],
                         lv > 0 ? (lv == 1 ? Y : C.rhs[lv - 1]) : nullptr, lv > 0 ? C.G[lv - 1].c : 0,
                         lv > 0 ? C.cL[lv - 1] : nullptr, lv > 0 ? C.cR[lv - 1] : nullptr, C.cL[lv], C.cR[lv]);
        if (top) backward_chunk<B>(G, k, col, r, C.bwd[lv], lv == 0 ? Y : C.sol[lv], nullptr);
      }
      ph_end(c, PH_CH_FWD);
      grid_sync(c);
    }
    for (int lv = C.nl - 2; lv >= 0; --lv) {
      const ChunkGeo G = C.G[lv];
      const long long items = (long long)G.K * r;
      for (long long t = gtid; t < items; t += gsize) {
        const int k = (int)(t / r), col = (int)(t - (long long)k * r);
        backward_chunk<B>(G, k, col, r, C.bwd[lv], lv == 0 ? Y : C.sol[lv], C.sol[lv + 1]);
      }
      ph_end(c, PH_CH_BWD);
      grid_sync(c);
    }
  }
  // ---- border: partial sums of sum_B B[row, j] y[row] per (landmark, chunk)  (k_chain_border) ----
  if (l > 0) {
    for (int item = c.b; item < l * C.max_bl_chunks; item += c.G) {
      const int j = item / C.max_bl_chunks, ch = item - j * C.max_bl_chunks;
      const int q0 = C.bl_ptr[j] + ch * kLmChunk, q1 = min(C.bl_ptr[j + 1], q0 + kLmChunk);
      if (q0 >= q1) continue;
      cta_sparse_dot(c, q0, q1, C.bl_row, C.bl_val, nullptr, Y, 0, r, c.sW + c.nth);
      if (c.tid < r) C.part_bl[((size_t)j * C.max_bl_chunks + ch) * r + c.tid] = c.sW[c.nth + c.tid];
      __syncthreads();
    }
    ph_end(c, PH_CH_BORDER);
    grid_sync(c);
  }
  // ---- every CTA: u_j = (v_L[j] - range partials) - border partials ; z_L = S_L^-1 u  (l x r values each,
  // fixed summation order) ----
  double *uS = c.sW, *zL = c.sW + l * r;
  for (int i = c.tid; i < l * r; i += c.nth) {
    const int jj = i / r, cc = i - jj * r;
    double y = 0.0;
    if (jj != C.pinned_landmark) {
      y = V[(np + jj) * r + cc];
      const int nc = (C.rinc_ptr[n + jj + 1] - C.rinc_ptr[n + jj] + kLmChunk - 1) / kLmChunk;
      for (int ch = 0; ch < nc; ++ch) y -= C.part_pre[((size_t)jj * C.max_rinc_chunks + ch) * r + cc];
    }
    const int nb = (C.bl_ptr[jj + 1] - C.bl_ptr[jj] + kLmChunk - 1) / kLmChunk;
    for (int ch = 0; ch < nb; ++ch) y -= C.part_bl[((size_t)jj * C.max_bl_chunks + ch) * r + cc];
    uS[i] = y;
  }
  __syncthreads();
  for (int i = c.tid; i < l * r; i += c.nth) {
    const int jj = i / r, cc = i - jj * r;
    double sacc = 0.0;
    for (int j2 = 0; j2 < l; ++j2) sacc = fma(C.SLinv[(size_t)jj * l + j2], uS[j2 * r + cc], sacc);
    zL[i] = (jj == C.pinned_landmark) ? 0.0 : sacc;
  }
  __syncthreads();
  // ---- post: z_P = y_P - W z_L ; z_L ; ranges back-substituted  (k_chain_post) ----
  {
    const unsigned nE = (unsigned)((rg0 + m) * r), ur = (unsigned)r;
    auto zpose = [&](size_t row, int cc) -> double {
      if ((int)row == C.pinned_pose_row) return 0.0;
      double s = Y[row * r + cc];
      for (int j = 0; j < l; ++j) s = fma(-C.W[row * l + j], zL[(size_t)j * r + cc], s);
      return s;
    };
    for (unsigned e = (unsigned)gtid; e < nE; e += (unsigned)gsize) {
      const unsigned row = e / ur;
      const int cc = (int)(e - row * ur);
      double out;
      if (row < np) {
        out = zpose(row, cc);
      } else if (row < rg0) {
        out = zL[(row - np) * r + cc];
      } else {
        const size_t k = row - rg0;
        double s = V[e];
        for (int p = 0; p < 2; ++p) {
          const int x = C.rend_x[k * 2 + p];
          if (x < 0) continue;
          const double zx = x < n ? zpose((size_t)x * B + (B - 1), cc) : zL[(size_t)(x - n) * r + cc];
          s = fma(-C.rend_e[k * 2 + p], zx, s);
        }
        out = s * C.rdinv[k];
      }
      Z[e] = out;
    }
  }
  ph_end(c, PH_CH_POST);
  grid_sync(c);
}

}  // namespace cora_b200
