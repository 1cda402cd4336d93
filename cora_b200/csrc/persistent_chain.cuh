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
  int smem_doubles;  // dynamic shared memory of the launch: the chunk staging below uses everything behind sW
};
constexpr int kLmChunk = 256;

// CTA-wide reduction of one landmark's sparse column: out[c] = sum_q val[q] * X[row[q]*r + c]
__device__ __forceinline__ void cta_sparse_dot(PCtx &c, int q0, int q1, const int *idx, const double *val,
                                               const double *scale_by_idx, const double *X, size_t row_off, int r,
                                               double *out_smem, const double *X2 = nullptr, double a2 = 0.0) {
  const int per = c.nth / r, e = c.tid / r, cc = c.tid - e * r;
  double acc = 0.0;
  if (e < per)
    for (int q = q0 + e; q < q1; q += per) {
      const int k = idx[q];
      double v = val[q];
      if (scale_by_idx != nullptr) v *= scale_by_idx[k];
      double x = X[(row_off + (size_t)k) * r + cc];
      if (X2 != nullptr) x = fma(a2, X2[(row_off + (size_t)k) * r + cc], x);
      acc = fma(v, x, acc);
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


// Two-stage TMA pipeline over the CTA's contiguous share of the pose rows.  Loads through the LSU -- plain,
// 16-byte or cp.async, any depth -- top out at ~2.3 TB/s in this kernel (every LSU-path variant of the phases
// below measured 29-30 us for 65 MB: the SM's outstanding-miss capacity, not the request depth, is the limit
// at 2 CTAs x 256 threads); bulk copies are not subject to it.  Stage s uses mbarrier c.mbar[s] with the
// CTA-wide parity bits c.mpar0 / c.mpar1 (every thread waits, every thread flips).
__device__ __forceinline__ void chain_tile_issue(PCtx &c, int buf, double *dst0, const double *src0, unsigned bytes0,
                                                 double *dst1, const double *src1, unsigned bytes1) {
  if (c.tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last touched by generic accesses
    mbar_expect_tx(c.mbar + buf, bytes0 + bytes1);
    if (bytes0) bulk_g2s(dst0, src0, bytes0, c.mbar + buf);  // (no landmarks: the W tile is empty)
    if (bytes1) bulk_g2s(dst1, src1, bytes1, c.mbar + buf);
  }
}
__device__ __forceinline__ void chain_tile_wait(PCtx &c, int buf) {
  mbar_wait(c.mbar + buf, buf ? c.mpar1 : c.mpar0);
  if (buf) c.mpar1 ^= 1u; else c.mpar0 ^= 1u;
}

// ---- chunk substitutions out of shared memory ------------------------------------------------------------
// forward_chunk / backward_chunk (chain_chol.cuh) walk a chunk with one global round trip per block step
// (48 / 32 coefficients + the right-hand side): 7 steps x 6 + 5 levels were 220 us of a 400 us CG iteration at
// 100k poses, whatever the level's size.  Here a CTA takes a contiguous range of the level's chunks in batches,
// copies the batch's coefficients (coalesced runs of the [j][e][K] layout) and every thread's own right-hand
// sides into shared memory with 8-byte cp.async -- all of a batch's loads in flight at once, ONE round trip --
// and the serial steps then read shared memory only; results leave with plain stores.  Arithmetic and its
// order are those of forward_chunk / backward_chunk, so the host routines keep pinning the device path.
struct ChainStage {
  double *buf;  // behind sW, 16-byte aligned
  int avail;    // doubles
};

// fills k0, k1 of the CTA's contiguous share of K chunks
// (levels with an even K are split on even boundaries: their coefficient runs can then move as 16-byte aligned
// bulk copies)
__device__ __forceinline__ void chain_cta_chunks(const PCtx &c, int K, int &k0, int &k1) {
  const int sh = (K & 1) ? 0 : 1, Kh = K >> sh;
  k0 = (int)((unsigned)Kh * (unsigned)c.b / (unsigned)c.G) << sh;
  k1 = (int)((unsigned)Kh * (unsigned)(c.b + 1) / (unsigned)c.G) << sh;
}

// rows [0, nrows) x chunks [k0, k0 + KCb) of a [row][K] coefficient array -> dst[row * KCb + kk].  Even K, k0 and
// KCb: one bulk copy per row on mbarrier c.mbar[0] (the LSU path tops out at ~2.3 TB/s in this kernel, see
// chain_tile_issue below); otherwise 8-byte cp.async.  Returns whether chain_tile_wait(c, 0) must follow.
__device__ __forceinline__ bool chain_stage_coeff(PCtx &c, double *dst, const double *src, int nrows, int K, int k0, int KCb) {
  const bool bulk = ((K | k0 | KCb) & 1) == 0 && (unsigned)nrows * (unsigned)KCb * 8u < (1u << 20);
  if (bulk) {
    if (c.tid == 0) mbar_expect_tx(c.mbar, (unsigned)nrows * (unsigned)KCb * 8u);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int row = c.tid; row < nrows; row += c.nth)
      bulk_g2s(dst + (size_t)row * KCb, src + (size_t)row * K + k0, (unsigned)KCb * 8u, c.mbar);
  } else {
    const unsigned nC = (unsigned)(nrows * KCb), uk = (unsigned)KCb;
    for (unsigned idx = c.tid; idx < nC; idx += c.nth) {
      const unsigned row = idx / uk, kk = idx - row * uk;
      cp_async8(dst + idx, src + (size_t)row * K + k0 + kk);
    }
  }
  return bulk;
}

// Shared-memory layout of a batch: the global [j][e][K] layout restricted to the batch's chunks ([row][kk]),
// filled by a flat index loop (every lane of every cp.async instruction active).  A chunk-major copy with 16-byte
// loads and compile-time offsets made the block steps 20 % faster but its warp-per-row staging made every small
// level twice as slow; this one measured best.
template <int B>
__device__ __forceinline__ void chain_forward_level_staged(const ChainDev &C, PCtx &c, int lv, ChainStage st) {
  constexpr int BB = B * B;
  const ChunkGeo G = C.G[lv];
  const int ld = c.r, K = G.K;
  const double *__restrict__ fwd = C.fwd[lv];
  const double *__restrict__ bwd = C.bwd[lv];
  const double *__restrict__ UR = C.UR[lv];
  double *sol = lv == 0 ? C.Y : C.sol[lv];
  double *rhs_cur = lv == 0 ? C.Y : C.rhs[lv];
  const double *rhs_prev = lv > 0 ? (lv == 1 ? C.Y : C.rhs[lv - 1]) : nullptr;
  const int c_prev = lv > 0 ? C.G[lv - 1].c : 0;
  const double *cL_prev = lv > 0 ? C.cL[lv - 1] : nullptr, *cR_prev = lv > 0 ? C.cR[lv - 1] : nullptr;
  double *cL = C.cL[lv], *cR = C.cR[lv];
  const int nslot = lv == 0 ? 1 : 3;
  const bool top = G.top;
  int kb0, kb1;
  chain_cta_chunks(c, K, kb0, kb1);
  if (kb1 <= kb0) return;
  const int Lmax = G.c;
  const int per_chunk = Lmax * (top ? 5 : 3) * BB + BB + nslot * Lmax * B * ld;
  int KC = min(min(st.avail / per_chunk, c.nth / ld), kb1 - kb0);
  if (KC < 1) {  // no room to stage (very large rank): the direct routines
    for (int t = c.tid; t < (kb1 - kb0) * ld; t += c.nth) {
      const int k = kb0 + t / ld, col = t % ld;
      forward_chunk<B>(G, k, col, ld, fwd, UR, sol, rhs_cur, rhs_prev, c_prev, cL_prev, cR_prev, cL, cR);
      if (top) backward_chunk<B>(G, k, col, ld, bwd, sol, nullptr);
    }
    return;
  }
  const int nbatch = (kb1 - kb0 + KC - 1) / KC;
  KC = (kb1 - kb0 + nbatch - 1) / nbatch;
  if (!(K & 1) && (KC & 1) && (KC + 1) * per_chunk <= st.avail && (KC + 1) * ld <= c.nth) ++KC;  // even batches: bulk copies
  for (int k0 = kb0; k0 < kb1; k0 += KC) {
    const int KCb = min(KC, kb1 - k0);
    const int TS = KCb * ld;  // active threads, stride of the per-thread slots
    // rows to stage: interior of the longest chunk of the batch
    int Lb = top ? G.n : G.c - 1;
    if (!top && k0 + KCb == K) Lb = max(Lb, G.interior(K - 1));
    double *sC = st.buf;                                  // [Lb*3*BB][KCb]  forward coefficients
    double *sBw = sC + (size_t)Lb * 3 * BB * KCb;         // [Lb*2*BB][KCb]  backward coefficients (top level only)
    double *sU = sBw + (top ? (size_t)Lb * 2 * BB * KCb : 0);  // [KCb][BB]
    double *sB = sU + (size_t)KCb * BB;                   // [(j*B+a)*nslot + s][TS]
    const bool bulk = !top && chain_stage_coeff(c, sC, fwd, Lb * 3 * BB, K, k0, KCb);
    if (top) {  // one chunk: cp.async
      const unsigned nC = (unsigned)(Lb * 3 * BB * KCb), nW = (unsigned)(Lb * 2 * BB * KCb), uk = (unsigned)KCb;
      for (unsigned idx = c.tid; idx < nC; idx += c.nth) {
        const unsigned row = idx / uk, kk = idx - row * uk;
        cp_async8(sC + idx, fwd + (size_t)row * K + k0 + kk);
      }
      for (unsigned idx = c.tid; idx < nW; idx += c.nth) {
        const unsigned row = idx / uk, kk = idx - row * uk;
        cp_async8(sBw + idx, bwd + (size_t)row * K + k0 + kk);
      }
    }
    for (int idx = c.tid; idx < KCb * BB; idx += c.nth) cp_async8(sU + idx, UR + (size_t)k0 * BB + idx);
    const int t = c.tid;
    const bool active = t < TS;
    const int kk = active ? t / ld : 0, col = t - kk * ld, k = k0 + kk;
    const int g0 = G.first(k), L = active ? G.interior(k) : 0;
    const bool hasR = G.has_right(k);
    if (active) {
      for (int j = 0; j < L; ++j) {
        const int g = g0 + j;
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          double *dst = sB + (size_t)((j * B + a) * nslot) * TS + t;
          if (lv == 0) {
            cp_async8(dst, sol + ((size_t)g * B + a) * ld + col);
          } else {
            const size_t s = (size_t)((g + 1) * c_prev - 1);
            cp_async8(dst, rhs_prev + (s * B + a) * ld + col);
            cp_async8(dst + TS, cR_prev + ((size_t)g * B + a) * ld + col);
            cp_async8(dst + 2 * TS, cL_prev + ((size_t)(g + 1) * B + a) * ld + col);
          }
        }
      }
      if (lv > 0 && hasR) {  // the separator: materialise its right-hand side for the next level
        const int g = g0 + L;
        const size_t s = (size_t)((g + 1) * c_prev - 1);
        double bs[B];
        CB_UNROLL
        for (int a = 0; a < B; ++a)
          bs[a] = rhs_prev[(s * B + a) * ld + col] - cR_prev[((size_t)g * B + a) * ld + col] -
                  cL_prev[((size_t)(g + 1) * B + a) * ld + col];
        CB_UNROLL
        for (int a = 0; a < B; ++a) rhs_cur[((size_t)g * B + a) * ld + col] = bs[a];
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    if (bulk) chain_tile_wait(c, 0);
    __syncthreads();
    if (active) {
      double w[B], acc[B];
      CB_UNROLL
      for (int a = 0; a < B; ++a) { w[a] = 0.0; acc[a] = 0.0; }
      for (int j = 0; j < L; ++j) {
        const int g = g0 + j;
        const double *f = sC + (size_t)j * 3 * BB * KCb + kk;
        double *bsl = sB + (size_t)(j * B * nslot) * TS + t;
        double y[B];
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          double s = bsl[(size_t)(a * nslot) * TS];
          if (lv > 0) s = s - bsl[(size_t)(a * nslot + 1) * TS] - bsl[(size_t)(a * nslot + 2) * TS];
          CB_UNROLL
          for (int q = 0; q < B; ++q) s -= f[(a * B + q) * KCb] * w[q];
          y[a] = s;
        }
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          double s = 0.0;
          CB_UNROLL
          for (int q = 0; q < B; ++q) s += f[(BB + a * B + q) * KCb] * y[q];
          w[a] = s;
        }
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          double s = acc[a];
          CB_UNROLL
          for (int q = 0; q < B; ++q) s += f[(2 * BB + a * B + q) * KCb] * w[q];
          acc[a] = s;
        }
        if (top) {
          CB_UNROLL
          for (int a = 0; a < B; ++a) bsl[(size_t)(a * nslot) * TS] = w[a];  // kept for the back substitution below
        } else {
          CB_UNROLL
          for (int a = 0; a < B; ++a) sol[((size_t)g * B + a) * ld + col] = w[a];
        }
      }
      CB_UNROLL
      for (int a = 0; a < B; ++a) {
        cL[((size_t)k * B + a) * ld + col] = acc[a];
        double s = 0.0;
        if (hasR)
          CB_UNROLL
          for (int q = 0; q < B; ++q) s += sU[kk * BB + q * B + a] * w[q];  // U^T w_last
        cR[((size_t)k * B + a) * ld + col] = s;
      }
      if (top) {  // backward_chunk with xsep == nullptr: xL = 0, xn starts at 0
        double xn[B];
        CB_UNROLL
        for (int a = 0; a < B; ++a) xn[a] = 0.0;
        for (int j = L - 1; j >= 0; --j) {
          const int g = g0 + j;
          const double *f = sBw + (size_t)j * 2 * BB * KCb + kk;
          const double *bsl = sB + (size_t)(j * B * nslot) * TS + t;
          double x[B];
          CB_UNROLL
          for (int a = 0; a < B; ++a) {
            double s = bsl[(size_t)(a * nslot) * TS];
            CB_UNROLL
            for (int q = 0; q < B; ++q) s -= f[(a * B + q) * KCb] * xn[q];  // xL = 0 at the top
            x[a] = s;
          }
          CB_UNROLL
          for (int a = 0; a < B; ++a) {
            sol[((size_t)g * B + a) * ld + col] = x[a];
            xn[a] = x[a];
          }
        }
      }
    }
    __syncthreads();  // the staging buffers are reused by the next batch / phase
  }
}

template <int B>
__device__ __forceinline__ void chain_backward_level_staged(const ChainDev &C, PCtx &c, int lv, ChainStage st) {
  constexpr int BB = B * B;
  const ChunkGeo G = C.G[lv];
  const int ld = c.r, K = G.K;
  const double *__restrict__ bwd = C.bwd[lv];
  double *sol = lv == 0 ? C.Y : C.sol[lv];
  const double *xsep = C.sol[lv + 1];
  int kb0, kb1;
  chain_cta_chunks(c, K, kb0, kb1);
  if (kb1 <= kb0) return;
  const int Lmax = G.c;
  const int per_chunk = Lmax * 2 * BB + Lmax * B * ld;
  int KC = min(min(st.avail / per_chunk, c.nth / ld), kb1 - kb0);
  if (KC < 1) {
    for (int t = c.tid; t < (kb1 - kb0) * ld; t += c.nth)
      backward_chunk<B>(G, kb0 + t / ld, t % ld, ld, bwd, sol, xsep);
    return;
  }
  const int nbatch = (kb1 - kb0 + KC - 1) / KC;
  KC = (kb1 - kb0 + nbatch - 1) / nbatch;
  if (!(K & 1) && (KC & 1) && (KC + 1) * per_chunk <= st.avail && (KC + 1) * ld <= c.nth) ++KC;  // even batches: bulk copies
  for (int k0 = kb0; k0 < kb1; k0 += KC) {
    const int KCb = min(KC, kb1 - k0);
    const int TS = KCb * ld;
    int Lb = G.c - 1;
    if (k0 + KCb == K) Lb = max(Lb, G.interior(K - 1));
    double *sC = st.buf;                            // [Lb*2*BB][KCb]
    double *sB = sC + (size_t)Lb * 2 * BB * KCb;    // [j*B+a][TS]
    const bool bulk = chain_stage_coeff(c, sC, bwd, Lb * 2 * BB, K, k0, KCb);
    const int t = c.tid;
    const bool active = t < TS;
    const int kk = active ? t / ld : 0, col = t - kk * ld, k = k0 + kk;
    const int g0 = G.first(k), L = active ? G.interior(k) : 0;
    double xL[B], xn[B];
    CB_UNROLL
    for (int a = 0; a < B; ++a) { xL[a] = 0.0; xn[a] = 0.0; }
    if (active) {
      for (int j = 0; j < L; ++j)
        CB_UNROLL
        for (int a = 0; a < B; ++a)
          cp_async8(sB + (size_t)(j * B + a) * TS + t, sol + ((size_t)(g0 + j) * B + a) * ld + col);
      CB_UNROLL
      for (int a = 0; a < B; ++a) {
        if (G.has_left(k)) xL[a] = xsep[((size_t)(k - 1) * B + a) * ld + col];
        if (G.has_right(k)) xn[a] = xsep[((size_t)k * B + a) * ld + col];
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    if (bulk) chain_tile_wait(c, 0);
    __syncthreads();
    if (active) {
      if (G.has_right(k))
        CB_UNROLL
        for (int a = 0; a < B; ++a) sol[((size_t)(g0 + L) * B + a) * ld + col] = xn[a];
      for (int j = L - 1; j >= 0; --j) {
        const int g = g0 + j;
        const double *f = sC + (size_t)j * 2 * BB * KCb + kk;
        const double *bsl = sB + (size_t)(j * B) * TS + t;
        double x[B];
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          double s = bsl[(size_t)a * TS];
          CB_UNROLL
          for (int q = 0; q < B; ++q) s -= f[(a * B + q) * KCb] * xn[q] + f[(BB + a * B + q) * KCb] * xL[q];
          x[a] = s;
        }
        CB_UNROLL
        for (int a = 0; a < B; ++a) {
          sol[((size_t)g * B + a) * ld + col] = x[a];
          xn[a] = x[a];
        }
      }
    }
    __syncthreads();
  }
}

template <int D>
__device__ __forceinline__ void chain_apply_persistent(const ChainDev &C, PCtx &c, const double *V, double *Z,
                                                       const double *HP = nullptr, double alpha = 0.0,
                                                       double *Vnew = nullptr) {
  constexpr int B = D + 1;
  const int r = c.r, n = C.n, l = C.l, m = C.m;
  const size_t np = (size_t)n * B, rg0 = np + l;
  const long long gtid = (long long)c.b * c.nth + c.tid, gsize = (long long)c.G * c.nth;
  double *Y = C.Y;
  // HP != nullptr: the right-hand side is V + alpha HP (the residual update r += alpha Hp of STPCG,
  // IterativeSolvers.h:377, folded into this phase: one pass and one grid barrier less per CG iteration); it is
  // evaluated on the fly wherever this phase reads it -- V and HP are complete and not written here -- and stored
  // to Vnew, which the later phases (and the caller) read
  auto veff = [&](size_t i) -> double {
    double x = V[i];
    if (HP != nullptr) x = fma(alpha, HP[i], x);
    return x;
  };
  ph_begin(c);
  // ---- pre: Y = V on pose rows minus the range elimination (k_chain_pre); the landmark rows' reductions
  // over their incident ranges are split in chunks over all CTAs and combined after the border phase ----
  // the CTA's share of the pose rows: even boundaries (16-byte aligned tiles for any r), the last CTA to the end
  const size_t row0c = (np * (size_t)c.b / c.G) & ~(size_t)1;
  const size_t row1c = c.b == c.G - 1 ? np : ((np * (size_t)(c.b + 1) / c.G) & ~(size_t)1);
  const size_t row1e = row1c & ~(size_t)1;  // tiles cover [row0c, row1e); an odd last row is handled on its own
  const int stage_avail = C.smem_doubles - (int)(c.sW - c.smem) - 2 * c.nth;  // behind the scratch of cta_sparse_dot
  double *const stage_base = c.sW + 2 * c.nth;
  // pass 1 streams V (and HP) through two shared-memory tiles; the first two are requested before anything else
  const int nin = HP != nullptr ? 2 : 1;
  const int TRP = min((stage_avail / 2) / (nin * r), 2048) & ~1;
  const bool pre_tiles = TRP >= 2 && row1e > row0c;
  const int pre_tstride = TRP * nin * r;
  const int pre_ntile = pre_tiles ? (int)((row1e - row0c + TRP - 1) / TRP) : 0;
  auto pre_issue = [&](int t) {
    double *sV = stage_base + (t & 1) * pre_tstride, *sH = sV + TRP * r;
    const size_t ra = row0c + (size_t)t * TRP;
    const unsigned by = (unsigned)(min((size_t)TRP, row1e - ra) * r * sizeof(double));
    chain_tile_issue(c, t & 1, sV, V + ra * r, by, sH, HP != nullptr ? HP + ra * r : nullptr, HP != nullptr ? by : 0u);
  };
  for (int t = 0; t < min(2, pre_ntile); ++t) pre_issue(t);
  for (int item = c.b; item < l * C.max_rinc_chunks; item += c.G) {
    const int j = item / C.max_rinc_chunks, ch = item - j * C.max_rinc_chunks;
    const int q0 = C.rinc_ptr[n + j] + ch * kLmChunk, q1 = min(C.rinc_ptr[n + j + 1], q0 + kLmChunk);
    if (q0 >= q1) continue;  // uniform per CTA
    cta_sparse_dot(c, q0, q1, C.rinc_k, C.rinc_e, C.rdinv, V, rg0, r, c.sW + c.nth, HP, alpha);
    if (c.tid < r) C.part_pre[((size_t)j * C.max_rinc_chunks + ch) * r + c.tid] = c.sW[c.nth + c.tid];
    __syncthreads();
  }
  {
    // pass 1: Vnew = V + alpha HP, Y = Vnew on the pose rows
    const long long pin0 = (long long)C.pinned_pose_row * r, pin1 = pin0 + r;  // (no pinned row: [-r, 0))
    auto emit = [&](size_t e, double vv) {
      if (HP != nullptr) Vnew[e] = vv;
      Y[e] = ((long long)e >= pin0 && (long long)e < pin1) ? 0.0 : vv;
    };
    if (pre_tiles) {
      for (int t = 0; t < pre_ntile; ++t) {
        chain_tile_wait(c, t & 1);
        const double *sV = stage_base + (t & 1) * pre_tstride, *sH = sV + TRP * r;
        const size_t ra = row0c + (size_t)t * TRP;
        const int ne = (int)(min((size_t)TRP, row1e - ra) * r);
        for (int i = c.tid; i < ne; i += c.nth) {
          double vv = sV[i];
          if (HP != nullptr) vv = fma(alpha, sH[i], vv);
          emit(ra * r + i, vv);
        }
        __syncthreads();
        if (t + 2 < pre_ntile) pre_issue(t + 2);  // into the buffer this barrier released
      }
    } else {
      for (size_t e = row0c * r + c.tid; e < row1e * r; e += c.nth) emit(e, veff(e));
    }
    for (size_t e = row1e * r + c.tid; e < row1c * r; e += c.nth) emit(e, veff(e));
    __syncthreads();  // pass 2 reads this CTA's rows of Y
    // pass 2: range elimination on the translation rows of the CTA's poses, one thread per pose (1 round trip
    // for the list bounds; the fifth of the poses that has ranges pays two more)
    const int x_lo = (int)(row0c / B), x_hi = (int)(row1c / B);
    for (int x = x_lo + c.tid; x < x_hi; x += c.nth) {
      const int q0 = C.rinc_ptr[x], q1 = C.rinc_ptr[x + 1];
      const size_t row = (size_t)x * B + (B - 1);
      if (q1 > q0 && (int)row != C.pinned_pose_row) {
        for (int cc = 0; cc < r; ++cc) {
          double vv = Y[row * r + cc];
          for (int q = q0; q < q1; ++q)
            vv -= C.rinc_e[q] * C.rdinv[C.rinc_k[q]] * veff((rg0 + C.rinc_k[q]) * r + cc);
          Y[row * r + cc] = vv;
        }
      }
    }
  }
  if (HP != nullptr) {  // landmark and range rows of the updated right-hand side
    const size_t nAll = (rg0 + m) * r;
    for (size_t e = np * r + (size_t)gtid; e < nAll; e += (size_t)gsize) Vnew[e] = veff(e);
    V = Vnew;  // complete after the next grid barrier
  }
  ph_end(c, PH_CH_PRE);
  grid_sync(c);
  // ---- the pose chain: forward levels, top solve, backward levels ----
  if (n > 0) {
    ChainStage st;
    st.buf = c.sW;
    st.avail = C.smem_doubles - (int)(c.sW - c.smem);
    // (Running the small levels back to back on one CTA, block barriers instead of grid barriers, measured slower:
    // 36 us for three of them against ~25 us as grid phases.)
    for (int lv = 0; lv < C.nl; ++lv) {
      chain_forward_level_staged<B>(C, c, lv, st);  // the top level is back-substituted in the same phase
      ph_end(c, PH_CH_FWD);
      grid_sync(c);
    }
    for (int lv = C.nl - 2; lv >= 0; --lv) {
      chain_backward_level_staged<B>(C, c, lv, st);
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
  // the pose rows of the post phase (z_P = y_P - W z_L: W is np x l, 65 MB with Y at 100k poses) stream through the
  // two-stage TMA tile pipeline; its first two tiles are requested before the landmark solve and the range rows
  const int post_avail = C.smem_doubles - (int)(c.sW - c.smem) - ((2 * l * r + 1) & ~1);
  const int TRW = min((post_avail / 2) / (l + r), 2048) & ~1;
  double *const tb = c.sW + ((2 * l * r + 1) & ~1);
  const bool post_tiles = TRW >= 2 && row1e > row0c;
  const int post_tstride = TRW * (l + r);
  const int post_ntile = post_tiles ? (int)((row1e - row0c + TRW - 1) / TRW) : 0;
  auto post_issue = [&](int t) {
    double *sWt = tb + (t & 1) * post_tstride, *sYt = sWt + TRW * l;
    const size_t ra = row0c + (size_t)t * TRW;
    const size_t nr = min((size_t)TRW, row1e - ra);
    chain_tile_issue(c, t & 1, sWt, C.W + ra * l, (unsigned)(nr * l * sizeof(double)), sYt, Y + ra * r,
                     (unsigned)(nr * r * sizeof(double)));
  };
  for (int t = 0; t < min(2, post_ntile); ++t) post_issue(t);
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
    // range and landmark rows first (two dependent round trips, few elements), then the pose rows kIlp at a time
    for (unsigned e = (unsigned)(np * r) + (unsigned)gtid; e < nE; e += (unsigned)gsize) {
      const unsigned row = e / ur;
      const int cc = (int)(e - row * ur);
      double out;
      if (row < rg0) {
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
    if (post_tiles) {
      for (int t = 0; t < post_ntile; ++t) {
        chain_tile_wait(c, t & 1);
        const double *sWt = tb + (t & 1) * post_tstride, *sYt = sWt + TRW * l;
        const size_t ra = row0c + (size_t)t * TRW;
        const int nr = (int)min((size_t)TRW, row1e - ra);
        for (int i = c.tid; i < nr * r; i += c.nth) {
          const int lr = i / r, cc = i - lr * r;
          double sacc = sYt[i];
          const double *w = sWt + lr * l;
          for (int j = 0; j < l; ++j) sacc = fma(-w[j], zL[(size_t)j * r + cc], sacc);
          Z[ra * r + i] = ((int)(ra + lr) == C.pinned_pose_row) ? 0.0 : sacc;
        }
        __syncthreads();
        if (t + 2 < post_ntile) post_issue(t + 2);
      }
    } else {
      for (size_t e = row0c * r + c.tid; e < row1e * r; e += c.nth) {
        const size_t row = e / ur;
        Z[e] = zpose(row, (int)(e - row * ur));
      }
    }
    for (size_t e = row1e * r + c.tid; e < row1c * r; e += c.nth) {
      const size_t row = e / ur;
      Z[e] = zpose(row, (int)(e - row * ur));
    }
  }
  ph_end(c, PH_CH_POST);
  grid_sync(c);
}

}  // namespace cora_b200
