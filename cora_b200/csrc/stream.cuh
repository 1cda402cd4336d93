// stream.cuh -- warp-autonomous streaming phases of the persistent TNT kernel (sm_100a).
//
// The tile pipeline of persistent.cuh moves a 192-row tile through wait -> Q x -> epilogue -> store with the
// whole CTA in lock step: a 4.4 us dependent chain per tile that neither tile size nor the synchronisation
// flavour shortens (profiles/README.md, r01c-r01f: 0.51 of HBM in the Hessian phase, 0.43 eligible warps per
// cycle, 30 % of the stalls on block barriers).  Here the unit of work is a STRIP (stream_layout.hpp): one warp
// takes 32 / (d+1) poses (lane = (pose, row of the pose block)) or 32 scalar rows from TMA to TMA on its own:
//   * every warp owns a private ring of NS shared-memory stages with one mbarrier each; lane 0 issues the TMA
//     bulk copies of strip u+NS-1 (the strip's record of the data matrix, its diagonal slot, the rows of the
//     dense operands incl. one pose of halo either side) before the warp starts on strip u -- no block barrier
//     anywhere in the phase, 16 independent pipelines per SM instead of 2;
//   * the rank R is a template parameter: the r columns of a row live in registers, the block row of lane
//     (p, a) is two 16-byte shared loads, the operand block of the column pose (d+1) x R doubles read as
//     16-byte loads that the lanes of a pose share (broadcast);
//   * the Riemannian epilogue (tangent projection, src/CORA_problem.cpp:782-867) needs sym(Y_p W_p^T): lane
//     (p, a) forms column a of Y W^T from its own row of W and gets row a with d-1 shuffles;
//   * results go to a staging row block in the stage and leave with one TMA bulk store per strip.
// Reference semantics exactly as persistent.cuh: dataMatrixProduct src/CORA_problem.cpp:742-757, Riemannian
// gradient :772-780, Hessian-vector product :822-867 (Lambda hoisted), preconditioner closure src/CORA.cpp:89-92,
// STPCG updates IterativeSolvers.h:374-386.
#pragma once
#include "persistent.cuh"

#ifndef CORA_STREAM_STORE
#define CORA_STREAM_STORE 0  // 0: rows staged in the stage + one TMA bulk store per strip; 1: each lane stores its row
#endif

namespace cora_b200 {

struct StreamDev {
  int SP, CP, GP, nPS, nSS, nStrips;
  int nstage;                          // stages per warp ring
  int interleave;                      // 1: strip ranges per CTA, warps take them round-robin; 0: ranges per warp
  int stage_doubles;                   // doubles per stage
  int xw_off, yw_off, dg_off, rc_off, gw_off;  // doubles from the stage base
  int ring_base;                       // doubles from the start of the dynamic shared memory
  int lm_base, lm_rows;                // landmark cache of the CTA (doubles from the start of shared memory; rows)
  const uint4 *info;                   // [nStrips]: {record offset, record size (16-byte units), range window start, rows}
  const unsigned char *rec;
  const double *diagQ;                 // [nPS][128]   diagonal slot of Q
  const double *sdiagP;                // [nSS][32]    diagonal of the scalar rows of Q
  double *diagL[2];                    // Q - Lambda(Y) (current / proposal), same form
  double *sdiagL[2];                   // diag(Q) - lambda_k
  const int4 *warp_strip;              // [grid * warps]: pose strips [x, y) and scalar strips [z, w) of the warp
};

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
// one lane of the (converged) warp; the same lane on every call
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// geometry of one strip's operand windows (elements of an N x R row-major vector)
struct StripWin {
  long long x_lo, x_al;  // first element of the operand window / aligned down to 16 bytes
  long long y_lo, y_al;  // first own element / aligned down
  int x_n, y_n;          // elements copied (multiples of 2)
  int own_n;             // own elements
  int np;                // poses (pose strip) or rows (scalar strip)
  int wp0, wp1;          // poses of the window [wp0, wp1) (pose strips)
};

template <int D, int R>
__device__ __forceinline__ StripWin strip_window(const DevLayout &L, const StreamDev &SD, int u) {
  constexpr int D1 = D + 1;
  StripWin W;
  long long x_hi;
  if (u < SD.nPS) {
    const int p0 = u * SD.SP;
    W.np = min(SD.SP, L.n - p0);
    W.wp0 = max(p0 - 1, 0);
    W.wp1 = min(p0 + W.np + 1, L.n);
    W.x_lo = (long long)W.wp0 * D1 * R;
    x_hi = (long long)W.wp1 * D1 * R;
    W.y_lo = (long long)p0 * D1 * R;
    W.own_n = W.np * D1 * R;
  } else {
    const int k0 = (u - SD.nPS) * kStripScalarRows;
    W.np = min(kStripScalarRows, L.l + L.m - k0);
    W.wp0 = W.wp1 = 0;
    W.x_lo = (long long)(L.nPoseRows + k0) * R;
    x_hi = W.x_lo + (long long)W.np * R;
    W.y_lo = W.x_lo;
    W.own_n = W.np * R;
  }
  W.x_al = W.x_lo & ~1LL;
  W.x_n = (int)(((x_hi + 1) & ~1LL) - W.x_al);
  W.y_al = W.y_lo & ~1LL;
  W.y_n = (int)(((W.y_lo + W.own_n + 1) & ~1LL) - W.y_al);
  return W;
}

// One warp's view of its ring
struct Ring {
  double *base;             // stage 0
  unsigned long long *bar;  // [nstage]
  unsigned par;             // phase parity bit per stage
  // issue-time info (record range, range window) of the warp's strips, 32 at a time: lane j holds that of strip
  // 32 * ro_batch + j of the warp's list.  (A global load per strip inside the strip loop shares a scoreboard with the shared-memory
  // loads of the block products and stalls the first DFMA of every strip for an L2 round trip.)
  uint4 inf;
  int ro_batch;
};

// The strips of one warp: a contiguous range of pose strips, then a contiguous range of scalar strips (every
// CTA gets the same mix of the two kinds: the phase ends when its slowest CTA does).
struct StripList {
  int p0, np, s0, ns, stride;
  // v: pose strips [x, y) and scalar strips [z, w).  stride 1: the range belongs to this warp alone;
  // stride W > 1: the range belongs to the CTA and warp `first` takes every W-th strip of it, so that the warps
  // of a CTA stream neighbouring strips at the same time (DRAM page locality).
  __device__ __forceinline__ StripList(const int4 v, int first, int W) {
    stride = W;
    p0 = v.x + first; np = v.y > p0 ? (v.y - p0 + W - 1) / W : 0;
    s0 = v.z + first; ns = v.w > s0 ? (v.w - s0 + W - 1) / W : 0;
  }
  __device__ __forceinline__ int count() const { return np + ns; }
  __device__ __forceinline__ int at(int k) const { return k < np ? p0 + k * stride : s0 + (k - np) * stride; }
};

// lane 0: arm the stage's mbarrier and issue the bulk copies of strip u.  NV = dense operand windows staged:
// 1 = X window only, 2 = X window + own rows of Y.  dgP / dgS: diagonal slot sources (pose / scalar strips).
template <int D, int R, int NV>
__device__ __forceinline__ void strip_issue(const DevLayout &L, const StreamDev &SD, int u, double *stage,
                                            unsigned long long *bar, const double *X, const double *Y,
                                            const double *dgP, const double *dgS, const uint4 inf) {
  const StripWin W = strip_window<D, R>(L, SD, u);
  const unsigned xb = (unsigned)W.x_n * 8u, yb = (unsigned)W.y_n * 8u, rb = inf.y * 16u;
  const unsigned db = u < SD.nPS ? 1024u : 256u;
  // range rows of X attached to the strip's poses: [g_lo, g_hi) elements, aligned to 16 bytes
  const long long g_lo = ((long long)(L.nPoseRows + (int)inf.z) * R) & ~1LL;
  const long long g_hi = ((long long)(L.nPoseRows + (int)inf.z + (int)inf.w) * R + 1) & ~1LL;
  const unsigned gb = inf.w ? (unsigned)(g_hi - g_lo) * 8u : 0u;
  if (CORA_STREAM_STORE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (1: the stage is only ever read)
  mbar_expect_tx(bar, xb + (NV > 1 ? yb : 0u) + rb + db + gb);
  bulk_g2s(stage + SD.xw_off, X + W.x_al, xb, bar);
  if (NV > 1) bulk_g2s(stage + SD.yw_off, Y + W.y_al, yb, bar);
  bulk_g2s(stage + SD.rc_off, SD.rec + (size_t)inf.x * 16, rb, bar);
  if (u < SD.nPS) bulk_g2s(stage + SD.dg_off, dgP + (size_t)u * 128, db, bar);
  else bulk_g2s(stage + SD.dg_off, dgS + (size_t)(u - SD.nPS) * kStripScalarRows, db, bar);
  if (gb) bulk_g2s(stage + SD.gw_off, X + g_lo, gb, bar);
}

// (d+1) x R operand block from shared or global memory (16-byte loads when the block size is even)
template <int D, int R, bool GLOBAL>
__device__ __forceinline__ void load_block(const double *src, double (&x)[(D + 1) * R]) {
  constexpr int NB = (D + 1) * R;
  if constexpr (NB % 2 == 0) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int i = 0; i < NB / 2; ++i) {
      const double2 v = GLOBAL ? __ldcg(s2 + i) : s2[i];
      x[2 * i] = v.x;
      x[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = GLOBAL ? __ldcg(src + i) : src[i];
  }
}

// Tangent projection of row a of a pose block held one row per lane: on entry w = row a of W (lanes a < D of
// the pose), y = Y_p (D x R, every lane of the pose).  Returns P[b] = sym(Y W^T)[a][b] and w <- w - P Y.
// Every lane of the warp must call (shuffles); lanes with `act` false pass garbage that nobody reads.
template <int D, int R>
__device__ __forceinline__ void lane_tangent(const double (&y)[(D + 1) * R], double (&w)[R], int a, int lane, bool act,
                                             double (&P)[D]) {
  double col[D];  // col[b] = Y_b . W_a = (Y W^T)[b][a]
#pragma unroll
  for (int b = 0; b < D; ++b) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < R; ++c) s = fma(y[b * R + c], w[c], s);
    col[b] = s;
    P[b] = s;
  }
  const bool rot = act && a < D;
  // sym(Y W^T)[a][b] = (col_a[b] + col_b[a]) / 2: the lanes a = i and a = j of a pose swap col[j] <-> col[i]
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j < D; ++j) {
      const double snd = (a == i) ? col[j] : col[i];
      const int src = !rot ? lane : (a == i ? lane + (j - i) : (a == j ? lane - (j - i) : lane));
      const double t = 0.5 * (snd + __shfl_sync(0xffffffffu, snd, src));
      if (a == i) P[j] = t;
      if (a == j) P[i] = t;
    }
  if (rot) {
#pragma unroll
    for (int c = 0; c < R; ++c) {
      double s = w[c];
#pragma unroll
      for (int b = 0; b < D; ++b) s = fma(-P[b], y[b * R + c], s);
      w[c] = s;
    }
  }
}

// Leave the strip's own rows: staged block -> global.  One TMA bulk store when the destination is 16-byte
// aligned and whole, coalesced stores otherwise.  The caller has __syncwarp()ed after writing `stg`.
__device__ __forceinline__ void strip_store(double *dst, const double *stg, int n_el, int lane) {
  if ((((unsigned long long)dst | (unsigned long long)(n_el * 8)) & 15ull) == 0ull) {
    if (elect_one()) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_s2g(dst, stg, (unsigned)n_el * 8u);
      bulk_commit();
    }
  } else {
    for (int i = lane; i < n_el; i += 32) dst[i] = stg[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Q-products over the warp's strips.
//   SPMM: out = Q X
//   GRAD: out2 = Q X, out = proj_X(Q X); writes Q - Lambda(X) to dgLw / sdLw;  acc[0] += <X,QX>, acc[1] += <out,out>
//   HESS: out = proj_Y((Q - Lambda) X) with the diagonal slots streamed from dgL / sdL;
//         acc[0] += <X,out>, acc[1] += <out,out>, acc[2] += <X,X>
template <int D, int R, int MODE>
__device__ __forceinline__ void stream_qprod(const DevLayout &L, const StreamDev &SD, PCtx &c, Ring &rg,
                                             const double *X, const double *Y, double *out, double *out2,
                                             const double *longpart, const double *dgL, const double *sdL,
                                             double *dgLw, double *sdLw, double *acc) {
  constexpr int D1 = D + 1;
  constexpr int NV = (MODE == QM_HESS) ? 2 : 1;
  constexpr int NB = D1 * R;
  const int lane = c.tid & 31, warp = c.tid >> 5;
  const int gw = c.b * (c.nth >> 5) + warp;
  const int nwarp = c.nth >> 5;
  const StripList SLst(SD.interleave ? SD.warp_strip[c.b] : SD.warp_strip[gw], SD.interleave ? warp : 0, SD.interleave ? nwarp : 1);
  const int nk = SLst.count();
  const int NS = SD.nstage;
  const double *dgP = (MODE == QM_HESS) ? dgL : SD.diagQ;
  const double *dgS = (MODE == QM_HESS) ? sdL : SD.sdiagP;
  ph_begin(c);
  // issue-time info of the k-th strip of the warp; every lane calls
  auto strip_info = [&](int k) -> uint4 {
    const int bt = k >> 5;
    if (bt != rg.ro_batch) {
      rg.ro_batch = bt;
      const int kk = (bt << 5) + lane;
      if (kk < nk) rg.inf = __ldg(SD.info + SLst.at(kk));
    }
    uint4 v;
    v.x = __shfl_sync(0xffffffffu, rg.inf.x, k & 31);
    v.y = __shfl_sync(0xffffffffu, rg.inf.y, k & 31);
    v.z = __shfl_sync(0xffffffffu, rg.inf.z, k & 31);
    v.w = __shfl_sync(0xffffffffu, rg.inf.w, k & 31);
    return v;
  };
  // landmark rows of X: one bulk copy per CTA and phase into the landmark cache
  const double *lmb = c.smem + SD.lm_base;
  if (SD.lm_rows > 0) {
    const long long l_lo = ((long long)L.nPoseRows * R) & ~1LL;
    const long long l_hi = ((long long)(L.nPoseRows + SD.lm_rows) * R + 1) & ~1LL;
    lmb += (long long)L.nPoseRows * R - l_lo;
    if (c.tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(c.lmbar, (unsigned)(l_hi - l_lo) * 8u);
      bulk_g2s(c.smem + SD.lm_base, X + l_lo, (unsigned)(l_hi - l_lo) * 8u, c.lmbar);
    }
  }
  bool lm_ready = SD.lm_rows == 0;
  int k_next = 0;  // next strip (position in the warp's list) to issue
  for (int k = 0; k < NS - 1 && k_next < nk; ++k) {
    const uint4 inf = strip_info(k_next);
    if (elect_one())
      strip_issue<D, R, NV>(L, SD, SLst.at(k_next), rg.base + (size_t)k * SD.stage_doubles, rg.bar + k, X, Y, dgP, dgS, inf);
    ++k_next;
  }
  int st = 0;  // stage of strip u
  for (int kcur = 0; kcur < nk; ++kcur) {
    const int u = SLst.at(kcur);
    if (k_next < nk) {
      int sn = st + NS - 1;
      if (sn >= NS) sn -= NS;
      const uint4 inf = strip_info(k_next);
      if (elect_one()) {
        if (CORA_STREAM_STORE == 0) bulk_wait_read0();  // the stage's staging rows may still be the source of a bulk store
        strip_issue<D, R, NV>(L, SD, SLst.at(k_next), rg.base + (size_t)sn * SD.stage_doubles, rg.bar + sn, X, Y, dgP, dgS, inf);
      }
      ++k_next;
    }
    double *stage = rg.base + (size_t)st * SD.stage_doubles;
    sub_begin(c);
    mbar_wait(rg.bar + st, (rg.par >> st) & 1u);
    if (!lm_ready) {
      mbar_wait(c.lmbar, c.lm_par);
      lm_ready = true;
    }
    sub_end(c, PH_Q_WAIT);  // (thread 0 = warp 0 of the CTA only: time this warp waited for its strip)
    rg.par ^= 1u << st;
    const StripWin W = strip_window<D, R>(L, SD, u);
    const double *xw = stage + SD.xw_off;
    double *yw = stage + SD.yw_off;
    const double *dg = stage + SD.dg_off;
    const int *hdr = reinterpret_cast<const int *>(stage + SD.rc_off);
    const int nsp = hdr[1], nlong = hdr[2];  // (hdr[3]: first scalar index of the staged range window)
    double w[R], xo[R];
    bool act;
    if (u < SD.nPS) {
      // ------------------------------------------------------------ pose strip ----
#ifdef CORA_EXP_ONESLOT
      const int S = 1;
      const int *cols = hdr + 4;
      const int Sreal = hdr[0];
#define S_LAYOUT Sreal
#else
      const int S = hdr[0];
      const int *cols = hdr + 4;
#define S_LAYOUT S
#endif
      const int *gptr = cols + S_LAYOUT * SD.CP;  // [36]: per lane (pose, row)
      const int *lq = gptr + 36;
      const unsigned *pk = reinterpret_cast<const unsigned *>(lq + SD.CP);
      const double *val = reinterpret_cast<const double *>(pk + ((nsp + 3) & ~3));
      const double *qv = val + ((nsp + 1) & ~1);
      const int p = lane / D1, a = lane - p * D1;
      act = p < W.np;
      const int pc = act ? p : 0;
      const double *xbase = xw + (W.x_lo - W.x_al);  // element (pose wp0, row 0, column 0)
      // spill entries of this ROW (pointers per lane)
#ifdef CORA_EXP_NOGATHER
      const int k0 = gptr[lane], k1 = k0;
#else
      const int k0 = gptr[lane], k1 = gptr[lane + 1];
#endif
      const double *gwb = stage + SD.gw_off + (((long long)(L.nPoseRows + hdr[3]) * R) & 1LL);
#pragma unroll
      for (int cc = 0; cc < R; ++cc) w[cc] = 0.0;
#ifdef CORA_STREAM_SUBPROF
      sub_end(c, PH_CH_PRE);
#endif
      double d4[4];  // this lane's row of the diagonal block (kept for the Lambda patch)
      const int winLo = W.wp0 * D1, winHi = W.wp1 * D1;
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const double2 *q2 = reinterpret_cast<const double2 *>(s == 0 ? dg : qv + (size_t)(s - 1) * 128);
        const double2 qa = q2[lane], qb = q2[32 + lane];
        const double q[4] = {qa.x, qa.y, qb.x, qb.y};
        if (s == 0) { d4[0] = q[0]; d4[1] = q[1]; d4[2] = q[2]; d4[3] = q[3]; }
        const int jb = cols[s * SD.CP + pc];
        double x[NB];
#ifdef CORA_EXP_NOFAR
        load_block<D, R, false>(xbase + (size_t)(jb - winLo) * R, x);
#else
        if (jb >= winLo && jb + D1 <= winHi) load_block<D, R, false>(xbase + (size_t)(jb - winLo) * R, x);
        else load_block<D, R, true>(X + (size_t)jb * R, x);
#endif
#pragma unroll
        for (int b = 0; b < D1; ++b)
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(q[b], x[b * R + cc], w[cc]);
      }
#ifdef CORA_STREAM_SUBPROF
      sub_end(c, PH_CH_FWD);
#endif
      // couplings outside the block slots: range rows from the staged window, landmark rows from the CTA's cache,
      // anything else (pose-pose ranges, blocks beyond the slots) gathered from L2
      for (int kk = k0; kk < k1; ++kk) {
        const unsigned e = pk[kk], kind = e >> 30, idx = e & kColMask;
        const double v = val[kk];
        if (kind == kSpGlobal) {
          const double *xp = X + (size_t)idx * R;
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(v, __ldcg(xp + cc), w[cc]);
        } else {
          const double *xp = (kind == kSpRange ? gwb : lmb) + idx * R;
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(v, xp[cc], w[cc]);
        }
      }
#ifdef CORA_STREAM_SUBPROF
      sub_end(c, PH_CH_BWD);
#endif
      if (nlong > 0 && act && lq[pc] >= 0) {  // pose hub group: chunk partials of this row
        const int q = lq[pc];
        for (int ch = L.long_chunk_ptr[q]; ch < L.long_chunk_ptr[q + 1]; ++ch)
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] += __ldcg(longpart + ((size_t)ch * D1 + a) * R + cc);
      }
      {
        const double *xop = xbase + (size_t)((u * SD.SP - W.wp0) * D1 + lane) * R;
#pragma unroll
        for (int cc = 0; cc < R; ++cc) xo[cc] = act ? xop[cc] : 0.0;
      }
      if (!act) {
#pragma unroll
        for (int cc = 0; cc < R; ++cc) w[cc] = 0.0;
      }
      if (MODE != QM_SPMM) {
        double y[NB];
        const double *yp = (MODE == QM_HESS) ? yw + (W.y_lo - W.y_al) + (size_t)pc * NB
                                             : xbase + (size_t)(u * SD.SP - W.wp0 + pc) * NB;
        load_block<D, R, false>(yp, y);
        if (MODE == QM_GRAD) {
#pragma unroll
          for (int cc = 0; cc < R; ++cc) acc[0] = fma(xo[cc], w[cc], acc[0]);
          if (CORA_STREAM_STORE == 1) {
            if (act)
#pragma unroll
              for (int cc = 0; cc < R; ++cc) out2[W.y_lo + lane * R + cc] = w[cc];
          } else {
            __syncwarp();  // every lane has read the operand window: its own rows become the staging of Q X
            double *stg2 = const_cast<double *>(xbase) + (size_t)(u * SD.SP - W.wp0) * NB;
            if (act)
#pragma unroll
              for (int cc = 0; cc < R; ++cc) stg2[lane * R + cc] = w[cc];
            __syncwarp();
            strip_store(out2 + W.y_lo, stg2, W.own_n, lane);
          }
        }
        double P[D];
        lane_tangent<D, R>(y, w, a, lane, act, P);
        if (MODE == QM_GRAD && act && a < D) {  // row a of the diagonal block of Q - Lambda
          double2 *o2 = reinterpret_cast<double2 *>(dgLw + (size_t)u * 128);
          o2[lane] = make_double2(d4[0] - P[0], d4[1] - P[1]);
          if constexpr (D == 3) o2[32 + lane] = make_double2(d4[2] - P[D - 1], d4[3]);
        }
      }
    } else {
      // ---------------------------------------------------------- scalar strip ----
      const int *gptr = hdr + 4;
      const int *lq = gptr + 36;
      const unsigned *pk = reinterpret_cast<const unsigned *>(lq + 32);
      const double *val = reinterpret_cast<const double *>(pk + ((nsp + 3) & ~3));
      act = lane < W.np;
      const int lc = act ? lane : 0;
      const int sidx = (u - SD.nPS) * kStripScalarRows + lane;
      const bool is_range = sidx >= L.l;
      const double *xbase = xw + (W.x_lo - W.x_al);
      const int k0 = gptr[lc], k1 = act ? gptr[lc + 1] : k0;
      const double dgv = act ? dg[lane] : 0.0;
#pragma unroll
      for (int cc = 0; cc < R; ++cc) {
        xo[cc] = act ? xbase[lc * R + cc] : 0.0;
        w[cc] = dgv * xo[cc];
      }
      for (int k = k0; k < k1; ++k) {
        const unsigned e = pk[k], kind = e >> 30, idx = e & kColMask;
        const double v = val[k];
        if (kind == kSpLandmark) {
          const double *xp = lmb + idx * R;
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(v, xp[cc], w[cc]);
        } else {
          const double *xp = X + (size_t)idx * R;
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(v, __ldcg(xp + cc), w[cc]);
        }
      }
      if (nlong > 0) {
        // hub rows: the warp sums the chunk partials of each row, chunks strided over the lanes, HB rows per
        // round so that HB * R loads per lane are in flight before the first dependent add / shuffle
        constexpr int HB = 4;
        unsigned mask = __ballot_sync(0xffffffffu, act && lq[lc] >= 0);
        while (mask) {
          int src[HB], c0[HB], c1[HB];
          int cmax = 0;
#pragma unroll
          for (int hb = 0; hb < HB; ++hb) {
            src[hb] = -1; c0[hb] = 0; c1[hb] = 0;
            if (mask) {
              src[hb] = __ffs(mask) - 1;
              mask &= mask - 1;
              const int q = __shfl_sync(0xffffffffu, lq[lc], src[hb]);
              c0[hb] = L.long_chunk_ptr[q];
              c1[hb] = L.long_chunk_ptr[q + 1];
              cmax = max(cmax, c1[hb] - c0[hb]);
            }
          }
          double s[HB][R];
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
#pragma unroll
            for (int cc = 0; cc < R; ++cc) s[hb][cc] = 0.0;
          for (int r0 = 0; r0 < cmax; r0 += 32) {
            double t[HB][R];
#pragma unroll
            for (int hb = 0; hb < HB; ++hb) {
              const int ch = c0[hb] + r0 + lane;
              const bool ok = ch < c1[hb];
#pragma unroll
              for (int cc = 0; cc < R; ++cc) t[hb][cc] = ok ? __ldcg(longpart + (size_t)ch * D1 * R + cc) : 0.0;
            }
#pragma unroll
            for (int hb = 0; hb < HB; ++hb)
#pragma unroll
              for (int cc = 0; cc < R; ++cc) s[hb][cc] += t[hb][cc];
          }
#pragma unroll
          for (int hb = 0; hb < HB; ++hb)
#pragma unroll
            for (int cc = 0; cc < R; ++cc) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s[hb][cc] += __shfl_xor_sync(0xffffffffu, s[hb][cc], o);
              if (lane == src[hb]) w[cc] += s[hb][cc];
            }
        }
      }
      if (MODE != QM_SPMM) {
        if (MODE == QM_GRAD) {
#pragma unroll
          for (int cc = 0; cc < R; ++cc) acc[0] = fma(xo[cc], w[cc], acc[0]);
          if (CORA_STREAM_STORE == 1) {
            if (act)
#pragma unroll
              for (int cc = 0; cc < R; ++cc) out2[W.y_lo + lane * R + cc] = w[cc];
          } else {
            __syncwarp();
            double *stg2 = const_cast<double *>(xbase);
            if (act)
#pragma unroll
              for (int cc = 0; cc < R; ++cc) stg2[lane * R + cc] = w[cc];
            __syncwarp();
            strip_store(out2 + W.y_lo, stg2, W.own_n, lane);
          }
        }
        double sdot = 0.0;
        if (act && is_range) {  // ObliqueManifold.cpp:16-27
          // GRAD: the base point is X itself (held in xo; its staged rows were just overwritten by Q X)
          const double *yp = yw + (W.y_lo - W.y_al) + (size_t)lane * R;
          double yv[R];
#pragma unroll
          for (int cc = 0; cc < R; ++cc) { yv[cc] = (MODE == QM_HESS) ? yp[cc] : xo[cc]; sdot = fma(yv[cc], w[cc], sdot); }
#pragma unroll
          for (int cc = 0; cc < R; ++cc) w[cc] = fma(-sdot, yv[cc], w[cc]);
        }
        if (MODE == QM_GRAD && act) sdLw[sidx] = dgv - sdot;  // diag(Q) - lambda_k (landmark rows: lambda = 0)
      }
    }
#ifdef CORA_STREAM_SUBPROF
    if (__double_as_longlong(w[0]) == 0x7ff8dead00000002LL) asm volatile("trap;");
    sub_end(c, PH_CH_BORDER);
#endif
    // ---- dots, staging, store ----
    if (MODE == QM_GRAD) {
#pragma unroll
      for (int cc = 0; cc < R; ++cc) acc[1] = fma(w[cc], w[cc], acc[1]);
    } else if (MODE == QM_HESS) {
#pragma unroll
      for (int cc = 0; cc < R; ++cc) {
        acc[0] = fma(xo[cc], w[cc], acc[0]);
        acc[1] = fma(w[cc], w[cc], acc[1]);
        acc[2] = fma(xo[cc], xo[cc], acc[2]);
      }
    }
#ifdef CORA_EXP_NOSTORE
    if (__double_as_longlong(w[0]) == 0x7ff8dead00000003LL) out[0] = w[1] + w[2] + w[3] + w[4];
    __syncwarp();
    if (false) {
#else
    if (CORA_STREAM_STORE == 1) {
      if (act)
#pragma unroll
        for (int cc = 0; cc < R; ++cc) out[W.y_lo + lane * R + cc] = w[cc];
      __syncwarp();  // every lane is done reading the stage before it is refilled
#endif
    } else {
      __syncwarp();  // all reads of the Y rows are done: they become the staging of the result
      double *stg = yw + (W.y_lo - W.y_al);
      if (act)
#pragma unroll
        for (int cc = 0; cc < R; ++cc) stg[lane * R + cc] = w[cc];
      __syncwarp();
      strip_store(out + W.y_lo, stg, W.own_n, lane);
    }
    sub_end(c, PH_Q_QX);
    ++st;
    if (st == NS) st = 0;
  }
  sub_begin(c);
  if (elect_one()) bulk_wait0();  // the bulk stores are complete before the CTA arrives at the grid barrier
  if (MODE == QM_GRAD) asm volatile("fence.proxy.async.global;" ::: "memory");  // patched diagonal slots are read by TMA later
  sub_end(c, PH_Q_STORE);
  if (SD.lm_rows > 0) {
    if (!lm_ready) mbar_wait(c.lmbar, c.lm_par);  // (a warp without strips: the copy must land before the cache is reused)
    c.lm_par ^= 1u;
  }
  __syncthreads();
  sub_end(c, PH_Q_EPI);  // wait for the other warps of the CTA
  ph_end(c, MODE == QM_HESS ? PH_HESS : PH_GRAD);
}


// ------------------------------------------------------------------------------------------------
// STPCG update + preconditioner closure over the warp's strips (IterativeSolvers.h:374-386, src/CORA.cpp:89-92):
//   AXPY: Rv += alpha HP          then   V = proj_Y(z),  z = Rv * dinv (zsrc 0) | Rv (1) | Z (2)
//   acc[0] += <Rv, V>, acc[1] += <V, V>
// Stage: [Rv rows | HP or Z rows | Y rows | dinv rows]; Rv' is staged in place, V over the second operand.
template <int D, int R, bool AXPY>
__device__ __forceinline__ void stream_update(const DevLayout &L, const StreamDev &SD, PCtx &c, Ring &rg,
                                              const double *Y, const double *HP, double *Rv, const double *Z,
                                              double *V, double alpha, int zsrc, double *acc) {
  constexpr int D1 = D + 1;
  constexpr int NB = D1 * R;
  constexpr int YW = 32 * R + 2;
  const int lane = c.tid & 31, warp = c.tid >> 5;
  const int gw = c.b * (c.nth >> 5) + warp;
  const int nwarp = c.nth >> 5;
  const StripList SLst(SD.interleave ? SD.warp_strip[c.b] : SD.warp_strip[gw], SD.interleave ? warp : 0, SD.interleave ? nwarp : 1);
  const int nk = SLst.count();
  const int NS = SD.nstage;
  const double *second = AXPY ? HP : (zsrc == 2 ? Z : nullptr);
  ph_begin(c);
  auto issue = [&](int u, int k) {
    const StripWin W = strip_window<D, R>(L, SD, u);
    double *stage = rg.base + (size_t)k * SD.stage_doubles;
    unsigned long long *bar = rg.bar + k;
    const unsigned yb = (unsigned)W.y_n * 8u;
    const int row0 = (int)(W.y_lo / R), nrow = W.own_n / R;
    const int d_al = row0 & ~1;
    const unsigned dbytes = (unsigned)(((row0 + nrow + 1) & ~1) - d_al) * 8u;
    if (CORA_STREAM_STORE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, yb * (second != nullptr ? 3u : 2u) + (zsrc == 0 ? dbytes : 0u));
    bulk_g2s(stage, Rv + W.y_al, yb, bar);
    if (second != nullptr) bulk_g2s(stage + YW, second + W.y_al, yb, bar);
    bulk_g2s(stage + 2 * YW, Y + W.y_al, yb, bar);
    if (zsrc == 0) bulk_g2s(stage + 3 * YW, L.dinv + d_al, dbytes, bar);
  };
  int k_next = 0;
  for (int k = 0; k < NS - 1 && k_next < nk; ++k, ++k_next)
    if (elect_one()) issue(SLst.at(k_next), k);
  int st = 0;
  for (int kcur = 0; kcur < nk; ++kcur) {
    const int u = SLst.at(kcur);
    if (k_next < nk) {
      int sn = st + NS - 1;
      if (sn >= NS) sn -= NS;
      if (elect_one()) {
        if (CORA_STREAM_STORE == 0) bulk_wait_read0();
        issue(SLst.at(k_next), sn);
      }
      ++k_next;
    }
    double *stage = rg.base + (size_t)st * SD.stage_doubles;
    mbar_wait(rg.bar + st, (rg.par >> st) & 1u);
    rg.par ^= 1u << st;
    const StripWin W = strip_window<D, R>(L, SD, u);
    const int sh = (int)(W.y_lo - W.y_al);
    double *sR = stage + sh, *s2 = stage + YW + sh;
    const double *sY = stage + 2 * YW + sh;
    const int row0 = (int)(W.y_lo / R);
    const double *sD = stage + 3 * YW + (row0 & 1);
    const bool pose = u < SD.nPS;
    const int nrow = W.own_n / R;
    const bool act = lane < nrow;
    double rr[R], z[R];
#pragma unroll
    for (int cc = 0; cc < R; ++cc) {
      rr[cc] = act ? sR[lane * R + cc] : 0.0;
      if (AXPY) rr[cc] = fma(alpha, act ? s2[lane * R + cc] : 0.0, rr[cc]);
    }
    if (zsrc == 0) {
      const double dv = act ? sD[lane] : 0.0;
#pragma unroll
      for (int cc = 0; cc < R; ++cc) z[cc] = rr[cc] * dv;
    } else if (zsrc == 1) {
#pragma unroll
      for (int cc = 0; cc < R; ++cc) z[cc] = rr[cc];
    } else {
#pragma unroll
      for (int cc = 0; cc < R; ++cc) z[cc] = act ? s2[lane * R + cc] : 0.0;
    }
    if (pose) {
      const int p = lane / D1, a = lane - p * D1;
      double y[NB];
      load_block<D, R, false>(sY + (size_t)(act ? p : 0) * NB, y);
      double P[D];
      lane_tangent<D, R>(y, z, a, lane, act, P);
    } else {
      const int sidx = (u - SD.nPS) * kStripScalarRows + lane;
      if (act && sidx >= L.l) {
        double yv[R], sdot = 0.0;
#pragma unroll
        for (int cc = 0; cc < R; ++cc) { yv[cc] = sY[lane * R + cc]; sdot = fma(yv[cc], z[cc], sdot); }
#pragma unroll
        for (int cc = 0; cc < R; ++cc) z[cc] = fma(-sdot, yv[cc], z[cc]);
      }
    }
#pragma unroll
    for (int cc = 0; cc < R; ++cc) {
      acc[0] = fma(rr[cc], z[cc], acc[0]);
      acc[1] = fma(z[cc], z[cc], acc[1]);
    }
    if (CORA_STREAM_STORE == 1) {
      if (act) {
#pragma unroll
        for (int cc = 0; cc < R; ++cc) {
          if (AXPY) Rv[W.y_lo + lane * R + cc] = rr[cc];
          V[W.y_lo + lane * R + cc] = z[cc];
        }
      }
      __syncwarp();
    } else {
      __syncwarp();
      if (act) {
#pragma unroll
        for (int cc = 0; cc < R; ++cc) {
          if (AXPY) sR[lane * R + cc] = rr[cc];
          s2[lane * R + cc] = z[cc];
        }
      }
      __syncwarp();
      if (AXPY) strip_store(Rv + W.y_lo, sR, W.own_n, lane);
      strip_store(V + W.y_lo, s2, W.own_n, lane);
    }
    ++st;
    if (st == NS) st = 0;
  }
  if (elect_one()) bulk_wait0();
  __syncthreads();
  ph_end(c, AXPY ? PH_UPDATE : PH_PRECOND);
}


// ------------------------------------------------------------------------------------------------
// STPCG direction update over the warp's strips (IterativeSolvers.h:374,420), one pass:
//   S += alpha P ;  Pn = -V + beta P
// Stage: [S rows | P rows | V rows]; S' is staged in place, Pn over the V rows.
template <int D, int R>
__device__ __forceinline__ void stream_pupdate(const DevLayout &L, const StreamDev &SD, PCtx &c, Ring &rg, double alpha,
                                               double beta, double *S, const double *P, const double *V, double *Pn) {
  constexpr int YW = 32 * R + 2;
  const int lane = c.tid & 31, warp = c.tid >> 5;
  const int gw = c.b * (c.nth >> 5) + warp;
  const int nwarp = c.nth >> 5;
  const StripList SLst(SD.interleave ? SD.warp_strip[c.b] : SD.warp_strip[gw], SD.interleave ? warp : 0, SD.interleave ? nwarp : 1);
  const int nk = SLst.count();
  const int NS = SD.nstage;
  ph_begin(c);
  auto issue = [&](int u, int k) {
    const StripWin W = strip_window<D, R>(L, SD, u);
    double *stage = rg.base + (size_t)k * SD.stage_doubles;
    unsigned long long *bar = rg.bar + k;
    const unsigned yb = (unsigned)W.y_n * 8u;
    if (CORA_STREAM_STORE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, 3u * yb);
    bulk_g2s(stage, S + W.y_al, yb, bar);
    bulk_g2s(stage + YW, P + W.y_al, yb, bar);
    bulk_g2s(stage + 2 * YW, V + W.y_al, yb, bar);
  };
  int k_next = 0;
  for (int k = 0; k < NS - 1 && k_next < nk; ++k, ++k_next)
    if (elect_one()) issue(SLst.at(k_next), k);
  int st = 0;
  for (int kcur = 0; kcur < nk; ++kcur) {
    const int u = SLst.at(kcur);
    if (k_next < nk) {
      int sn = st + NS - 1;
      if (sn >= NS) sn -= NS;
      if (elect_one()) {
        if (CORA_STREAM_STORE == 0) bulk_wait_read0();
        issue(SLst.at(k_next), sn);
      }
      ++k_next;
    }
    double *stage = rg.base + (size_t)st * SD.stage_doubles;
    mbar_wait(rg.bar + st, (rg.par >> st) & 1u);
    rg.par ^= 1u << st;
    const StripWin W = strip_window<D, R>(L, SD, u);
    const int sh = (int)(W.y_lo - W.y_al);
    double *sS = stage + sh, *sV = stage + 2 * YW + sh;
    const double *sP = stage + YW + sh;
    // the strip's own rows are one contiguous block: flat, conflict-free, no row structure needed
    for (int i = lane; i < W.own_n; i += 32) {
      const double pv = sP[i];
      sS[i] = fma(alpha, pv, sS[i]);
      sV[i] = fma(beta, pv, -sV[i]);
    }
    __syncwarp();
    strip_store(S + W.y_lo, sS, W.own_n, lane);
    strip_store(Pn + W.y_lo, sV, W.own_n, lane);
    ++st;
    if (st == NS) st = 0;
  }
  if (elect_one()) bulk_wait0();
  __syncthreads();
  ph_end(c, PH_PUPDATE);
}

}  // namespace cora_b200
