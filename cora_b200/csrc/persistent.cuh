// persistent.cuh -- the whole truncated-Newton trust-region solve as ONE persistent cooperative
// kernel (one launch per TNT call; sm_100a, 148 SMs x resident CTAs).
//
//   TNT    libs/Optimization/include/Optimization/Riemannian/TNT.h:242-689
//   STPCG  libs/Optimization/include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426
//   closures f / QM / metric / retract / precon: src/CORA.cpp:52-122
//
// Why: the multi-launch path (solver.cuh) spends ~170 us per CG iteration and ~1 ms per outer
// iteration on launch latency, gated no-op launches and a one-CTA-per-hub-row kernel, against a
// ~47 us HBM floor (profiles/README.md, r01a).  Here every CTA owns the tiles b, b+G, b+2G, ... for
// the whole solve; phases are separated by a grid barrier (one atomic counter in L2), every inner
// product is reduced through per-CTA partials that EVERY CTA sums in the same fixed order after the
// barrier (deterministic, no broadcast needed), and all scalar logic of STPCG and TNT runs
// redundantly and identically in every CTA.  One CG iteration is three phases:
//   A  Hp = Hess[p] (Q p from the block-ELL slice staged in shared memory, neighbours of in-tile
//      poses served from the staged tile, Riemannian epilogue) + <p,Hp>, <Hp,Hp>, <p,p>
//   B  s += alpha p ; r += alpha Hp ; v = proj_Y(M^-1 r) ; <r,v>
//   C  p' = -v + beta p  (double buffered) fused with the landmark hub-row partial sums of Q p',
//      split in chunks over all CTAs, evaluated on the fly as -v[j] + beta p[j]
// Vectors written by other CTAs during the kernel are read with ld.global.cg (L2-coherent); only
// the data matrix goes through the non-coherent path.
#pragma once
#include "ops.cuh"

namespace cora_b200 {

constexpr int kPPart = 8;  // partial sums per CTA and reduction
constexpr int kMaxTilesPerCta = 64;  // tile metadata cached in shared memory for the whole solve

// trace rows in the device trace buffer (TNTResult, TNT.h:168-194)
enum TraceRow { TR_F = 0, TR_G, TR_PG, TR_DELTA, TR_TIME, TR_HNORM, TR_HM, TR_RHO, TR_INNER, TR_ROWS };

struct TntDev {  // results of one persistent TNT call (device -> host)
  double f, gnorm, pgnorm, Delta, elapsed;
  int status, num_outer, n_state, pad;
  long long total_inner, barriers;
  int perm[V_COUNT];  // perm[role] = index of the original buffer now playing that role
  unsigned long long prof_ns[24];  // CTA 0's time per phase kind (PhaseId)
  unsigned int prof_cnt[24];
};
enum PhaseId { PH_HUB = 0, PH_GRAD, PH_HESS, PH_UPDATE, PH_PUPDATE, PH_RETRACT, PH_PRECOND, PH_CGINIT, PH_SYNC, PH_MISC, PH_Q_WAIT, PH_Q_QX, PH_Q_EPI, PH_Q_STORE, PH_CH_PRE, PH_CH_FWD, PH_CH_BWD, PH_CH_BORDER, PH_CH_POST, PH_SMID, PH_REDUCE, PH_COUNT };


__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Geometry of a staged tile in the persistent kernels: UNPADDED, element (local row, column) at lrow * r + c,
// i.e. exactly the global row-major layout -- so a tile's rows (plus one pose block of halo either side) of a
// dense vector are ONE contiguous range and arrive by a single TMA bulk copy.  (kernels.cuh's Geo pads rows
// and poses for its thread-per-pose epilogue; the mappings used here tolerate the occasional 2-way conflict.)
template <int D>
struct PGeo {
  static constexpr int D1 = D + 1;
  static constexpr int PADP = 0;
  int r, RS;
  __device__ __forceinline__ PGeo(int r_) : r(r_), RS(r_) {}
  __device__ __forceinline__ int soff(int lrow, int c) const { return lrow * RS + c; }
  __device__ __forceinline__ int pose_base(int p) const { return p * D1 * RS; }
};

// ---------------------------------------------------------------- async copies ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One staging buffer of the tile pipeline: the tile's slice of the data matrix (block-ELL values and
// columns, spill group pointers / entries) filled by bulk copies, and three tile-row slots of dense
// vectors (padded layout of PGeo<D>) filled by 8-byte cp.async.
struct TileBuf {
  double *sval, *spv, *slam, *slot[3];
  int *scol, *gptr;
  unsigned *spk;
  unsigned long long *mbar;
  int *meta;  // [0] = slots S of the tile
};
struct TileMeta {  // per owned tile, cached in shared memory at kernel start
  long long boff, coff, spoff;
  int S, nsp, lq0, lq1;
};

// Per-CTA context of the persistent kernel.
struct PCtx {
  int b, G, tid, nth, r, nbuf;
  int t0, t1;           // the CTA's contiguous tile range
  long long e0, e1;     // ... and its flat element range [e0, e1) of every N x r vector
  unsigned long long *bar;
  unsigned long long target;
  double *partials;
  int parity;
  int nbar;
  unsigned mpar0, mpar1;  // mbarrier phase parity per buffer
  unsigned long long *lmbar;  // mbarrier of the landmark cache (streaming phases)
  unsigned lm_par;
  // shared memory (everything that lives for the whole solve is kept THERE, not in registers: the
  // hot loops need the register file)
  unsigned long long *prof_ns, *tph;   // [PH_COUNT], [2] written by thread 0 only
  unsigned int *prof_cnt;
  double *sred, *sbc, *sW;
  const TileMeta *tmeta;
  unsigned long long *mbar;  // [2]
  int *meta;                 // [2][4]
  // carve-up of the dynamic shared memory in doubles from `smem`: buffer k of the data-matrix slice
  // starts at qbase + k*qstride, vector slot j of buffer k at vbase + (k*3 + j)*vstride + pstride
  double *smem;
  int qbase, qstride, vbase, vstride, pstride, hpad, nbv, spcap, ncol, TRP, nlam;
  __device__ __forceinline__ TileBuf pick(int buf) const {
    TileBuf B;
    double *q = smem + qbase + (nbuf == 2 ? buf : 0) * qstride;
    B.sval = q;
    B.spv = q + nbv;
    B.slam = nullptr;
    int *qi = (int *)(q + nbv + spcap);
    B.scol = qi;
    B.gptr = qi + ncol;
    B.spk = (unsigned *)(qi + ncol + TRP);
    double *vb = smem + vbase + (nbuf == 2 ? buf : 0) * 3 * vstride + hpad;
    B.slot[0] = vb;
    B.slot[1] = vb + vstride;
    B.slot[2] = vb + 2 * vstride;
    B.mbar = mbar + buf;
    B.meta = meta + 4 * buf;
    return B;
  }
};

// In-kernel phase clocks (thread 0 of every CTA, %globaltimer).  CORA_NO_PHASE_TIMERS compiles them out (A/B of
// their own cost); CORA_STRIP_TIMERS adds the per-strip wait / compute split of the streaming phases.
__device__ __forceinline__ void ph_begin(PCtx &c) {
#ifndef CORA_NO_PHASE_TIMERS
  if (c.tid == 0) c.tph[0] = global_timer_ns();
#endif
}
__device__ __forceinline__ void ph_end(PCtx &c, int id) {
#ifdef CORA_NO_PHASE_TIMERS
  return;
#endif
  if (c.tid == 0) {
    const unsigned long long t = global_timer_ns();
    c.prof_ns[id] += t - c.tph[0];
    c.prof_cnt[id] += 1;
    c.tph[0] = t;
  }
}

__device__ __forceinline__ void sub_begin(PCtx &c) {
#ifdef CORA_STRIP_TIMERS
  if (c.tid == 0) c.tph[1] = global_timer_ns();
#endif
}
__device__ __forceinline__ void sub_end(PCtx &c, int id) {
#ifndef CORA_STRIP_TIMERS
  return;
#endif
  if (c.tid == 0) {
    const unsigned long long t = global_timer_ns();
    c.prof_ns[id] += t - c.tph[1];
    c.prof_cnt[id] += 1;
    c.tph[1] = t;
  }
}

__device__ __forceinline__ void grid_sync(PCtx &c) {
  // global data written with ordinary stores in this phase is read by TMA bulk copies (async proxy) in the next
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  ph_begin(c);
  if (c.tid == 0) {
    c.target += (unsigned long long)c.G;
    // arrive: release-add (MEMBAR.ALL.GPU + REDG; the block barrier above makes it cumulative over the CTA's
    // writes); wait: acquire-load spin (LDG.STRONG.GPU + CCTL.IVALL: the L1 invalidation that makes plain loads
    // of other CTAs' data coherent after the barrier).  No MEMBAR.SC on either side.
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(c.bar), "l"(1ULL) : "memory");
    while (ld_acquire_u64(c.bar) < c.target) { }
  }
  ++c.nbar;
  ph_end(c, PH_SYNC);
  __syncthreads();
}

// Sum K per-thread accumulators over the whole grid.  On return every thread of every CTA holds
// the same totals (summed in a fixed order) and *now_ns = CTA 0's clock at the barrier.
// Partials are stored [k][cta] so that, after the barrier, every WARP re-reduces them on its own: coalesced
// loads (the acquire of the barrier invalidated L1; the eight warps' requests for the same lines merge there),
// a fixed per-lane order and a fixed xor tree -- no block barrier and no shared memory after the grid barrier.
template <int K>
__device__ __forceinline__ void grid_reduce(double (&acc)[K], PCtx &c, unsigned long long *now_ns) {
  static_assert(K <= kPPart, "too many partials");
  ph_begin(c);
  const int lane = c.tid & 31, warp = c.tid >> 5, nw = c.nth >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) c.sred[warp * K + k] = acc[k];
  __syncthreads();
  double *part = c.partials + (size_t)c.parity * ((size_t)c.G * kPPart + 8);
  if (c.tid < K) {
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += c.sred[i * K + c.tid];
    part[(size_t)c.tid * c.G + c.b] = s;
  }
  if (c.tid == 0 && c.b == 0) ((unsigned long long *)part)[(size_t)c.G * kPPart] = global_timer_ns();
  ph_end(c, PH_REDUCE);
  grid_sync(c);
  double t[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (int i = lane; i < c.G; i += 32) s += part[(size_t)k * c.G + i];
    t[k] = s;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
    acc[k] = t[k];
  }
  if (now_ns) *now_ns = ((const unsigned long long *)part)[(size_t)c.G * kPPart];
  ph_end(c, PH_REDUCE);
  c.parity ^= 1;
}

// ------------------------------------------------------------------ hub chunks ----
// Partial sums of the landmark hub rows of Q x for x = a*X + b*V (V may be nullptr), one chunk of
// at most kHubChunk entries per work item: longpart[chunk][row-in-group][column].
template <int D>
__device__ __forceinline__ void hub_phase(const DevLayout &L, PCtx &c, const double *X, double a, const double *V,
                                          double bcoef, double *longpart) {
  constexpr int D1 = D + 1;
  const int r = c.r;
  const int per = c.nth / r;
  const int e = c.tid / r, cc = c.tid - e * r;
  double *sm = c.sW;  // D1 * nth doubles: spans sW and the vector slots that follow it
  ph_begin(c);
  for (int ch = c.b; ch < L.numChunks; ch += c.G) {
    double acc[D1];
#pragma unroll
    for (int q = 0; q < D1; ++q) acc[q] = 0.0;
    if (e < per) {
      const int k0 = L.chunk_beg[ch], k1 = L.chunk_end[ch];
#pragma unroll 4
      for (int k = k0 + e; k < k1; k += per) {
        const unsigned pk = __ldg(L.long_pk + k);
        const int lr = (int)(pk >> 30);
        const size_t o = (size_t)(pk & kColMask) * r + cc;
        double x = a * __ldcg(X + o);
        if (V != nullptr) x = fma(bcoef, __ldcg(V + o), x);
        const double xv = __ldg(L.long_val + k) * x;
#pragma unroll
        for (int q = 0; q < D1; ++q) acc[q] += (lr == q) ? xv : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < D1; ++q) sm[q * c.nth + c.tid] = acc[q];
    __syncthreads();
    if (c.tid < D1 * r) {
      const int q = c.tid / r, c2 = c.tid - q * r;
      double s = 0.0;
      for (int i = 0; i < per; ++i) s += sm[q * c.nth + i * r + c2];
      longpart[((size_t)ch * D1 + q) * r + c2] = s;
    }
  }
  __syncthreads();
  ph_end(c, PH_HUB);
}

// tile geometry without touching global memory
template <int D>
__device__ __forceinline__ TileInfo tile_geom(const DevLayout &L, int t, int r) {
  TileInfo T;
  T.row0 = t * L.TR;
  T.nR = min(L.TR, L.N - T.row0);
  T.nP = max(0, min(L.TP, L.n - t * L.TP));
  T.nS = T.nR - T.nP * (D + 1);
  T.S = 0;
  T.ebase = (long long)T.row0 * r;
  return T;
}

// metadata of tile t: cached in shared memory for the CTA's first kMaxTilesPerCta tiles, read from
// global memory beyond that (very large problems: 1M poses = 74 tiles per CTA)
__device__ __forceinline__ TileMeta tile_meta(const DevLayout &L, const PCtx &c, int t) {
  if (t - c.t0 < kMaxTilesPerCta) return c.tmeta[t - c.t0];
  TileMeta M;
  M.boff = __ldg(L.tile_boff + t); M.coff = __ldg(L.tile_coff + t); M.spoff = __ldg(L.tile_sp_off + t);
  M.S = __ldg(L.tile_slots + t); M.nsp = __ldg(L.tile_sp_cnt + t);
  M.lq0 = __ldg(L.tile_long_ptr + t); M.lq1 = __ldg(L.tile_long_ptr + t + 1);
  return M;
}

// ---------------------------------------------------------------- tile pipeline ----
// Issue the asynchronous loads of tile t into buffer `buf`: NV dense vectors (tile rows, padded
// layout) by cp.async, and -- when NEEDQ -- the tile's data-matrix slice by TMA bulk copies.
template <int D, bool NEEDQ, int NV, bool HALO2 = false>
__device__ __forceinline__ void tile_prefetch(const DevLayout &L, PCtx &c, int t, int buf, const double *v0,
                                              const double *v1, const double *v2, const double *bsrc = nullptr) {
  constexpr int D1 = D + 1;
  if (c.tid != 0) return;  // everything is TMA: one thread arms the mbarrier and issues the bulk copies
  const int r = c.r;
  const TileBuf B = c.pick(buf);
  // dense vectors: tile rows + one pose block of halo either side = one contiguous, 16-byte aligned range
  const long long e_row0 = (long long)t * L.TR * r;
  const int nR = min(L.TR, L.N - t * L.TR);
  const long long e_lo = t > 0 ? e_row0 - c.hpad : 0;
  long long e_hi = min(((long long)t * L.TR + nR + D1) * r, (long long)L.N * r);
  e_hi = (e_hi + 1) & ~1LL;  // the work vectors are allocated with slack beyond N * r
  const unsigned vbytes = (unsigned)(e_hi - e_lo) * 8u;
  unsigned total = NV * vbytes;
  unsigned bq = 0, bc = 0, bg = 0, bp = 0, bv = 0;
  TileMeta M;
  if (NEEDQ) {
    M = tile_meta(L, c, t);
    B.meta[0] = M.S;
    bq = (unsigned)M.S * D1 * D1 * L.TP * 8u; bc = (unsigned)M.S * L.TP * 4u;
    bg = (unsigned)L.TRP * 4u; bp = (unsigned)M.nsp * 4u; bv = (unsigned)M.nsp * 8u;
    total += bq + bc + bg + bp + bv;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to the buffer vs the copies
  mbar_expect_tx(B.mbar, total);
  const long long doff = e_lo - e_row0;  // <= 0: the halo before the tile sits in front of the slot pointer
  if (NV > 0) bulk_g2s(B.slot[0] + doff, v0 + e_lo, vbytes, B.mbar);
  if (NV > 1) bulk_g2s(B.slot[1] + doff, v1 + e_lo, vbytes, B.mbar);
  if (NV > 2) bulk_g2s(B.slot[2] + doff, v2 + e_lo, vbytes, B.mbar);
  if (NEEDQ) {
    if (M.S > 0) {
      bulk_g2s(B.sval, (bsrc != nullptr ? bsrc : L.bval) + M.boff, bq, B.mbar);
      bulk_g2s(B.scol, L.bcol + M.coff, bc, B.mbar);
    }
    bulk_g2s(B.gptr, L.sp_gptr + (size_t)t * L.TRP, bg, B.mbar);
    if (M.nsp > 0) {
      bulk_g2s(B.spk, L.sp_pk + M.spoff, bp, B.mbar);
      bulk_g2s(B.spv, L.sp_val + M.spoff, bv, B.mbar);
    }
  }
}
template <int D, bool NEEDQ, int NV, bool HALO2 = false>
__device__ __forceinline__ void tile_acquire(const DevLayout &L, PCtx &c, int t, int buf, const double *v0,
                                             const double *v1, const double *v2, const double *bsrc = nullptr) {
  if (c.nbuf == 2 && t + 1 < c.t1) tile_prefetch<D, NEEDQ, NV, HALO2>(L, c, t + 1, buf ^ 1, v0, v1, v2, bsrc);
  mbar_wait(c.mbar + buf, buf ? c.mpar1 : c.mpar0);
  if (buf) c.mpar1 ^= 1u; else c.mpar0 ^= 1u;
}
template <int D, bool NEEDQ, int NV, bool HALO2 = false>
__device__ __forceinline__ void tile_release(const DevLayout &L, PCtx &c, int t, int &buf, const double *v0,
                                             const double *v1, const double *v2, const double *bsrc = nullptr) {
  __syncthreads();
  if (c.nbuf == 2) buf ^= 1;
  else if (t + 1 < c.t1) tile_prefetch<D, NEEDQ, NV, HALO2>(L, c, t + 1, 0, v0, v1, v2, bsrc);
}

// Sums of the hub-row chunk partials of the tile's long groups, one warp per (group row, column) with the
// chunks strided over the lanes and a fixed-order xor tree: hub[(q - lq0) * hub_stride + a * r + cc].
// (A landmark row couples to thousands of poses: 33 chunks per row at 100k poses -- summed by one thread
// this was a 16 us straggler on the CTA that owns the landmark rows.)
// rows per long group stored in the hub scratch: d+1 if the tile has pose hubs, 1 if only scalar rows
template <int D>
__device__ __forceinline__ int hub_stride(const DevLayout &L, const TileMeta &M, int r) {
  return (M.lq1 > M.lq0 && L.long_grp[M.lq0] < L.n) ? (D + 1) * r : r;
}

template <int D>
__device__ __forceinline__ void tile_hub_sums(const DevLayout &L, PCtx &c, const TileMeta &M, const double *longpart,
                                              double *hub) {
  constexpr int D1 = D + 1;
  const int r = c.r;
  const int lane = c.tid & 31, warp = c.tid >> 5, nwarps = c.nth >> 5;
  const int hs = hub_stride<D>(L, M, r);
  const int npairs = (M.lq1 - M.lq0) * hs;
  for (int pair = warp; pair < npairs; pair += nwarps) {
    const int q = M.lq0 + pair / hs;
    const int rem = pair - (q - M.lq0) * hs;
    const int a = rem / r, cc = rem - a * r;
    const int c0 = L.long_chunk_ptr[q], c1 = L.long_chunk_ptr[q + 1];
    double sacc = 0.0;
    for (int ch = c0 + lane; ch < c1; ch += 32) sacc += __ldcg(longpart + ((size_t)ch * D1 + a) * r + cc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (lane == 0) hub[pair] = sacc;
  }
}

// sW <- (Q X)[tile]; sX holds the tile rows of X plus one pose block of halo on either side
// (staged), X is the global vector (columns outside the staged window, spill columns).
template <int D>
__device__ __forceinline__ void tile_qx(const DevLayout &L, int t, const TileInfo &T, const PGeo<D> &geo, PCtx &c,
                                        const TileBuf &B, const double *X, const double *sX,
                                        const double *longpart, const double *slam, const double *lamS) {
  auto gx = [&](size_t off) -> double { return __ldcg(X + off); };
  constexpr int D1 = D + 1;
  const int r = c.r, TP = L.TP;
  const int S = B.meta[0];
  const int nPoseItems = T.nP * r;
  const int winLo = max(T.row0 - D1, 0);
  const int winHi = min(min(T.row0 + T.nR + D1, L.N), L.nPoseRows);  // exclusive, pose rows only
  const int pstride = D1 * geo.RS + geo.PADP;
  for (int it = c.tid; it < nPoseItems; it += c.nth) {
    const int p = it / r, cc = it - p * r;
    // spill columns first: their L2 round trip overlaps the block products below
    const int k0 = B.gptr[p], k1 = B.gptr[p + 1];
    double xs0 = 0.0, xs1 = 0.0;
    if (k0 < k1) xs0 = gx((size_t)(B.spk[k0] & kColMask) * r + cc);
    if (k0 + 1 < k1) xs1 = gx((size_t)(B.spk[k0 + 1] & kColMask) * r + cc);
    double acc[D1];
#pragma unroll
    for (int a = 0; a < D1; ++a) acc[a] = 0.0;
    for (int s = 0; s < S; ++s) {
      const int jb = B.scol[s * TP + p];
      double x[D1];
      if (jb >= winLo && jb + D1 <= winHi) {
        const int lp = (jb - T.row0 + D1) / D1 - 1;  // local pose index, -1 and nP are the halo
        const double *xp = sX + lp * pstride + cc;
#pragma unroll
        for (int q = 0; q < D1; ++q) x[q] = xp[q * geo.RS];
      } else {
        const size_t xo = (size_t)jb * r + cc;
#pragma unroll
        for (int q = 0; q < D1; ++q) x[q] = gx(xo + (size_t)q * r);
      }
      const double *bv = B.sval + (size_t)s * D1 * D1 * TP + p;
#pragma unroll
      for (int a = 0; a < D1; ++a)
#pragma unroll
        for (int q = 0; q < D1; ++q) acc[a] = fma(bv[(a * D1 + q) * TP], x[q], acc[a]);
    }
    for (int k = k0; k < k1; ++k) {
      const unsigned pk = B.spk[k];
      const int lr = (int)(pk >> 30);
      const double xg = k == k0 ? xs0 : (k == k0 + 1 ? xs1 : gx((size_t)(pk & kColMask) * r + cc));
      const double xv = B.spv[k] * xg;
#pragma unroll
      for (int a = 0; a < D1; ++a) acc[a] += (lr == a) ? xv : 0.0;
    }
    const int o = geo.pose_base(p) + cc;
#pragma unroll
    for (int a = 0; a < D1; ++a) c.sW[o + a * geo.RS] = acc[a];
  }
  // scalar rows (landmarks, ranges): diagonal + spill.  The gathers of UQ items (2 each) are issued before the first
  // dependent use and before any store to sW (a shared-memory store between the items would pin the order of the
  // next item's shared loads), so a 192-row scalar tile is one L2 round trip instead of two or more.
  const int nSI = T.nS * r;
  constexpr int UQ = 4;
  for (int j0 = c.tid; j0 < nSI; j0 += UQ * c.nth) {
    int k0v[UQ], k1v[UQ], off[UQ];
    double xs0[UQ], xs1[UQ], dgv[UQ], xd[UQ];
#pragma unroll
    for (int q = 0; q < UQ; ++q) {
      const int j = j0 + q * c.nth;
      k0v[q] = 0; k1v[q] = -1; off[q] = 0;
      xs0[q] = 0.0; xs1[q] = 0.0; dgv[q] = 0.0; xd[q] = 0.0;
      if (j < nSI) {
        const int sr = j / r, cc = j - sr * r;
        const int lrow = T.nP * D1 + sr;
        const int sidx = T.row0 + lrow - L.nPoseRows;
        const int u = T.nP + sr;
        const int k0 = B.gptr[u], k1 = B.gptr[u + 1];
        k0v[q] = k0; k1v[q] = k1;
        if (k0 < k1) xs0[q] = gx((size_t)(B.spk[k0] & kColMask) * r + cc);
        if (k0 + 1 < k1) xs1[q] = gx((size_t)(B.spk[k0 + 1] & kColMask) * r + cc);
        dgv[q] = lamS != nullptr ? __ldcg(lamS + sidx) : __ldg(L.sdiag + sidx);  // lamS: diag(Q) - lambda_k
        off[q] = geo.soff(lrow, cc);
        xd[q] = sX[off[q]];
      }
    }
    double accq[UQ];
#pragma unroll
    for (int q = 0; q < UQ; ++q) {
      const int k0 = k0v[q], k1 = k1v[q];
      double acc = dgv[q] * xd[q];
      if (k0 < k1) acc = fma(B.spv[k0], xs0[q], acc);
      if (k0 + 1 < k1) acc = fma(B.spv[k0 + 1], xs1[q], acc);
      if (k0 + 2 < k1) {
        const int cc = (j0 + q * c.nth) % r;
        for (int k = k0 + 2; k < k1; ++k)
          acc = fma(B.spv[k], gx((size_t)(B.spk[k] & kColMask) * r + cc), acc);
      }
      accq[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < UQ; ++q)
      if (j0 + q * c.nth < nSI) c.sW[off[q]] = accq[q];
  }
  __syncthreads();
  const TileMeta M = tile_meta(L, c, t);
  if (M.lq1 > M.lq0) {
    double *hub = B.slot[2];  // free at this point in every mode (the epilogue output is written later)
    tile_hub_sums<D>(L, c, M, longpart, hub);
    __syncthreads();
    const int hs = hub_stride<D>(L, M, r);
    for (int i = c.tid; i < (M.lq1 - M.lq0) * hs; i += c.nth) {
      const int ql = i / hs;
      const int rem = i - ql * hs;
      const int a = rem / r, cc = rem - a * r;
      const int g = L.long_grp[M.lq0 + ql];
      const int nrow = g < L.n ? D1 : 1;
      if (a >= nrow) continue;
      const int lrow0 = (g < L.n ? g * D1 : L.nPoseRows + (g - L.n)) - T.row0;
      c.sW[geo.soff(lrow0 + a, cc)] += hub[i];
    }
    __syncthreads();
  }
}

// Riemannian epilogue over a staged tile with D threads per pose (thread (p, a) owns row a of the
// pose block): W <- proj_Y(W - [CURV] sym(Y G^T) Dd), result written to sOut (a different buffer, so
// that no thread overwrites a row its neighbours still read).  Row form of
// src/CORA_problem.cpp:782-867 / StiefelProduct.cpp:38-55 / ObliqueManifold.cpp:16-27.
// Contains block barriers: every thread of the CTA must call it.
template <int D, bool CURV>
__device__ __forceinline__ void tile_epilogue2(const DevLayout &L, const TileInfo &T, const PGeo<D> &geo, PCtx &c,
                                               const double *sY, const double *sG, const double *sDd, double *sW,
                                               double *sOut, double *lam_out = nullptr, double *lamS_out = nullptr,
                                               const double *sv0 = nullptr) {
  const int r = c.r, RS = geo.RS;
  const int nPU = T.nP * D;
  if (CURV) {
    for (int u = c.tid; u < nPU + T.nS; u += c.nth) {
      if (u < nPU) {
        const int p = u / D, a = u - p * D;
        const int o = geo.pose_base(p);
        const double *y = sY + o, *g = sG + o, *dd = sDd + o;
        double Pr[D], Pc[D];  // P[a][b], P[b][a] with P = Y G^T
#pragma unroll
        for (int b = 0; b < D; ++b) { Pr[b] = 0.0; Pc[b] = 0.0; }
        for (int cc = 0; cc < r; ++cc) {
          const double ya = y[a * RS + cc], ga = g[a * RS + cc];
#pragma unroll
          for (int b = 0; b < D; ++b) {
            Pr[b] = fma(ya, g[b * RS + cc], Pr[b]);
            Pc[b] = fma(y[b * RS + cc], ga, Pc[b]);
          }
        }
#pragma unroll
        for (int b = 0; b < D; ++b) Pr[b] = 0.5 * (Pr[b] + Pc[b]);
        double *w = sW + o + a * RS;
        for (int cc = 0; cc < r; ++cc) {
          double s = w[cc];
#pragma unroll
          for (int b = 0; b < D; ++b) s = fma(-Pr[b], dd[b * RS + cc], s);
          w[cc] = s;
        }
      } else {
        const int lrow = T.nP * (D + 1) + (u - nPU);
        if (T.row0 + lrow >= L.nPoseRows + L.l) {
          const int o = geo.soff(lrow, 0);
          oblique_curvature(sY + o, sG + o, sDd + o, sW + o, r);
        }
      }
    }
    __syncthreads();
  }
  for (int u = c.tid; u < nPU + T.nS; u += c.nth) {
    if (u < nPU) {
      const int p = u / D, a = u - p * D;
      const int o = geo.pose_base(p);
      const double *y = sY + o, *w = sW + o;
      double Pr[D], Pc[D];  // P[a][b], P[b][a] with P = Y W^T
#pragma unroll
      for (int b = 0; b < D; ++b) { Pr[b] = 0.0; Pc[b] = 0.0; }
      for (int cc = 0; cc < r; ++cc) {
        const double ya = y[a * RS + cc], wa = w[a * RS + cc];
#pragma unroll
        for (int b = 0; b < D; ++b) {
          Pr[b] = fma(ya, w[b * RS + cc], Pr[b]);
          Pc[b] = fma(y[b * RS + cc], wa, Pc[b]);
        }
      }
#pragma unroll
      for (int b = 0; b < D; ++b) Pr[b] = 0.5 * (Pr[b] + Pc[b]);
      if (lam_out != nullptr)  // W = Q Y here: row a of Lambda_p = sym(Y_p (QY)_p^T); store (Q - Lambda) diag block
#pragma unroll
        for (int b = 0; b < D; ++b) {
          const int e = (a * (D + 1) + b) * L.TP + p;
          lam_out[e] = sv0[e] - Pr[b];
        }
      double *out = sOut + o + a * RS;
      for (int cc = 0; cc < r; ++cc) {
        double s = w[a * RS + cc];
#pragma unroll
        for (int b = 0; b < D; ++b) s = fma(-Pr[b], y[b * RS + cc], s);
        out[cc] = s;
      }
    } else {
      const int lrow = T.nP * (D + 1) + (u - nPU);
      const int o = geo.soff(lrow, 0);
      if (T.row0 + lrow >= L.nPoseRows + L.l) {
        double s = 0.0;
        for (int cc = 0; cc < r; ++cc) s = fma(sY[o + cc], sW[o + cc], s);
        for (int cc = 0; cc < r; ++cc) sOut[o + cc] = fma(-s, sY[o + cc], sW[o + cc]);
        if (lamS_out != nullptr) lamS_out[T.row0 + lrow - L.nPoseRows] = __ldg(L.sdiag + T.row0 + lrow - L.nPoseRows) - s;
      } else {
        for (int cc = 0; cc < r; ++cc) sOut[o + cc] = sW[o + cc];  // landmark rows: Euclidean
        if (lamS_out != nullptr) lamS_out[T.row0 + lrow - L.nPoseRows] = __ldg(L.sdiag + T.row0 + lrow - L.nPoseRows);
      }
    }
  }
  // translation rows of the poses are Euclidean: copy them through
  for (int p = c.tid; p < T.nP; p += c.nth) {
    const int o = geo.pose_base(p) + D * RS;
    for (int cc = 0; cc < r; ++cc) sOut[o + cc] = sW[o + cc];
  }
  __syncthreads();
}

// --------------------------------------------------------------------- phases ----
// GRAD: out2 = Q X, out = proj_X(Q X); acc[0] += <X, QX>, acc[1] += <grad, grad>
// HESS: out = Hess_Y[X] (G = Q Y);     acc[0] += <X, out>, acc[1] += <out, out>, acc[2] += <X, X>
template <int D, int MODE>
__device__ __forceinline__ void qprod_phase(const DevLayout &L, PCtx &c, const double *X, const double *Y,
                                            double *out, double *out2, const double *longpart, double *lam,
                                            double *lamS, double *acc) {
  // lam: a full copy of the block-ELL values whose diagonal blocks hold Q - Lambda(X); lamS: diag(Q) - lambda_k
  // of the scalar rows.  GRAD (X is the point itself, one slot) WRITES them for its tiles; HESS (X the
  // tangent vector, Y the base point, two slots) streams them instead of Q:
  //       Hess_Y[X] = proj_Y((Q - Lambda(Y)) X)   (src/CORA_problem.cpp:822-867 with the
  //       SymBlockDiagProduct of Y and grad F hoisted out of the CG loop: it does not depend on X)
  constexpr int NV = (MODE == QM_HESS) ? 2 : 1;
  const double *lam_in = (MODE == QM_HESS) ? lam : nullptr;
  const int r = c.r;
  const PGeo<D> geo(r);
  ph_begin(c);
  int buf = 0;
  if (c.t0 < c.t1) tile_prefetch<D, true, NV>(L, c, c.t0, 0, X, Y, nullptr, lam_in);
  for (int t = c.t0; t < c.t1; ++t) {
    sub_begin(c);
    tile_acquire<D, true, NV>(L, c, t, buf, X, Y, nullptr, lam_in);
    sub_end(c, PH_Q_WAIT);
    const TileBuf B = c.pick(buf);
    TileInfo T = tile_geom<D>(L, t, r);
    const int nE = T.nR * r;
    const double *sX = B.slot[0], *sY = B.slot[1];
    double *sO = B.slot[2];
    tile_qx<D>(L, t, T, geo, c, B, X, sX, longpart, nullptr, MODE == QM_HESS ? lamS : nullptr);
    sub_end(c, PH_Q_QX);
    if (MODE == QM_SPMM) {
      for (int le = c.tid; le < nE; le += c.nth) {
        const int lrow = le / r, cc = le - lrow * r;
        out[T.ebase + le] = c.sW[geo.soff(lrow, cc)];
      }
    } else if (MODE == QM_GRAD) {
      for (int le = c.tid; le < nE; le += c.nth) {
        const int lrow = le / r, cc = le - lrow * r;
        const int so = geo.soff(lrow, cc);
        const double w = c.sW[so];
        out2[T.ebase + le] = w;
        acc[0] = fma(sX[so], w, acc[0]);
      }
      tile_epilogue2<D, false>(L, T, geo, c, sX, nullptr, nullptr, c.sW, sO, lam + tile_meta(L, c, t).boff, lamS, B.sval);
      asm volatile("fence.proxy.async.global;" ::: "memory");  // the patched blocks are read by TMA bulk copies later
      for (int le = c.tid; le < nE; le += c.nth) {
        const int lrow = le / r, cc = le - lrow * r;
        const double w = sO[geo.soff(lrow, cc)];
        out[T.ebase + le] = w;
        acc[1] = fma(w, w, acc[1]);
      }
    } else {
      tile_epilogue2<D, false>(L, T, geo, c, sY, nullptr, nullptr, c.sW, sO);
      sub_end(c, PH_Q_EPI);
      for (int le = c.tid; le < nE; le += c.nth) {
        const int lrow = le / r, cc = le - lrow * r;
        const int so = geo.soff(lrow, cc);
        const double w = sO[so], dd = sX[so];
        out[T.ebase + le] = w;
        acc[0] = fma(dd, w, acc[0]);
        acc[1] = fma(w, w, acc[1]);
        acc[2] = fma(dd, dd, acc[2]);
      }
    }
    tile_release<D, true, NV>(L, c, t, buf, X, Y, nullptr, lam_in);
    sub_end(c, PH_Q_STORE);
  }
  ph_end(c, MODE == QM_HESS ? PH_HESS : PH_GRAD);
}



// out = a*X + b*Y over the CTA's contiguous element range (Y may be nullptr); flat and unrolled
__device__ __forceinline__ void axpby_flat(PCtx &c, double a, const double *X, double bcoef, const double *Y,
                                           double *out) {
  ph_begin(c);
#pragma unroll 4
  for (long long e = c.e0 + c.tid; e < c.e1; e += c.nth) {
    double v = a * __ldcg(X + e);
    if (Y != nullptr) v = fma(bcoef, __ldcg(Y + e), v);
    out[e] = v;
  }
  ph_end(c, PH_PUPDATE);
}

// s += alpha p ; p' = -v + beta p   (IterativeSolvers.h:374,420) in one pass over the CTA's elements,
// 16-byte accesses (the CTA's element range starts at a multiple of TR*r, which is even)
__device__ __forceinline__ void cg_pupdate_flat(PCtx &c, double alpha, double beta, double *S, const double *P,
                                                const double *V, double *Pn) {
  ph_begin(c);
  const long long n2 = (c.e1 - c.e0) >> 1;
  const double2 *P2 = reinterpret_cast<const double2 *>(P + c.e0), *V2 = reinterpret_cast<const double2 *>(V + c.e0);
  double2 *S2 = reinterpret_cast<double2 *>(S + c.e0), *Pn2 = reinterpret_cast<double2 *>(Pn + c.e0);
#pragma unroll 4
  for (long long i = c.tid; i < n2; i += c.nth) {
    const double2 p = __ldcg(P2 + i), v = __ldcg(V2 + i);
    double2 s = __ldcg(S2 + i), pn;
    s.x = fma(alpha, p.x, s.x); s.y = fma(alpha, p.y, s.y);
    pn.x = fma(beta, p.x, -v.x); pn.y = fma(beta, p.y, -v.y);
    S2[i] = s;
    Pn2[i] = pn;
  }
  if (c.tid == 0 && ((c.e1 - c.e0) & 1)) {  // odd tail (last CTA only)
    const long long e = c.e1 - 1;
    const double p = __ldcg(P + e);
    S[e] = fma(alpha, p, __ldcg(S + e));
    Pn[e] = fma(beta, p, -__ldcg(V + e));
  }
  ph_end(c, PH_PUPDATE);
}

// STPCG initialisation (IterativeSolvers.h:207-279): S = 0, R = grad, P = -v
__device__ __forceinline__ void cg_init_flat(PCtx &c, const double *Gr, const double *Vv, double *S, double *R,
                                             double *P) {
  ph_begin(c);
#pragma unroll 4
  for (long long e = c.e0 + c.tid; e < c.e1; e += c.nth) {
    S[e] = 0.0;
    R[e] = __ldcg(Gr + e);
    P[e] = Vv != nullptr ? -__ldcg(Vv + e) : 0.0;
  }
  ph_end(c, PH_CGINIT);
}

// acc[0] += <A, B> over the CTA's element range
__device__ __forceinline__ void dot_flat(PCtx &c, const double *A, const double *B, double *acc) {
#pragma unroll 4
  for (long long e = c.e0 + c.tid; e < c.e1; e += c.nth) acc[0] = fma(__ldcg(A + e), __ldcg(B + e), acc[0]);
}

// XP = projectToManifold(X + S) (src/CORA_problem.cpp:905-938); acc[0] += <S,S>, acc[1] += <Gr,S>
template <int D>
__device__ __forceinline__ void retract_phase(const DevLayout &L, PCtx &c, const double *X, const double *S,
                                              const double *Gr, double *XP, double *acc) {
  constexpr int NV = 3;
  const int r = c.r;
  const PGeo<D> geo(r);
  double *sW = c.sW;
  ph_begin(c);
  // S was last written by THIS CTA with ordinary stores (s += alpha p on the way out of STPCG) and no grid barrier
  // lies in between: make those stores visible to the TMA reads below (async proxy reads L2, it does not see the
  // SM's in-flight stores)
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  int buf = 0;
  if (c.t0 < c.t1) tile_prefetch<D, false, NV>(L, c, c.t0, 0, X, S, Gr);
  for (int t = c.t0; t < c.t1; ++t) {
    tile_acquire<D, false, NV>(L, c, t, buf, X, S, Gr);
    const TileBuf B = c.pick(buf);
    const TileInfo T = tile_geom<D>(L, t, r);
    const int nE = T.nR * r;
    for (int le = c.tid; le < nE; le += c.nth) {
      const int lrow = le / r, cc = le - lrow * r;
      const int so = geo.soff(lrow, cc);
      const double s = B.slot[1][so];
      acc[0] = fma(s, s, acc[0]);
      acc[1] = fma(B.slot[2][so], s, acc[1]);
      sW[so] = B.slot[0][so] + s;
    }
    __syncthreads();
    for (int u = c.tid; u < T.nP + T.nS; u += c.nth) {
      if (u < T.nP) {
        stiefel_polar<D>(sW + geo.pose_base(u), r, geo.RS);
      } else {
        const int lrow = T.nP * (D + 1) + (u - T.nP);
        if (T.row0 + lrow >= L.nPoseRows + L.l) {
          double *w = sW + geo.soff(lrow, 0);
          double s = 0.0;
          for (int cc = 0; cc < r; ++cc) s = fma(w[cc], w[cc], s);
          const double inv = 1.0 / sqrt(s);
          for (int cc = 0; cc < r; ++cc) w[cc] *= inv;
        }
      }
    }
    __syncthreads();
    for (int le = c.tid; le < nE; le += c.nth) {
      const int lrow = le / r, cc = le - lrow * r;
      XP[T.ebase + le] = sW[geo.soff(lrow, cc)];
    }
    tile_release<D, false, NV>(L, c, t, buf, X, S, Gr);
  }
  ph_end(c, PH_RETRACT);
}

}  // namespace cora_b200
