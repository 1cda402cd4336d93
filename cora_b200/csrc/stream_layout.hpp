// stream_layout.hpp -- host-side construction of the STRIP layout of the data matrix: the form the
// warp-autonomous streaming phases of the persistent kernel (stream.cuh) consume.
//
// A strip is the unit one WARP processes on its own: SP = 32 / (d+1) consecutive poses (lane = (pose, row of
// the pose block): 8 poses x 4 rows at d = 3, 10 x 3 at d = 2) or 32 consecutive scalar rows (landmark / range
// rows, lane = row).  Everything a warp needs of the data matrix for one strip is ONE contiguous, 16-byte
// aligned RECORD, so that it arrives with a single TMA bulk copy into the warp's private shared-memory ring:
//
//   pose strip record                                  scalar strip record
//     int  hdr[4]  = {S, nsp, nlong, rw0}                int  hdr[4] = {0, nsp, nlong, 0}
//     int  cols[S][CP]    base row of the column pose    int  gptr[36]   spill pointers (33 used)
//     int  gptr[36]       spill pointers per (pose, row) int  lq[32]     hub (long-group) index or -1
//     int  lq[CP]         hub group of the pose or -1    uint pk[nsp4], double val[nsp2]
//     uint pk[nsp4]       kind << 30 | index: a global row to gather, a row of the strip's staged range
//                         window (the range rows attached to the strip's poses are contiguous: layout.hpp),
//                         or a row of the CTA's landmark cache
//     double val[nsp2]
//     double qv[S-1][2][32][2]   off-diagonal block slots: lane (p, a) reads its row a of the
//                                (d+1) x (d+1) block as two conflict-free 16-byte loads ([half][lane][2])
//
// The DIAGONAL block slot lives outside the record in its own array (same [half][lane][2] form, 1 KB per
// strip): the Hessian product streams Q - Lambda(Y) there (written by the gradient phase once per outer
// iteration, src/CORA_problem.cpp:822-867 with the SymBlockDiagProduct hoisted), the plain product streams Q.
// Same for the diagonal of the scalar rows (sdiag / sdiag - lambda_k, padded to whole strips).
//
// Built from the tile layout (layout.hpp), i.e. from the reference's CSR data matrix
// (src/CORA_problem.cpp:625-712) after the pose-major permutation.
#pragma once
#include "layout.hpp"

namespace cora_b200 {

constexpr int kStripScalarRows = 32;
constexpr int kStreamMaxWarps = 16;  // warps per CTA the ring bookkeeping of the kernels is sized for
constexpr int kStreamMaxStages = 4;
constexpr int kRangeWindow = 8;     // range rows of x staged per pose strip (rows attached to the strip's poses)
constexpr int kLmCacheMax = 128;    // landmark rows of x cached in shared memory per CTA (all of them, or none)
// packed spill entry: kind << 30 | index
constexpr uint32_t kSpGlobal = 0u;  // index = internal row: gathered from global memory (L2)
constexpr uint32_t kSpRange = 1u;   // index = row within the strip's staged range window
constexpr uint32_t kSpLandmark = 2u;  // index = landmark: row of the CTA's landmark cache

struct StreamHost {
  int D1 = 0, SP = 0, CP = 0, GP = 0;
  int nPS = 0, nSS = 0, nStrips = 0;
  int max_rec_bytes = 0;
  int max_slots = 1;
  bool has_pose_hubs = false;
  std::vector<uint32_t> rec_off;  // nStrips + 1, in 16-byte units
  std::vector<uint32_t> info;     // nStrips x 4: {record offset, record size (16-byte units), first scalar index of the
                                  // staged range window, rows of the window}
  bool lm_cache = false;          // landmark entries are encoded as rows of the landmark cache
  std::vector<unsigned char> rec;
  std::vector<double> diagQ;   // nPS x 128: diagonal block slot, [half][lane][2]
  std::vector<double> sdiagP;  // nSS x 32: diagonal of the scalar rows, zero padded
  std::vector<float> cost;     // per strip, for the warp partition
};

inline int strip_poses(int d) { return 32 / (d + 1); }

inline void build_stream_layout(const HostLayout &L, StreamHost &S) {
  const int D1 = L.D1, n = L.n;
  const int SP = strip_poses(L.d);
  S.D1 = D1; S.SP = SP;
  S.CP = (SP + 3) & ~3;
  S.GP = (SP + 1 + 3) & ~3;
  S.nPS = (n + SP - 1) / SP;
  const int nScal = L.l + L.m;
  S.nSS = (nScal + kStripScalarRows - 1) / kStripScalarRows;
  S.nStrips = S.nPS + S.nSS;
  S.rec_off.assign((size_t)S.nStrips + 1, 0);
  S.info.assign((size_t)std::max(S.nStrips, 1) * 4, 0u);
  S.lm_cache = L.l > 0 && L.l <= kLmCacheMax;
  S.rec.clear();
  S.diagQ.assign((size_t)std::max(S.nPS, 1) * 128, 0.0);
  S.sdiagP.assign((size_t)std::max(S.nSS, 1) * kStripScalarRows, 0.0);
  S.cost.assign((size_t)S.nStrips, 1.0f);
  for (int k = 0; k < nScal; ++k) S.sdiagP[k] = L.sdiag[k];
  std::vector<int32_t> long_of((size_t)L.G, -1);
  for (size_t q = 0; q < L.long_grp.size(); ++q) {
    long_of[L.long_grp[q]] = (int32_t)q;
    if (L.long_grp[q] < n) S.has_pose_hubs = true;
  }
  auto bv = [&](int i, int s, int a, int b) -> double {
    const int t = i / L.TP, p = i % L.TP;
    return L.bval[L.tile_boff[t] + (((int64_t)s * D1 + a) * D1 + b) * L.TP + p];
  };
  auto bc = [&](int i, int s) -> int32_t {
    const int t = i / L.TP, p = i % L.TP;
    return L.bcol[L.tile_coff[t] + (int64_t)s * L.TP + p];
  };
  auto append = [&](const void *src, size_t bytes) {
    const unsigned char *b = (const unsigned char *)src;
    S.rec.insert(S.rec.end(), b, b + bytes);
  };
  auto pad16 = [&]() { while (S.rec.size() % 16) S.rec.push_back(0); };
  std::vector<int32_t> ibuf;
  std::vector<uint32_t> pk;
  std::vector<double> val, qv;
  for (int u = 0; u < S.nStrips; ++u) {
    S.rec_off[u] = (uint32_t)(S.rec.size() / 16);
    pk.clear(); val.clear();
    if (u < S.nPS) {
      const int i0 = u * SP, np = std::min(SP, n - i0);
      // slots of the strip = largest degree among its poses (padding slots of a pose point at the pose itself)
      int Smax = 1;
      for (int p = 0; p < np; ++p) {
        const int i = i0 + p, St = L.tile_slots[i / L.TP];
        int deg = 1;
        for (int s = 1; s < St; ++s)
          if (bc(i, s) != i * D1) deg = s + 1;
        Smax = std::max(Smax, deg);
      }
      S.max_slots = std::max(S.max_slots, Smax);
      int nlong = 0;
      std::vector<int32_t> gptr(36, 0), lq((size_t)S.CP, -1), cols((size_t)Smax * S.CP, 0);
      // range window of the strip: the first kRangeWindow range rows among its spill columns
      const int64_t rgrow0 = L.nPoseRows + L.l;
      int64_t rw0 = -1;
      for (int p = 0; p < np; ++p)
        for (int32_t k = L.grp_ptr[i0 + p]; k < L.grp_ptr[i0 + p + 1]; ++k) {
          const int64_t ci = L.rem_pk[k] & kColMask;
          if (ci >= rgrow0 && (rw0 < 0 || ci < rw0)) rw0 = ci;
        }
      int rwn = 0;
      // spill pointers per lane = (pose, row of the pose block); the entries of a pose are stored by row
      for (int p = 0; p < SP; ++p) {
        if (p < np) {
          const int i = i0 + p;
          lq[p] = long_of[i];
          if (lq[p] >= 0) ++nlong;
        }
        for (int a = 0; a < D1; ++a) {
          gptr[p * D1 + a] = (int32_t)pk.size();
          if (p >= np) continue;
          const int i = i0 + p;
          for (int32_t k = L.grp_ptr[i]; k < L.grp_ptr[i + 1]; ++k) {
            if ((int)(L.rem_pk[k] >> 30) != a) continue;
            const int64_t ci = L.rem_pk[k] & kColMask;
            uint32_t e = (kSpGlobal << 30) | (uint32_t)ci;
            if (ci >= rgrow0 && ci - rw0 < kRangeWindow) {
              e = (kSpRange << 30) | (uint32_t)(ci - rw0);
              rwn = std::max(rwn, (int)(ci - rw0) + 1);
            } else if (S.lm_cache && ci >= L.nPoseRows && ci < rgrow0) {
              e = (kSpLandmark << 30) | (uint32_t)(ci - L.nPoseRows);
            }
            pk.push_back(e); val.push_back(L.rem_val[k]);
          }
        }
      }
      S.info[(size_t)u * 4 + 2] = rwn > 0 ? (uint32_t)(rw0 - L.nPoseRows) : 0u;
      S.info[(size_t)u * 4 + 3] = (uint32_t)rwn;
      for (int j = SP * D1; j < 36; ++j) gptr[j] = (int32_t)pk.size();
      const int nsp = (int)pk.size();
      for (int s = 0; s < Smax; ++s)
        for (int p = 0; p < S.CP; ++p) {
          const int i = i0 + std::min(p, np - 1);
          const int St = L.tile_slots[i / L.TP];
          cols[(size_t)s * S.CP + p] = (p < np && s < St) ? bc(i, s) : i * D1;
        }
      const int32_t hdr[4] = {Smax, nsp, nlong, (int32_t)S.info[(size_t)u * 4 + 2]};
      append(hdr, sizeof(hdr));
      append(cols.data(), cols.size() * 4);
      append(gptr.data(), gptr.size() * 4);
      append(lq.data(), lq.size() * 4);
      while (pk.size() % 4) pk.push_back(0);
      while (val.size() % 2) val.push_back(0.0);
      append(pk.data(), pk.size() * 4);
      append(val.data(), val.size() * 8);
      // block values [slot][half][lane][2]; slot 0 goes to diagQ
      qv.assign((size_t)Smax * 128, 0.0);
      for (int s = 0; s < Smax; ++s)
        for (int p = 0; p < np; ++p) {
          const int i = i0 + p;
          if (s >= L.tile_slots[i / L.TP]) continue;
          for (int a = 0; a < D1; ++a)
            for (int b = 0; b < D1; ++b)
              qv[(size_t)s * 128 + (size_t)(b >> 1) * 64 + (size_t)(p * D1 + a) * 2 + (b & 1)] = bv(i, s, a, b);
        }
      std::copy(qv.begin(), qv.begin() + 128, S.diagQ.begin() + (size_t)u * 128);
      if (Smax > 1) append(qv.data() + 128, (size_t)(Smax - 1) * 128 * 8);
      pad16();
      S.cost[u] = 1.0f + 0.01f * (float)nsp + 0.25f * (float)(Smax - 3 > 0 ? Smax - 3 : 0) + 1.0f * (float)nlong;
    } else {
      const int k0 = (u - S.nPS) * kStripScalarRows, ns = std::min(kStripScalarRows, nScal - k0);
      std::vector<int32_t> gptr(36, 0), lq(32, -1);
      int nlong = 0;
      for (int j = 0; j < 32; ++j) {
        gptr[j] = (int32_t)pk.size();
        if (j >= ns) continue;
        const int64_t g = (int64_t)n + k0 + j;
        for (int32_t k = L.grp_ptr[g]; k < L.grp_ptr[g + 1]; ++k) {
          const int64_t ci = L.rem_pk[k] & kColMask;
          uint32_t e = (kSpGlobal << 30) | (uint32_t)ci;
          if (S.lm_cache && ci >= L.nPoseRows && ci < L.nPoseRows + L.l) e = (kSpLandmark << 30) | (uint32_t)(ci - L.nPoseRows);
          pk.push_back(e); val.push_back(L.rem_val[k]);
        }
        lq[j] = long_of[g];
        if (lq[j] >= 0) ++nlong;
      }
      for (int j = 32; j < 36; ++j) gptr[j] = (int32_t)pk.size();
      const int nsp = (int)pk.size();
      const int32_t hdr[4] = {0, nsp, nlong, 0};
      append(hdr, sizeof(hdr));
      append(gptr.data(), gptr.size() * 4);
      append(lq.data(), lq.size() * 4);
      while (pk.size() % 4) pk.push_back(0);
      while (val.size() % 2) val.push_back(0.0);
      append(pk.data(), pk.size() * 4);
      append(val.data(), val.size() * 8);
      pad16();
      S.cost[u] = 1.0f + 0.8f * (float)nlong;
    }
    S.max_rec_bytes = std::max<int>(S.max_rec_bytes, (int)(S.rec.size() - (size_t)S.rec_off[u] * 16));
    S.info[(size_t)u * 4 + 0] = S.rec_off[u];
    S.info[(size_t)u * 4 + 1] = (uint32_t)(S.rec.size() / 16) - S.rec_off[u];
  }
  S.rec_off[S.nStrips] = (uint32_t)(S.rec.size() / 16);
  if (S.rec.size() / 16 >= (size_t)0xffffffffu) throw std::invalid_argument("strip records exceed 64 GB");
}

// Strips of `nw` warps, 4 ints per warp: pose strips [p0, p1) and scalar strips [s0, s1).  The scalar strips are
// spread evenly over ALL warps (gathers from L2 make them latency-bound: concentrated in a few CTAs they -- and
// the CTAs sharing an SM with them -- finish the phase late), the pose strips fill every warp up to the same
// total cost.
inline void partition_strips(const StreamHost &S, int nw, double scalar_weight, std::vector<int32_t> &out) {
  out.assign((size_t)nw * 4, 0);
  double total = 0.0;
  for (int u = 0; u < S.nStrips; ++u) total += u < S.nPS ? S.cost[u] : S.cost[u] * scalar_weight;
  double cum_target = 0.0, cum_done = 0.0;
  int pu = 0;
  for (int w = 0; w < nw; ++w) {
    const int s0 = (int)((int64_t)w * S.nSS / nw), s1 = (int)((int64_t)(w + 1) * S.nSS / nw);
    for (int u = s0; u < s1; ++u) cum_done += S.cost[S.nPS + u] * scalar_weight;
    cum_target = total * (w + 1) / nw;
    const int p0 = pu;
    while (pu < S.nPS && (w == nw - 1 || cum_done + 0.5 * S.cost[pu] <= cum_target)) cum_done += S.cost[pu++];
    out[(size_t)w * 4 + 0] = p0; out[(size_t)w * 4 + 1] = pu;
    out[(size_t)w * 4 + 2] = S.nPS + s0; out[(size_t)w * 4 + 3] = S.nPS + s1;
  }
}

// Rebuild a reference-ordered CSR from the STRIP layout (records + diagonal slots + hub groups of the tile
// layout; exact zeros dropped).  Test hook only -- no product path uses it.
inline void stream_to_csr(const HostLayout &L, const StreamHost &S, std::vector<int32_t> &rowptr,
                          std::vector<int32_t> &col, std::vector<double> &val) {
  struct T { int32_t r, c; double v; };
  std::vector<T> tr;
  const int D1 = L.D1;
  auto emit = [&](int64_t ri, int64_t ci, double v) {
    if (v != 0.0) tr.push_back({L.int2ref[ri], L.int2ref[ci], v});
  };
  auto column = [&](int u, uint32_t e) -> int64_t {  // internal column row of a packed spill entry of strip u
    const uint32_t kind = e >> 30, idx = e & kColMask;
    if (kind == kSpRange) return L.nPoseRows + (int64_t)S.info[(size_t)u * 4 + 2] + idx;
    if (kind == kSpLandmark) return L.nPoseRows + idx;
    return idx;
  };
  for (int u = 0; u < S.nStrips; ++u) {
    const unsigned char *rec = S.rec.data() + (size_t)S.rec_off[u] * 16;
    const int32_t *hdr = (const int32_t *)rec;
    const int nsp = hdr[1];
    if (u < S.nPS) {
      const int Sl = hdr[0];
      const int32_t *cols = hdr + 4, *gptr = cols + (size_t)Sl * S.CP, *lq = gptr + 36;
      const uint32_t *pk = (const uint32_t *)(lq + S.CP);
      const double *sv = (const double *)(pk + ((nsp + 3) & ~3));
      const double *qv = sv + ((nsp + 1) & ~1);
      const int i0 = u * S.SP, np = std::min(S.SP, L.n - i0);
      for (int s = 0; s < Sl; ++s) {
        const double *q = s == 0 ? S.diagQ.data() + (size_t)u * 128 : qv + (size_t)(s - 1) * 128;
        for (int p = 0; p < np; ++p)
          for (int a = 0; a < D1; ++a)
            for (int b = 0; b < D1; ++b)
              emit((int64_t)(i0 + p) * D1 + a, (int64_t)cols[(size_t)s * S.CP + p] + b,
                   q[(size_t)(b >> 1) * 64 + (size_t)(p * D1 + a) * 2 + (b & 1)]);
      }
      for (int j = 0; j < np * D1; ++j)
        for (int k = gptr[j]; k < gptr[j + 1]; ++k)
          emit((int64_t)i0 * D1 + j, column(u, pk[k]), sv[k]);
    } else {
      const int32_t *gptr = hdr + 4, *lq = gptr + 36;
      const uint32_t *pk = (const uint32_t *)(lq + 32);
      const double *sv = (const double *)(pk + ((nsp + 3) & ~3));
      const int k0 = (u - S.nPS) * kStripScalarRows, ns = std::min(kStripScalarRows, L.l + L.m - k0);
      for (int j = 0; j < ns; ++j) {
        const int64_t ri = L.nPoseRows + k0 + j;
        emit(ri, ri, S.sdiagP[(size_t)k0 + j]);
        for (int k = gptr[j]; k < gptr[j + 1]; ++k) emit(ri, column(u, pk[k]), sv[k]);
      }
    }
  }
  for (size_t q = 0; q < L.long_grp.size(); ++q)
    for (int32_t k = L.long_ptr[q]; k < L.long_ptr[q + 1]; ++k) {
      const int64_t g = L.long_grp[q];
      const int64_t ri = g < L.n ? g * D1 + (L.long_pk[k] >> 30) : L.nPoseRows + (g - L.n);
      emit(ri, L.long_pk[k] & kColMask, L.long_val[k]);
    }
  std::sort(tr.begin(), tr.end(), [](const T &x, const T &y) { return x.r != y.r ? x.r < y.r : x.c < y.c; });
  rowptr.assign((size_t)L.N + 1, 0);
  col.clear(); val.clear();
  for (size_t k = 0; k < tr.size(); ++k) {
    if (k > 0 && tr[k].r == tr[k - 1].r && tr[k].c == tr[k - 1].c) {
      val.back() += tr[k].v;
      continue;
    }
    col.push_back(tr[k].c); val.push_back(tr[k].v);
    ++rowptr[tr[k].r + 1];
  }
  for (int64_t i = 0; i < L.N; ++i) rowptr[i + 1] += rowptr[i];
}

}  // namespace cora_b200
