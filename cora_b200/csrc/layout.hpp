// layout.hpp -- host-side construction of the device layout of the data matrix.
//
// Input: the reference's data matrix, CSR int32/f64 in the reference row order
// [d rows per pose | m range rows | n pose translations | l landmark translations]
// (src/CORA_problem.cpp:964-1021; block diagram include/CORA/CORA_problem.h:147-184).
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   * rows are permuted POSE-MAJOR: pose i owns the D1 = d+1 consecutive internal
//     rows [D1*i, D1*i+d) (rotation rows) and D1*i+d (its translation); then the l
//     landmark rows, then the m range rows.  A chain of odometry factors becomes a
//     block-tridiagonal matrix of D1 x D1 blocks.
//   * pose x pose couplings are stored as BLOCK-ELL, sliced per tile of TR rows
//     (TP = TR/D1 poses): values [slot][a][b][pose-in-tile] (pose fastest, so a warp
//     reading consecutive poses is conflict-free / coalesced), one int32 column
//     (internal base row of the column pose) per [slot][pose].  Slot 0 is always the
//     diagonal block.  Slots per tile = max degree in the tile, capped at SMAX.
//   * everything else (couplings to landmark and range rows, and pose blocks beyond
//     SMAX) is the CSR SPILL: one "group" per pose (entries carry their row within
//     the pose in the two top bits of the packed column) or per scalar row; the
//     diagonal of scalar rows lives in its own dense array `sdiag` so that
//     S = Q - Lambda + eta*I only patches slot 0 and sdiag.
//   * groups with more than LONG_GROUP entries (landmark hub rows) are moved to a
//     separate list that a one-CTA-per-group kernel reduces.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace cora_b200 {

constexpr int kSMAX = 8;         // max block-ELL slots per pose
constexpr int kLongGroup = 64;   // spill groups longer than this go to the hub kernel
constexpr uint32_t kColMask = 0x3fffffffu;
constexpr int kHubChunk = 128;   // smallest number of hub-row entries per work item of the persistent kernel
constexpr int kHubItems = 296;   // work items aimed at (one per resident CTA of the persistent kernel on a 148-SM part)

struct HostLayout {
  int d = 0, n = 0, m = 0, l = 0, D1 = 0;
  int64_t N = 0;
  int TR = 0, TP = 0, numTiles = 0;
  int64_t nPoseRows = 0;  // D1*n
  int64_t G = 0;          // groups = n + l + m
  std::vector<int32_t> int2ref, ref2int;
  std::vector<int32_t> range_pos;  // internal position (0..m-1) of reference range row k: ranges sorted by their pose
  std::vector<int32_t> tile_slots;
  std::vector<int64_t> tile_boff, tile_coff;
  std::vector<double> bval;
  std::vector<int32_t> bcol;
  std::vector<double> sdiag;  // l + m
  std::vector<int32_t> grp_ptr;
  std::vector<uint32_t> rem_pk;
  std::vector<double> rem_val;
  std::vector<int32_t> tile_long_ptr, long_grp, long_ptr;
  std::vector<uint32_t> long_pk;
  std::vector<double> long_val;
  std::vector<int32_t> chunk_beg, chunk_end, long_chunk_ptr;
  // per-tile copy of the spill (persistent kernel: one bulk copy per tile): tile-local group
  // pointers [numTiles][TRP], entries padded to a multiple of 4 per tile
  int TRP = 0;
  int64_t max_tile_spill = 0;
  std::vector<int64_t> tile_sp_off;
  std::vector<int32_t> tile_sp_cnt, sp_gptr;
  std::vector<uint32_t> sp_pk;
  std::vector<double> sp_val;
  std::vector<double> diag;  // N, internal order
  int64_t nnz_in = 0, nnz_block = 0, nnz_rem = 0, nnz_long = 0, max_slots = 0;

  inline int64_t ref_to_int(int64_t rr) const {
    const int64_t dn = (int64_t)d * n;
    if (rr < dn) return (rr / d) * D1 + (rr % d);
    if (rr < dn + m) return nPoseRows + l + (range_pos.empty() ? rr - dn : (int64_t)range_pos[rr - dn]);
    const int64_t t = rr - dn - m;
    if (t < n) return t * D1 + d;
    return nPoseRows + (t - n);
  }
};

inline void build_layout(HostLayout &L, int d, int n, int m, int nt, const int32_t *rowptr,
                         const int32_t *col, const double *val, int64_t nnz, int TR) {
  if (d != 2 && d != 3) throw std::invalid_argument("dimension must be 2 or 3");
  if (n < 0 || m < 0 || nt < n) throw std::invalid_argument("inconsistent problem sizes");
  const int D1 = d + 1;
  if (TR % (3 * 4) != 0 || TR <= 0) throw std::invalid_argument("tile rows must be a multiple of 12");
  L.d = d; L.n = n; L.m = m; L.l = nt - n; L.D1 = D1;
  L.N = (int64_t)d * n + m + nt;
  if (L.N >= (int64_t)kColMask) throw std::invalid_argument("problem too large for int32 packed columns");
  if (rowptr[L.N] != nnz || rowptr[0] != 0) throw std::invalid_argument("rowptr[0] != 0 or rowptr[N] != nnz");
  for (int64_t i = 0; i < L.N; ++i)
    if (rowptr[i + 1] < rowptr[i]) throw std::invalid_argument("rowptr is not monotone");
  for (int64_t k = 0; k < nnz; ++k)
    if (col[k] < 0 || col[k] >= L.N) throw std::invalid_argument("column index out of range");
  L.TR = TR; L.TP = TR / D1;
  L.numTiles = (int)((L.N + TR - 1) / TR);
  L.nPoseRows = (int64_t)D1 * n;
  L.G = (int64_t)n + L.l + m;
  L.nnz_in = nnz;
  const int64_t N = L.N;
  {
    // Range rows are ordered by the (first) pose they are attached to, ties in reference order: the range rows
    // coupled to a run of consecutive poses are then one contiguous block of every N x r vector, which the
    // streaming kernels stage with a single bulk copy instead of gathering them row by row from L2.
    const int64_t dn = (int64_t)d * n;
    std::vector<int64_t> key((size_t)m);
    for (int64_t k = 0; k < m; ++k) {
      int64_t best = (int64_t)n + k;  // no pose among the columns (landmark-landmark range): after all poses
      for (int64_t q = rowptr[dn + k]; q < rowptr[dn + k + 1]; ++q) {
        if (col[q] < 0 || col[q] >= N) throw std::invalid_argument("column index out of range");
        const int64_t t = (int64_t)col[q] - dn - m;
        if (t >= 0 && t < n) best = std::min(best, t);
      }
      key[k] = best;
    }
    std::vector<int32_t> order((size_t)m);
    for (int64_t k = 0; k < m; ++k) order[k] = (int32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    L.range_pos.assign((size_t)m, 0);
    for (int64_t pos = 0; pos < m; ++pos) L.range_pos[order[pos]] = (int32_t)pos;
  }
  L.int2ref.resize(N); L.ref2int.resize(N);
  for (int64_t rr = 0; rr < N; ++rr) {
    const int64_t ii = L.ref_to_int(rr);
    L.ref2int[rr] = (int32_t)ii;
    L.int2ref[ii] = (int32_t)rr;
  }
  L.diag.assign(N, 0.0);
  L.sdiag.assign((size_t)L.l + m, 0.0);

  // ---- pass 1: neighbour lists of every pose (sorted, unique, self excluded) ----
  std::vector<int64_t> nbr_ptr((size_t)n + 1, 0);
  std::vector<int32_t> nbr;
  nbr.reserve((size_t)n * 3);
  std::vector<int32_t> tmp;
  auto pose_ref_rows = [&](int i, int a) -> int64_t {  // reference row of internal row D1*i+a
    return a < d ? (int64_t)d * i + a : (int64_t)d * n + m + i;
  };
  for (int i = 0; i < n; ++i) {
    tmp.clear();
    for (int a = 0; a < D1; ++a) {
      const int64_t rr = pose_ref_rows(i, a);
      for (int64_t k = rowptr[rr]; k < rowptr[rr + 1]; ++k) {
        const int64_t ci = L.ref2int[col[k]];
        if (ci < L.nPoseRows) {
          const int32_t j = (int32_t)(ci / D1);
          if (j != i) tmp.push_back(j);
        }
      }
    }
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    nbr.insert(nbr.end(), tmp.begin(), tmp.end());
    nbr_ptr[i + 1] = (int64_t)nbr.size();
  }
  // ---- tiles ----
  L.tile_slots.assign(L.numTiles, 0);
  L.tile_boff.assign((size_t)L.numTiles + 1, 0);
  L.tile_coff.assign((size_t)L.numTiles + 1, 0);
  for (int t = 0; t < L.numTiles; ++t) {
    const int64_t p0 = (int64_t)t * L.TP, p1 = std::min<int64_t>(n, p0 + L.TP);
    int S = 0;
    for (int64_t i = p0; i < p1; ++i) {
      const int deg = 1 + (int)(nbr_ptr[i + 1] - nbr_ptr[i]);
      S = std::max(S, std::min(deg, kSMAX));
    }
    L.tile_slots[t] = S;
    L.max_slots = std::max<int64_t>(L.max_slots, S);
    L.tile_boff[t + 1] = L.tile_boff[t] + (int64_t)S * D1 * D1 * L.TP;
    L.tile_coff[t + 1] = L.tile_coff[t] + (int64_t)S * L.TP;
  }
  L.bval.assign((size_t)L.tile_boff[L.numTiles], 0.0);
  L.bcol.assign((size_t)L.tile_coff[L.numTiles], 0);
  // padding slots point at the pose itself (a valid row) with zero values
  for (int t = 0; t < L.numTiles; ++t) {
    const int S = L.tile_slots[t];
    for (int s = 0; s < S; ++s)
      for (int p = 0; p < L.TP; ++p) {
        const int64_t i = (int64_t)t * L.TP + p;
        int64_t j = i < n ? i : (n > 0 ? n - 1 : 0);
        if (i < n && s >= 1) {
          const int64_t k = nbr_ptr[i] + (s - 1);
          if (k < nbr_ptr[i + 1]) j = nbr[k];
        }
        L.bcol[L.tile_coff[t] + (int64_t)s * L.TP + p] = (int32_t)(j * D1);
      }
  }
  // ---- pass 2: split entries into block-ELL / spill; count spill per group ----
  struct Ent { uint32_t pk; double v; };
  std::vector<int32_t> gcount((size_t)L.G + 1, 0);
  // first count
  auto slot_of = [&](int i, int32_t j) -> int {  // -1: not in ELL
    if (j == i) return 0;
    const int32_t *b = nbr.data() + nbr_ptr[i], *e = nbr.data() + nbr_ptr[i + 1];
    const int32_t *it = std::lower_bound(b, e, j);
    const int s = 1 + (int)(it - b);
    return s < kSMAX ? s : -1;
  };
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<int32_t> fill;
    if (pass == 1) {
      L.grp_ptr.assign((size_t)L.G + 1, 0);
      for (int64_t g = 0; g < L.G; ++g) L.grp_ptr[g + 1] = L.grp_ptr[g] + gcount[g];
      L.rem_pk.assign((size_t)L.grp_ptr[L.G], 0);
      L.rem_val.assign((size_t)L.grp_ptr[L.G], 0.0);
      fill.assign(L.grp_ptr.begin(), L.grp_ptr.end() - 1);
    }
    for (int64_t ii = 0; ii < N; ++ii) {
      const int64_t rr = L.int2ref[ii];
      const bool pose_row = ii < L.nPoseRows;
      const int i = pose_row ? (int)(ii / D1) : -1;
      const int a = pose_row ? (int)(ii % D1) : 0;
      const int64_t g = pose_row ? i : (int64_t)n + (ii - L.nPoseRows);
      for (int64_t k = rowptr[rr]; k < rowptr[rr + 1]; ++k) {
        const int64_t ci = L.ref2int[col[k]];
        const double v = val[k];
        if (pass == 0 && ci == ii) L.diag[ii] += v;
        if (pose_row && ci < L.nPoseRows) {
          const int32_t j = (int32_t)(ci / D1);
          const int b = (int)(ci % D1);
          const int s = slot_of(i, j);
          if (s >= 0) {
            if (pass == 0) {
              const int t = i / L.TP, p = i % L.TP;
              L.bval[L.tile_boff[t] + (((int64_t)s * D1 + a) * D1 + b) * L.TP + p] += v;
              if (v != 0.0) ++L.nnz_block;
            }
            continue;
          }
        }
        if (!pose_row && ci == ii) {
          if (pass == 0) L.sdiag[ii - L.nPoseRows] += v;
          continue;
        }
        if (pass == 0) {
          ++gcount[g];
        } else {
          const int32_t pos = fill[g]++;
          L.rem_pk[pos] = ((uint32_t)a << 30) | (uint32_t)ci;
          L.rem_val[pos] = v;
        }
      }
    }
  }
  // ---- move long groups out of the spill ----
  {
    std::vector<int32_t> new_ptr((size_t)L.G + 1, 0);
    std::vector<uint32_t> new_pk;
    std::vector<double> new_val;
    new_pk.reserve(L.rem_pk.size());
    new_val.reserve(L.rem_val.size());
    L.long_ptr.assign(1, 0);
    L.tile_long_ptr.assign((size_t)L.numTiles + 1, 0);
    for (int64_t g = 0; g < L.G; ++g) {
      const int32_t b = L.grp_ptr[g], e = L.grp_ptr[g + 1];
      if (e - b > kLongGroup) {
        L.long_grp.push_back((int32_t)g);
        L.long_pk.insert(L.long_pk.end(), L.rem_pk.begin() + b, L.rem_pk.begin() + e);
        L.long_val.insert(L.long_val.end(), L.rem_val.begin() + b, L.rem_val.begin() + e);
        L.long_ptr.push_back((int32_t)L.long_pk.size());
        const int64_t row0 = g < n ? g * D1 : L.nPoseRows + (g - n);
        ++L.tile_long_ptr[row0 / L.TR + 1];
      } else {
        new_pk.insert(new_pk.end(), L.rem_pk.begin() + b, L.rem_pk.begin() + e);
        new_val.insert(new_val.end(), L.rem_val.begin() + b, L.rem_val.begin() + e);
      }
      new_ptr[g + 1] = (int32_t)new_pk.size();
    }
    for (int t = 0; t < L.numTiles; ++t) L.tile_long_ptr[t + 1] += L.tile_long_ptr[t];
    L.grp_ptr.swap(new_ptr);
    L.rem_pk.swap(new_pk);
    L.rem_val.swap(new_val);
    L.nnz_rem = (int64_t)L.rem_pk.size();
    L.nnz_long = (int64_t)L.long_pk.size();
    // per-tile spill slices
    L.TRP = L.TR + 4;
    L.tile_sp_off.assign((size_t)L.numTiles + 1, 0);
    L.tile_sp_cnt.assign((size_t)L.numTiles, 0);
    L.sp_gptr.assign((size_t)L.numTiles * L.TRP, 0);
    for (int t = 0; t < L.numTiles; ++t) {
      const int64_t row0 = (int64_t)t * L.TR;
      const int nR = (int)std::min<int64_t>(L.TR, N - row0);
      const int nP = (int)std::max<int64_t>(0, std::min<int64_t>(L.TP, (int64_t)n - (int64_t)t * L.TP));
      const int nS = nR - nP * D1;
      int32_t *gp = L.sp_gptr.data() + (size_t)t * L.TRP;
      int32_t cnt = 0;
      for (int u = 0; u < nP + nS; ++u) {
        const int64_t g = u < nP ? (int64_t)t * L.TP + u : (int64_t)n + (row0 + (int64_t)nP * D1 + (u - nP) - L.nPoseRows);
        gp[u] = cnt;
        for (int32_t k = L.grp_ptr[g]; k < L.grp_ptr[g + 1]; ++k) {
          L.sp_pk.push_back(L.rem_pk[k]);
          L.sp_val.push_back(L.rem_val[k]);
          ++cnt;
        }
      }
      for (int u = nP + nS; u < L.TRP; ++u) gp[u] = cnt;
      L.max_tile_spill = std::max<int64_t>(L.max_tile_spill, cnt);
      while (cnt % 4) { L.sp_pk.push_back(0); L.sp_val.push_back(0.0); ++cnt; }
      L.tile_sp_cnt[t] = cnt;
      L.tile_sp_off[t + 1] = L.tile_sp_off[t] + cnt;
    }
    // hub rows are split in chunks of equal size, about one chunk per resident CTA of the persistent kernel: the
    // chunk partial sums are produced by all CTAs in one round and a hub row is left with few partials to add up
    L.long_chunk_ptr.assign(1, 0);
    int64_t chunk = std::max<int64_t>(kHubChunk, ((int64_t)L.long_pk.size() + kHubItems - 1) / kHubItems);
    for (;; chunk += std::max<int64_t>(1, chunk / 64)) {  // rounding up per row must not push the count past one round
      int64_t cnt = 0;
      for (size_t q = 0; q < L.long_grp.size(); ++q) cnt += std::max<int64_t>(1, (L.long_ptr[q + 1] - L.long_ptr[q] + chunk - 1) / chunk);
      if (cnt <= kHubItems || chunk >= (int64_t)L.long_pk.size()) break;
    }
    for (size_t q = 0; q < L.long_grp.size(); ++q) {
      const int64_t e = L.long_ptr[q + 1] - L.long_ptr[q];
      const int64_t nch = std::max<int64_t>(1, (e + chunk - 1) / chunk), per = (e + nch - 1) / nch;
      for (int64_t k = L.long_ptr[q]; k < L.long_ptr[q + 1]; k += per) {
        L.chunk_beg.push_back((int32_t)k);
        L.chunk_end.push_back((int32_t)std::min<int64_t>(k + per, L.long_ptr[q + 1]));
      }
      L.long_chunk_ptr.push_back((int32_t)L.chunk_beg.size());
    }
  }
}

// Rebuild a reference-ordered CSR from the layout (exact zeros dropped; duplicates
// were summed at build time).  Test hook only -- no product path uses it.
inline void layout_to_csr(const HostLayout &L, std::vector<int32_t> &rowptr,
                          std::vector<int32_t> &col, std::vector<double> &val) {
  struct T { int32_t r, c; double v; };
  std::vector<T> tr;
  const int D1 = L.D1;
  for (int t = 0; t < L.numTiles; ++t) {
    const int S = L.tile_slots[t];
    for (int s = 0; s < S; ++s)
      for (int p = 0; p < L.TP; ++p) {
        const int64_t i = (int64_t)t * L.TP + p;
        if (i >= L.n) continue;
        const int64_t jb = L.bcol[L.tile_coff[t] + (int64_t)s * L.TP + p];
        for (int a = 0; a < D1; ++a)
          for (int b = 0; b < D1; ++b) {
            const double v = L.bval[L.tile_boff[t] + (((int64_t)s * D1 + a) * D1 + b) * L.TP + p];
            if (v != 0.0) tr.push_back({L.int2ref[i * D1 + a], L.int2ref[jb + b], v});
          }
      }
  }
  for (int64_t k = 0; k < (int64_t)L.sdiag.size(); ++k)
    if (L.sdiag[k] != 0.0) {
      const int32_t rr = L.int2ref[L.nPoseRows + k];
      tr.push_back({rr, rr, L.sdiag[k]});
    }
  auto emit = [&](int64_t g, uint32_t pk, double v) {
    if (v == 0.0) return;
    const int a = (int)(pk >> 30);
    const int64_t ci = pk & kColMask;
    const int64_t ri = g < L.n ? g * D1 + a : L.nPoseRows + (g - L.n);
    tr.push_back({L.int2ref[ri], L.int2ref[ci], v});
  };
  for (int64_t g = 0; g < L.G; ++g)
    for (int32_t k = L.grp_ptr[g]; k < L.grp_ptr[g + 1]; ++k) emit(g, L.rem_pk[k], L.rem_val[k]);
  for (size_t q = 0; q < L.long_grp.size(); ++q)
    for (int32_t k = L.long_ptr[q]; k < L.long_ptr[q + 1]; ++k)
      emit(L.long_grp[q], L.long_pk[k], L.long_val[k]);
  std::sort(tr.begin(), tr.end(), [](const T &x, const T &y) { return x.r != y.r ? x.r < y.r : x.c < y.c; });
  rowptr.assign((size_t)L.N + 1, 0);
  col.clear(); val.clear();
  for (size_t k = 0; k < tr.size(); ++k) {
    if (!col.empty() && k > 0 && tr[k].r == tr[k - 1].r && tr[k].c == tr[k - 1].c) {
      val.back() += tr[k].v;
      continue;
    }
    col.push_back(tr[k].c); val.push_back(tr[k].v);
    ++rowptr[tr[k].r + 1];
  }
  for (int64_t i = 0; i < L.N; ++i) rowptr[i + 1] += rowptr[i];
}

}  // namespace cora_b200
