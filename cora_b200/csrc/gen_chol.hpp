// gen_chol.hpp -- general sparse block Cholesky of the reduced POSE system for graphs that are not a single
// odometry chain (loop closures, several robots: TIERS, MR.CLAM).
//
// The reference hands (Q + lambda I)[:-1,:-1] and S + eta I to CHOLMOD (src/CORA_problem.cpp:544-614,
// src/CORA_preconditioners.cpp:16-83, src/CORA_utils.cpp:33-57).  chain_chol.cuh eliminates the range rows
// first and closes the landmark columns with a dense Schur complement; what remains is a symmetric positive
// definite matrix over the n poses with uniform (d+1) x (d+1) blocks whose pattern is the pose graph.  For a
// chain it is block tridiagonal (chain_chol.cuh); this file factors the general pattern:
//
//   ordering   nested dissection by breadth-first level structures (George): a pseudo-peripheral start, the
//              level nearest the middle is the separator, recursion on the connected components, small parts by
//              minimum degree.  Multi-robot graphs are long and thin (robots x time), so separators are a
//              handful of poses and the elimination tree is O(log n) separators high instead of O(n)
//              (minimum degree alone: ~1000 levels on TIERS, nested dissection: ~50).
//   symbolic   elimination tree, postorder, column structures by child merging; clusters = subtrees / tree
//              paths of at most 12 consecutive poses that one warp (or a group of warps) eliminates at once;
//              clusters are levelled by their dependencies: one level = one sweep between two grid barriers of the
//              cooperative solve kernel (gen_chol_dev.cuh).
//   numeric    left-looking block Cholesky, L_vv^-1 kept explicitly, and the inverse of the in-cluster part of L
//              for every cluster (the device eliminates a cluster by two products, never by a serial chain).
//   solve      y_v = L_vv^-1 (b_v - sum_{u<v} L_vu y_u);  x_v = L_vv^-T (y_v - sum_{w>v} L_wv^T x_w)
//              (gen_solve_host: the plain recurrences, the host reference of the device kernel).
//
// Everything here is plain host C++ (the CPU test hook pins it against the oracle's sparse LU); the device
// solve lives in gen_chol_dev.cuh and consumes the arrays of GenSym unchanged.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace cora_b200 {

constexpr int kGenClusterMax = 12;  // poses per cluster

struct GenSym {
  int n = 0, ne = 0;
  std::vector<int32_t> perm, iperm;              // perm[position] = pose, iperm[pose] = position
  std::vector<int32_t> colptr, rowidx;           // strictly-lower pattern of L by column, rows ascending
  std::vector<int32_t> rowptr, colidx, rowslot;  // the same pattern by row (columns ascending) + slot in column storage
  std::vector<int32_t> edge_slot;                // per input edge (i < j): slot of block (max pos, min pos)
  std::vector<uint8_t> edge_tr;                  // 1: that slot holds the transpose of the input block M_ij
  std::vector<int32_t> cl_ptr;                   // clusters of consecutive positions (one warp each)
  std::vector<int32_t> lvl_ptr, lvl_cl;          // clusters grouped by dependency level
  // schedule of the device solves (gen_chol_dev.cuh): per off-diagonal block the cluster-local pose it belongs to
  // (row storage: its row; column storage: its column), per cluster one descriptor the warp loads at once
  std::vector<uint8_t> erow_f, erow_b;           // (bit 7: the other pose of the block is inside the same cluster)
  std::vector<int32_t> col2row;                  // column-storage slot -> index of the same block in row storage
  std::vector<int32_t> desc;                     // 8 ints per cluster: {first position, poses, fwd first block, blocks,
                                                 //  bwd first block, blocks, offset of the cluster inverse (blocks), 0}
  int64_t linv_blocks = 0;                       // total blocks of the per-cluster inverses (sum of poses^2)
  int etree_height = 0;
  int64_t nnzL() const { return (int64_t)rowidx.size(); }
  int levels() const { return (int)lvl_ptr.size() - 1; }
};

namespace gen_detail {

struct Graph {
  int n = 0;
  std::vector<int32_t> xadj, adj;
};

inline Graph build_graph(int n, const std::vector<int32_t> &ei, const std::vector<int32_t> &ej) {
  Graph g;
  g.n = n;
  g.xadj.assign((size_t)n + 1, 0);
  for (size_t e = 0; e < ei.size(); ++e) { ++g.xadj[ei[e] + 1]; ++g.xadj[ej[e] + 1]; }
  for (int i = 0; i < n; ++i) g.xadj[i + 1] += g.xadj[i];
  g.adj.assign((size_t)g.xadj[n], 0);
  std::vector<int32_t> fill(g.xadj.begin(), g.xadj.end() - 1);
  for (size_t e = 0; e < ei.size(); ++e) { g.adj[fill[ei[e]]++] = ej[e]; g.adj[fill[ej[e]]++] = ei[e]; }
  for (int i = 0; i < n; ++i) std::sort(g.adj.begin() + g.xadj[i], g.adj.begin() + g.xadj[i + 1]);
  return g;
}

struct Dissector {
  const Graph &g;
  int leaf;
  std::vector<int32_t> mark, order, queue, lvl_off;
  std::vector<uint8_t> inset;
  int stamp = 0;
  Dissector(const Graph &gr, int leaf_size) : g(gr), leaf(leaf_size), mark((size_t)gr.n, 0), inset((size_t)gr.n, 0) {
    order.reserve((size_t)gr.n);
  }
  // breadth-first level structure from `start` inside `inset`; levels are queue[lvl_off[k] .. lvl_off[k+1])
  void bfs(int start) {
    ++stamp;
    queue.clear(); lvl_off.clear();
    queue.push_back(start); mark[start] = stamp;
    lvl_off.push_back(0);
    size_t head = 0;
    while (head < queue.size()) {
      const size_t end = queue.size();
      lvl_off.push_back((int32_t)end);
      for (; head < end; ++head) {
        const int v = queue[head];
        for (int32_t q = g.xadj[v]; q < g.xadj[v + 1]; ++q) {
          const int u = g.adj[q];
          if (inset[u] && mark[u] != stamp) { mark[u] = stamp; queue.push_back(u); }
        }
      }
    }
    // lvl_off has one entry per level start plus the final end (the loop pushes `end` before scanning a level)
    if (lvl_off.back() != (int32_t)queue.size()) lvl_off.push_back((int32_t)queue.size());
  }
  int num_levels() const { return (int)lvl_off.size() - 1; }
  int degree_in(int v) const {
    int dg = 0;
    for (int32_t q = g.xadj[v]; q < g.xadj[v + 1]; ++q) dg += inset[g.adj[q]] ? 1 : 0;
    return dg;
  }
  // minimum degree on the subgraph induced by `nodes` (small parts; couplings to separators ignored)
  void min_degree(const std::vector<int32_t> &nodes) {
    const int k = (int)nodes.size();
    if (k == 0) return;
    std::vector<int32_t> loc((size_t)k);
    std::vector<std::vector<int32_t>> a((size_t)k);
    ++stamp;
    for (int i = 0; i < k; ++i) { mark[nodes[i]] = stamp; }
    // local index through a sorted copy
    std::vector<int32_t> sorted(nodes);
    std::sort(sorted.begin(), sorted.end());
    auto local = [&](int v) { return (int)(std::lower_bound(sorted.begin(), sorted.end(), v) - sorted.begin()); };
    for (int i = 0; i < k; ++i) {
      const int v = sorted[i];
      for (int32_t q = g.xadj[v]; q < g.xadj[v + 1]; ++q)
        if (mark[g.adj[q]] == stamp) a[i].push_back(local(g.adj[q]));
    }
    std::vector<uint8_t> gone((size_t)k, 0);
    std::vector<int32_t> merged;
    for (int step = 0; step < k; ++step) {
      int best = -1;
      for (int i = 0; i < k; ++i)
        if (!gone[i] && (best < 0 || a[i].size() < a[best].size())) best = i;
      gone[best] = 1;
      order.push_back(sorted[best]);
      const std::vector<int32_t> nb = a[best];
      for (int u : nb) {
        merged.clear();
        std::set_union(a[u].begin(), a[u].end(), nb.begin(), nb.end(), std::back_inserter(merged));
        merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int x) { return x == u || x == best; }), merged.end());
        a[u] = merged;
      }
      a[best].clear();
    }
    (void)loc;
  }
  void dissect(std::vector<int32_t> nodes) {
    if ((int)nodes.size() <= leaf) { min_degree(nodes); return; }
    for (int v : nodes) inset[v] = 1;
    // connected components
    std::vector<std::vector<int32_t>> comps;
    {
      const int base = stamp + 1;
      std::vector<int32_t> seen_from;  // stamps used by the component sweeps are base, base+1, ...
      for (int v : nodes) {
        if (mark[v] >= base && mark[v] <= stamp) continue;
        bfs(v);
        comps.emplace_back(queue.begin(), queue.end());
      }
    }
    if (comps.size() > 1) {
      for (int v : nodes) inset[v] = 0;
      for (auto &c : comps) dissect(std::move(c));
      return;
    }
    // pseudo-peripheral start: repeat from a minimum-degree node of the last level
    int s = nodes[0];
    for (int it = 0; it < 4; ++it) {
      bfs(s);
      const int nl = num_levels();
      int s2 = queue[lvl_off[nl - 1]], bd = degree_in(s2);
      for (int32_t q = lvl_off[nl - 1]; q < lvl_off[nl]; ++q) {
        const int dg = degree_in(queue[q]);
        if (dg < bd) { bd = dg; s2 = queue[q]; }
      }
      if (s2 == s) break;
      s = s2;
    }
    bfs(s);
    const int nl = num_levels();
    if (nl < 3) {  // a blob: no level separates it
      for (int v : nodes) inset[v] = 0;
      min_degree(nodes);
      return;
    }
    const double tot = (double)nodes.size();
    int bj = -1;
    double bkey0 = 0, bkey1 = 0, bkey2 = 0;
    for (int j = 1; j < nl - 1; ++j) {
      const double sz = lvl_off[j + 1] - lvl_off[j];
      const double frac = (lvl_off[j] + 0.5 * sz) / tot;
      const bool mid = frac >= 0.3 && frac <= 0.7;
      const double k0 = mid ? 0 : 1, k1 = mid ? sz : std::fabs(frac - 0.5), k2 = std::fabs(frac - 0.5);
      if (bj < 0 || k0 < bkey0 || (k0 == bkey0 && (k1 < bkey1 || (k1 == bkey1 && k2 < bkey2)))) {
        bj = j; bkey0 = k0; bkey1 = k1; bkey2 = k2;
      }
    }
    std::vector<int32_t> A(queue.begin(), queue.begin() + lvl_off[bj]);
    std::vector<int32_t> S(queue.begin() + lvl_off[bj], queue.begin() + lvl_off[bj + 1]);
    std::vector<int32_t> Bp(queue.begin() + lvl_off[bj + 1], queue.end());
    for (int v : nodes) inset[v] = 0;
    nodes.clear(); nodes.shrink_to_fit();
    dissect(std::move(A));
    dissect(std::move(Bp));
    for (int v : S) order.push_back(v);
  }
};

// elimination tree of the graph under the order `perm` (Liu's algorithm with path compression)
inline void etree(const Graph &g, const std::vector<int32_t> &perm, const std::vector<int32_t> &iperm,
                  std::vector<int32_t> &parent) {
  const int n = g.n;
  parent.assign((size_t)n, -1);
  std::vector<int32_t> anc((size_t)n, -1);
  for (int v = 0; v < n; ++v) {
    const int i = perm[v];
    for (int32_t q = g.xadj[i]; q < g.xadj[i + 1]; ++q) {
      int u = iperm[g.adj[q]];
      while (u != -1 && u < v) {
        const int next = anc[u];
        anc[u] = v;
        if (next == -1) parent[u] = v;
        u = next;
      }
    }
  }
}

}  // namespace gen_detail

// Ordering + symbolic factorisation + clustering.  Edges are pose pairs (i < j); duplicates are allowed.
inline void gen_symbolic(GenSym &S, int n, const std::vector<int32_t> &ei, const std::vector<int32_t> &ej,
                         int leaf_size = 12, int cluster_max = kGenClusterMax) {
  using namespace gen_detail;
  S.n = n; S.ne = (int)ei.size();
  const Graph g = build_graph(n, ei, ej);
  // ---- nested dissection order ----
  std::vector<int32_t> perm;
  {
    Dissector D(g, leaf_size);
    std::vector<int32_t> all((size_t)n);
    std::iota(all.begin(), all.end(), 0);
    D.dissect(std::move(all));
    perm.swap(D.order);
  }
  std::vector<int32_t> iperm((size_t)n);
  for (int p = 0; p < n; ++p) iperm[perm[p]] = p;
  // ---- elimination tree, postorder (children in ascending order), composed permutation ----
  std::vector<int32_t> parent;
  etree(g, perm, iperm, parent);
  {
    std::vector<int32_t> head((size_t)n, -1), next((size_t)n, -1), post;
    post.reserve((size_t)n);
    for (int v = n - 1; v >= 0; --v)
      if (parent[v] >= 0) { next[v] = head[parent[v]]; head[parent[v]] = v; }
    std::vector<int32_t> stack;
    for (int root = 0; root < n; ++root) {
      if (parent[root] >= 0) continue;
      stack.push_back(root);
      while (!stack.empty()) {
        const int v = stack.back();
        const int c = head[v];
        if (c >= 0) { head[v] = next[c]; stack.push_back(c); }
        else { post.push_back(v); stack.pop_back(); }
      }
    }
    std::vector<int32_t> perm2((size_t)n);
    for (int p = 0; p < n; ++p) perm2[p] = perm[post[p]];
    perm.swap(perm2);
    for (int p = 0; p < n; ++p) iperm[perm[p]] = p;
    etree(g, perm, iperm, parent);
  }
  S.perm = perm; S.iperm = iperm;
  // ---- column structures by child merging ----
  std::vector<std::vector<int32_t>> cols((size_t)n);
  {
    std::vector<int32_t> flag((size_t)n, -1);
    std::vector<std::vector<int32_t>> kids((size_t)n);
    for (int v = 0; v < n; ++v) if (parent[v] >= 0) kids[parent[v]].push_back(v);
    for (int v = 0; v < n; ++v) {
      std::vector<int32_t> &c = cols[v];
      flag[v] = v;
      const int i = perm[v];
      for (int32_t q = g.xadj[i]; q < g.xadj[i + 1]; ++q) {
        const int u = iperm[g.adj[q]];
        if (u > v && flag[u] != v) { flag[u] = v; c.push_back(u); }
      }
      for (int k : kids[v])
        for (int u : cols[k])
          if (u > v && flag[u] != v) { flag[u] = v; c.push_back(u); }
      std::sort(c.begin(), c.end());
    }
  }
  S.colptr.assign((size_t)n + 1, 0);
  for (int v = 0; v < n; ++v) S.colptr[v + 1] = S.colptr[v] + (int32_t)cols[v].size();
  S.rowidx.assign((size_t)S.colptr[n], 0);
  for (int v = 0; v < n; ++v) std::copy(cols[v].begin(), cols[v].end(), S.rowidx.begin() + S.colptr[v]);
  // by row
  S.rowptr.assign((size_t)n + 1, 0);
  for (int32_t w : S.rowidx) ++S.rowptr[w + 1];
  for (int v = 0; v < n; ++v) S.rowptr[v + 1] += S.rowptr[v];
  S.colidx.assign(S.rowidx.size(), 0);
  S.rowslot.assign(S.rowidx.size(), 0);
  {
    std::vector<int32_t> fill(S.rowptr.begin(), S.rowptr.end() - 1);
    for (int v = 0; v < n; ++v)
      for (int32_t q = S.colptr[v]; q < S.colptr[v + 1]; ++q) {
        const int w = S.rowidx[q];
        S.colidx[fill[w]] = v; S.rowslot[fill[w]] = q; ++fill[w];
      }
  }
  // ---- input edges -> slots ----
  S.edge_slot.assign(ei.size(), -1);
  S.edge_tr.assign(ei.size(), 0);
  for (size_t e = 0; e < ei.size(); ++e) {
    const int pi = iperm[ei[e]], pj = iperm[ej[e]];
    const int c = std::min(pi, pj), w = std::max(pi, pj);
    const auto b = S.rowidx.begin() + S.colptr[c], en = S.rowidx.begin() + S.colptr[c + 1];
    const auto it = std::lower_bound(b, en, w);
    S.edge_slot[e] = (int32_t)(it - S.rowidx.begin());
    // stored block is (row w, column c) of the permuted matrix; the input block is M_ij = (rows i, columns j)
    S.edge_tr[e] = (pi < pj) ? 1 : 0;  // pi < pj: row = j, column = i -> M_ji = M_ij^T
  }
  // ---- clusters: whole subtrees of at most cluster_max poses; tree paths above them ----
  std::vector<int32_t> size((size_t)n, 1), height((size_t)n, 0);
  for (int v = 0; v < n; ++v)
    if (parent[v] >= 0) {
      size[parent[v]] += size[v];
      height[parent[v]] = std::max(height[parent[v]], height[v] + 1);
    }
  S.etree_height = n ? *std::max_element(height.begin(), height.end()) + 1 : 0;
  std::vector<int32_t> cl((size_t)n, -1);
  S.cl_ptr.clear();
  S.cl_ptr.push_back(0);
  {
    int v = 0;
    while (v < n) {
      // the largest subtree that ends at a position >= v, starts at v and has at most cluster_max poses
      // (postorder: the subtree of w is [w - size[w] + 1, w])
      int w = v;  // v is a leaf of the postorder here or an upper node
      if (size[v] == 1) {
        while (parent[w] >= 0 && parent[w] - size[parent[w]] + 1 == v && size[parent[w]] <= cluster_max) w = parent[w];
        // [v, w] is the subtree of w
      } else {
        // upper node: extend along the tree path while the next position is the parent of the current one
        int len = 1;
        while (w + 1 < n && parent[w] == w + 1 && len < cluster_max) { ++w; ++len; }
      }
      const int id = (int)S.cl_ptr.size() - 1;
      for (int q = v; q <= w; ++q) cl[q] = id;
      S.cl_ptr.push_back(w + 1);
      v = w + 1;
    }
  }
  const int nc = (int)S.cl_ptr.size() - 1;
  std::vector<int32_t> clevel((size_t)nc, 0);
  for (int v = 0; v < n; ++v)
    for (int32_t q = S.rowptr[v]; q < S.rowptr[v + 1]; ++q) {
      const int u = S.colidx[q];
      if (cl[u] != cl[v]) clevel[cl[v]] = std::max(clevel[cl[v]], clevel[cl[u]] + 1);
    }
  // (clusters are numbered in position order and dependencies point to lower positions, so one ascending
  //  pass over v sees every clevel[cl[u]] final before it is used)
  const int nl = nc ? *std::max_element(clevel.begin(), clevel.end()) + 1 : 0;
  S.lvl_ptr.assign((size_t)nl + 1, 0);
  for (int c = 0; c < nc; ++c) ++S.lvl_ptr[clevel[c] + 1];
  for (int t = 0; t < nl; ++t) S.lvl_ptr[t + 1] += S.lvl_ptr[t];
  S.lvl_cl.assign((size_t)nc, 0);
  {
    std::vector<int32_t> fill(S.lvl_ptr.begin(), S.lvl_ptr.end() - 1);
    for (int c = 0; c < nc; ++c) S.lvl_cl[fill[clevel[c]]++] = c;
  }
  // ---- device schedule ----
  S.erow_f.assign(S.colidx.size(), 0); S.erow_b.assign(S.rowidx.size(), 0);
  S.desc.clear();
  S.linv_blocks = 0;
  S.col2row.assign(S.rowidx.size(), 0);
  for (size_t q = 0; q < S.rowslot.size(); ++q) S.col2row[S.rowslot[q]] = (int32_t)q;
  for (int c = 0; c < nc; ++c) {
    const int p0 = S.cl_ptr[c], p1 = S.cl_ptr[c + 1];
    for (int p = p0; p < p1; ++p) {
      for (int32_t q = S.rowptr[p]; q < S.rowptr[p + 1]; ++q)
        S.erow_f[q] = (uint8_t)((p - p0) | (S.colidx[q] >= p0 ? 0x80 : 0));
      for (int32_t q = S.colptr[p]; q < S.colptr[p + 1]; ++q)
        S.erow_b[q] = (uint8_t)((p - p0) | (S.rowidx[q] < p1 ? 0x80 : 0));
    }
    const int32_t dsc[8] = {p0, p1 - p0, S.rowptr[p0], S.rowptr[p1] - S.rowptr[p0],
                            S.colptr[p0], S.colptr[p1] - S.colptr[p0], (int32_t)S.linv_blocks, 0};
    S.desc.insert(S.desc.end(), dsc, dsc + 8);
    S.linv_blocks += ((int64_t)(p1 - p0) * (p1 - p0) + 1) & ~(int64_t)1;  // even: 16-byte aligned for any block size
  }
}

// B x B helpers (row-major)
template <int B>
inline void gblk_submul_nt(double *Z, const double *X, const double *Y) {  // Z -= X * Y^T
  for (int a = 0; a < B; ++a)
    for (int b = 0; b < B; ++b) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += X[a * B + k] * Y[b * B + k];
      Z[a * B + b] -= s;
    }
}

// Left-looking numeric factorisation.  A: n diagonal blocks in POSE order; E: one block M_ij (rows of pose i,
// columns of pose j, i < j) per input edge, duplicates summed.  Lval: nnzL blocks in column storage, Dinv: n
// blocks L_vv^-1 in POSITION order.  Returns false when a pivot is not positive.
template <int B>
inline bool gen_numeric(const GenSym &S, const double *A, const double *E, std::vector<double> &Lval,
                        std::vector<double> &Dinv) {
  constexpr int BB = B * B;
  const int n = S.n;
  Lval.assign((size_t)S.nnzL() * BB, 0.0);
  Dinv.assign((size_t)std::max(n, 1) * BB, 0.0);
  for (int e = 0; e < S.ne; ++e) {
    double *dst = Lval.data() + (size_t)S.edge_slot[e] * BB;
    const double *src = E + (size_t)e * BB;
    if (S.edge_tr[e])
      for (int a = 0; a < B; ++a) for (int b = 0; b < B; ++b) dst[a * B + b] += src[b * B + a];
    else
      for (int q = 0; q < BB; ++q) dst[q] += src[q];
  }
  bool ok = true;
  std::vector<int32_t> where((size_t)n, -1);
  for (int v = 0; v < n; ++v) {
    double P[BB];
    const double *Av = A + (size_t)S.perm[v] * BB;
    for (int a = 0; a < B; ++a) for (int b = 0; b < B; ++b) P[a * B + b] = 0.5 * (Av[a * B + b] + Av[b * B + a]);
    for (int32_t q = S.colptr[v]; q < S.colptr[v + 1]; ++q) where[S.rowidx[q]] = q;
    for (int32_t q = S.rowptr[v]; q < S.rowptr[v + 1]; ++q) {
      const int u = S.colidx[q];
      const int32_t s0 = S.rowslot[q];
      const double *Lvu = Lval.data() + (size_t)s0 * BB;
      gblk_submul_nt<B>(P, Lvu, Lvu);
      for (int32_t s = s0 + 1; s < S.colptr[u + 1]; ++s)
        gblk_submul_nt<B>(Lval.data() + (size_t)where[S.rowidx[s]] * BB, Lval.data() + (size_t)s * BB, Lvu);
    }
    // P = Lc Lc^T, Li = Lc^-1
    double Lc[BB] = {0}, Li[BB] = {0};
    for (int j = 0; j < B; ++j) {
      double s = P[j * B + j];
      for (int k = 0; k < j; ++k) s -= Lc[j * B + k] * Lc[j * B + k];
      if (!(s > 0.0)) { ok = false; s = 1.0; }
      const double dj = std::sqrt(s);
      Lc[j * B + j] = dj;
      for (int i = j + 1; i < B; ++i) {
        double t = P[i * B + j];
        for (int k = 0; k < j; ++k) t -= Lc[i * B + k] * Lc[j * B + k];
        Lc[i * B + j] = t / dj;
      }
    }
    for (int j = 0; j < B; ++j) {
      Li[j * B + j] = 1.0 / Lc[j * B + j];
      for (int i = j + 1; i < B; ++i) {
        double t = 0.0;
        for (int k = j; k < i; ++k) t -= Lc[i * B + k] * Li[k * B + j];
        Li[i * B + j] = t / Lc[i * B + i];
      }
    }
    for (int q = 0; q < BB; ++q) Dinv[(size_t)v * BB + q] = Li[q];
    for (int32_t q = S.colptr[v]; q < S.colptr[v + 1]; ++q) {  // L_wv = X Li^T
      double *X = Lval.data() + (size_t)q * BB, T[BB];
      for (int a = 0; a < B; ++a)
        for (int b = 0; b < B; ++b) {
          double s = 0.0;
          for (int k = 0; k <= b; ++k) s += X[a * B + k] * Li[b * B + k];
          T[a * B + b] = s;
        }
      for (int e = 0; e < BB; ++e) X[e] = T[e];
    }
  }
  return ok;
}

// Inverse of the in-cluster part of L for every cluster K (poses p0..p1-1): Linv_K = L_KK^-1, a block lower
// triangular (poses x poses) matrix stored dense as [j][j'][B*B].  With it a cluster is eliminated by two products
// instead of a serial chain over its poses:  y_K = Linv_K (b_K - L_K,ext y_ext)  and  x_K = Linv_K^T (y_K - ...).
template <int B>
inline void gen_cluster_inverses(const GenSym &S, const std::vector<double> &Lval, const std::vector<double> &Dinv,
                                 std::vector<double> &Linv) {
  constexpr int BB = B * B;
  Linv.assign((size_t)std::max<int64_t>(S.linv_blocks, 1) * BB, 0.0);
  const int nc = (int)S.cl_ptr.size() - 1;
  for (int c = 0; c < nc; ++c) {
    const int p0 = S.cl_ptr[c], k = S.cl_ptr[c + 1] - p0;
    double *X = Linv.data() + (size_t)S.desc[(size_t)c * 8 + 6] * BB;
    auto blk = [&](int j, int jc) { return X + ((size_t)j * k + jc) * BB; };
    for (int j = 0; j < k; ++j) {
      const double *Dj = Dinv.data() + (size_t)(p0 + j) * BB;
      // row j of the inverse: X[j][jc] = Dj * (delta_{j,jc} I - sum_{j'' in row j, in cluster} L[j][j''] X[j''][jc])
      for (int jc = 0; jc <= j; ++jc) {
        double T[BB];
        for (int e = 0; e < BB; ++e) T[e] = 0.0;
        if (jc == j) for (int a = 0; a < B; ++a) T[a * B + a] = 1.0;
        for (int32_t q = S.rowptr[p0 + j]; q < S.rowptr[p0 + j + 1]; ++q) {
          const int u = S.colidx[q] - p0;
          if (u < jc) continue;  // outside the cluster (u < 0) or X[u][jc] = 0
          const double *Lb = Lval.data() + (size_t)S.rowslot[q] * BB, *Xu = blk(u, jc);
          for (int a = 0; a < B; ++a)
            for (int b = 0; b < B; ++b) {
              double s = 0.0;
              for (int e = 0; e < B; ++e) s += Lb[a * B + e] * Xu[e * B + b];
              T[a * B + b] -= s;
            }
        }
        double *O = blk(j, jc);
        for (int a = 0; a < B; ++a)
          for (int b = 0; b < B; ++b) {
            double s = 0.0;
            for (int e = 0; e <= a; ++e) s += Dj[a * B + e] * T[e * B + b];
            O[a * B + b] = s;
          }
      }
    }
  }
}

// X: [n][B][ld] in POSE order, solved in place (host reference of the device kernels of gen_chol_dev.cuh).
template <int B>
inline void gen_solve_host(const GenSym &S, const double *Lval, const double *Dinv, double *X, int ld, int ncols) {
  constexpr int BB = B * B;
  const int n = S.n;
  std::vector<double> y((size_t)std::max(n, 1) * B * ncols);
  double acc[B];
  for (int v = 0; v < n; ++v)
    for (int c = 0; c < ncols; ++c) {
      for (int a = 0; a < B; ++a) acc[a] = X[((size_t)S.perm[v] * B + a) * ld + c];
      for (int32_t q = S.rowptr[v]; q < S.rowptr[v + 1]; ++q) {
        const double *Lb = Lval + (size_t)S.rowslot[q] * BB;
        const double *yu = y.data() + (size_t)S.colidx[q] * B * ncols;
        for (int a = 0; a < B; ++a)
          for (int b = 0; b < B; ++b) acc[a] -= Lb[a * B + b] * yu[(size_t)b * ncols + c];
      }
      const double *Li = Dinv + (size_t)v * BB;
      for (int a = 0; a < B; ++a) {
        double s = 0.0;
        for (int b = 0; b <= a; ++b) s += Li[a * B + b] * acc[b];
        y[((size_t)v * B + a) * ncols + c] = s;
      }
    }
  for (int v = n - 1; v >= 0; --v)
    for (int c = 0; c < ncols; ++c) {
      for (int a = 0; a < B; ++a) acc[a] = y[((size_t)v * B + a) * ncols + c];
      for (int32_t q = S.colptr[v]; q < S.colptr[v + 1]; ++q) {
        const double *Lb = Lval + (size_t)q * BB;
        const double *xw = y.data() + (size_t)S.rowidx[q] * B * ncols;
        for (int a = 0; a < B; ++a)
          for (int b = 0; b < B; ++b) acc[a] -= Lb[b * B + a] * xw[(size_t)b * ncols + c];
      }
      const double *Li = Dinv + (size_t)v * BB;
      for (int a = 0; a < B; ++a) {
        double s = 0.0;
        for (int b = a; b < B; ++b) s += Li[b * B + a] * acc[b];
        y[((size_t)v * B + a) * ncols + c] = s;
      }
      for (int a = 0; a < B; ++a) X[((size_t)S.perm[v] * B + a) * ld + c] = y[((size_t)v * B + a) * ncols + c];
    }
}

}  // namespace cora_b200
