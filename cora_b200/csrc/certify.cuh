// certify.cuh -- certification, saddle escape, rounding and the staircase driver.
//
//   Problem::certify_solution   src/CORA_problem.cpp:1030-1103
//   fast_verification           src/CORA_utils.cpp:17-186
//   saddleEscape                src/CORA.cpp:245-350
//   projectSolution             src/CORA.cpp:352-441, projectToSOd src/CORA_utils.cpp:188-202
//   solveCORA                   src/CORA.cpp:26-243
//
// Certification mirrors the reference's decision procedure: S + eta I is tested for positive
// definiteness by a Cholesky-type factorisation (the chain Cholesky of chain_chol.cuh in place of
// CholmodSupernodalLLT); only when that fails is an eigen-solver run to find a direction of
// negative curvature, with the reference's early exit x'Sx < -eta/2.  Where the reference runs
// SYM-ILDL-preconditioned LOBPCG (third-party, PARITY UNPINNED), this implementation runs Lanczos
// with full re-orthogonalisation on the device: shift-and-invert through the chain Cholesky of
// S + sigma I (sigma the first eta*4^k that makes it positive definite), or plain Lanczos on S with
// the fused product Q x - Lambda x when the graph is not an odometry chain.
#pragma once
#include <random>

#include "lanczos.cuh"

namespace cora_b200 {

// --------------------------------------------------------------- tall kernels ---
// out[j] = <Q_j, w>, j < k; Q holds k vectors of length n back to back.
__global__ void __launch_bounds__(kThreads) k_tall_dots(const double *__restrict__ Q, int k, long long n,
                                                        const double *__restrict__ w, double *partials,
                                                        unsigned *counter, double *out) {
  __shared__ double sred[64];
  __shared__ int s_last;
  const long long slab = (n + gridDim.x - 1) / gridDim.x;
  const long long e0 = (long long)blockIdx.x * slab, e1 = e0 + slab < n ? e0 + slab : n;
  for (int j = 0; j < k; ++j) {
    double acc[1] = {0.0};
    const double *q = Q + (size_t)j * n;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) acc[0] = fma(q[e], w[e], acc[0]);
    block_sum<1>(acc, sred);
    if (threadIdx.x == 0) partials[(size_t)blockIdx.x * k + j] = acc[0];
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicInc(counter, gridDim.x - 1);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + (size_t)b * k + j);
    out[j] = s;
  }
}

// w -= sum_j c[j] Q_j
__global__ void __launch_bounds__(kThreads) k_tall_axpy(const double *__restrict__ Q, int k, long long n,
                                                        const double *__restrict__ c, double *w) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    double s = w[e];
    for (int j = 0; j < k; ++j) s = fma(-c[j], Q[(size_t)j * n + e], s);
    w[e] = s;
  }
}

// out = sum_j c[j] Q_j
__global__ void __launch_bounds__(kThreads) k_tall_combine(const double *__restrict__ Q, int k, long long n,
                                                           const double *__restrict__ c, double *out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < k; ++j) s = fma(c[j], Q[(size_t)j * n + e], s);
    out[e] = s;
  }
}

// out_c = sum_k Y[:, k] M[k, c]: nd orthonormal vectors spanning the columns of the row-major iterate Y (N x r),
// stored back to back like the Lanczos basis (M = V diag(ev^-1/2) of the Gram matrix Y^T Y)
__global__ void __launch_bounds__(kThreads) k_span_basis(const double *__restrict__ Y, int r, const double *__restrict__ M,
                                                         int nd, long long n, double *out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n * nd;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e / n);
    const long long row = e - (long long)c * n;
    double s = 0.0;
    for (int k = 0; k < r; ++k) s = fma(Y[row * r + k], M[(size_t)k * nd + c], s);
    out[e] = s;
  }
}

// G = A^T B for row-major A (N x ra), B (N x rb): partials per CTA, last CTA reduces.
__global__ void __launch_bounds__(kThreads) k_gram(const double *__restrict__ A, int ra,
                                                   const double *__restrict__ Bm, int rb, long long N,
                                                   double *partials, unsigned *counter, double *out) {
  __shared__ int s_last;
  const long long slab = (N + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * slab, r1 = r0 + slab < N ? r0 + slab : N;
  const int np = ra * rb;
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    const int i = p / rb, j = p - i * rb;
    double s = 0.0;
    for (long long row = r0; row < r1; ++row) s = fma(A[row * ra + i], Bm[row * rb + j], s);
    partials[(size_t)blockIdx.x * np + p] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicInc(counter, gridDim.x - 1);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + (size_t)b * np + p);
    out[p] = s;
  }
}

// out (N x rb) = A (N x ra) * M (ra x rb, row-major, device)
__global__ void __launch_bounds__(kThreads) k_right_multiply(const double *__restrict__ A, int ra,
                                                             const double *__restrict__ M, int rb, long long N,
                                                             double *out) {
  const long long nE = N * rb;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / rb;
    const int j = (int)(e - row * rb);
    double s = 0.0;
    for (int i = 0; i < ra; ++i) s = fma(A[row * ra + i], M[i * rb + j], s);
    out[e] = s;
  }
}

// Rounding of projectSolution (src/CORA.cpp:381-418): count the pose blocks with positive
// determinant (count kernel), then project every d x d block to SO(d) (projectToSOd,
// src/CORA_utils.cpp:188-202: U V^T with the last column of U flipped when det U det V <= 0) and
// normalise the range rows.  Y is N x d row-major internal; one thread per pose / range row.
template <int D>
__device__ __forceinline__ double det_block(const double *m, int ld) {
  if (D == 2) return m[0] * m[ld + 1] - m[1] * m[ld];
  return m[0] * (m[ld + 1] * m[2 * ld + 2] - m[ld + 2] * m[2 * ld + 1]) -
         m[1] * (m[ld] * m[2 * ld + 2] - m[ld + 2] * m[2 * ld]) +
         m[2] * (m[ld] * m[2 * ld + 1] - m[ld + 1] * m[2 * ld]);
}
template <int D>
__global__ void __launch_bounds__(kThreads) k_count_positive_det(const DevLayout L, const double *__restrict__ Y,
                                                                 unsigned *count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  if (det_block<D>(Y + (size_t)i * (D + 1) * D, D) > 0.0) atomicAdd(count, 1u);
}
template <int D>
__global__ void __launch_bounds__(kThreads) k_round_solution(const DevLayout L, double *Y, int reflect) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < L.n) {
    double *m = Y + (size_t)u * (D + 1) * D;
    double M[D][D], G[D][D], V[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) M[a][b] = m[a * D + b] * ((reflect && b == D - 1) ? -1.0 : 1.0);
    const double det = (D == 2) ? M[0][0] * M[1][1] - M[0][1] * M[1][0]
                                : M[0][0] * (M[1][1] * M[2 % D][2 % D] - M[1][2 % D] * M[2 % D][1]) -
                                      M[0][1] * (M[1][0] * M[2 % D][2 % D] - M[1][2 % D] * M[2 % D][0]) +
                                      M[0][2 % D] * (M[1][0] * M[2 % D][1] - M[1][1] * M[2 % D][0]);
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(M[a][k], M[b][k], s);
        G[a][b] = s;
      }
    jacobi_eig<D>(G, V);  // M M^T = V diag(G) V^T, singular values sqrt(G[a][a])
    int imin = 0;
#pragma unroll
    for (int a = 1; a < D; ++a)
      if (G[a][a] < G[imin][imin]) imin = a;
    double sc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) sc[a] = 1.0 / sqrt(fmax(G[a][a], 1e-300));
    if (!(det > 0.0)) {
#pragma unroll
      for (int a = 0; a < D; ++a)
        if (a == imin) sc[a] = -sc[a];
    }
    double T[D][D];  // V diag(sc) V^T
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(V[a][k] * sc[k], V[b][k], s);
        T[a][b] = s;
      }
    double R[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(T[a][k], M[k][b], s);
        R[a][b] = s;
      }
    // one Newton-Schulz step restores orthonormality to rounding level for ill-conditioned blocks
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(R[a][k], R[b][k], s);
        G[a][b] = ((a == b) ? 1.5 : 0.0) - 0.5 * s;
      }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(G[a][k], R[k][b], s);
        m[a * D + b] = s;
      }
    if (reflect) m[D * D + D - 1] = -m[D * D + D - 1];  // translation row of the pose
  } else if (u < L.n + L.l + L.m) {
    double *w = Y + ((size_t)L.nPoseRows + (u - L.n)) * D;
    if (reflect) w[D - 1] = -w[D - 1];
    if (u >= L.n + L.l) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) s = fma(w[c], w[c], s);
      const double inv = 1.0 / sqrt(s);
#pragma unroll
      for (int c = 0; c < D; ++c) w[c] *= inv;
    }
  }
}

// ------------------------------------------------------------ host dense eig ---
// Cyclic Jacobi for a small dense symmetric matrix (row-major n x n); eigenvalues ascending,
// eigenvectors in the columns of V.
inline void sym_eig_jacobi(int n, std::vector<double> A, std::vector<double> &evals, std::vector<double> &V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; ++i) {
      dg += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-32 * dg || off == 0.0) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double th = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  evals.resize(n);
  std::vector<double> V2((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    evals[i] = A[(size_t)idx[i] * n + idx[i]];
    for (int k = 0; k < n; ++k) V2[(size_t)k * n + i] = V[(size_t)k * n + idx[i]];
  }
  V.swap(V2);
}

// G (ra x rb, host) = A^T B on the device
inline std::vector<double> gram_host(H *h, const double *A, int ra, const double *B, int rb) {
  const int np = ra * rb;
  const int grid = std::max(1, std::min(h->sm_count * 4, (int)((h->DL.N + 255) / 256)));
  DevBuf<double> part, out;
  part.alloc((size_t)grid * np);
  out.alloc(np);
  k_gram<<<grid, kThreads, 0, h->stream>>>(A, ra, B, rb, (long long)h->DL.N, part.p, h->d_counter.p + 2, out.p);
  check_launch(h);
  std::vector<double> G(np);
  CUDA_CHECK(cudaMemcpyAsync(G.data(), out.p, np * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return G;
}

// ------------------------------------------------------------- certification ---
struct CertOut {
  bool certified = false;
  int branch = CORA_B200_CERT_NONE;
  double theta = 0.0;
  int64_t iters = 0;
  bool have_x = false;  // direction of negative curvature left in ws[V_T1] (N x 1 internal)
};

// Device Lanczos with full re-orthogonalisation on the operator `op` (w = op(q)); returns the
// extreme Ritz pair wanted by `pick_largest` after each step and stops when accept(theta_S) says so.
// Basis vectors live in `basis` (kmax x N).  On exit the chosen Ritz vector is in xout (unit norm).
struct LanczosResult {
  int steps = 0;
  double ritz = 0.0;
  double theta_S = 0.0;  // x' S x of the returned vector
  bool accepted = false;
};

// defl / ndefl: orthonormal vectors (stored like the basis) that are projected out of the start vector and of every
// operator product: the search runs in their orthogonal complement.  certify_resident passes an orthonormal basis of
// span(Y): at a critical point S Y = 0, so these r zero-eigenvalue directions are the closest competitors of a
// slightly negative lambda_min under shift-and-invert and carry no negative curvature.  (The reference seeds its
// block eigen-solver with Y -- the bootstrap of src/CORA.cpp:155-168 -- for the same reason: to dispose of them.)
template <typename Op, typename Rayleigh, typename Accept>
inline LanczosResult device_lanczos(H *h, Op op, Rayleigh rayleigh, Accept accept, bool pick_largest, int kmax,
                                    DevBuf<double> &basis, double *w, double *xout, unsigned seed,
                                    const double *defl = nullptr, int ndefl = 0) {
  const long long N = h->DL.N;
  const int grid = std::max(1, std::min(h->sm_count * 4, (int)((N + 255) / 256)));
  DevBuf<double> part, coef;
  part.alloc((size_t)grid * (std::max(kmax, ndefl) + 1));
  coef.alloc(std::max(kmax, ndefl) + 1);
  auto deflate = [&](double *v) {
    if (ndefl <= 0) return;
    for (int pass = 0; pass < 2; ++pass) {
      k_tall_dots<<<grid, kThreads, 0, h->stream>>>(defl, ndefl, N, v, part.p, h->d_counter.p + 2, coef.p);
      check_launch(h);
      k_tall_axpy<<<flat_grid(h, N), kThreads, 0, h->stream>>>(defl, ndefl, N, coef.p, v);
      check_launch(h);
    }
  };
  if (basis.n < (size_t)(kmax + 1) * N) basis.alloc((size_t)(kmax + 1) * N);
  {  // random start (the reference seeds LOBPCG with Matrix::Random; any start is admissible)
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> x((size_t)N);
    double nrm = 0.0;
    for (auto &v : x) { v = U(gen); nrm += v * v; }
    nrm = std::sqrt(nrm);
    for (auto &v : x) v /= nrm;
    CUDA_CHECK(cudaMemcpyAsync(basis.p, x.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (ndefl > 0) {
      deflate(basis.p);
      launch_dot2(h, basis.p, basis.p, nullptr, nullptr, N, SC_TMP);
      read_scal(h);
      const double n2 = std::sqrt(std::max(h->h_scal[SC_TMP], 1e-300));
      launch_axpby(h, 1.0 / n2, basis.p, 0.0, nullptr, basis.p, N);
    }
  }
  std::vector<double> al, be, hc(kmax + 1);
  LanczosResult R;
  for (int k = 0; k < kmax; ++k) {
    double *q = basis.p + (size_t)k * N;
    op(q, w);
    deflate(w);
    // two passes of classical Gram-Schmidt against the whole basis (full re-orthogonalisation)
    double alpha = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      k_tall_dots<<<grid, kThreads, 0, h->stream>>>(basis.p, k + 1, N, w, part.p, h->d_counter.p + 2, coef.p);
      check_launch(h);
      k_tall_axpy<<<flat_grid(h, N), kThreads, 0, h->stream>>>(basis.p, k + 1, N, coef.p, w);
      check_launch(h);
      CUDA_CHECK(cudaMemcpyAsync(hc.data(), coef.p, (k + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      alpha += hc[k];
    }
    al.push_back(alpha);
    launch_dot2(h, w, w, nullptr, nullptr, N, SC_TMP);
    read_scal(h);
    const double beta = std::sqrt(std::max(0.0, h->h_scal[SC_TMP]));
    R.steps = k + 1;
    // Ritz pair
    std::vector<double> dd = al, ee = be, Z;
    tridiag_eig(k + 1, dd, ee, &Z);
    const int pick = pick_largest ? k : 0;
    R.ritz = dd[pick];
    const bool last = (k == kmax - 1) || beta <= 1e-13 * (std::fabs(R.ritz) + 1e-300);
    // residual of the Ritz pair: |beta * z_k|
    const double resid = std::fabs(beta * Z[(size_t)k * (k + 1) + pick]);
    const bool converged = resid <= 1e-3 * std::fabs(R.ritz);
    if (converged || last || (k % 4 == 3)) {
      std::vector<double> c(k + 1);
      for (int j = 0; j <= k; ++j) c[j] = Z[(size_t)j * (k + 1) + pick];
      CUDA_CHECK(cudaMemcpyAsync(coef.p, c.data(), (k + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      k_tall_combine<<<flat_grid(h, N), kThreads, 0, h->stream>>>(basis.p, k + 1, N, coef.p, xout);
      check_launch(h);
      R.theta_S = rayleigh(xout);
      if (accept(R.theta_S)) { R.accepted = true; return R; }
      if (converged || last) return R;
    }
    be.push_back(beta);
    launch_axpby(h, 1.0 / beta, w, 0.0, nullptr, basis.p + (size_t)(k + 1) * N, N);
  }
  return R;
}

// Certify the iterate in ws[V_X] (rank r).
inline CertOut certify_resident(H *h, int r, double eta, int max_iters, int verbose) {
  CertOut out;
  ensure_workspace(h, r);
  const long long N = h->DL.N;
  double *X = h->ws[V_X].p;
  // singular-value ratio of Y (src/CORA_problem.cpp:1039-1049) from the r x r Gram matrix
  std::vector<double> spanM;  // r x r: V diag(ev^-1/2), columns map Y to an orthonormal basis of its span
  {
    std::vector<double> G = gram_host(h, X, r, X, r), ev, V;
    sym_eig_jacobi(r, G, ev, V);
    if (ev[r - 1] / std::max(ev[0], 0.0) > 1e12 || !(ev[0] > 0.0)) {
      out.certified = true;
      out.branch = CORA_B200_CERT_SV_RATIO;
      return out;
    }
    spanM.assign((size_t)r * r, 0.0);
    for (int k = 0; k < r; ++k)
      for (int c = 0; c < r; ++c) spanM[(size_t)k * r + c] = V[(size_t)k * r + c] / std::sqrt(ev[c]);
  }
  compute_lambda(h, X, r);
  DevLayout LS = build_certificate_layout(h, 0.0);  // values of S on Q's structure
  auto Sx = [&](const double *x, double *y) {
    launch_qprod(h, QM_SPMM, x, nullptr, nullptr, y, nullptr, 1, POST_STORE, SC_TMP, nullptr, &LS);
  };
  double *x = h->ws[V_T1].p, *w = h->ws[V_T0].p, *t = h->ws[V_Z].p;
  auto rayleigh = [&](const double *v) -> double {  // v' S v / v' v
    Sx(v, t);
    launch_dot2(h, v, t, v, v, N, SC_TMP);
    read_scal(h);
    return h->h_scal[SC_TMP] / h->h_scal[SC_TMP + 1];
  };
  auto accept = [&](double th) { return th < -eta / 2; };  // src/CORA_utils.cpp:90-99
  DevBuf<double> basis;
  bool chain = true;
  ChainChol *C = nullptr;
  try {
    bool pd = false;
    C = build_chain_chol(h, LS.bval, LS.sdiag, eta, /*pin_last=*/false, &pd, /*want_solve=*/false);
    release_chain_chol(h, C);
    C = nullptr;
    if (pd) {  // S + eta I > 0  (src/CORA_utils.cpp:33-57)
      out.certified = true;
      out.branch = CORA_B200_CERT_PSD;
      return out;
    }
  } catch (const Error &e) {
    if (e.code != CORA_B200_ENOTIMPL) throw;
    chain = false;
  }
  if (N <= 1) {
    out.theta = rayleigh(X);
    return out;
  }
  const int kmax = (int)std::min<long long>(std::max(8, max_iters), N - 1);
  // orthonormal basis of span(Y), projected out of the eigen-search (see device_lanczos)
  DevBuf<double> spanY, spanMd;
  int ndefl = 0;
  if (N > 2 * (long long)r + 2) {
    spanMd.upload(spanM, h->stream);
    spanY.alloc((size_t)N * r);
    k_span_basis<<<flat_grid(h, N * r), kThreads, 0, h->stream>>>(X, r, spanMd.p, r, N, spanY.p);
    check_launch(h);
    ndefl = r;
  }
  if (chain) {
    // shift-and-invert: the largest eigenvalue of (S + sigma I)^-1 is 1 / (lambda_min(S) + sigma)
    double sigma = std::max(eta, 1e-12);
    for (int tries = 0; tries < 60; ++tries) {
      sigma *= 4.0;
      bool pd = false;
      C = build_chain_chol(h, LS.bval, LS.sdiag, sigma, false, &pd, true);
      if (pd) break;
      release_chain_chol(h, C);
      C = nullptr;
    }
    if (!C) throw Error(CORA_B200_ERUNTIME, "certification: could not shift S to positive definiteness");
    auto op = [&](const double *q, double *y) { chain_solve(h, C, q, y, 1, nullptr); };
    LanczosResult L = device_lanczos(h, op, rayleigh, accept, /*pick_largest=*/true, std::min(kmax, 80), basis, w, x, 12345u,
                                     spanY.p, ndefl);
    if (!accept(L.theta_S) && ndefl > 0) {
      // Y is not a critical point (TNT stopped on the relative decrease): S Y != 0 and the negative curvature may
      // live partly in span(Y).  S + eta I failed the Cholesky test, so a direction exists: search the whole space.
      const int64_t s0 = L.steps;
      L = device_lanczos(h, op, rayleigh, accept, /*pick_largest=*/true, std::min(kmax, 80), basis, w, x, 12345u);
      L.steps += (int)s0;
    }
    release_chain_chol(h, C);
    out.iters = L.steps;
    out.theta = L.theta_S;
    out.have_x = true;
    // (x' S x >= -eta/2 after a full search: the direction is not a verified descent direction; the caller must not
    //  escape along it)
    out.branch = accept(L.theta_S) ? CORA_B200_CERT_EIGENPAIR : CORA_B200_CERT_INCONCLUSIVE;
    if (verbose) std::printf("  certify: shift-invert Lanczos sigma=%.3e steps=%d theta=%.6e\n", sigma, L.steps, L.theta_S);
  } else {
    auto op = [&](const double *q, double *y) { Sx(q, y); };
    LanczosResult L = device_lanczos(h, op, rayleigh, accept, /*pick_largest=*/false, std::min(kmax, 400), basis, w, x, 12345u,
                                     spanY.p, ndefl);
    if (!accept(L.theta_S) && ndefl > 0) {
      const int64_t s0 = L.steps;
      L = device_lanczos(h, op, rayleigh, accept, /*pick_largest=*/false, std::min(kmax, 400), basis, w, x, 12345u);
      L.steps += (int)s0;
    }
    out.iters = L.steps;
    out.theta = L.theta_S;
    out.have_x = true;
    // No factorisation of S + eta I for this graph (loop closures / several robots), so positive semidefiniteness
    // cannot be PROVEN: never certified here.  With x' S x < -eta/2 the vector is a valid direction of negative
    // curvature (the reference's own early exit, src/CORA_utils.cpp:90-99); otherwise the verdict is inconclusive
    // and the caller must neither certify nor escape along x.
    out.certified = false;
    out.branch = accept(L.theta_S) ? CORA_B200_CERT_EIGENPAIR : CORA_B200_CERT_INCONCLUSIVE;
    if (verbose) std::printf("  certify: plain Lanczos steps=%d theta=%.6e\n", L.steps, L.theta_S);
  }
  if (h->formulation == CORA_B200_FORMULATION_IMPLICIT && out.have_x && !out.certified) {
    // src/CORA_problem.cpp:1085-1100: keep the rotation / range part of the direction, normalised, and report its
    // Rayleigh quotient with the simplified certificate matrix  x' (Q_implicit x - Lambda x)
    zero_translation_rows(h, x, 1, nullptr);
    launch_dot2(h, x, x, x, x, N, SC_TMP);
    read_scal(h);
    const double nrm = std::sqrt(h->h_scal[SC_TMP]);
    if (!(nrm > 0.0)) throw Error(CORA_B200_ERUNTIME, "NaN in theta -- result not certified and implicit form");
    launch_axpby(h, 1.0 / nrm, x, 0.0, nullptr, x, N);
    const double *xf = implicit_complete(h, x, 1, nullptr);
    Sx(xf, t);  // rows of S [x; t*(x)]: the translation block of Lambda is zero
    launch_dot2(h, x, t, x, x, N, SC_TMP);
    read_scal(h);
    out.theta = h->h_scal[SC_TMP];
    if (std::isnan(out.theta)) throw Error(CORA_B200_ERUNTIME, "NaN in theta -- result not certified and implicit form");
  }
  return out;
}

// The PSD half of fast_verification on its own (src/CORA_utils.cpp:33-57): is S(Y) + eta I positive definite?
// Lambda blocks, certificate values on Q's structure and the chain Cholesky all run on the device; no sv-ratio
// short-circuit, no eigen-search.  Throws ENOTIMPL on graphs without a device factorisation.
inline bool psd_test_resident(H *h, int r, double eta) {
  ensure_workspace(h, r);
  compute_lambda(h, h->ws[V_X].p, r);
  DevLayout LS = build_certificate_layout(h, 0.0);
  bool pd = false;
  ChainChol *C = build_chain_chol(h, LS.bval, LS.sdiag, eta, /*pin_last=*/false, &pd, /*want_solve=*/false);
  release_chain_chol(h, C);
  return pd;
}

// Test hook: smallest eigenpair of the handle's own matrix by the device Lanczos of the certification (plain mode,
// full re-orthogonalisation).  Lets the reference's eigenpair known answers (I - 2 x x^T -> (-1, +-x),
// tests/test_certification.cpp:45-79) pin the CUDA eigen-search directly: any symmetric matrix can be loaded as a
// "problem" made of landmark rows only.
inline void debug_min_eigenpair(H *h, int max_iters, double *theta, double *x_out, int *steps) {
  ensure_workspace(h, 1);
  h->resident_r = 0;
  const long long N = h->DL.N;
  double *x = h->ws[V_T1].p, *w = h->ws[V_T0].p, *t = h->ws[V_Z].p;
  auto Qx = [&](const double *q, double *y) {
    launch_qprod_raw(h, QM_SPMM, q, nullptr, nullptr, y, nullptr, 1, POST_STORE, SC_TMP, nullptr);
  };
  auto rayleigh = [&](const double *v) -> double {
    Qx(v, t);
    launch_dot2(h, v, t, v, v, N, SC_TMP);
    read_scal(h);
    return h->h_scal[SC_TMP] / h->h_scal[SC_TMP + 1];
  };
  auto never = [](double) { return false; };
  DevBuf<double> basis;
  const int kmax = (int)std::min<long long>(std::max(2, max_iters), N - 1);
  LanczosResult L = device_lanczos(h, Qx, rayleigh, never, /*pick_largest=*/false, kmax, basis, w, x, 12345u);
  *theta = L.theta_S;
  if (steps) *steps = L.steps;
  export_matrix(h, x, 1, x_out);
}

inline void certify_host(H *h, int r, const double *Y, double eta, int nx, const double *bootstrap,
                         int bootstrap_cols, int max_iters, int *is_certified, double *theta, double *x,
                         double *all_eigvecs, int cap, int *ncols, int64_t *num_iters) {
  // nx / bootstrap: block size and initial block of the reference's LOBPCG (src/CORA.cpp:155-168: Y itself on the
  // first loop, the last eigenvector block afterwards).  The single-vector Lanczos here has no block to seed; what the
  // bootstrap buys the reference -- getting the zero-eigenvalue directions span(Y) out of the way -- is done by
  // projecting span(Y) out of the search (certify_resident), for every call, from the iterate itself.
  (void)nx; (void)bootstrap; (void)bootstrap_cols;
  ensure_workspace(h, r);
  h->resident_r = 0;
  import_matrix(h, Y, r, h->ws[V_X].p, r);
  h->resident_r = r;
  CertOut c = certify_resident(h, r, eta, max_iters > 0 ? max_iters : 500, 0);
  h->last_cert_branch = c.branch;
  *is_certified = c.certified ? 1 : 0;
  *theta = c.theta;
  if (num_iters) *num_iters = c.iters;
  const size_t N = (size_t)h->io_rows();  // (implicit formulation: the truncated direction, :1085-1100)
  if (c.have_x) {
    export_matrix(h, h->ws[V_T1].p, 1, x);
    if (all_eigvecs && cap >= 1) std::memcpy(all_eigvecs, x, N * sizeof(double));
    if (ncols) *ncols = (all_eigvecs && cap >= 1) ? 1 : 0;
  } else {
    std::memset(x, 0, N * sizeof(double));
    if (ncols) *ncols = 0;
  }
}

// ------------------------------------------------------------- saddle escape ---
// Resident version: Y (rank r-1) is in ws[V_X] with ld r-1; v (N x 1 internal) in ws[V_T1].
// On return ws[V_X] holds the escaped iterate at rank r.  src/CORA.cpp:245-350.
__global__ void __launch_bounds__(kThreads) k_augment(const double *__restrict__ Yold, const double *__restrict__ v,
                                                      double *Ya, double *Yd, long long N, int r) {
  const long long nE = N * r;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nE;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / r;
    const int c = (int)(e - row * r);
    Ya[e] = c < r - 1 ? Yold[row * (r - 1) + c] : 0.0;
    Yd[e] = c < r - 1 ? 0.0 : v[row];
  }
}

inline void saddle_escape_resident(H *h, int r, double theta, double gtol, double pgtol, int verbose) {
  ensure_workspace(h, r);
  const long long N = h->DL.N, nE = N * r;
  auto v = [&](int i) { return h->ws[i].p; };
  // Y_augmented -> V_XP (then swapped into V_X), Ydot -> V_S
  k_augment<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(v(V_X), v(V_T1), v(V_XP), v(V_S), N, r);
  check_launch(h);
  swap_vec(h, V_X, V_XP);
  launch_qprod(h, QM_GRAD, v(V_X), v(V_X), nullptr, v(V_GRAD), v(V_G), r, POST_STORE, SC_XG, nullptr);
  read_scal(h);
  const double FY = 0.5 * h->h_scal[SC_XG];
  const double alpha_min = 1e-6;
  double alpha = std::max(16 * alpha_min, 100 * gtol / std::fabs(theta));
  std::vector<double> alphas, fvals;
  bool found = false;
  while (alpha >= alpha_min) {
    launch_retract(h, v(V_X), v(V_S), alpha, nullptr, v(V_XP), r, -1);
    launch_qprod(h, QM_GRAD, v(V_XP), v(V_XP), nullptr, v(V_GRADP), v(V_GP), r, POST_STORE, SC_XG2, nullptr);
    precondition_project(h, v(V_XP), v(V_GRADP), v(V_T0), r, SC_RV2);
    read_scal(h);
    const double Ft = 0.5 * h->h_scal[SC_XG2];
    const double gn = std::sqrt(h->h_scal[SC_GG2]);
    const double pgn = std::sqrt(h->h_scal[SC_RV2 + 1]);
    alphas.push_back(alpha);
    fvals.push_back(Ft);
    if (verbose) std::printf("  saddle escape: alpha=%.3e F=%.9e (FY=%.9e) |g|=%.3e |Pg|=%.3e\n", alpha, Ft, FY, gn, pgn);
    if (Ft < FY && gn > gtol && pgn > pgtol) { found = true; break; }
    alpha /= 2;
  }
  if (found) {
    swap_vec(h, V_X, V_XP);
  } else {
    size_t k = 0;
    for (size_t i = 1; i < fvals.size(); ++i)
      if (fvals[i] < fvals[k]) k = i;
    if (!fvals.empty() && fvals[k] < FY) {
      launch_retract(h, v(V_X), v(V_S), alphas[k], nullptr, v(V_XP), r, -1);
      swap_vec(h, V_X, V_XP);
    } else {
      std::printf("WARNING! BACKTRACKING LINE SEARCH FAILED TO ESCAPE FROM SADDLE POINT!\n");
    }
  }
  h->resident_r = r;
}

inline void saddle_escape_host(H *h, int r_new, const double *Y, double theta, const double *vdir, double gtol,
                               double pgtol, double *Y_out) {
  ensure_workspace(h, r_new);
  h->resident_r = 0;
  import_matrix(h, Y, r_new - 1, h->ws[V_X].p, r_new - 1);
  import_matrix(h, vdir, 1, h->ws[V_T1].p, 1);
  saddle_escape_resident(h, r_new, theta, gtol, pgtol, 0);
  export_matrix(h, h->ws[V_X].p, r_new, Y_out);
}

// ---------------------------------------------------------- projectSolution ----
// ws[V_X] (rank r) -> ws[V_X] (rank d).  src/CORA.cpp:352-441.
inline void project_solution_resident(H *h, int r) {
  const int d = h->DL.d;
  const long long N = h->DL.N;
  ensure_workspace(h, std::max(r, d));
  std::vector<double> G = gram_host(h, h->ws[V_X].p, r, h->ws[V_X].p, r), ev, V;
  sym_eig_jacobi(r, G, ev, V);
  // Yd = U_d Sigma_d = Y V_d with V_d the right singular vectors of the d largest singular values
  std::vector<double> Vd((size_t)r * d);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < d; ++j) Vd[(size_t)i * d + j] = V[(size_t)i * r + (r - 1 - j)];
  DevBuf<double> dV;
  dV.upload(Vd, h->stream);
  k_right_multiply<<<flat_grid(h, N * d), kThreads, 0, h->stream>>>(h->ws[V_X].p, r, dV.p, d, N, h->ws[V_XP].p);
  check_launch(h);
  swap_vec(h, V_X, V_XP);
  unsigned *cnt = h->d_counter.p + 3;
  CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned), h->stream));
  const int n = h->DL.n;
  if (n > 0) {
    DISPATCH_D(h, k_count_positive_det<DD><<<(n + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(
                      h->DL, h->ws[V_X].p, cnt));
    check_launch(h);
  }
  unsigned ng0 = 0;
  CUDA_CHECK(cudaMemcpyAsync(&ng0, cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned), h->stream));
  const int reflect = (n > 0 && (int)ng0 < n / 2) ? 1 : 0;  // :403
  const int tot = h->DL.n + h->DL.l + h->DL.m;
  DISPATCH_D(h, k_round_solution<DD><<<(tot + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(
                    h->DL, h->ws[V_X].p, reflect));
  check_launch(h);
  h->resident_r = d;
}

inline void project_solution_host(H *h, int r, const double *Y, double *Y_out) {
  ensure_workspace(h, r);
  h->resident_r = 0;
  import_matrix(h, Y, r, h->ws[V_X].p, r);
  h->resident_r = r;
  project_solution_resident(h, r);
  export_matrix(h, h->ws[V_X].p, h->DL.d, Y_out);
}

// ------------------------------------------------------------------ staircase ---
inline void solve_staircase(H *h, int r0, const double *X0, int max_rank, const cora_b200_tnt_params &p,
                            int verbose, double *X_out, cora_b200_solve_result *res) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto since = [&](clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); };
  const int d = h->DL.d;
  if (r0 < d) throw Error(CORA_B200_EINVAL, "relaxation rank must be >= dim");
  ensure_workspace(h, std::max(r0, std::min(max_rank, kMaxGeomRank)));
  h->resident_r = 0;
  import_matrix(h, X0, r0, h->ws[V_T0].p, r0);
  launch_retract(h, h->ws[V_T0].p, nullptr, 0.0, nullptr, h->ws[V_X].p, r0, -1);  // src/CORA.cpp:128
  h->resident_r = r0;
  int rank = r0, ns = 0;
  int64_t total_cg = 0;
  cora_b200_tnt_result tr{};
  CertOut cert;
  double eta = 0.0;
  res->lifted_f = 0.0; res->lifted_rank = 0; res->certified = 0; res->refined_certified = 0;
  auto record = [&](double tnt_s, double cert_s) {
    if (res->stages && ns < res->stage_capacity) {
      cora_b200_stage &s = res->stages[ns];
      s.rank = rank; s.status = tr.status; s.num_outer = tr.num_outer; s.certified = cert.certified;
      s.cg_iterations = tr.total_inner; s.f = tr.f; s.gradfx_norm = tr.gradfx_norm; s.theta = cert.theta;
      s.eta = eta; s.tnt_seconds = tnt_s; s.cert_seconds = cert_s;
      s.cert_branch = cert.branch; s.reserved = 0;
    }
    ++ns;
  };
  while (rank <= max_rank) {  // src/CORA.cpp:134
    if (rank > kMaxGeomRank) throw Error(CORA_B200_EINVAL, "relaxation rank above 24 is not supported");  // (r0 itself)
    auto ta = clk::now();
    std::memset(&tr, 0, sizeof(tr));
    tnt_resident(h, rank, p, &tr);
    const double tnt_s = since(ta);
    total_cg += tr.total_inner;
    eta = std::min(std::max(tr.f * 5e-6, 1e-7), 1e-1);  // :154
    auto tb = clk::now();
    cert = certify_resident(h, rank, eta, 500, verbose);
    h->last_cert_branch = cert.branch;
    const double cert_s = since(tb);
    if (verbose)
      std::printf("rank %d: f=%.9e |g|=%.3e status=%d outer=%d cg=%lld certified=%d theta=%.3e eta=%.3e (tnt %.3fs, cert %.3fs)\n",
                  rank, tr.f, tr.gradfx_norm, tr.status, tr.num_outer, (long long)tr.total_inner, (int)cert.certified,
                  cert.theta, eta, tnt_s, cert_s);
    record(tnt_s, cert_s);
    if (std::isnan(cert.theta)) throw Error(CORA_B200_ERUNTIME, "Theta is NaN");
    res->lifted_f = tr.f;
    res->lifted_rank = rank;
    res->certified = cert.certified ? 1 : 0;
    if (cert.certified) break;
    // no verdict and no descent direction (graphs without a factorisation of S + eta I): lifting the rank along an
    // unverified vector could end in a spurious sv-ratio "certificate" at the next rank -- stop and round instead
    if (cert.branch == CORA_B200_CERT_INCONCLUSIVE) break;
    if (rank + 1 > max_rank || rank + 1 > kMaxGeomRank) break;  // staircase exhausted: round the current iterate
    ++rank;  // problem.incrementRank()
    saddle_escape_resident(h, rank, cert.theta, 1e-4, 1e-4, verbose);
  }
  const int cur = h->resident_r;
  if (cur > d) {  // :200-233
    project_solution_resident(h, cur);
    rank = d;
    auto ta = clk::now();
    std::memset(&tr, 0, sizeof(tr));
    tnt_resident(h, d, p, &tr);
    const double tnt_s = since(ta);
    total_cg += tr.total_inner;
    eta = std::min(std::max(tr.f * 5e-6, 1e-7), 1e-1);
    auto tb = clk::now();
    cert = certify_resident(h, d, eta, 500, verbose);
    const double cert_s = since(tb);
    if (verbose)
      std::printf("refine rank %d: f=%.9e |g|=%.3e status=%d outer=%d cg=%lld certified=%d theta=%.3e\n", d, tr.f,
                  tr.gradfx_norm, tr.status, tr.num_outer, (long long)tr.total_inner, (int)cert.certified, cert.theta);
    record(tnt_s, cert_s);
    res->refined_certified = cert.certified ? 1 : 0;
  } else {
    res->refined_certified = res->certified;
  }
  res->f = tr.f;
  res->final_rank = h->resident_r;
  res->num_stages = ns;
  res->total_cg_iterations = total_cg;
  if (h->resident_r != d) {
    // certified at rank d (no rounding needed) or staircase exhausted: return the leading d columns
    // only when the rank is d; otherwise the caller receives the rank-d rounding
    project_solution_resident(h, h->resident_r);
  }
  export_matrix(h, h->ws[V_X].p, d, X_out);
  res->seconds = since(t0);
}

}  // namespace cora_b200
