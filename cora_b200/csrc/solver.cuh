// solver.cuh -- device-resident STPCG and the TNT outer loop.
//
//   STPCG  libs/Optimization/include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426
//   TNT    libs/Optimization/include/Optimization/Riemannian/TNT.h:242-689
//   closures (f, QM, metric, retract, precon): src/CORA.cpp:52-122
//
// One CG iteration is three launches (k_qprod<HESS>, k_cg_update, k_cg_pupdate); all
// scalars of the recurrences live in CgCtrl on the device and are advanced by the last
// CTA of each reduction, so the host only polls a "done" flag once per chunk of
// iterations (with one chunk of look-ahead so the GPU never idles).
#pragma once
#include <cmath>

#include "chain_chol.cuh"
#include "ops.cuh"
#include "persistent_kernel.cuh"

namespace cora_b200 {

// Z = M^-1 V  (Problem::precondition, src/CORA_problem.cpp:869-903), no projection.
inline void apply_preconditioner(H *h, const double *V, double *Z, int r, const CgCtrl *ctrl) {
  const long long nE = (long long)h->DL.N * r;
  if (h->precond == CORA_B200_PRECON_JACOBI) {
    k_jacobi<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(h->DL.dinv, V, Z, r, nE);
    check_launch(h);
  } else if (h->precond == CORA_B200_PRECON_REG_CHOLESKY) {
    if (!h->chol) throw Error(CORA_B200_ERUNTIME, "RegularizedCholesky factor missing");
    chain_solve(h, h->chol, V, Z, r, ctrl);
  } else {
    throw Error(CORA_B200_EINVAL, "The desired preconditioner is not implemented");  // :892-894
  }
  // Formulation::Implicit (:878-885): V is lifted with zero translations (its translation rows ARE zero here),
  // solved with the full factor, and only the rotation / range rows of the result are kept
  if (h->formulation == CORA_B200_FORMULATION_IMPLICIT && h->precond != CORA_B200_PRECON_JACOBI)
    zero_translation_rows(h, Z, r, ctrl);
}

// V = proj_Y(M^-1 R) with <R,V>, <V,V> -> scal[slot..slot+1]   (src/CORA.cpp:89-92)
inline void precondition_project(H *h, const double *Y, double *R, double *Vout, int r, int slot) {
  UArgs A{};
  A.Y = Y; A.R = R; A.V = Vout; A.r = r; A.do_axpy = 0; A.do_proj = 1; A.post = POST_STORE; A.slot = slot;
  A.gated = 0; A.ctrl = nullptr;
  if (h->precond == CORA_B200_PRECON_JACOBI) {
    A.zsrc = 0;
  } else {
    apply_preconditioner(h, R, h->ws[V_Z].p, r, nullptr);
    A.zsrc = 2;
    A.Z = h->ws[V_Z].p;
  }
  launch_update(h, A);
}

inline void update_preconditioner(H *h) {
  release_chain_chol(h, h->chol);
  h->chol = nullptr;
  h->precond_requested = h->precond;
  if (h->precond == CORA_B200_PRECON_REG_CHOLESKY) {
    if (!h->lambda_user) h->lambda_reg = estimate_spectral_norm(h) / (h->reg_max_cond - 1.0);  // :556-591
    bool pd = false;
    try {
      h->chol = build_chain_chol(h, h->d_bval.p, h->d_sdiag.p, h->lambda_reg, /*pin_last=*/true, &pd);
    } catch (const Error &e) {
      if (e.code != CORA_B200_ENOTIMPL) throw;
      // A coupling neither the chain nor the general pose-graph factorisation covers (a range row tied to a
      // rotation variable: not produced by any CORA measurement type).  RegularizedCholesky is the reference's
      // DEFAULT, so a drop-in caller must still get a working handle: apply Jacobi instead and report it
      // (cora_b200_effective_preconditioner).
      h->precond = CORA_B200_PRECON_JACOBI;
      return;
    }
    if (!pd) throw Error(CORA_B200_ERUNTIME, "RegularizedCholesky: Q + lambda I is not positive definite");
  }
}

inline void compute_lambda(H *h, const double *X, int r) {
  launch_qprod(h, QM_SPMM, X, nullptr, nullptr, h->ws[V_G].p, nullptr, r, POST_STORE, SC_TMP, nullptr);
  const size_t ns = (size_t)h->DL.n * h->DL.d * h->DL.d;
  if (h->d_lam_st.n < std::max<size_t>(ns, 1)) h->d_lam_st.alloc(std::max<size_t>(ns, 1));
  if (h->d_lam_ob.n < (size_t)std::max(h->DL.m, 1)) h->d_lam_ob.alloc((size_t)std::max(h->DL.m, 1));
  const int tot = h->DL.n + h->DL.m;
  if (tot > 0) {
    DISPATCH_D(h, k_lambda<DD><<<(tot + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(
                      h->DL, X, h->ws[V_G].p, h->d_lam_st.p, h->d_lam_ob.p, r));
    check_launch(h);
  }
}

// Values of S + eta I on Q's structure (needs compute_lambda first).
inline DevLayout build_certificate_layout(H *h, double eta) {
  const size_t nb = (size_t)h->HL.tile_boff[h->HL.numTiles];
  const size_t ns = (size_t)h->DL.l + h->DL.m;
  if (h->d_bvalS.n < std::max<size_t>(nb, 1)) h->d_bvalS.alloc(std::max<size_t>(nb, 1));
  if (h->d_sdiagS.n < std::max<size_t>(ns, 1)) h->d_sdiagS.alloc(std::max<size_t>(ns, 1));
  if (nb) CUDA_CHECK(cudaMemcpyAsync(h->d_bvalS.p, h->d_bval.p, nb * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  const int tot = h->DL.n + h->DL.l + h->DL.m;
  if (tot > 0) {
    DISPATCH_D(h, k_patch_certificate<DD><<<(tot + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(
                      h->DL, h->d_lam_st.p, h->d_lam_ob.p, eta, h->d_bvalS.p, h->d_sdiagS.p));
    check_launch(h);
  }
  DevLayout LS = h->DL;
  LS.bval = h->d_bvalS.p;
  LS.sdiag = h->d_sdiagS.p;
  return LS;
}

// ------------------------------------------------------------------- STPCG -----
struct StpcgOut {
  double hM = 0.0;
  int iterations = 0;
  int exit_reason = 0;
};

inline void enqueue_cg_iteration(H *h, int r) {
  double *X = h->ws[V_X].p, *G = h->ws[V_G].p, *P = h->ws[V_P].p, *HP = h->ws[V_HP].p;
  CgCtrl *ctrl = h->d_ctrl.p;
  launch_qprod(h, QM_HESS, P, X, G, HP, nullptr, r, POST_CG_HESS, 0, ctrl);
  UArgs A{};
  A.Y = X; A.P = P; A.HP = HP; A.S = h->ws[V_S].p; A.R = h->ws[V_R].p; A.V = h->ws[V_V].p;
  A.ctrl = ctrl; A.r = r; A.gated = 1; A.post = POST_CG_UPDATE;
  if (h->precond == CORA_B200_PRECON_JACOBI) {
    A.do_axpy = 1; A.zsrc = 0; A.do_proj = 1;
    launch_update(h, A);
  } else {
    A.do_axpy = 1; A.do_proj = 0;
    launch_update(h, A);
    apply_preconditioner(h, h->ws[V_R].p, h->ws[V_Z].p, r, ctrl);
    A.do_axpy = 0; A.zsrc = 2; A.Z = h->ws[V_Z].p; A.do_proj = 1;
    launch_update(h, A);
  }
  const long long nE = (long long)h->DL.N * r;
  k_cg_pupdate<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(ctrl, h->ws[V_V].p, P, nE);
  check_launch(h);
}

// Solve for the step S at the current iterate (V_X, V_G, V_GRAD, V_PG = P(grad)).
inline StpcgOut run_stpcg(H *h, int r, double Delta, const cora_b200_tnt_params &p, int rv_slot) {
  if (!(Delta > 0)) throw Error(CORA_B200_EINVAL, "Trust-region radius (Delta) must be a positive real value");
  const long long nE = (long long)h->DL.N * r;
  const int max_it = p.max_TPCG_iterations;
  k_cg_init<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(h->d_ctrl.p, h->d_scal.p, rv_slot, Delta, max_it,
                                                         p.kappa_fgr, p.theta, 1e-8, h->ws[V_GRAD].p,
                                                         h->ws[V_PG].p, h->ws[V_S].p, h->ws[V_R].p,
                                                         h->ws[V_P].p, nE);
  check_launch(h);
  int launched = 0, nchunks = 0;
  auto enqueue_chunk = [&]() {
    for (int i = 0; i < h->cg_chunk && launched < max_it; ++i, ++launched) enqueue_cg_iteration(h, r);
    const int b = nchunks & 1;
    CUDA_CHECK(cudaMemcpyAsync(&h->h_ctrl[b], h->d_ctrl.p, sizeof(CgCtrl), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaEventRecord(h->ev_chunk[b], h->stream));
    ++nchunks;
  };
  enqueue_chunk();
  int k = 0;
  CgCtrl fin;
  for (;;) {
    if (nchunks == k + 1 && launched < max_it) enqueue_chunk();  // one chunk of look-ahead
    CUDA_CHECK(cudaEventSynchronize(h->ev_chunk[k & 1]));
    fin = h->h_ctrl[k & 1];
    if (fin.state != 0) break;
    ++k;
    if (k >= nchunks) {
      if (launched >= max_it) throw Error(CORA_B200_ERUNTIME, "STPCG did not terminate (internal error)");
      enqueue_chunk();
    }
  }
  if (nchunks > k + 1) CUDA_CHECK(cudaEventSynchronize(h->ev_chunk[(k + 1) & 1]));  // drain look-ahead
  StpcgOut out;
  out.iterations = fin.it;
  out.hM = fin.hM;
  out.exit_reason = fin.exit_reason;
  if (fin.exit_reason == CG_EXIT_KERNEL) {
    // IterativeSolvers.h:305-338, finished from the host: p lies in ker(H)
    launch_dot2(h, h->ws[V_P].p, h->ws[V_R].p, nullptr, nullptr, nE, SC_TMP);
    read_scal(h);
    double sMp = fin.sMp, sgn = 1.0;
    if (h->h_scal[SC_TMP] < 0) { sgn = -1.0; sMp = -sMp; }
    const double sigma = (-sMp + std::sqrt(sMp * sMp + fin.pM2 * (fin.Delta2 - fin.sM2))) / fin.pM2;
    launch_axpby(h, 1.0, h->ws[V_S].p, sigma * sgn, h->ws[V_P].p, h->ws[V_S].p, nE);
    out.hM = fin.Delta;
  }
  return out;
}

// --------------------------------------------------------------------- TNT -----
struct TraceWriter {
  cora_b200_tnt_result *res;
  int n_state = 0, n_iter = 0;
  void state(double t, double f, double g, double pg, double Delta) {
    if (n_state < res->trace_capacity) {
      if (res->time) res->time[n_state] = t;
      if (res->objective_values) res->objective_values[n_state] = f;
      if (res->gradient_norms) res->gradient_norms[n_state] = g;
      if (res->preconditioned_gradient_norms) res->preconditioned_gradient_norms[n_state] = pg;
      if (res->trust_region_radius) res->trust_region_radius[n_state] = Delta;
    }
    ++n_state;
  }
  void iter(int inner, double hn, double hM, double rho) {
    if (n_iter < res->trace_capacity) {
      if (res->inner_iterations) res->inner_iterations[n_iter] = inner;
      if (res->update_step_norms) res->update_step_norms[n_iter] = hn;
      if (res->update_step_M_norms) res->update_step_M_norms[n_iter] = hM;
      if (res->gain_ratios) res->gain_ratios[n_iter] = rho;
    }
    ++n_iter;
  }
};

inline void swap_vec(H *h, int a, int b) {
  std::swap(h->ws[a].p, h->ws[b].p);
  std::swap(h->ws[a].n, h->ws[b].n);
}

// ------------------------------------------------------- persistent TNT (host) ---
// dynamic shared memory of the tile pipeline (persistent.cuh); with_q = false: only the dense-vector slots
// (retraction, hub scratch, chain apply), the data-matrix slice buffers are not needed
template <int D>
inline size_t persistent_smem(const H *h, int r, int nbuf, bool with_q) {
  const int D1 = D + 1;
  const size_t nbv = (size_t)h->DL.maxSlots * D1 * D1 * h->DL.TP;
  const size_t ncol = ((size_t)h->DL.maxSlots * h->DL.TP + 3) & ~(size_t)3;
  const size_t spcap = ((size_t)h->DL.maxTileSpill + 3) & ~(size_t)3;
  const size_t pstride = (size_t)D1 * r, hpad = (pstride + 1) & ~(size_t)1;
  const size_t vstride = (hpad + (size_t)h->DL.TR * r + pstride + 2 + 1) & ~(size_t)1;
  const size_t qbuf = with_q ? (nbv + spcap) * sizeof(double) + (ncol + h->DL.TRP + spcap) * sizeof(int) : 0;
  return 144 * sizeof(double) + nbuf * qbuf + (1 + 3 * (size_t)nbuf) * vstride * sizeof(double);
}

// grid / shared-memory configuration of the persistent kernels at rank r (cached per rank).  Ranks with a
// compiled streaming kernel (build.py: PK_LIST) run the data-matrix products and the preconditioned update on
// per-warp strip rings (stream.cuh), every other rank on the any-rank tile pipeline.
inline void persistent_configure(H *h, int r) {
  if (h->persistent_grid_r == r) return;
  const int d = h->DL.d, D1 = h->DL.D1;
  const int threads = h->persistent_threads, warps = threads / 32;
  void *kfn = h->allow_stream ? persistent_tnt_kernel(d, r) : nullptr;
  bool stream = kfn != nullptr && warps <= kStreamMaxWarps;
  size_t smem = 0;
  int nbuf = 2;
  int stage = 0, xw = 0, yw = 0;
  if (stream) {
    const StreamHost &S = h->SH;
    xw = ((S.SP + 2) * D1 * r + 2 + 1) & ~1;
    yw = 32 * r + 2;
    const int recd = ((S.max_rec_bytes + 15) / 16) * 2;
    stage = xw + yw + 128 + recd + (kRangeWindow * r + 2);
    stage = std::max(stage, 3 * yw + 36);
    stage = (stage + 1) & ~1;
    const size_t ring = (size_t)warps * h->stream_stages * stage * sizeof(double);
    const size_t lmc = S.lm_cache ? ((size_t)h->DL.l * r + 4) * sizeof(double) : 0;  // landmark cache behind the ring
    size_t vec = 0;
    DISPATCH_D(h, vec = persistent_smem<DD>(h, r, 2, false));
    smem = std::max(vec, 144 * sizeof(double) + ring + lmc);
    if (smem > 227 * 1024) stream = false;  // very large ranks / records: tile pipeline
  }
  if (!stream) {
    kfn = persistent_tnt_kernel(d, 0);
    if (!kfn) throw Error(CORA_B200_ERUNTIME, "persistent TNT kernel not compiled for this dimension");
    // double-buffered tile pipeline when two CTAs of it fit on an SM, single-buffered otherwise
    DISPATCH_D(h, smem = persistent_smem<DD>(h, r, 2, true));
    if (smem > 113 * 1024) nbuf = 1;
    DISPATCH_D(h, smem = persistent_smem<DD>(h, r, nbuf, true));
    if (smem > 227 * 1024)
      throw Error(CORA_B200_ERUNTIME, "persistent TNT kernel: tile working set exceeds shared memory at this rank");
    // the hub-row scratch of a tile lives in one vector slot
    const HostLayout &HL = h->HL;
    const size_t vstride = (size_t)HL.TR * r;
    for (int t = 0; t < HL.numTiles; ++t) {
      const int q0 = HL.tile_long_ptr[t], q1 = HL.tile_long_ptr[t + 1];
      if (q1 == q0) continue;
      const size_t hs = HL.long_grp[q0] < HL.n ? (size_t)HL.D1 * r : (size_t)r;
      if ((size_t)(q1 - q0) * hs > vstride)
        throw Error(CORA_B200_ERUNTIME, "persistent TNT kernel: too many hub rows in one tile for the shared-memory scratch");
    }
  }
  h->persistent_stream = stream;
  h->persistent_kfn = kfn;
  h->persistent_spmm_kfn = persistent_spmm_kernel(d, stream ? r : 0);
  h->persistent_nbuf = nbuf;
  h->persistent_smem = smem;
  h->stream_stage_doubles = stage; h->stream_xw = xw; h->stream_yw = yw;
  int per_sm = 0;
  CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem));
  if (per_sm < 1) throw Error(CORA_B200_ERUNTIME, "persistent TNT kernel does not fit on an SM at this rank");
  if (const char *e = getenv("CORA_B200_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
  const int G0 = std::max(1, std::min(h->sm_count * per_sm, h->DL.numTiles));
  h->persistent_grid = G0;
  h->persistent_grid_r = r;
  {  // cost-balanced contiguous partition of the tiles: scalar-row tiles (two L2 gathers per element) weigh more
    const double ws = 0.4;
    const HostLayout &HL = h->HL;
    std::vector<double> cost(HL.numTiles);
    double total = 0.0;
    for (int t = 0; t < HL.numTiles; ++t) {
      const int64_t row0 = (int64_t)t * HL.TR;
      const int nR = (int)std::min<int64_t>(HL.TR, HL.N - row0);
      const int nP = (int)std::max<int64_t>(0, std::min<int64_t>(HL.TP, (int64_t)HL.n - (int64_t)t * HL.TP));
      const int nS = nR - nP * HL.D1;
      cost[t] = 1.0 + ws * (double)nS / HL.TR;
      total += cost[t];
    }
    std::vector<int> t0(G0 + 1, HL.numTiles);
    t0[0] = 0;
    double accum = 0.0;
    int b = 1;
    for (int t = 0; t < HL.numTiles && b < G0; ++t) {
      accum += cost[t];
      while (b < G0 && accum >= total * b / G0) t0[b++] = t + 1;
    }
    for (int i = 1; i <= G0; ++i) t0[i] = std::max(t0[i], t0[i - 1]);
    t0[G0] = HL.numTiles;
    h->d_cta_t0.upload(t0, h->stream);
    if (stream) {
      std::vector<int32_t> ws0;
      partition_strips(h->SH, h->stream_interleave ? G0 : G0 * warps, h->stream_scalar_weight, ws0);
      std::vector<int> wsi(ws0.begin(), ws0.end());
      h->d_warp_strip.upload(wsi, h->stream);  // cudaMalloc: 256-byte aligned, read as int4
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
}

inline void fill_stream_args(const H *h, StreamDev &sd) {
  const StreamHost &S = h->SH;
  sd.SP = S.SP; sd.CP = S.CP; sd.GP = S.GP; sd.nPS = S.nPS; sd.nSS = S.nSS; sd.nStrips = S.nStrips;
  sd.nstage = h->stream_stages;
  sd.interleave = h->stream_interleave ? 1 : 0;
  sd.stage_doubles = h->stream_stage_doubles;
  sd.xw_off = 0; sd.yw_off = h->stream_xw; sd.dg_off = h->stream_xw + h->stream_yw; sd.rc_off = sd.dg_off + 128;
  sd.gw_off = sd.rc_off + ((S.max_rec_bytes + 15) / 16) * 2;
  sd.ring_base = 144;
  sd.lm_rows = S.lm_cache ? h->DL.l : 0;
  sd.lm_base = 144 + (h->persistent_threads / 32) * h->stream_stages * h->stream_stage_doubles;
  sd.info = reinterpret_cast<const uint4 *>(h->d_rec_off.p); sd.rec = h->d_rec.p;
  sd.diagQ = h->d_diagQ.p; sd.sdiagP = h->d_sdiagP.p;
  sd.diagL[0] = h->d_diagL.p; sd.diagL[1] = h->d_diagL.p + h->d_diagL.n / 2;
  sd.sdiagL[0] = h->d_sdiagL.p; sd.sdiagL[1] = h->d_sdiagL.p + h->d_sdiagL.n / 2;
  sd.warp_strip = reinterpret_cast<const int4 *>(h->d_warp_strip.p);
}

inline void tnt_persistent(H *h, int r, const cora_b200_tnt_params &p, cora_b200_tnt_result *res) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  const int64_t launches0 = h->launches;
  persistent_configure(h, r);
  const int G = h->persistent_grid;
  const size_t smem = h->persistent_smem;
  const size_t npart = 2 * ((size_t)G * kPPart + 8);
  if (h->d_ppartials.n < npart) h->d_ppartials.alloc(npart);
  const size_t nlp = 2 * (size_t)std::max(h->DL.numChunks, 1) * h->DL.D1 * h->ws_r;
  if (h->d_longpart.n < nlp) h->d_longpart.alloc(nlp);
  if (!h->d_bar.p) h->d_bar.alloc(2);
  if (!h->persistent_stream) {
    // two copies (current / proposal) of the block values with Q - Lambda on the diagonal blocks, and of
    // diag(Q) - lambda_k of the scalar rows; everything but those entries is Q and is copied once
    const size_t nl = std::max<size_t>((size_t)h->HL.tile_boff[h->HL.numTiles], 1), ns = (size_t)h->DL.l + h->DL.m + 1;
    if (h->d_lamT.n < 2 * nl) {
      h->d_lamT.alloc(2 * nl);
      h->d_lamS.alloc(2 * ns);
      for (int k = 0; k < 2; ++k) {
        CUDA_CHECK(cudaMemcpyAsync(h->d_lamT.p + k * nl, h->d_bval.p, nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(h->d_lamS.p + k * ns, h->d_sdiag.p, (ns - 1) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      }
    }
  }
  if (!h->d_tntdev.p) {
    h->d_tntdev.alloc(sizeof(TntDev));
    CUDA_CHECK(cudaMallocHost(&h->h_tntdev, sizeof(TntDev)));
  }
  const int cap = std::max(p.max_iterations + 2, 4);
  if (h->trace_cap < cap) {
    h->d_trace.alloc((size_t)TR_ROWS * cap);
    h->h_trace.resize((size_t)TR_ROWS * cap);
    h->trace_cap = cap;
  }
  PArgs A{};
  for (int i = 0; i < V_COUNT; ++i) A.v[i] = h->ws[i].p;
  A.longpart = h->d_longpart.p;
  A.partials = h->d_ppartials.p;
  A.bar = h->d_bar.p;
  A.trace = h->d_trace.p;
  A.out = (TntDev *)h->d_tntdev.p;
  A.p = p;
  A.r = r;
  A.trace_cap = h->trace_cap;
  A.precond = h->precond;
  A.nbuf = h->persistent_nbuf;
  A.cta_t0 = h->d_cta_t0.p;
  if (h->persistent_stream) fill_stream_args(h, A.sd);
  if (h->precond == CORA_B200_PRECON_REG_CHOLESKY) {
    ChainChol *C = h->chol;
    if (!C) throw Error(CORA_B200_ERUNTIME, "RegularizedCholesky factor missing");
    chain_ensure_ws(h, C, r);
    ChainDev &cd = A.chain;
    cd.nl = (int)C->levels.size();
    if (cd.nl > kMaxChainLevels) throw Error(CORA_B200_ERUNTIME, "chain factor has too many levels");
    cd.n = C->host.n; cd.l = C->host.l; cd.m = C->host.m;
    cd.smem_doubles = (int)(smem / sizeof(double));
    cd.pinned_pose_row = C->host.pinned_pose_row; cd.pinned_landmark = C->host.pinned_landmark;
    for (int lv = 0; lv < cd.nl; ++lv) {
      ChainLevelDev *Dv = C->levels[lv];
      cd.G[lv] = Dv->G; cd.fwd[lv] = Dv->fwd.p; cd.bwd[lv] = Dv->bwd.p; cd.UR[lv] = Dv->UR.p;
      cd.sol[lv] = Dv->sol.p; cd.rhs[lv] = Dv->rhs.p; cd.cL[lv] = Dv->cL.p; cd.cR[lv] = Dv->cR.p;
    }
    cd.rinc_ptr = C->rinc_ptr.p; cd.rinc_k = C->rinc_k.p; cd.rend_x = C->rend_x.p; cd.bl_ptr = C->bl_ptr.p;
    cd.bl_row = C->bl_row.p; cd.rinc_e = C->rinc_e.p; cd.rdinv = C->rdinv.p; cd.rend_e = C->rend_e.p;
    cd.bl_val = C->bl_val.p; cd.W = C->W.p; cd.SLinv = C->SLinv.p; cd.u = C->u.p; cd.Y = C->Y.p;
    {  // chunked landmark reductions
      const ChainFactorHost &F = C->host;
      int mr = 1, mb = 1;
      for (int j = 0; j < cd.l; ++j) {
        mr = std::max(mr, (int)((F.rinc_ptr[cd.n + j + 1] - F.rinc_ptr[cd.n + j] + kLmChunk - 1) / kLmChunk));
        mb = std::max(mb, (int)((F.bl_ptr[j + 1] - F.bl_ptr[j] + kLmChunk - 1) / kLmChunk));
      }
      cd.max_rinc_chunks = mr; cd.max_bl_chunks = mb;
      const size_t need = (size_t)std::max(cd.l, 1) * (mr + mb) * h->ws_r;
      if (h->d_lmpart.n < need) h->d_lmpart.alloc(need);
      cd.part_pre = h->d_lmpart.p;
      cd.part_bl = h->d_lmpart.p + (size_t)std::max(cd.l, 1) * mr * h->ws_r;
      if ((size_t)cd.n * h->DL.D1 * r >= (1ull << 31) || ((size_t)h->DL.N * r) >= (1ull << 31))
        throw Error(CORA_B200_ERUNTIME, "persistent chain apply: problem too large for 32-bit element indices");
    }
    const size_t vstride = (size_t)h->DL.TR * r;
    if (2 * (size_t)cd.l * r > 6 * vstride)
      throw Error(CORA_B200_ERUNTIME, "persistent TNT kernel: too many landmarks for the shared-memory border solve");
  }
  if (!h->persistent_stream) {
    A.lam[0] = h->d_lamT.p; A.lam[1] = h->d_lamT.p + h->d_lamT.n / 2;
    A.lamS[0] = h->d_lamS.p; A.lamS[1] = h->d_lamS.p + h->d_lamS.n / 2;
  }
  const bool phase_prof = getenv("CORA_B200_PHASE_PROFILE") != nullptr;
  if (h->d_prof_all.n < (size_t)G * PH_COUNT) h->d_prof_all.alloc((size_t)G * PH_COUNT);
  A.prof_all = h->d_prof_all.p;
  A.calibrate = phase_prof ? 1 : 0;
  CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  CUDA_CHECK(cudaMemsetAsync(h->d_bar.p, 0, 2 * sizeof(unsigned long long), h->stream));
  DevLayout Lc = h->DL;
  void *args[] = {(void *)&Lc, (void *)&A};
  // the attribute belongs to the function, not to this handle: set it before every launch
  CUDA_CHECK(cudaFuncSetAttribute(h->persistent_kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaLaunchCooperativeKernel(h->persistent_kfn, dim3(G), dim3(h->persistent_threads), args, smem, h->stream));
  check_launch(h);
  CUDA_CHECK(cudaMemcpyAsync(h->h_tntdev, h->d_tntdev.p, sizeof(TntDev), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->h_trace.data(), h->d_trace.p, (size_t)TR_ROWS * h->trace_cap * sizeof(double),
                             cudaMemcpyDeviceToHost, h->stream));
  h->h_prof_all.resize((size_t)G * PH_COUNT);
  h->prof_all_grid = G;
  CUDA_CHECK(cudaMemcpyAsync(h->h_prof_all.data(), h->d_prof_all.p, h->h_prof_all.size() * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  CUDA_CHECK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  const TntDev &o = *(const TntDev *)h->h_tntdev;
  {  // mirror the kernel's buffer rotation
    double *ptr[V_COUNT];
    size_t cnt[V_COUNT];
    for (int i = 0; i < V_COUNT; ++i) { ptr[i] = h->ws[o.perm[i]].p; cnt[i] = h->ws[o.perm[i]].n; }
    for (int i = 0; i < V_COUNT; ++i) { h->ws[i].p = ptr[i]; h->ws[i].n = cnt[i]; }
  }
  const int cap_d = h->trace_cap;
  const double *T = h->h_trace.data();
  const int ns = std::min(o.n_state, res->trace_capacity), ni = std::min(o.num_outer, res->trace_capacity);
  for (int i = 0; i < ns; ++i) {
    if (res->time) res->time[i] = T[(size_t)TR_TIME * cap_d + i];
    if (res->objective_values) res->objective_values[i] = T[(size_t)TR_F * cap_d + i];
    if (res->gradient_norms) res->gradient_norms[i] = T[(size_t)TR_G * cap_d + i];
    if (res->preconditioned_gradient_norms) res->preconditioned_gradient_norms[i] = T[(size_t)TR_PG * cap_d + i];
    if (res->trust_region_radius) res->trust_region_radius[i] = T[(size_t)TR_DELTA * cap_d + i];
  }
  for (int i = 0; i < ni; ++i) {
    if (res->inner_iterations) res->inner_iterations[i] = (int32_t)T[(size_t)TR_INNER * cap_d + i];
    if (res->update_step_norms) res->update_step_norms[i] = T[(size_t)TR_HNORM * cap_d + i];
    if (res->update_step_M_norms) res->update_step_M_norms[i] = T[(size_t)TR_HM * cap_d + i];
    if (res->gain_ratios) res->gain_ratios[i] = T[(size_t)TR_RHO * cap_d + i];
  }
  if (p.verbose) {
    for (int i = 0; i < o.num_outer && i + 1 < cap_d; ++i)
      std::printf("Iter: %4d, time: %.3e, f: %.8e, |g|: %.3e, |M^{-1}g|: %.3e, Delta: %.3e, inner iters: %3d, "
                  "|h|: %.3e, |h|_M: %.3e, df: %.6e, rho: %.3e. %s\n",
                  i, T[(size_t)TR_TIME * cap_d + i], T[(size_t)TR_F * cap_d + i], T[(size_t)TR_G * cap_d + i],
                  T[(size_t)TR_PG * cap_d + i], T[(size_t)TR_DELTA * cap_d + i], (int)T[(size_t)TR_INNER * cap_d + i],
                  T[(size_t)TR_HNORM * cap_d + i], T[(size_t)TR_HM * cap_d + i],
                  T[(size_t)TR_F * cap_d + i] - T[(size_t)TR_F * cap_d + i + 1], T[(size_t)TR_RHO * cap_d + i],
                  T[(size_t)TR_RHO * cap_d + i] > p.eta1 ? "Step accepted" : "Step REJECTED!");
    std::printf("\n");
  }
  if (getenv("CORA_B200_PHASE_PROFILE")) {
    static const char *names[PH_COUNT] = {"hub", "grad", "hess", "update", "pupdate", "retract", "precond", "cginit", "sync", "misc", "q.wait", "q.qx", "q.epi", "q.store", "ch.pre", "ch.fwd", "ch.bwd", "ch.border", "ch.post", "smid", "reduce"};
    std::printf("[persistent] %s grid %d, nbuf %d, smem %zu, barriers %lld, outer %d, CG %lld, device %.3f ms\n", h->persistent_stream ? "stream" : "tile", G, h->persistent_nbuf, smem, o.barriers, o.num_outer, o.total_inner, ms);
    for (int i = 0; i < PH_COUNT; ++i)
      if (o.prof_cnt[i]) std::printf("  %-8s n=%6u total %9.1f us  avg %8.2f us\n", names[i], o.prof_cnt[i], o.prof_ns[i] * 1e-3, o.prof_ns[i] * 1e-3 / o.prof_cnt[i]);
    const std::vector<unsigned long long> &pa = h->h_prof_all;
    for (int i = 0; i < PH_COUNT; ++i) {
      if (!o.prof_cnt[i]) continue;
      std::vector<double> col(G);
      for (int b = 0; b < G; ++b) col[b] = pa[(size_t)b * PH_COUNT + i] * 1e-3 / o.prof_cnt[i];
      std::vector<double> srt(col);
      std::sort(srt.begin(), srt.end());
      int amax = 0;
      for (int b = 0; b < G; ++b) if (col[b] > col[amax]) amax = b;
      std::printf("  per-CTA avg %-8s min %7.2f  med %7.2f  p90 %7.2f  max %7.2f (cta %d)  last-cta %7.2f\n", names[i], srt[0], srt[G / 2], srt[(G * 9) / 10], srt[G - 1], amax, col[G - 1]);
      if (atoi(getenv("CORA_B200_PHASE_PROFILE")) >= 2 && (i == PH_HESS || i == PH_UPDATE)) {
        std::printf("    all CTAs (time@sm):");
        for (int b = 0; b < G; ++b) std::printf("%s%.1f@%d", b % 12 == 0 ? "\n     " : " ", col[b], (int)pa[(size_t)b * PH_COUNT + PH_SMID]);
        std::printf("\n");
      }
    }
  }
  res->f = o.f;
  res->gradfx_norm = o.gnorm;
  res->preconditioned_gradfx_norm = o.pgnorm;
  res->elapsed_time = std::chrono::duration<double>(clk::now() - t0).count();
  res->device_time = ms * 1e-3;
  res->status = o.status;
  res->num_outer = o.num_outer;
  res->total_inner = o.total_inner;
  res->kernel_launches = h->launches - launches0;
  h->resident_r = r;
}

// reps x (out = Q X) through the tile pipeline of the persistent kernel; returns the CUDA-event milliseconds
inline float spmm_persistent(H *h, int r, const double *X, double *out, int reps, bool sync = true) {
  persistent_configure(h, r);
  const int G = h->persistent_grid;
  const size_t nlp = 2 * (size_t)std::max(h->DL.numChunks, 1) * h->DL.D1 * h->ws_r;
  if (h->d_longpart.n < nlp) h->d_longpart.alloc(nlp);
  if (!h->d_bar.p) h->d_bar.alloc(2);
  PArgs A{};
  A.longpart = h->d_longpart.p;
  A.bar = h->d_bar.p;
  A.r = r;
  A.nbuf = h->persistent_nbuf;
  A.cta_t0 = h->d_cta_t0.p;
  if (h->persistent_stream) fill_stream_args(h, A.sd);
  CUDA_CHECK(cudaMemsetAsync(h->d_bar.p, 0, sizeof(unsigned long long), h->stream));
  DevLayout Lc = h->DL;
  void *args[] = {(void *)&Lc, (void *)&A, (void *)&X, (void *)&out, (void *)&reps};
  CUDA_CHECK(cudaFuncSetAttribute(h->persistent_spmm_kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->persistent_smem));
  CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  CUDA_CHECK(cudaLaunchCooperativeKernel(h->persistent_spmm_kfn, dim3(G), dim3(h->persistent_threads), args,
                                         h->persistent_smem, h->stream));
  check_launch(h);
  if (!sync) return 0.f;  // enqueued only (peer_product.cuh chains it between its exchange kernels)
  CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  CUDA_CHECK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  return ms;
}

inline void tnt_resident(H *h, int r, const cora_b200_tnt_params &p, cora_b200_tnt_result *res) {
  check_geom_rank(r);
  if (h->precond != CORA_B200_PRECON_JACOBI && h->precond != CORA_B200_PRECON_REG_CHOLESKY)
    throw Error(CORA_B200_EINVAL, "The desired preconditioner is not implemented");
  ensure_workspace(h, r);
  // the persistent kernel applies Jacobi or the CHAIN factor in its own phases; the general sparse factor of graphs
  // with loop closures / several robots (gen_chol_dev.cuh) is applied by level-scheduled launches on the multi-launch path
  const bool general_factor = h->precond == CORA_B200_PRECON_REG_CHOLESKY && h->chol && h->chol->general;
  if (h->use_persistent && !general_factor && h->formulation == CORA_B200_FORMULATION_EXPLICIT) {
    tnt_persistent(h, r, p, res);
    return;
  }
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto elapsed = [&]() { return std::chrono::duration<double>(clk::now() - t0).count(); };
  const int64_t launches0 = h->launches;
  CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  TraceWriter tw{res};
  const double sqrt_eps = std::sqrt(2.220446049250313e-16);
  auto v = [&](int i) { return h->ws[i].p; };

  // TNT.h:372-392: f(x), QM(x), gradient norms
  launch_qprod(h, QM_GRAD, v(V_X), v(V_X), nullptr, v(V_GRAD), v(V_G), r, POST_STORE, SC_XG, nullptr);
  int rv_slot = SC_RV, rv_slot_prop = SC_RV2;
  precondition_project(h, v(V_X), v(V_GRAD), v(V_PG), r, rv_slot);
  read_scal(h);
  double fx = 0.5 * h->h_scal[SC_XG];
  double gnorm = std::sqrt(h->h_scal[SC_GG]);
  double pgnorm = std::sqrt(h->h_scal[rv_slot + 1]);
  double Delta = p.Delta0;
  int status = CORA_B200_TNT_ITERATION_LIMIT;
  int64_t total_inner = 0;
  int iteration = 0;
  for (; iteration < p.max_iterations; ++iteration) {
    const double el = elapsed();
    if (p.max_computation_time > 0 && el > p.max_computation_time) {  // TNT.h:447-452
      status = CORA_B200_TNT_ELAPSED_TIME;
      break;
    }
    tw.state(el, fx, gnorm, pgnorm, Delta);
    if (p.verbose)
      std::printf("Iter: %4d, time: %.3e, f: %.8e, |g|: %.3e, |M^{-1}g|: %.3e", iteration, el, fx, gnorm, pgnorm);
    if (gnorm < p.gradient_tolerance) { status = CORA_B200_TNT_GRADIENT; break; }  // :474-481
    if (pgnorm < p.preconditioned_gradient_tolerance) { status = CORA_B200_TNT_PRECONDITIONED_GRADIENT; break; }

    const StpcgOut cg = run_stpcg(h, r, Delta, p, rv_slot);  // :489-492
    total_inner += cg.iterations;
    // proposed point, its objective / gradient, model decrease (TNT.h:503-512); the
    // preconditioned gradient at the proposal is computed speculatively so that one
    // host synchronisation per outer iteration suffices
    launch_retract(h, v(V_X), v(V_S), 1.0, v(V_GRAD), v(V_XP), r, SC_HH);
    launch_qprod(h, QM_GRAD, v(V_XP), v(V_XP), nullptr, v(V_GRADP), v(V_GP), r, POST_STORE, SC_XG2, nullptr);
    launch_qprod(h, QM_HESS, v(V_S), v(V_X), v(V_G), v(V_HP), nullptr, r, POST_STORE, SC_HHH, nullptr);
    precondition_project(h, v(V_XP), v(V_GRADP), v(V_T0), r, rv_slot_prop);
    read_scal(h);
    const double hnorm = std::sqrt(h->h_scal[SC_HH]);
    const double fxp = 0.5 * h->h_scal[SC_XG2];
    const double dm = -h->h_scal[SC_GH] - 0.5 * h->h_scal[SC_HHH];
    const double df = fx - fxp;
    const double rel = df / (sqrt_eps + std::fabs(fx));
    const double rho = df / dm;
    const bool accepted = !std::isnan(rho) && rho > p.eta1;  // :532
    tw.iter(cg.iterations, hnorm, cg.hM, rho);
    if (p.verbose)
      std::printf(", Delta: %.3e, inner iters: %3d, |h|: %.3e, |h|_M: %.3e, df: %.6e, rho: %.3e. %s\n", Delta,
                  cg.iterations, hnorm, cg.hM, df, rho, accepted ? "Step accepted" : "Step REJECTED!");
    if (accepted) {
      swap_vec(h, V_X, V_XP);
      fx = fxp;
      if (rel < p.relative_decrease_tolerance) {  // :561-564 (gradient of the old point is reported)
        status = CORA_B200_TNT_RELATIVE_DECREASE;
        ++iteration;
        break;
      }
      if (hnorm < p.stepsize_tolerance) {  // :567-570
        status = CORA_B200_TNT_STEPSIZE;
        ++iteration;
        break;
      }
      swap_vec(h, V_G, V_GP);  // QM(x) :573
      swap_vec(h, V_GRAD, V_GRADP);
      swap_vec(h, V_PG, V_T0);
      std::swap(rv_slot, rv_slot_prop);
      gnorm = std::sqrt(h->h_scal[SC_GG2]);
      pgnorm = std::sqrt(h->h_scal[rv_slot + 1]);
    }
    if (!std::isnan(rho) && rho >= p.eta2) {  // :590-603
      Delta = std::max(p.alpha2 * cg.hM, Delta);
    } else if (std::isnan(rho) || rho < p.eta1) {
      Delta = p.alpha1 * cg.hM;
      if (Delta < p.Delta_tolerance) {
        status = CORA_B200_TNT_TRUST_REGION;
        ++iteration;
        break;
      }
    }
  }
  CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  CUDA_CHECK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  const double el = elapsed();
  tw.state(el, fx, gnorm, pgnorm, Delta);
  if (p.verbose) std::printf("\n");
  res->f = fx;
  res->gradfx_norm = gnorm;
  res->preconditioned_gradfx_norm = pgnorm;
  res->elapsed_time = el;
  res->device_time = ms * 1e-3;
  res->status = status;
  res->num_outer = tw.n_iter;
  res->total_inner = total_inner;
  res->kernel_launches = h->launches - launches0;
  h->resident_r = r;
}

}  // namespace cora_b200
