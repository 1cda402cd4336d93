// persistent_kernel.cuh -- k_tnt_persistent: the whole TNT solve in one cooperative launch
// (see persistent.cuh for the design and the phase functions, persistent_chain.cuh for the
// RegularizedCholesky apply).
#pragma once
#include "persistent.cuh"
#include "persistent_chain.cuh"
#include "persistent_reg.cuh"
#include "stream.cuh"

#ifndef CORA_PERSIST_THREADS
#define CORA_PERSIST_THREADS 256  // CTA size the persistent kernels are compiled for ...
#define CORA_PERSIST_MINB 2       // ... and resident CTAs per SM (register budget = 65536 / (threads * CTAs))
#endif

namespace cora_b200 {

struct PArgs {
  double *v[V_COUNT];
  double *longpart;             // 2 x numChunks x D1 x r
  double *partials;             // 2 x (G*kPPart + 8)
  unsigned long long *bar;      // grid barrier counter (zeroed by the host before the launch)
  double *trace;                // TR_ROWS x trace_cap
  TntDev *out;
  unsigned long long *prof_all;  // [G][PH_COUNT] per-CTA phase times
  int calibrate;                 // 1: time 16 empty grid barriers at kernel start (profiling runs)
  const int *cta_t0;             // [G+1] cost-balanced contiguous tile ranges
  double *lam[2];                // Lambda blocks sym(Y_i (QY)_i^T) per tile [a][b][pose] (current / proposal)
  double *lamS[2];               // lambda_k = (QY)_k . y_k per scalar row (0 for landmark rows)
  cora_b200_tnt_params p;
  int r, trace_cap, precond, nbuf;
  StreamDev sd;                  // strip layout + ring geometry (kernels compiled for a fixed rank, R > 0)
  ChainDev chain;                // RegularizedCholesky factor (precond == CORA_B200_PRECON_REG_CHOLESKY)
};

// Set-up shared by the persistent kernels: contiguous tile range of the CTA, carve-up of the dynamic shared
// memory (tile pipeline of persistent.cuh; with STREAM the data-matrix slice buffers are not needed: the
// products run on the strip rings), per-warp ring of the streaming phases.
template <int D, bool STREAM>
__device__ __forceinline__ void persistent_setup(const DevLayout &L, const PArgs &A, PCtx &c, double *smem,
                                                 TileMeta *s_tmeta, unsigned long long *s_mbar,
                                                 unsigned long long *s_rbar, Ring &rg) {
  constexpr int D1 = D + 1;
  const int r = A.r;
  c.t0 = A.cta_t0[c.b];
  c.t1 = A.cta_t0[c.b + 1];
  c.e0 = (long long)c.t0 * L.TR * r;
  c.e1 = min((long long)c.t1 * L.TR, (long long)L.N) * r;
  if (c.e0 > c.e1) c.e0 = c.e1;
  c.nbv = STREAM ? 0 : L.maxSlots * D1 * D1 * L.TP;              // doubles
  c.ncol = STREAM ? 0 : (L.maxSlots * L.TP + 3) & ~3;            // ints
  c.spcap = STREAM ? 0 : (L.maxTileSpill + 3) & ~3;              // entries
  c.TRP = STREAM ? 0 : L.TRP;
  c.pstride = D1 * r;
  c.hpad = (c.pstride + 1) & ~1;                              // halo in front of the tile rows, kept 16-byte aligned
  c.vstride = (c.hpad + L.TR * r + c.pstride + 2 + 1) & ~1;  // + halo behind, + one element of copy rounding
  c.nlam = D * D * L.TP;
  c.qstride = c.nbv + c.spcap + (c.ncol + c.TRP + c.spcap) / 2;  // doubles (int regions are multiples of 4)
  c.smem = smem;
  c.sred = smem;
  c.sbc = smem + 128;  // sred: up to 16 warps x 8 partial sums
  c.qbase = 144;
  const int after_q = c.qbase + c.nbuf * c.qstride;
  c.sW = smem + after_q;
  c.vbase = after_q + c.vstride;
  c.tmeta = s_tmeta;
  c.mbar = s_mbar;
  if (!STREAM)
    for (int i = c.tid; i < min(c.t1 - c.t0, kMaxTilesPerCta); i += c.nth) {
      const int t = c.t0 + i;
      TileMeta M;
      M.boff = L.tile_boff[t]; M.coff = L.tile_coff[t]; M.spoff = L.tile_sp_off[t];
      M.S = L.tile_slots[t]; M.nsp = L.tile_sp_cnt[t];
      M.lq0 = L.tile_long_ptr[t]; M.lq1 = L.tile_long_ptr[t + 1];
      s_tmeta[i] = M;
    }
  if (c.tid == 0) {
    mbar_init(&s_mbar[0], 1);
    mbar_init(&s_mbar[1], 1);
    if (STREAM)
      for (int i = 0; i < kStreamMaxWarps * kStreamMaxStages + 1; ++i) mbar_init(&s_rbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int warp = c.tid >> 5;
  rg.base = smem + A.sd.ring_base + (size_t)warp * A.sd.nstage * A.sd.stage_doubles;
  rg.bar = s_rbar + warp * kStreamMaxStages;
  rg.par = 0u;
  rg.inf = make_uint4(0u, 0u, 0u, 0u);
  rg.ro_batch = -1;
  c.lmbar = s_rbar + (STREAM ? kStreamMaxWarps * kStreamMaxStages : 0);
  c.lm_par = 0u;
}

// ============================================================ k_tnt_persistent ====
// R > 0: the kernel is compiled for rank R and the data-matrix products and the preconditioned update run as
// warp-autonomous streaming phases (stream.cuh); R == 0: any rank, tile pipeline of persistent.cuh.
template <int D, int R>
__global__ void __launch_bounds__(CORA_PERSIST_THREADS, CORA_PERSIST_MINB) k_tnt_persistent(const DevLayout L, const PArgs A) {
  constexpr int D1 = D + 1;
  constexpr bool STREAM = R > 0;
  extern __shared__ __align__(16) double smem[];
  __shared__ CgCtrl cg;
  __shared__ __align__(8) unsigned long long s_mbar[2];
  __shared__ __align__(8) unsigned long long s_rbar[STREAM ? kStreamMaxWarps * kStreamMaxStages + 1 : 1];
  __shared__ int s_meta[2][4];
  __shared__ TileMeta s_tmeta[STREAM ? 1 : kMaxTilesPerCta];
  __shared__ unsigned long long s_prof_ns[PH_COUNT], s_tph[2];
  __shared__ unsigned int s_prof_cnt[PH_COUNT];
  __shared__ double *s_v[V_COUNT];  // the work vectors by role; rotated by thread 0 (swp)
  __shared__ int s_perm[V_COUNT];
  PCtx c;
  c.b = blockIdx.x; c.G = gridDim.x; c.tid = threadIdx.x; c.nth = blockDim.x; c.r = A.r; c.nbuf = A.nbuf;
  c.bar = A.bar; c.target = 0; c.partials = A.partials; c.parity = 0; c.nbar = 0;
  c.prof_ns = s_prof_ns; c.prof_cnt = s_prof_cnt; c.tph = s_tph;
  c.mpar0 = c.mpar1 = 0u;
  const int r = STREAM ? R : A.r;
  Ring rg;
  {
    c.meta = &s_meta[0][0];
    persistent_setup<D, STREAM>(L, A, c, smem, s_tmeta, s_mbar, s_rbar, rg);
    if (c.tid < PH_COUNT) { s_prof_ns[c.tid] = 0; s_prof_cnt[c.tid] = 0; }
    if (c.tid < V_COUNT) { s_v[c.tid] = A.v[c.tid]; s_perm[c.tid] = c.tid; }
    if (c.tid == 0) s_tph[0] = s_tph[1] = 0;
    __syncthreads();
  }
  double *const *v = s_v;
  // rotate two roles; callers guarantee that every thread is past its last use of the old pointers
  auto swp = [&](int a, int b2) {
    __syncthreads();
    if (c.tid == 0) {
      double *tp = s_v[a]; s_v[a] = s_v[b2]; s_v[b2] = tp;
      const int ti = s_perm[a]; s_perm[a] = s_perm[b2]; s_perm[b2] = ti;
    }
    __syncthreads();
  };
  const cora_b200_tnt_params &P = A.p;
  const size_t lpstride = (size_t)max(L.numChunks, 1) * D1 * r;
  double *lp0 = A.longpart, *lp1 = A.longpart + lpstride;
  int lcur = 0;  // which copy of Q - Lambda belongs to the current iterate (the other one: the proposal)
  // gradient phase at X: out = grad, out2 = Q X, writes copy `wl` of Q - Lambda;  acc: <X,QX>, <grad,grad>
  auto grad_product = [&](const double *X, double *out, double *out2, const double *lp, int wl, double *acc) {
    if constexpr (STREAM)
      stream_qprod<D, R, QM_GRAD>(L, A.sd, c, rg, X, nullptr, out, out2, lp, nullptr, nullptr, A.sd.diagL[wl],
                                  A.sd.sdiagL[wl], acc);
    else
      qprod_phase<D, QM_GRAD>(L, c, X, nullptr, out, out2, lp, A.lam[wl], A.lamS[wl], acc);
  };
  // Hessian product at base point Y with copy `wl` of Q - Lambda;  acc: <X,HX>, <HX,HX>, <X,X>
  auto hess_product = [&](const double *X, const double *Y, double *out, const double *lp, int wl, double *acc) {
    if constexpr (STREAM)
      stream_qprod<D, R, QM_HESS>(L, A.sd, c, rg, X, Y, out, nullptr, lp, A.sd.diagL[wl], A.sd.sdiagL[wl], nullptr,
                                  nullptr, acc);
    else
      qprod_phase<D, QM_HESS>(L, c, X, Y, out, nullptr, lp, A.lam[wl], A.lamS[wl], acc);
  };
  const bool master = (c.b == 0 && c.tid == 0);
  const double sqrt_eps = 1.4901161193847656e-08;
  unsigned long long now = 0, t0 = 0;
  int n_state = 0, n_iter = 0;
  auto tr_state = [&](double el, double f, double g, double pg, double Dl) {
    if (master && n_state < A.trace_cap) {
      A.trace[(size_t)TR_TIME * A.trace_cap + n_state] = el;
      A.trace[(size_t)TR_F * A.trace_cap + n_state] = f;
      A.trace[(size_t)TR_G * A.trace_cap + n_state] = g;
      A.trace[(size_t)TR_PG * A.trace_cap + n_state] = pg;
      A.trace[(size_t)TR_DELTA * A.trace_cap + n_state] = Dl;
    }
    ++n_state;
  };
  auto tr_iter = [&](int inner, double hn, double hM, double rho) {
    if (master && n_iter < A.trace_cap) {
      A.trace[(size_t)TR_INNER * A.trace_cap + n_iter] = (double)inner;
      A.trace[(size_t)TR_HNORM * A.trace_cap + n_iter] = hn;
      A.trace[(size_t)TR_HM * A.trace_cap + n_iter] = hM;
      A.trace[(size_t)TR_RHO * A.trace_cap + n_iter] = rho;
    }
    ++n_iter;
  };
  // preconditioned, projected vector: Vout = proj_Y(M^-1 Rin); sums <Rin,Vout>, <Vout,Vout>
  const bool use_chain = (A.precond == CORA_B200_PRECON_REG_CHOLESKY);
  // Rin must be complete grid-wide (a barrier lies between its producer and this call)
  // HPf != nullptr (chain factor only): Rnew = Rin + alphaf HPf is formed inside the apply and takes Rin's place
  auto precond_project = [&](const double *Y, double *Rin, double *Vout, double *acc2, const double *HPf = nullptr,
                             double alphaf = 0.0, double *Rnew = nullptr) {
    const double *Zin = nullptr;
    int zsrc = A.precond == CORA_B200_PRECON_JACOBI ? 0 : 1;
    if (use_chain) {
      chain_apply_persistent<D>(A.chain, c, Rin, v[V_Z], HPf, alphaf, Rnew);
      if (HPf != nullptr) Rin = Rnew;
      Zin = v[V_Z];
      zsrc = 2;
    }
    if constexpr (STREAM) stream_update<D, R, false>(L, A.sd, c, rg, Y, nullptr, Rin, Zin, Vout, 0.0, zsrc, acc2);
    else update_reg<D, false>(L, c, Y, nullptr, Rin, Zin, Vout, 0.0, zsrc, acc2);
  };

  if (A.calibrate) {  // barrier latency calibration (profiling runs only)
    for (int i = 0; i < 16; ++i) grid_sync(c);
    if (c.tid == 0) { s_prof_ns[PH_MISC] = s_prof_ns[PH_SYNC]; s_prof_cnt[PH_MISC] = s_prof_cnt[PH_SYNC]; s_prof_ns[PH_SYNC] = 0; s_prof_cnt[PH_SYNC] = 0; }
    __syncthreads();
  }
  // ---- TNT.h:372-392: f(x), QM(x), gradient norms ----
  if (L.numChunks > 0) {
    hub_phase<D>(L, c, v[V_X], 1.0, nullptr, 0.0, lp0);
    grid_sync(c);
  }
  double fx, gnorm, pgnorm, rv_cur;
  {
    double acc[3] = {0.0, 0.0, 0.0};
    grad_product(v[V_X], v[V_GRAD], v[V_G], lp0, lcur, acc);
    grid_reduce<3>(acc, c, &t0);
    fx = 0.5 * acc[0];
    gnorm = sqrt(acc[1]);
    double a2[2] = {0.0, 0.0};
    precond_project(v[V_X], v[V_GRAD], v[V_PG], a2);
    grid_reduce<2>(a2, c, &now);
    rv_cur = a2[0];
    pgnorm = sqrt(a2[1]);
  }
  double Delta = P.Delta0;
  int status = CORA_B200_TNT_ITERATION_LIMIT;
  long long total_inner = 0;
  int iteration = 0;
  double el = 0.0;
  for (; iteration < P.max_iterations; ++iteration) {
    el = (double)(now - t0) * 1e-9;
    if (P.max_computation_time > 0 && el > P.max_computation_time) {  // TNT.h:447-452
      status = CORA_B200_TNT_ELAPSED_TIME;
      break;
    }
    tr_state(el, fx, gnorm, pgnorm, Delta);
    if (gnorm < P.gradient_tolerance) { status = CORA_B200_TNT_GRADIENT; break; }  // :474-481
    if (pgnorm < P.preconditioned_gradient_tolerance) { status = CORA_B200_TNT_PRECONDITIONED_GRADIENT; break; }

    // ---------------- STPCG (IterativeSolvers.h:207-426) ----------------
    __syncthreads();
    if (c.tid == 0) {
      cg.mode = CG_MODE_STEP; cg.it = 0; cg.max_it = P.max_TPCG_iterations; cg.exit_reason = CG_EXIT_NONE;
      cg.rv = rv_cur; cg.Delta = Delta; cg.Delta2 = Delta * Delta;
      cg.sMp = 0.0; cg.sM2 = 0.0; cg.pM2 = rv_cur; cg.sM2_next = 0.0;
      cg.alpha = cg.beta = cg.kappa = cg.sigma = 0.0; cg.hM = 0.0;
      cg.eps = 1e-8; cg.kappa_fgr = P.kappa_fgr; cg.theta = P.theta;
      const double r0 = sqrt(rv_cur);
      cg.target = r0 * fmin(P.kappa_fgr, pow(r0, P.theta));  // :278-279
      int done = 0;
      if (cg.max_it <= 0) { cg.exit_reason = CG_EXIT_MAXIT; done = 1; }
      else if (r0 <= cg.target) { cg.exit_reason = CG_EXIT_TARGET; done = 1; }
      cg.state = done;
    }
    __syncthreads();
    cg_init_flat(c, v[V_GRAD], v[V_PG], v[V_S], v[V_R], v[V_P]);
    if (L.numChunks > 0) hub_phase<D>(L, c, v[V_PG], -1.0, nullptr, 0.0, lp0);
    grid_sync(c);
    while (cg.state == 0) {
      double acc[3] = {0.0, 0.0, 0.0};
      hess_product(v[V_P], v[V_X], v[V_HP], lp0, lcur, acc);
      grid_reduce<3>(acc, c, nullptr);
      if (c.tid == 0) cg_post_hess(&cg, acc[0], acc[1], acc[2]);
      __syncthreads();
      if (cg.state != 0) break;  // p in ker(H): finished below
      if (cg.mode == CG_MODE_BOUNDARY) {  // :355-361  s += sigma p
        axpby_flat(c, 1.0, v[V_S], cg.sigma, v[V_P], v[V_S]);
        grid_sync(c);
        __syncthreads();
        if (c.tid == 0) cg.state = 1;
        __syncthreads();
        break;
      }
      double a2[2] = {0.0, 0.0};
      const double alpha = cg.alpha;
      if (use_chain) {
        // r += alpha Hp (:377) inside the factor's first phase, written to the spare vector
        precond_project(v[V_X], v[V_R], v[V_V], a2, v[V_HP], alpha, v[V_T1]);
        swp(V_R, V_T1);
      } else if constexpr (STREAM) {
        stream_update<D, R, true>(L, A.sd, c, rg, v[V_X], v[V_HP], v[V_R], nullptr, v[V_V], alpha,
                                  A.precond == CORA_B200_PRECON_JACOBI ? 0 : 1, a2);
      } else {
        update_reg<D, true>(L, c, v[V_X], v[V_HP], v[V_R], nullptr, v[V_V], alpha,
                            A.precond == CORA_B200_PRECON_JACOBI ? 0 : 1, a2);
      }
      grid_reduce<2>(a2, c, nullptr);
      if (c.tid == 0) cg_post_update(&cg, a2[0]);
      __syncthreads();
      if (cg.state != 0) {
        // s += alpha p of the last iteration (:374); s is read next by this CTA only (retraction)
        axpby_flat(c, 1.0, v[V_S], alpha, v[V_P], v[V_S]);
        break;
      }
      const double beta = cg.beta;
      if constexpr (STREAM) stream_pupdate<D, R>(L, A.sd, c, rg, alpha, beta, v[V_S], v[V_P], v[V_V], v[V_T1]);
      else cg_pupdate_flat(c, alpha, beta, v[V_S], v[V_P], v[V_V], v[V_T1]);  // s += alpha p ; p' = -v + beta p
      if (L.numChunks > 0) hub_phase<D>(L, c, v[V_P], beta, v[V_V], -1.0, lp0);
      grid_sync(c);
      swp(V_P, V_T1);
    }
    double hM = cg.hM;
    const int inner = cg.it;
    if (cg.exit_reason == CG_EXIT_KERNEL) {  // :305-338
      double a1[1] = {0.0};
      dot_flat(c, v[V_P], v[V_R], a1);
      grid_reduce<1>(a1, c, nullptr);
      double sMp = cg.sMp, sgn = 1.0;
      if (a1[0] < 0) { sgn = -1.0; sMp = -sMp; }
      const double sigma = (-sMp + sqrt(sMp * sMp + cg.pM2 * (cg.Delta2 - cg.sM2))) / cg.pM2;
      axpby_flat(c, 1.0, v[V_S], sigma * sgn, v[V_P], v[V_S]);
      grid_sync(c);
      hM = cg.Delta;
    }
    total_inner += inner;

    // ------------- proposed point, model decrease (TNT.h:503-512) -------------
    double hnorm, gh;
    {
      double a2[2] = {0.0, 0.0};
      retract_phase<D>(L, c, v[V_X], v[V_S], v[V_GRAD], v[V_XP], a2);
      grid_reduce<2>(a2, c, nullptr);
      hnorm = sqrt(a2[0]);
      gh = a2[1];
    }
    if (L.numChunks > 0) {
      hub_phase<D>(L, c, v[V_XP], 1.0, nullptr, 0.0, lp0);
      hub_phase<D>(L, c, v[V_S], 1.0, nullptr, 0.0, lp1);
      grid_sync(c);
    }
    double fxp, gnorm_p, hHh;
    {
      double a6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      grad_product(v[V_XP], v[V_GRADP], v[V_GP], lp0, lcur ^ 1, a6);
      hess_product(v[V_S], v[V_X], v[V_HP], lp1, lcur, a6 + 3);
      grid_reduce<6>(a6, c, nullptr);
      fxp = 0.5 * a6[0];
      gnorm_p = sqrt(a6[1]);
      hHh = a6[3];
    }
    double rv_prop, pgnorm_p;
    {
      double a2[2] = {0.0, 0.0};
      precond_project(v[V_XP], v[V_GRADP], v[V_T0], a2);
      grid_reduce<2>(a2, c, &now);
      rv_prop = a2[0];
      pgnorm_p = sqrt(a2[1]);
    }
    const double dm = -gh - 0.5 * hHh;
    const double df = fx - fxp;
    const double rel = df / (sqrt_eps + fabs(fx));
    const double rho = df / dm;
    const bool accepted = !isnan(rho) && rho > P.eta1;  // :532
    tr_iter(inner, hnorm, hM, rho);
    if (accepted) {
      swp(V_X, V_XP);
      fx = fxp;
      if (rel < P.relative_decrease_tolerance) {  // :561-564
        status = CORA_B200_TNT_RELATIVE_DECREASE;
        ++iteration;
        break;
      }
      if (hnorm < P.stepsize_tolerance) {  // :567-570
        status = CORA_B200_TNT_STEPSIZE;
        ++iteration;
        break;
      }
      swp(V_G, V_GP);
      swp(V_GRAD, V_GRADP);
      lcur ^= 1;
      swp(V_PG, V_T0);
      rv_cur = rv_prop;
      gnorm = gnorm_p;
      pgnorm = pgnorm_p;
    }
    if (!isnan(rho) && rho >= P.eta2) {  // :590-603
      Delta = fmax(P.alpha2 * hM, Delta);
    } else if (isnan(rho) || rho < P.eta1) {
      Delta = P.alpha1 * hM;
      if (Delta < P.Delta_tolerance) {
        status = CORA_B200_TNT_TRUST_REGION;
        ++iteration;
        break;
      }
    }
  }
  el = (double)((master ? global_timer_ns() : now) - t0) * 1e-9;
  tr_state(el, fx, gnorm, pgnorm, Delta);
  if (A.prof_all != nullptr && c.tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s_prof_ns[PH_SMID] = smid;
#pragma unroll
    for (int i = 0; i < PH_COUNT; ++i) A.prof_all[(size_t)c.b * PH_COUNT + i] = s_prof_ns[i];
  }
  if (master) {
    TntDev *o = A.out;
    o->f = fx; o->gnorm = gnorm; o->pgnorm = pgnorm; o->Delta = Delta; o->elapsed = el;
    o->status = status; o->num_outer = n_iter; o->n_state = n_state;
    o->total_inner = total_inner; o->barriers = (long long)c.nbar;
#pragma unroll
    for (int i = 0; i < V_COUNT; ++i) o->perm[i] = s_perm[i];
#pragma unroll
    for (int i = 0; i < PH_COUNT; ++i) { o->prof_ns[i] = s_prof_ns[i]; o->prof_cnt[i] = s_prof_cnt[i]; }
  }
}


// ============================================================ k_spmm_persistent ====
// `reps` data-matrix products out = Q X through the same phases (roofline leg of bench.py;
// Problem::dataMatrixProduct, src/CORA_problem.cpp:742-757).
template <int D, int R>
__global__ void __launch_bounds__(CORA_PERSIST_THREADS, CORA_PERSIST_MINB) k_spmm_persistent(const DevLayout L, const PArgs A, const double *X,
                                                                 double *out, int reps) {
  constexpr bool STREAM = R > 0;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long s_mbar[2];
  __shared__ __align__(8) unsigned long long s_rbar[STREAM ? kStreamMaxWarps * kStreamMaxStages + 1 : 1];
  __shared__ int s_meta[2][4];
  __shared__ TileMeta s_tmeta[STREAM ? 1 : kMaxTilesPerCta];
  __shared__ unsigned long long s_prof_ns[PH_COUNT], s_tph[2];
  __shared__ unsigned int s_prof_cnt[PH_COUNT];
  PCtx c;
  c.b = blockIdx.x; c.G = gridDim.x; c.tid = threadIdx.x; c.nth = blockDim.x; c.r = A.r; c.nbuf = A.nbuf;
  c.bar = A.bar; c.target = 0; c.partials = A.partials; c.parity = 0; c.nbar = 0;
  c.prof_ns = s_prof_ns; c.prof_cnt = s_prof_cnt; c.tph = s_tph;
  c.mpar0 = c.mpar1 = 0u;
  Ring rg;
  c.meta = &s_meta[0][0];
  persistent_setup<D, STREAM>(L, A, c, smem, s_tmeta, s_mbar, s_rbar, rg);
  if (c.tid < PH_COUNT) { s_prof_ns[c.tid] = 0; s_prof_cnt[c.tid] = 0; }
  if (c.tid == 0) s_tph[0] = s_tph[1] = 0;
  __syncthreads();
  double *lp0 = A.longpart;
  for (int rep = 0; rep < reps; ++rep) {
    if (L.numChunks > 0) {
      hub_phase<D>(L, c, X, 1.0, nullptr, 0.0, lp0);
      grid_sync(c);
    }
    if constexpr (STREAM)
      stream_qprod<D, R, QM_SPMM>(L, A.sd, c, rg, X, nullptr, out, nullptr, lp0, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
      qprod_phase<D, QM_SPMM>(L, c, X, nullptr, out, nullptr, lp0, nullptr, nullptr, nullptr);
    grid_sync(c);
  }
}

// Kernel entry points by (d, rank), one instantiation per translation unit (pk_*.cu); nullptr: not compiled
void *persistent_tnt_kernel(int d, int R);
void *persistent_spmm_kernel(int d, int R);

}  // namespace cora_b200
