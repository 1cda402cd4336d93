// capi_solver.cu -- tier-2 (solver level) entry points of the C-ABI.
#include "certify.cuh"
#include "solver.cuh"

using namespace cora_b200;

#ifndef API_BEGIN
#error "capi_solver.cu is compiled through unity.cu (after capi_core.cu)"
#endif

extern "C" int cora_b200_tnt_default_params(cora_b200_tnt_params *p) {
  API_BEGIN
  require(p != nullptr, "NULL params");
  // src/CORA.cpp:95-109 over the library defaults of TNT.h:76-130
  p->Delta0 = 5;
  p->eta1 = 0.05;
  p->eta2 = 0.9;
  p->alpha1 = 0.25;
  p->alpha2 = 3.0;
  p->max_TPCG_iterations = 80;
  p->max_iterations = 250;
  p->kappa_fgr = 0.1;
  p->theta = 0.8;
  p->preconditioned_gradient_tolerance = 1e-6;
  p->gradient_tolerance = 1e-6;
  p->relative_decrease_tolerance = 1e-6;
  p->stepsize_tolerance = 1e-6;
  p->Delta_tolerance = 1e-5;
  p->max_computation_time = 20;
  p->verbose = 0;
  p->reserved = 0;
  API_END
}

static void check_params(const cora_b200_tnt_params *p) {
  require(p != nullptr, "NULL params");
  require(p->Delta0 > 0, "Delta0 must be positive");
  require(p->max_TPCG_iterations >= 0 && p->max_iterations >= 0, "iteration limits must be nonnegative");
  require(p->kappa_fgr >= 0 && p->kappa_fgr < 1,
          "Target fractional reduction of the gradient norm (kappa_fgr) must be a real value in the range [0,1)");
  require(p->theta >= 0 && p->theta <= 1,
          "Target superlinear convergence rate (theta) must be a real value in the range [0,1]");
}

extern "C" int cora_b200_set_iterate(cora_b200_t *h, int r, const double *X) {
  API_BEGIN
  require(h && X, "NULL argument");
  check_geom_rank(r);
  CUDA_CHECK(cudaSetDevice(h->device));
  h->resident_r = 0;
  ensure_workspace(h, r);
  import_matrix(h, X, r, h->ws[V_X].p, r);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->resident_r = r;
  API_END
}

extern "C" int cora_b200_get_iterate(cora_b200_t *h, int r, double *X) {
  API_BEGIN
  require(h && X, "NULL argument");
  require(h->resident_r == r && r > 0, "no resident iterate of this rank");
  CUDA_CHECK(cudaSetDevice(h->device));
  export_matrix(h, h->ws[V_X].p, r, X);
  API_END
}

extern "C" int cora_b200_tnt_resident(cora_b200_t *h, const cora_b200_tnt_params *p, cora_b200_tnt_result *res) {
  API_BEGIN
  require(h && res, "NULL argument");
  check_params(p);
  require(h->resident_r > 0, "no resident iterate: call cora_b200_set_iterate first");
  CUDA_CHECK(cudaSetDevice(h->device));
  tnt_resident(h, h->resident_r, *p, res);
  API_END
}

extern "C" int cora_b200_tnt(cora_b200_t *h, int r, const double *X0, const cora_b200_tnt_params *p,
                             double *X_out, cora_b200_tnt_result *res) {
  API_BEGIN
  require(h && X0 && X_out && res, "NULL argument");
  check_params(p);
  check_geom_rank(r);
  CUDA_CHECK(cudaSetDevice(h->device));
  h->resident_r = 0;
  ensure_workspace(h, r);
  import_matrix(h, X0, r, h->ws[V_X].p, r);
  h->resident_r = r;
  tnt_resident(h, r, *p, res);
  export_matrix(h, h->ws[V_X].p, r, X_out);
  API_END
}

extern "C" int cora_b200_spmm_resident(cora_b200_t *h, int reps, float *ms_total) {
  API_BEGIN
  require(h && ms_total && reps > 0, "bad argument");
  require(h->resident_r > 0, "no resident iterate: call cora_b200_set_iterate first");
  CUDA_CHECK(cudaSetDevice(h->device));
  const int r = h->resident_r;
  if (h->use_persistent) {
    *ms_total = spmm_persistent(h, r, h->ws[V_X].p, h->ws[V_G].p, reps);
  } else {
    CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < reps; ++i)
      launch_qprod(h, QM_SPMM, h->ws[V_X].p, nullptr, nullptr, h->ws[V_G].p, nullptr, r, POST_STORE, SC_TMP, nullptr);
    CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
    CUDA_CHECK(cudaEventSynchronize(h->ev1));
    CUDA_CHECK(cudaEventElapsedTime(ms_total, h->ev0, h->ev1));
  }
  API_END
}

extern "C" int cora_b200_snapshot_iterate(cora_b200_t *h) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  require(h->resident_r > 0, "no resident iterate");
  CUDA_CHECK(cudaSetDevice(h->device));
  const size_t n = (size_t)h->DL.N * h->resident_r;
  if (h->d_snap.n < n) h->d_snap.alloc(n);
  CUDA_CHECK(cudaMemcpyAsync(h->d_snap.p, h->ws[V_X].p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->snap_r = h->resident_r;
  API_END
}

extern "C" int cora_b200_restore_iterate(cora_b200_t *h) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  require(h->snap_r > 0, "no snapshot");
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, h->snap_r);
  const size_t n = (size_t)h->DL.N * h->snap_r;
  CUDA_CHECK(cudaMemcpyAsync(h->ws[V_X].p, h->d_snap.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->resident_r = h->snap_r;
  API_END
}

extern "C" int cora_b200_profile_hessvec(cora_b200_t *h, int max_samples) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->prof_n = 0;
  h->prof_on = max_samples > 0;
  while ((int)h->prof_ev.size() < 2 * max_samples) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    h->prof_ev.push_back(e);
  }
  API_END
}

extern "C" int cora_b200_profile_read(cora_b200_t *h, int capacity, float *ms, int *count) {
  API_BEGIN
  require(h && ms && count, "NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  int n = 0;
  for (size_t i = 0; i + 1 < h->prof_n && n < capacity; i += 2, ++n)
    CUDA_CHECK(cudaEventElapsedTime(&ms[n], h->prof_ev[i], h->prof_ev[i + 1]));
  *count = n;
  h->prof_n = 0;
  API_END
}

extern "C" int cora_b200_get_work_vector(cora_b200_t *h, int which, int r, double *out) {
  API_BEGIN
  require(h && out, "NULL argument");
  require(which == 0 || which == 1 || (which >= 100 && which < 100 + V_COUNT),
          "which must be 0 (X), 1 (Q*X) or 100 + work vector index (test hook)");
  require(r > 0 && r <= h->ws_r, "no workspace of this rank");
  CUDA_CHECK(cudaSetDevice(h->device));
  export_matrix(h, h->ws[which >= 100 ? which - 100 : (which == 0 ? V_X : V_G)].p, r, out);
  API_END
}

// Device pointers of the resident iterate X and of Q X in the INTERNAL layout (N x r row-major, rows in the order
// cora_b200_row_order reports): lets a multi-GPU caller exchange rows GPU to GPU (row-partitioned product,
// cora_b200/rowpart.py) without a host round trip.  Marks rank r as resident.
extern "C" int cora_b200_device_vectors(cora_b200_t *h, int r, double **x, double **qx) {
  API_BEGIN
  require(h && x && qx, "NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, r);
  h->resident_r = r;
  *x = h->ws[V_X].p;
  *qx = h->ws[V_G].p;
  API_END
}

extern "C" int cora_b200_row_order(const cora_b200_t *h, int32_t *internal_to_reference) {
  API_BEGIN
  require(h && internal_to_reference, "NULL argument");
  std::copy(h->HL.int2ref.begin(), h->HL.int2ref.end(), internal_to_reference);
  API_END
}

extern "C" int cora_b200_phase_profile_ctas(cora_b200_t *h, int capacity, double *max_us, double *median_us) {
  API_BEGIN
  require(h && max_us && median_us, "NULL argument");
  require(h->h_tntdev != nullptr && h->prof_all_grid > 0, "no persistent TNT call has run on this handle");
  const TntDev &o = *(const TntDev *)h->h_tntdev;
  const int n = std::min(capacity, (int)PH_COUNT), G = h->prof_all_grid;
  std::vector<double> col((size_t)G);
  for (int i = 0; i < n; ++i) {
    max_us[i] = median_us[i] = 0.0;
    if (!o.prof_cnt[i]) continue;
    for (int b = 0; b < G; ++b) col[b] = h->h_prof_all[(size_t)b * PH_COUNT + i] * 1e-3 / o.prof_cnt[i];
    std::sort(col.begin(), col.end());
    max_us[i] = col[G - 1];
    median_us[i] = col[G / 2];
  }
  API_END
}

extern "C" int cora_b200_phase_profile(cora_b200_t *h, int capacity, double *total_us, int64_t *count,
                                       int *n_kinds, int *grid, int64_t *barriers) {
  API_BEGIN
  require(h && total_us && count && n_kinds, "NULL argument");
  require(h->h_tntdev != nullptr, "no persistent TNT call has run on this handle");
  const TntDev &o = *(const TntDev *)h->h_tntdev;
  const int n = std::min(capacity, (int)PH_COUNT);
  for (int i = 0; i < n; ++i) {
    total_us[i] = o.prof_ns[i] * 1e-3;
    count[i] = o.prof_cnt[i];
  }
  *n_kinds = n;
  if (grid) *grid = h->persistent_grid;
  if (barriers) *barriers = o.barriers;
  API_END
}

extern "C" int cora_b200_certify(cora_b200_t *h, int r, const double *Y, double eta, int nx,
                                 const double *bootstrap, int bootstrap_cols, int max_iters, int *is_certified,
                                 double *theta, double *x, double *all_eigvecs, int all_eigvecs_cols_capacity,
                                 int *all_eigvecs_cols, int64_t *num_iters) {
  API_BEGIN
  require(h && Y && is_certified && theta && x, "NULL argument");
  check_geom_rank(r);
  CUDA_CHECK(cudaSetDevice(h->device));
  certify_host(h, r, Y, eta, nx, bootstrap, bootstrap_cols, max_iters, is_certified, theta, x, all_eigvecs,
               all_eigvecs_cols_capacity, all_eigvecs_cols, num_iters);
  API_END
}

extern "C" int cora_b200_psd_test(cora_b200_t *h, int r, const double *Y, double eta, int *is_psd) {
  API_BEGIN
  require(h && is_psd, "NULL argument");
  check_geom_rank(r);
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, r);
  if (Y != nullptr) {  // NULL: test the resident iterate
    h->resident_r = 0;
    import_matrix(h, Y, r, h->ws[V_X].p, r);
    h->resident_r = r;
  }
  require(h->resident_r == r, "no resident iterate of this rank");
  *is_psd = psd_test_resident(h, r, eta) ? 1 : 0;
  API_END
}

extern "C" int cora_b200_debug_min_eigenpair(cora_b200_t *h, int max_iters, double *theta, double *x, int *steps) {
  API_BEGIN
  require(h && theta && x, "NULL argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  debug_min_eigenpair(h, max_iters, theta, x, steps);
  API_END
}

extern "C" int cora_b200_saddle_escape(cora_b200_t *h, int r_new, const double *Y, double theta, const double *v,
                                       double gradient_tolerance, double preconditioned_gradient_tolerance,
                                       double *Y_out) {
  API_BEGIN
  require(h && Y && v && Y_out, "NULL argument");
  check_geom_rank(r_new);
  CUDA_CHECK(cudaSetDevice(h->device));
  saddle_escape_host(h, r_new, Y, theta, v, gradient_tolerance, preconditioned_gradient_tolerance, Y_out);
  API_END
}

extern "C" int cora_b200_project_solution(cora_b200_t *h, int r, const double *Y, double *Y_out) {
  API_BEGIN
  require(h && Y && Y_out, "NULL argument");
  check_geom_rank(r);
  CUDA_CHECK(cudaSetDevice(h->device));
  project_solution_host(h, r, Y, Y_out);
  API_END
}

extern "C" int cora_b200_solve(cora_b200_t *h, int r0, const double *X0, int max_rank,
                               const cora_b200_tnt_params *p, int verbose, double *X_out,
                               cora_b200_solve_result *res) {
  API_BEGIN
  require(h && X0 && X_out && res, "NULL argument");
  check_params(p);
  check_geom_rank(r0);
  CUDA_CHECK(cudaSetDevice(h->device));
  solve_staircase(h, r0, X0, max_rank, *p, verbose, X_out, res);
  API_END
}
