// capi_core.cu -- lifecycle + tier-1 (operator level) entry points of the C-ABI.
#include <mutex>

#include "assemble.hpp"
#include "pose_io.hpp"
#include "pyfg.hpp"
#include "chain_chol.cuh"
#include "chain_factor_dev.cuh"
#include "ops.cuh"
#include "solver.cuh"
#include "implicit.cuh"
#include "lanczos.cuh"

namespace cora_b200 {
static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
}  // namespace cora_b200

using namespace cora_b200;

#define API_BEGIN try {
#define API_END                                                  \
  }                                                              \
  catch (const cora_b200::Error &e) {                            \
    set_last_error(e.what());                                    \
    return e.code;                                               \
  }                                                              \
  catch (const std::invalid_argument &e) {                       \
    set_last_error(e.what());                                    \
    return CORA_B200_EINVAL;                                     \
  }                                                              \
  catch (const std::exception &e) {                              \
    set_last_error(e.what());                                    \
    return CORA_B200_ERUNTIME;                                   \
  }                                                              \
  return CORA_B200_OK;

static void require(bool ok, const char *msg) {
  if (!ok) throw Error(CORA_B200_EINVAL, msg);
}

extern "C" const char *cora_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int cora_b200_version(void) { return 100; }

extern "C" int cora_b200_device_count(int *count) {
  API_BEGIN
  require(count != nullptr, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  API_END
}

namespace cora_b200 {

static void fill_dev_layout(H *h) {
  const HostLayout &L = h->HL;
  DevLayout &D = h->DL;
  D.d = L.d; D.D1 = L.D1; D.n = L.n; D.m = L.m; D.l = L.l;
  D.N = (int)L.N; D.TR = L.TR; D.TP = L.TP; D.numTiles = L.numTiles;
  D.nPoseRows = (int)L.nPoseRows; D.G = (int)L.G;
  D.maxSlots = (int)std::max<int64_t>(1, L.max_slots);
  D.numLong = (int)L.long_grp.size();
  D.tile_slots = h->d_tile_slots.p; D.tile_boff = h->d_tile_boff.p; D.tile_coff = h->d_tile_coff.p;
  D.bval = h->d_bval.p; D.bcol = h->d_bcol.p; D.sdiag = h->d_sdiag.p;
  D.grp_ptr = h->d_grp_ptr.p; D.rem_pk = h->d_rem_pk.p; D.rem_val = h->d_rem_val.p;
  D.tile_long_ptr = h->d_tile_long_ptr.p; D.long_grp = h->d_long_grp.p; D.long_ptr = h->d_long_ptr.p;
  D.long_pk = h->d_long_pk.p; D.long_val = h->d_long_val.p;
  D.dinv = h->d_dinv.p; D.int2ref = h->d_int2ref.p;
  D.numChunks = (int)L.chunk_beg.size();
  D.chunk_beg = h->d_chunk_beg.p; D.chunk_end = h->d_chunk_end.p; D.long_chunk_ptr = h->d_long_chunk_ptr.p;
  D.TRP = L.TRP; D.maxTileSpill = (int)L.max_tile_spill;
  D.tile_sp_off = h->d_tile_sp_off.p; D.tile_sp_cnt = h->d_tile_sp_cnt.p; D.sp_gptr = h->d_sp_gptr.p;
  D.sp_pk = h->d_sp_pk.p; D.sp_val = h->d_sp_val.p;
}

}  // namespace cora_b200

extern "C" int cora_b200_create(cora_b200_t **out, int device, void *stream, int d, int n_poses,
                                int n_ranges, int n_trans, const int32_t *rowptr, const int32_t *col,
                                const double *val, int64_t nnz, int preconditioner,
                                double reg_chol_max_cond) {
  H *h = nullptr;
  try {
    require(out != nullptr && rowptr != nullptr && (nnz == 0 || (col != nullptr && val != nullptr)),
            "NULL argument to cora_b200_create");
    require(preconditioner >= CORA_B200_PRECON_NONE && preconditioner <= CORA_B200_PRECON_REG_CHOLESKY,
            "unknown preconditioner");
    if (preconditioner == CORA_B200_PRECON_BLOCK_CHOLESKY)
      throw Error(CORA_B200_ENOTIMPL,
                  "Preconditioner::BlockCholesky is broken in the reference (src/CORA_problem.cpp:515-539) "
                  "and not implemented");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      throw Error(CORA_B200_ECUDA, "no CUDA device available: cora_b200 has no CPU fallback");
    }
    require(device >= 0 && device < ndev, "invalid CUDA device index");
    CUDA_CHECK(cudaSetDevice(device));
    h = new H();
    h->device = device;
    if (stream) {
      h->stream = (cudaStream_t)stream;
    } else {
      CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
      h->own_stream = true;
    }
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    int TR = 192;
    if (const char *e = getenv("CORA_B200_TILE_ROWS")) TR = atoi(e);
    if (const char *e = getenv("CORA_B200_CG_CHUNK")) h->cg_chunk = std::max(1, atoi(e));
    build_layout(h->HL, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, TR);
    HostLayout &L = h->HL;
    cudaStream_t s = h->stream;
    h->d_tile_slots.upload(L.tile_slots, s);
    { std::vector<long long> t(L.tile_boff.begin(), L.tile_boff.end()); h->d_tile_boff.upload(t, s);
      std::vector<long long> c(L.tile_coff.begin(), L.tile_coff.end()); h->d_tile_coff.upload(c, s);
      CUDA_CHECK(cudaStreamSynchronize(s)); }
    h->d_bval.upload(L.bval, s); h->d_bcol.upload(L.bcol, s); h->d_sdiag.upload(L.sdiag, s);
    h->d_grp_ptr.upload(L.grp_ptr, s); h->d_rem_pk.upload(L.rem_pk, s); h->d_rem_val.upload(L.rem_val, s);
    h->d_tile_long_ptr.upload(L.tile_long_ptr, s); h->d_long_grp.upload(L.long_grp, s);
    h->d_long_ptr.upload(L.long_ptr, s); h->d_long_pk.upload(L.long_pk, s); h->d_long_val.upload(L.long_val, s);
    h->d_int2ref.upload(L.int2ref, s);
    h->d_chunk_beg.upload(L.chunk_beg, s); h->d_chunk_end.upload(L.chunk_end, s);
    h->d_long_chunk_ptr.upload(L.long_chunk_ptr, s);
    { std::vector<long long> o(L.tile_sp_off.begin(), L.tile_sp_off.end()); h->d_tile_sp_off.upload(o, s);
      CUDA_CHECK(cudaStreamSynchronize(s)); }
    h->d_tile_sp_cnt.upload(L.tile_sp_cnt, s); h->d_sp_gptr.upload(L.sp_gptr, s);
    h->d_sp_pk.upload(L.sp_pk, s); h->d_sp_val.upload(L.sp_val, s);
    if (const char *e = getenv("CORA_B200_TNT_PATH")) h->use_persistent = std::string(e) != "launch";
    h->d_diag.upload(L.diag, s);
    { std::vector<double> inv(L.diag.size() + 2, 0.0);  // + slack: 16-byte bulk copies of odd row ranges
      for (size_t i = 0; i < L.diag.size(); ++i) inv[i] = 1.0 / L.diag[i];  // src/CORA_problem.cpp:616-618
      h->d_dinv.upload(inv, s);
      CUDA_CHECK(cudaStreamSynchronize(s)); }
    {  // strip layout for the streaming kernels (two copies of the diagonal slots: current / proposal point)
      build_stream_layout(L, h->SH);
      const StreamHost &S = h->SH;
      h->d_rec_off.upload(S.info, s);  // per strip {record offset, size, range window start, rows}: read as uint4
      h->d_rec.upload(S.rec, s);
      h->d_diagQ.upload(S.diagQ, s);
      h->d_sdiagP.upload(S.sdiagP, s);
      h->d_diagL.alloc(2 * S.diagQ.size());
      h->d_sdiagL.alloc(2 * S.sdiagP.size());
      for (int k = 0; k < 2; ++k) {
        CUDA_CHECK(cudaMemcpyAsync(h->d_diagL.p + k * S.diagQ.size(), S.diagQ.data(), S.diagQ.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(h->d_sdiagL.p + k * S.sdiagP.size(), S.sdiagP.data(), S.sdiagP.size() * sizeof(double), cudaMemcpyHostToDevice, s));
      }
      CUDA_CHECK(cudaStreamSynchronize(s));
      std::vector<unsigned char>().swap(h->SH.rec);
      std::vector<double>().swap(h->SH.diagQ);
      if (const char *e = getenv("CORA_B200_STREAM")) h->allow_stream = atoi(e) != 0;
      if (const char *e = getenv("CORA_B200_STREAM_STAGES")) h->stream_stages = std::max(2, std::min(kStreamMaxStages, atoi(e)));
      if (const char *e = getenv("CORA_B200_STREAM_SCALAR_WEIGHT")) h->stream_scalar_weight = atof(e);
      if (const char *e = getenv("CORA_B200_STREAM_INTERLEAVE")) h->stream_interleave = atoi(e) != 0;
    }
    h->d_partials.alloc((size_t)std::max(L.numTiles, h->sm_count * 8) * kNPart);
    h->d_scal.alloc(SC_COUNT);
    h->d_counter.alloc(4);
    h->d_ctrl.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(h->d_counter.p, 0, 4 * sizeof(unsigned), s));
    CUDA_CHECK(cudaMemsetAsync(h->d_scal.p, 0, SC_COUNT * sizeof(double), s));
    CUDA_CHECK(cudaMemsetAsync(h->d_ctrl.p, 0, sizeof(CgCtrl), s));
    CUDA_CHECK(cudaMallocHost((void **)&h->h_scal, SC_COUNT * sizeof(double)));
    CUDA_CHECK(cudaMallocHost((void **)&h->h_ctrl, 2 * sizeof(CgCtrl)));
    CUDA_CHECK(cudaEventCreate(&h->ev0));
    CUDA_CHECK(cudaEventCreate(&h->ev1));
    CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_chunk[0], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_chunk[1], cudaEventDisableTiming));
    CUDA_CHECK(cudaStreamSynchronize(s));
    fill_dev_layout(h);
    h->precond = preconditioner;
    h->reg_max_cond = reg_chol_max_cond > 0 ? reg_chol_max_cond : 1e6;
    // the big host copies of the values are no longer needed; the chain factorisation
    // reads them from the device
    update_preconditioner(h);
    *out = h;
  } catch (const cora_b200::Error &e) {
    set_last_error(e.what());
    if (h) cora_b200_destroy(h);
    return e.code;
  } catch (const std::invalid_argument &e) {
    set_last_error(e.what());
    if (h) cora_b200_destroy(h);
    return CORA_B200_EINVAL;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    if (h) cora_b200_destroy(h);
    return CORA_B200_ERUNTIME;
  }
  return CORA_B200_OK;
}

extern "C" int cora_b200_destroy(cora_b200_t *h) {
  if (!h) return CORA_B200_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  destroy_chain_chol(h->chol);
  destroy_chain_chol(h->chol_spare);
  destroy_chain_chol(h->ltrans);
  destroy_chain_sym(h->chain_sym);
  if (h->h_scal) cudaFreeHost(h->h_scal);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->h_tntdev) cudaFreeHost(h->h_tntdev);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 2; ++i)
    if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CORA_B200_OK;
}

extern "C" int cora_b200_size(const cora_b200_t *h, int64_t *N) {
  API_BEGIN
  require(h && N, "NULL argument");
  *N = h->HL.N;
  API_END
}

extern "C" int cora_b200_set_formulation(cora_b200_t *h, int formulation) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  CUDA_CHECK(cudaSetDevice(h->device));
  set_formulation(h, formulation);
  API_END
}

extern "C" int cora_b200_variable_rows(const cora_b200_t *h, int64_t *rows) {
  API_BEGIN
  require(h && rows, "NULL argument");
  *rows = h->io_rows();
  API_END
}

// Problem::getTranslationExplicitSolution (src/CORA_problem.cpp:1168-1197): Y ((d n + m) x r) -> [Y; t*] (N x r)
extern "C" int cora_b200_translation_explicit_solution(cora_b200_t *h, int r, const double *Y, double *Xfull) {
  API_BEGIN
  require(h && Y && Xfull, "NULL argument");
  require(h->formulation == CORA_B200_FORMULATION_IMPLICIT, "the problem is not in the implicit formulation");
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, r);
  h->resident_r = 0;
  import_matrix(h, Y, r, h->ws[V_X].p, r);
  const double *F = implicit_complete(h, h->ws[V_X].p, r, nullptr);
  export_matrix(h, F, r, Xfull, h->HL.N);
  API_END
}

extern "C" int cora_b200_set_preconditioner(cora_b200_t *h, int preconditioner, double reg_chol_max_cond) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  require(preconditioner >= CORA_B200_PRECON_NONE && preconditioner <= CORA_B200_PRECON_REG_CHOLESKY,
          "unknown preconditioner");
  if (preconditioner == CORA_B200_PRECON_BLOCK_CHOLESKY)
    throw Error(CORA_B200_ENOTIMPL, "Preconditioner::BlockCholesky not implemented");
  CUDA_CHECK(cudaSetDevice(h->device));
  h->precond = preconditioner;
  if (reg_chol_max_cond > 0) h->reg_max_cond = reg_chol_max_cond;
  update_preconditioner(h);
  API_END
}

extern "C" int cora_b200_effective_preconditioner(const cora_b200_t *h, int *preconditioner) {
  API_BEGIN
  require(h && preconditioner, "NULL argument");
  *preconditioner = h->precond;
  API_END
}

extern "C" int cora_b200_last_cert_branch(const cora_b200_t *h, int *branch) {
  API_BEGIN
  require(h && branch, "NULL argument");
  *branch = h->last_cert_branch;
  API_END
}

extern "C" int cora_b200_get_reg_lambda(const cora_b200_t *h, double *lambda) {
  API_BEGIN
  require(h && lambda, "NULL argument");
  *lambda = h->lambda_reg;
  API_END
}

extern "C" int cora_b200_set_reg_lambda(cora_b200_t *h, double lambda) {
  API_BEGIN
  require(h != nullptr, "NULL handle");
  require(lambda > 0, "lambda must be positive");
  CUDA_CHECK(cudaSetDevice(h->device));
  h->lambda_reg = lambda;
  h->lambda_user = true;
  update_preconditioner(h);
  API_END
}

// --------------------------------------------------------------------- tier 1 ---
namespace {
struct Tier1 {
  H *h;
  int r;
  Tier1(cora_b200_t *h_, int r_, bool geom = true) : h(h_), r(r_) {
    require(h != nullptr, "NULL handle");
    if (geom) check_geom_rank(r);
    CUDA_CHECK(cudaSetDevice(h->device));
    ensure_workspace(h, r);
    h->resident_r = 0;  // tier-1 calls clobber the resident iterate
  }
  double *v(int i) { return h->ws[i].p; }
  long long nE() const { return (long long)h->DL.N * r; }
};
}  // namespace

extern "C" int cora_b200_data_matrix_product(cora_b200_t *h, int r, const double *Y, double *out) {
  API_BEGIN
  require(Y && out, "NULL argument");
  Tier1 T(h, r, false);
  import_matrix(h, Y, r, T.v(V_X), r);
  launch_qprod(h, QM_SPMM, T.v(V_X), nullptr, nullptr, T.v(V_G), nullptr, r, POST_STORE, SC_TMP, nullptr);
  export_matrix(h, T.v(V_G), r, out);
  API_END
}

extern "C" int cora_b200_objective(cora_b200_t *h, int r, const double *Y, double *f) {
  API_BEGIN
  require(Y && f, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  launch_qprod(h, QM_GRAD, T.v(V_X), T.v(V_X), nullptr, T.v(V_GRAD), T.v(V_G), r, POST_STORE, SC_XG, nullptr);
  read_scal(h);
  *f = 0.5 * h->h_scal[SC_XG];
  API_END
}

extern "C" int cora_b200_egrad(cora_b200_t *h, int r, const double *Y, double *G) {
  return cora_b200_data_matrix_product(h, r, Y, G);
}

extern "C" int cora_b200_rgrad(cora_b200_t *h, int r, const double *Y, const double *G, double *out) {
  API_BEGIN
  require(Y && out, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  if (G) {
    import_matrix(h, G, r, T.v(V_G), r);
    launch_tangent(h, T.v(V_X), T.v(V_G), T.v(V_GRAD), r);
  } else {
    launch_qprod(h, QM_GRAD, T.v(V_X), T.v(V_X), nullptr, T.v(V_GRAD), T.v(V_G), r, POST_STORE, SC_XG, nullptr);
  }
  export_matrix(h, T.v(V_GRAD), r, out);
  API_END
}

extern "C" int cora_b200_hessvec(cora_b200_t *h, int r, const double *Y, const double *G,
                                 const double *Ydot, double *out) {
  API_BEGIN
  require(Y && Ydot && out, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  if (G) import_matrix(h, G, r, T.v(V_G), r);
  else launch_qprod(h, QM_SPMM, T.v(V_X), nullptr, nullptr, T.v(V_G), nullptr, r, POST_STORE, SC_TMP, nullptr);
  import_matrix(h, Ydot, r, T.v(V_P), r);
  launch_qprod(h, QM_HESS, T.v(V_P), T.v(V_X), T.v(V_G), T.v(V_HP), nullptr, r, POST_STORE, SC_HHH, nullptr);
  export_matrix(h, T.v(V_HP), r, out);
  API_END
}

extern "C" int cora_b200_tangent_proj(cora_b200_t *h, int r, const double *Y, const double *V, double *out) {
  API_BEGIN
  require(Y && V && out, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  import_matrix(h, V, r, T.v(V_T0), r);
  launch_tangent(h, T.v(V_X), T.v(V_T0), T.v(V_T1), r);
  export_matrix(h, T.v(V_T1), r, out);
  API_END
}

extern "C" int cora_b200_precondition(cora_b200_t *h, int r, const double *V, double *out) {
  API_BEGIN
  require(V && out, "NULL argument");
  Tier1 T(h, r, false);
  import_matrix(h, V, r, T.v(V_T0), r);
  apply_preconditioner(h, T.v(V_T0), T.v(V_Z), r, nullptr);
  export_matrix(h, T.v(V_Z), r, out);
  API_END
}

extern "C" int cora_b200_retract(cora_b200_t *h, int r, const double *Y, const double *V, double *out) {
  API_BEGIN
  require(Y && V && out, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  import_matrix(h, V, r, T.v(V_T0), r);
  launch_retract(h, T.v(V_X), T.v(V_T0), 1.0, nullptr, T.v(V_XP), r, -1);
  export_matrix(h, T.v(V_XP), r, out);
  API_END
}

extern "C" int cora_b200_project(cora_b200_t *h, int r, const double *A, double *out) {
  API_BEGIN
  require(A && out, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, A, r, T.v(V_X), r);
  launch_retract(h, T.v(V_X), nullptr, 0.0, nullptr, T.v(V_XP), r, -1);
  export_matrix(h, T.v(V_XP), r, out);
  API_END
}

extern "C" int cora_b200_lambda_blocks(cora_b200_t *h, int r, const double *Y, double *lam_st, double *lam_ob) {
  API_BEGIN
  require(Y && lam_st && lam_ob, "NULL argument");
  Tier1 T(h, r);
  import_matrix(h, Y, r, T.v(V_X), r);
  compute_lambda(h, T.v(V_X), r);
  const size_t ns = (size_t)h->DL.n * h->DL.d * h->DL.d;
  if (ns) CUDA_CHECK(cudaMemcpyAsync(lam_st, h->d_lam_st.p, ns * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> ob((size_t)h->DL.m);
  if (h->DL.m) CUDA_CHECK(cudaMemcpyAsync(ob.data(), h->d_lam_ob.p, (size_t)h->DL.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < h->DL.m; ++k) lam_ob[k] = ob[h->HL.range_pos[k]];  // internal range order -> reference order
  API_END
}

extern "C" int cora_b200_certificate_product(cora_b200_t *h, int r, const double *Y, int k, const double *x,
                                             double *out) {
  API_BEGIN
  require(Y && x && out && k > 0, "bad argument");
  Tier1 T(h, std::max(r, k), false);
  check_geom_rank(r);
  import_matrix(h, Y, r, T.v(V_X), r);
  compute_lambda(h, T.v(V_X), r);
  DevLayout LS = build_certificate_layout(h, 0.0);
  // the certificate matrix is the translation-explicit one in either formulation (src/CORA_problem.cpp:1055-1059)
  import_matrix(h, x, k, T.v(V_T0), k, h->HL.N);
  launch_qprod(h, QM_SPMM, T.v(V_T0), nullptr, nullptr, T.v(V_T1), nullptr, k, POST_STORE, SC_TMP, nullptr, &LS);
  export_matrix(h, T.v(V_T1), k, out, h->HL.N);
  API_END
}

namespace {
int layout_roundtrip_impl(bool strips, int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                          const int32_t *col, const double *val, int64_t nnz, int32_t *out_rowptr,
                          int32_t *out_col, double *out_val, int64_t *stats) {
  API_BEGIN
  require(rowptr && out_rowptr && stats, "NULL argument");
  HostLayout L;
  int TR = 192;
  if (const char *e = getenv("CORA_B200_TILE_ROWS")) TR = atoi(e);
  build_layout(L, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, TR);
  std::vector<int32_t> rp, ci;
  std::vector<double> v;
  if (strips) {  // through the strip layout of the streaming kernels as well
    StreamHost SH;
    build_stream_layout(L, SH);
    stream_to_csr(L, SH, rp, ci, v);
  } else {
    layout_to_csr(L, rp, ci, v);
  }
  if ((int64_t)ci.size() > nnz) throw Error(CORA_B200_ERUNTIME, "layout round trip produced more entries than the input");
  std::memcpy(out_rowptr, rp.data(), rp.size() * sizeof(int32_t));
  if (!ci.empty()) {
    std::memcpy(out_col, ci.data(), ci.size() * sizeof(int32_t));
    std::memcpy(out_val, v.data(), v.size() * sizeof(double));
  }
  stats[0] = L.numTiles; stats[1] = L.max_slots; stats[2] = L.nnz_block; stats[3] = L.nnz_rem;
  stats[4] = L.nnz_long; stats[5] = (int64_t)L.long_grp.size(); stats[6] = (int64_t)L.bval.size();
  stats[7] = (int64_t)ci.size();
  API_END
}
}  // namespace

extern "C" int cora_b200_layout_roundtrip(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                                          const int32_t *col, const double *val, int64_t nnz,
                                          int32_t *out_rowptr, int32_t *out_col, double *out_val,
                                          int64_t *stats) {
  return layout_roundtrip_impl(false, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, out_rowptr, out_col,
                               out_val, stats);
}

extern "C" int cora_b200_strip_layout_roundtrip(int d, int n_poses, int n_ranges, int n_trans,
                                                const int32_t *rowptr, const int32_t *col, const double *val,
                                                int64_t nnz, int32_t *out_rowptr, int32_t *out_col,
                                                double *out_val, int64_t *stats) {
  return layout_roundtrip_impl(true, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, out_rowptr, out_col,
                               out_val, stats);
}

// ------------------------------------------------- initialisation / export ----
extern "C" int cora_b200_odometry_initialization(int d, int n_poses, int n_landmarks, int64_t E, const int64_t *rp_i,
                                                 const int64_t *rp_j, const double *rp_t, int64_t Ep,
                                                 const int64_t *rot_i, const int64_t *rot_j, const double *rot_R,
                                                 int64_t m, const int64_t *rg_a, const int64_t *rg_b, int rank,
                                                 uint64_t seed, int reference_sign, double *X_out) {
  API_BEGIN
  require(X_out != nullptr && (E == 0 || (rp_i && rp_j && rp_t)) && (Ep == 0 || (rot_i && rot_j && rot_R)) &&
              (m == 0 || (rg_a && rg_b)), "NULL argument");
  require(n_poses >= 0 && n_landmarks >= 0, "negative size");
  odometry_initialization(d, n_poses, n_landmarks, E, rp_i, rp_j, rp_t, Ep, rot_i, rot_j, rot_R, m, rg_a, rg_b, rank,
                          seed, reference_sign, X_out);
  API_END
}

extern "C" int cora_b200_save_solution(const char *path, int format, int d, int n_poses, int n_ranges, int n_trans,
                                       const double *X, int64_t first_pose, int64_t count) {
  API_BEGIN
  require(path && X, "NULL argument");
  require(format == 0 || format == 1, "format: 0 = TUM, 1 = g2o");
  save_solution(path, format == 1, d, n_poses, n_ranges, n_trans, X, first_pose, count);
  API_END
}

// ------------------------------------------------------------------ assembly ----
namespace {
std::vector<int32_t> g_asm_rowptr, g_asm_col;
std::vector<double> g_asm_val;
std::mutex g_asm_mutex;
}  // namespace

extern "C" int cora_b200_assemble(int d, int n_poses, int n_landmarks, int64_t E, const int64_t *rp_i,
                                  const int64_t *rp_j, const double *rp_t, const double *rp_tau, int64_t Ep,
                                  const int64_t *rot_i, const int64_t *rot_j, const double *rot_R,
                                  const double *rot_kappa, int64_t m, const int64_t *rg_a, const int64_t *rg_b,
                                  const double *rg_r, const double *rg_w, int64_t *nnz, int32_t *rowptr,
                                  int32_t *col, double *val) {
  API_BEGIN
  require(nnz != nullptr, "NULL nnz");
  std::lock_guard<std::mutex> lock(g_asm_mutex);
  if (rowptr == nullptr) {  // phase 1: assemble, report nnz, keep the result for phase 2
    Measurements M{d, n_poses, n_landmarks, E, rp_i, rp_j, rp_t, rp_tau, Ep, rot_i, rot_j, rot_R, rot_kappa,
                   m, rg_a, rg_b, rg_r, rg_w};
    assemble_data_matrix(M, g_asm_rowptr, g_asm_col, g_asm_val);
    *nnz = (int64_t)g_asm_col.size();
  } else {  // phase 2: copy out
    require(*nnz == (int64_t)g_asm_col.size() && !g_asm_rowptr.empty(), "call with rowptr == NULL first");
    std::memcpy(rowptr, g_asm_rowptr.data(), g_asm_rowptr.size() * sizeof(int32_t));
    if (*nnz) {
      require(col && val, "NULL output");
      std::memcpy(col, g_asm_col.data(), g_asm_col.size() * sizeof(int32_t));
      std::memcpy(val, g_asm_val.data(), g_asm_val.size() * sizeof(double));
    }
    std::vector<int32_t>().swap(g_asm_rowptr);
    std::vector<int32_t>().swap(g_asm_col);
    std::vector<double>().swap(g_asm_val);
  }
  API_END
}

// ------------------------------------------------------------------- PyFG parser ----
struct cora_b200_pyfg {
  cora_b200::PyfgProblem P;
};

extern "C" int cora_b200_pyfg_parse(const char *path_or_text, int from_text, cora_b200_pyfg_t **out) {
  API_BEGIN
  require(path_or_text && out, "NULL argument");
  cora_b200_pyfg *g = new cora_b200_pyfg();
  try {
    if (from_text) {
      std::istringstream in{std::string(path_or_text)};
      parse_pyfg(in, g->P);
    } else {
      std::ifstream in(path_or_text);
      if (!in.good()) throw Error(CORA_B200_ERUNTIME, std::string("Could not open file ") + path_or_text);
      parse_pyfg(in, g->P);
    }
  } catch (...) {
    delete g;
    throw;
  }
  *out = g;
  API_END
}

extern "C" int cora_b200_pyfg_sizes(const cora_b200_pyfg_t *g, int *d, int *n_poses, int *n_landmarks, int64_t *E,
                                    int64_t *Ep, int64_t *m) {
  API_BEGIN
  require(g && d && n_poses && n_landmarks && E && Ep && m, "NULL argument");
  *d = g->P.d; *n_poses = (int)g->P.n(); *n_landmarks = (int)g->P.l();
  *E = (int64_t)g->P.rp_tau.size(); *Ep = (int64_t)g->P.rot_kappa.size(); *m = (int64_t)g->P.rg_w.size();
  API_END
}

extern "C" int cora_b200_pyfg_arrays(const cora_b200_pyfg_t *g, int64_t *rp_i, int64_t *rp_j, double *rp_t,
                                     double *rp_tau, int64_t *rot_i, int64_t *rot_j, double *rot_R,
                                     double *rot_kappa, int64_t *rg_a, int64_t *rg_b, double *rg_r, double *rg_w) {
  API_BEGIN
  require(g != nullptr, "NULL argument");
  const PyfgProblem &P = g->P;
  auto cp = [](auto &v, auto *dst) { if (!v.empty()) { if (!dst) throw Error(CORA_B200_EINVAL, "NULL output array"); std::copy(v.begin(), v.end(), dst); } };
  cp(P.rp_i, rp_i); cp(P.rp_j, rp_j); cp(P.rp_t, rp_t); cp(P.rp_tau, rp_tau);
  cp(P.rot_i, rot_i); cp(P.rot_j, rot_j); cp(P.rot_R, rot_R); cp(P.rot_kappa, rot_kappa);
  cp(P.rg_a, rg_a); cp(P.rg_b, rg_b); cp(P.rg_r, rg_r); cp(P.rg_w, rg_w);
  API_END
}

extern "C" int cora_b200_pyfg_free(cora_b200_pyfg_t *g) {
  delete g;
  return CORA_B200_OK;
}

// Test hook (CPU only): the chain factorisation + solve executed on the host through the SAME
// per-chunk functions the device kernels call, so the CPU suite can pin the algorithm against the
// oracle's sparse LU without a GPU.  V, out: N x r column-major in the reference row order.
extern "C" int cora_b200_debug_chain_host(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                                          const int32_t *col, const double *val, int64_t nnz, double shift,
                                          int pin_last, int r, const double *V, double *out, int *pos_def) {
  API_BEGIN
  require(rowptr && pos_def, "NULL argument");
  HostLayout L;
  build_layout(L, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, 192);
  ChainFactorHost F;
  const bool solve = (V != nullptr && out != nullptr && r > 0);
  if (L.D1 == 3) chain_factor_host<3>(F, L, L.bval.data(), L.sdiag.data(), shift, pin_last != 0, solve);
  else chain_factor_host<4>(F, L, L.bval.data(), L.sdiag.data(), shift, pin_last != 0, solve);
  *pos_def = F.pos_def ? 1 : 0;
  if (solve && F.pos_def) {
    const size_t N = (size_t)L.N;
    std::vector<double> Vi(N * r), Zi(N * r);
    for (size_t i = 0; i < N; ++i)
      for (int c = 0; c < r; ++c) Vi[i * r + c] = V[(size_t)c * N + L.int2ref[i]];
    if (L.D1 == 3) chain_apply_host<3>(F, L, Vi.data(), Zi.data(), r);
    else chain_apply_host<4>(F, L, Vi.data(), Zi.data(), r);
    for (size_t i = 0; i < N; ++i)
      for (int c = 0; c < r; ++c) out[(size_t)c * N + L.int2ref[i]] = Zi[i * r + c];
  }
  API_END
}

// Test hook (CPU only): structure of the pose-system factorisation the handle would build for this matrix.
// stats[0] = 1 if the pose graph is a chain (chain_chol.cuh levels), 0 if general (gen_chol.hpp);
// general: [1] pose couplings, [2] blocks of L below the diagonal, [3] elimination-tree height, [4] clusters,
// [5] cluster levels (grid barriers of one forward or backward sweep), [6] largest column, [7] poses,
// [8] blocks of the per-cluster inverses, [9] most row blocks read by one cluster, [10] longest row, [11] rows > 64.
extern "C" int cora_b200_debug_factor_stats(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                                            const int32_t *col, const double *val, int64_t nnz, int64_t *stats) {
  API_BEGIN
  require(rowptr && stats, "NULL argument");
  HostLayout L;
  build_layout(L, d, n_poses, n_ranges, n_trans, rowptr, col, val, nnz, 192);
  for (int i = 0; i < 12; ++i) stats[i] = 0;
  stats[7] = L.n;
  ChainSym S;
  try {
    if (L.D1 == 3) chain_symbolic_build<3>(S, L, false); else chain_symbolic_build<4>(S, L, false);
    stats[0] = 1;
  } catch (const Error &e) {
    if (e.code != CORA_B200_ENOTIMPL) throw;
    ChainSym G;
    if (L.D1 == 3) chain_symbolic_build<3>(G, L, true); else chain_symbolic_build<4>(G, L, true);
    stats[1] = (int64_t)G.e_i.size();
    stats[2] = G.gs.nnzL();
    stats[3] = G.gs.etree_height;
    stats[4] = (int64_t)G.gs.cl_ptr.size() - 1;
    stats[5] = G.gs.levels();
    int64_t mx = 0;
    for (int v = 0; v < G.gs.n; ++v) mx = std::max<int64_t>(mx, G.gs.colptr[v + 1] - G.gs.colptr[v]);
    stats[6] = mx;
    stats[8] = G.gs.linv_blocks;  // blocks of the per-cluster inverses
    for (size_t c = 0; c + 1 < G.gs.cl_ptr.size(); ++c)  // [9] most off-diagonal blocks one cluster reads in the forward sweep
      stats[9] = std::max<int64_t>(stats[9], G.gs.rowptr[G.gs.cl_ptr[c + 1]] - G.gs.rowptr[G.gs.cl_ptr[c]]);
    for (int v = 0; v < G.gs.n; ++v) {  // [10] longest row, [11] rows with more than 64 blocks
      const int64_t len = G.gs.rowptr[v + 1] - G.gs.rowptr[v];
      stats[10] = std::max(stats[10], len);
      stats[11] += len > 64 ? 1 : 0;
    }
  }
  API_END
}
