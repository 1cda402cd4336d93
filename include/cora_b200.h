/* cora_b200.h -- C-ABI of the B200-native CORA staircase inner loop.
 *
 * The reference (MarineRoboticsGroup/cora @ 015dc43) has no FFI seam: its solver
 * reaches the hot path only through the Riemannian methods of CORA::Problem and
 * through solveCORA().  This header is the seam a maintainer would bind instead;
 * every entry point names the reference interface it replaces (paths relative to
 * the reference root).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions (all pointers are HOST memory unless the name ends in _dev):
 *   dense matrices  column-major double, leading dimension = rows -- exactly the
 *                   buffer of an Eigen::MatrixXd (include/CORA/CORA_types.h:47-48)
 *   data matrix     CSR int32 rowptr[N+1], int32 col[nnz], double val[nnz] -- the
 *                   compressed Eigen::SparseMatrix<double,RowMajor>
 *                   (include/CORA/CORA_types.h:70); both triangles stored
 *   row order       [d rows per pose | one row per range factor | n pose
 *                   translations | l landmark translations]
 *                   (src/CORA_problem.cpp:964-1021)
 *   status          every function returns 0 on success, a CORA_B200_E* code
 *                   otherwise; cora_b200_last_error() returns the message the C++
 *                   shim rethrows as the matching reference exception.
 *   threading       one handle <-> one CUDA device + one stream; handles are
 *                   independent.  A handle is not re-entrant (the reference
 *                   Problem is not either: include/CORA/CORA_problem.h:270-280).
 */
#ifndef CORA_B200_H_
#define CORA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORA_B200_OK 0
#define CORA_B200_EINVAL 1   /* std::invalid_argument / MatrixShapeException   */
#define CORA_B200_ERUNTIME 2 /* std::runtime_error                              */
#define CORA_B200_ECUDA 3    /* CUDA failure or no CUDA device: never a CPU fallback */
#define CORA_B200_ENOTIMPL 4 /* NotImplementedException (CORA_types.h:15-19)    */

/* enum class Preconditioner (include/CORA/CORA_types.h:76-77) */
#define CORA_B200_PRECON_NONE 0
#define CORA_B200_PRECON_JACOBI 1
#define CORA_B200_PRECON_BLOCK_CHOLESKY 2 /* broken in the reference; ENOTIMPL */
#define CORA_B200_PRECON_REG_CHOLESKY 3

/* enum class TNTStatus (libs/Optimization/.../Riemannian/TNT.h:134-164) */
#define CORA_B200_TNT_GRADIENT 0
#define CORA_B200_TNT_PRECONDITIONED_GRADIENT 1
#define CORA_B200_TNT_RELATIVE_DECREASE 2
#define CORA_B200_TNT_STEPSIZE 3
#define CORA_B200_TNT_TRUST_REGION 4
#define CORA_B200_TNT_ITERATION_LIMIT 5
#define CORA_B200_TNT_ELAPSED_TIME 6
#define CORA_B200_TNT_USER_FUNCTION 7

typedef struct cora_b200_handle cora_b200_t;

const char *cora_b200_last_error(void);
int cora_b200_version(void);
/* number of visible CUDA devices (0 on a CPU-only host; not an error) */
int cora_b200_device_count(int *count);

/* ---- lifecycle: what Problem::updateProblemData() does once per problem -------
 * (src/CORA_problem.cpp:500-510: fillDataMatrix :625-712 has produced the CSR;
 * updatePreconditioner :512-623 is done inside create).  The handle copies Q, so
 * the caller may free its buffers.  `stream` is a cudaStream_t (NULL: the handle
 * creates its own non-blocking stream).  reg_chol_max_cond: the reference's
 * $CORA_REG_CHOLESKY_MAX_COND (:582), <= 0 selects its default 1e6. */
int cora_b200_create(cora_b200_t **h, int device, void *stream, int d, int n_poses,
                     int n_ranges, int n_trans, const int32_t *rowptr,
                     const int32_t *col, const double *val, int64_t nnz,
                     int preconditioner, double reg_chol_max_cond);
int cora_b200_destroy(cora_b200_t *h);
int cora_b200_size(const cora_b200_t *h, int64_t *N);
/* Formulation (include/CORA/CORA_types.h:52-56, Problem::setFormulation CORA_problem.h:338).  Implicit = the
 * translation-marginalised problem of src/CORA_problem.cpp:714-757: every dense matrix at this boundary then has
 * getExpectedVariableSize() = d n + m rows (rotations, then ranges; src/CORA_problem.cpp:944-954) instead of N,
 * the data-matrix product is Qmain Y - T_red chol(L_red)^-1 T_red^T Y, the preconditioner lifts with zero
 * translations (:878-885) and a failed certificate returns the rotation/range part of the direction with its
 * Rayleigh quotient in the implicit operator (:1085-1100).  The factor of L_red is built here (chain or general
 * pose-graph Cholesky).  Implicit solves run on the multi-launch TNT path. */
#define CORA_B200_FORMULATION_EXPLICIT 0
#define CORA_B200_FORMULATION_IMPLICIT 1
int cora_b200_set_formulation(cora_b200_t *h, int formulation);
/* Problem::getExpectedVariableSize (src/CORA_problem.cpp:944-954) */
int cora_b200_variable_rows(const cora_b200_t *h, int64_t *rows);
/* Problem::getTranslationExplicitSolution (src/CORA_problem.cpp:1168-1197): Y is (d n + m) x r, Xfull is N x r
 * = [Y; -chol(L_red)^-1 T_red^T Y; 0] (last translation pinned to zero). */
int cora_b200_translation_explicit_solution(cora_b200_t *h, int r, const double *Y, double *Xfull);

/* Problem::setPreconditioner (include/CORA/CORA_problem.h:335-337) + updatePreconditioner */
int cora_b200_set_preconditioner(cora_b200_t *h, int preconditioner, double reg_chol_max_cond);
/* The preconditioner actually applied.  Preconditioner::RegularizedCholesky (the reference's default,
 * src/pyfg_text_parser.cpp:116-120) is factored on the device for graphs made of one odometry chain plus
 * landmark / range factors (all BASELINE configurations); for any other graph (loop closures, several
 * robots) the handle is still created and applies Jacobi instead -- this call reports it. */
int cora_b200_effective_preconditioner(const cora_b200_t *h, int *preconditioner);

/* lambda of RegularizedCholesky actually used (src/CORA_problem.cpp:591) */
int cora_b200_get_reg_lambda(const cora_b200_t *h, double *lambda);
/* override lambda (parity runs pass the oracle's converged ||Q||_2/(c-1)) */
int cora_b200_set_reg_lambda(cora_b200_t *h, double lambda);

/* ---- tier 1: operator level (include/CORA/CORA_problem.h:340-350) ------------
 * Y, G, Ydot, V, A, out are N x r.  One upload + kernels + one download each; the
 * performance path is tier 2. */
/* Problem::dataMatrixProduct, src/CORA_problem.cpp:742-757 (Explicit) */
int cora_b200_data_matrix_product(cora_b200_t *h, int r, const double *Y, double *out);
/* Problem::evaluateObjective, :759-762 */
int cora_b200_objective(cora_b200_t *h, int r, const double *Y, double *f);
/* Problem::Euclidean_gradient, :764-770 */
int cora_b200_egrad(cora_b200_t *h, int r, const double *Y, double *G);
/* Problem::Riemannian_gradient(Y, NablaF_Y), :772-780; G may be NULL (recomputed) */
int cora_b200_rgrad(cora_b200_t *h, int r, const double *Y, const double *G, double *out);
/* Problem::Riemannian_Hessian_vector_product, :822-867 */
int cora_b200_hessvec(cora_b200_t *h, int r, const double *Y, const double *G,
                      const double *Ydot, double *out);
/* Problem::tangent_space_projection, :782-820 */
int cora_b200_tangent_proj(cora_b200_t *h, int r, const double *Y, const double *V,
                           double *out);
/* Problem::precondition, :869-903 (no tangent projection; src/CORA.cpp:89-92 adds it) */
int cora_b200_precondition(cora_b200_t *h, int r, const double *V, double *out);
/* Problem::retract, :936-938 */
int cora_b200_retract(cora_b200_t *h, int r, const double *Y, const double *V, double *out);
/* Problem::projectToManifold, :905-934 */
int cora_b200_project(cora_b200_t *h, int r, const double *A, double *out);
/* Problem::compute_Lambda_blocks, :1105-1131: lam_st is d x (d*n) column-major (the
 * reference's d x dn block row), lam_ob has n_ranges entries */
int cora_b200_lambda_blocks(cora_b200_t *h, int r, const double *Y, double *lam_st,
                            double *lam_ob);
/* S*x for S = get_certificate_matrix(Y) (:1162-1166) without forming S: x, out N x k */
int cora_b200_certificate_product(cora_b200_t *h, int r, const double *Y, int k,
                                  const double *x, double *out);

/* ---- tier 2: solver level ----------------------------------------------------- */

/* Optimization::Riemannian::TNTParams (TNT.h:76-130) with solveCORA's values as
 * the defaults cora_b200_tnt_default_params() fills (src/CORA.cpp:95-109). */
typedef struct cora_b200_tnt_params {
  double Delta0;                            /* 5      */
  double eta1;                              /* 0.05   */
  double eta2;                              /* 0.9    */
  double alpha1;                            /* 0.25   */
  double alpha2;                            /* 3.0    */
  int32_t max_TPCG_iterations;              /* 80     */
  int32_t max_iterations;                   /* 250    */
  double kappa_fgr;                         /* 0.1    */
  double theta;                             /* 0.8    */
  double preconditioned_gradient_tolerance; /* 1e-6   */
  double gradient_tolerance;                /* 1e-6   */
  double relative_decrease_tolerance;       /* 1e-6   */
  double stepsize_tolerance;                /* 1e-6   */
  double Delta_tolerance;                   /* 1e-5   */
  double max_computation_time;              /* 20 s (src/CORA.cpp:106); <=0: no cap */
  int32_t verbose;                          /* show_iterates */
  int32_t reserved;
} cora_b200_tnt_params;

/* TNTResult (TNT.h:168-194 + Riemannian/Concepts.h:136-148 + Base/Concepts.h:64-88).
 * Trace arrays are caller-allocated with `trace_capacity` entries each (NULL: not
 * recorded); per-iteration traces receive num_outer entries, the "state" traces
 * (objective_values, gradient_norms, preconditioned_gradient_norms,
 * trust_region_radius, time) num_outer + 1 as in the reference. */
typedef struct cora_b200_tnt_result {
  double f;
  double gradfx_norm;
  double preconditioned_gradfx_norm;
  double elapsed_time;   /* host wall-clock, seconds                         */
  double device_time;    /* CUDA-event time of the whole call on the stream  */
  int32_t status;        /* CORA_B200_TNT_*                                  */
  int32_t num_outer;     /* trust-region iterations performed                */
  int64_t total_inner;   /* sum of inner_iterations (CG iterations)          */
  int64_t kernel_launches;
  int32_t trace_capacity;
  int32_t reserved;
  double *objective_values;
  double *gradient_norms;
  double *preconditioned_gradient_norms;
  double *trust_region_radius;
  double *time;
  double *update_step_norms;
  double *update_step_M_norms;
  double *gain_ratios;
  int32_t *inner_iterations;
} cora_b200_tnt_result;

int cora_b200_tnt_default_params(cora_b200_tnt_params *p);

/* Optimization::Riemannian::TNT (TNT.h:242-689) with the closures of
 * src/CORA.cpp:52-122 (f, QM, metric, retract, precon), the STPCG loop
 * (IterativeSolvers.h:166-426) device resident.  X0 must be on the manifold (the
 * caller has applied projectToManifold, src/CORA.cpp:128). */
int cora_b200_tnt(cora_b200_t *h, int r, const double *X0, const cora_b200_tnt_params *p,
                  double *X_out, cora_b200_tnt_result *res);

/* Device-resident variants used by bench.py's `value` leg and by the staircase:
 * the iterate stays in HBM between calls. */
int cora_b200_set_iterate(cora_b200_t *h, int r, const double *X);   /* H2D + layout */
int cora_b200_get_iterate(cora_b200_t *h, int r, double *X);         /* D2H + layout */
int cora_b200_tnt_resident(cora_b200_t *h, const cora_b200_tnt_params *p,
                           cora_b200_tnt_result *res);
/* keep / restore a device copy of the resident iterate (bench.py restarts a solve without
 * touching the host) */
int cora_b200_snapshot_iterate(cora_b200_t *h);
int cora_b200_restore_iterate(cora_b200_t *h);
/* per-launch CUDA-event timing of the dominant kernel (the fused Hessian-vector product
 * inside STPCG): enable with max_samples > 0, read back the milliseconds of each launch */
int cora_b200_profile_hessvec(cora_b200_t *h, int max_samples);
int cora_b200_profile_read(cora_b200_t *h, int capacity, float *ms, int *count);
/* test hook: copy one of the handle's N x r work vectors out (reference layout).  which: 0 = resident
 * iterate X, 1 = Q*X as left by cora_b200_spmm_resident / the last gradient evaluation. */
int cora_b200_get_work_vector(cora_b200_t *h, int which, int r, double *out);
/* in-kernel phase profile of the last persistent TNT call (CTA 0's %globaltimer): for each phase
 * kind k < *n_kinds, total_us[k] and count[k].  Kinds, in order: hub, grad, hess, update, pupdate,
 * retract, precond, cginit, sync, misc, q.wait, q.qx, q.epi, q.store, ch.pre, ch.fwd, ch.bwd,
 * ch.border, ch.post (persistent.cuh PhaseId).
 * *grid / *barriers: CTAs of the cooperative launch and grid barriers executed. */
int cora_b200_phase_profile(cora_b200_t *h, int capacity, double *total_us, int64_t *count,
                            int *n_kinds, int *grid, int64_t *barriers);
/* Per phase kind: the average time of one execution of the phase on the slowest CTA and on the median CTA
 * (every CTA's clock; the persistent kernel runs as fast as its slowest CTA allows). */
int cora_b200_phase_profile_ctas(cora_b200_t *h, int capacity, double *max_us, double *median_us);

/* Device pointers of the resident iterate X and of Q*X (as left by cora_b200_spmm_resident) in the library's
 * internal layout: N x r row-major, row i of the buffers = reference row internal_to_reference[i]
 * (cora_b200_row_order).  For GPU-to-GPU row exchange in the row-partitioned product (SURVEY 8f-4,
 * cora_b200/rowpart.py); the pointers stay valid until the workspace grows (a larger rank) or the handle dies. */
int cora_b200_device_vectors(cora_b200_t *h, int r, double **x, double **qx);
int cora_b200_row_order(const cora_b200_t *h, int32_t *internal_to_reference /* N */);

/* Row-partitioned product of ONE problem across the GPUs of a node with the exchanges done by the library's own
 * kernels over peer-mapped memory (cudaIpc*, NVLink loads) instead of collectives -- cora_b200/csrc/peer_product.cuh.
 * Each rank: cora_b200_peer_create on its LOCAL handle (exports 3 x 64-byte IPC handles), ship the handles of all
 * ranks to everybody over any channel, cora_b200_peer_connect with the plan (which internal row of which peer's
 * operand each ghost row is; the rank's landmark rows), then cora_b200_peer_product: per product a cross-GPU flag
 * barrier + pull of the ghost rows, the persistent SpMM kernel on the rank's rows, a second barrier + the sum of all
 * ranks' partial landmark rows in rank order.  Replaces the one-core `data_matrix_ * Y` of
 * src/CORA_problem.cpp:742-757 for a problem sharded by rows. */
typedef struct cora_b200_peer cora_b200_peer_t;
int cora_b200_peer_create(cora_b200_t *h, int r, int n_landmark_rows, cora_b200_peer_t **out, void *handles /* 192 B */);
int cora_b200_peer_connect(cora_b200_peer_t *p, int world, int rank, const void *all_handles /* world x 192 B */,
                           int n_ghost_rows, const int32_t *ghost_peer, const int32_t *ghost_src_row,
                           const int32_t *ghost_dst_row, const int32_t *landmark_rows);
int cora_b200_peer_product(cora_b200_peer_t *p, int reps, float *ms_total);
int cora_b200_peer_destroy(cora_b200_peer_t *p);

/* timed data-matrix products on the resident iterate: reps launches of Q*X, returns
 * the CUDA-event milliseconds for all of them (roofline leg of bench.py) */
int cora_b200_spmm_resident(cora_b200_t *h, int reps, float *ms_total);

/* Problem::certify_solution (src/CORA_problem.cpp:1030-1103) + fast_verification
 * (src/CORA_utils.cpp:17-186).  x has N entries; all_eigvecs (may be NULL) is
 * N x all_eigvecs_cols_capacity, *all_eigvecs_cols receives the columns written. */
int cora_b200_certify(cora_b200_t *h, int r, const double *Y, double eta, int nx,
                      const double *bootstrap, int bootstrap_cols, int max_iters,
                      int *is_certified, double *theta, double *x, double *all_eigvecs,
                      int all_eigvecs_cols_capacity, int *all_eigvecs_cols,
                      int64_t *num_iters);
/* The PSD half of fast_verification alone (src/CORA_utils.cpp:33-57): *is_psd = 1 iff S(Y) + eta I is positive
 * definite (Cholesky on the device).  Y == NULL tests the resident iterate of rank r.  No sv-ratio short-circuit
 * (src/CORA_problem.cpp:1039-1049), no eigen-search.  ENOTIMPL on graphs without a device factorisation. */
int cora_b200_psd_test(cora_b200_t *h, int r, const double *Y, double eta, int *is_psd);

/* Test hook: smallest eigenpair of the handle's OWN matrix by the device Lanczos the certification uses (the
 * reference's eigenpair known answers, tests/test_certification.cpp:45-79, load I - 2 x x^T as a matrix of landmark
 * rows only: d any, n_poses = n_ranges = 0, n_trans = N). */
int cora_b200_debug_min_eigenpair(cora_b200_t *h, int max_iters, double *theta, double *x /* N */, int *steps);
/* Which test decided the last cora_b200_certify / staircase stage on this handle (CORA_B200_CERT_*). */
int cora_b200_last_cert_branch(const cora_b200_t *h, int *branch);


/* saddleEscape (src/CORA.cpp:245-350): Y is N x (r_new-1), v has N entries, Y_out
 * is N x r_new. */
int cora_b200_saddle_escape(cora_b200_t *h, int r_new, const double *Y, double theta,
                            const double *v, double gradient_tolerance,
                            double preconditioned_gradient_tolerance, double *Y_out);

/* projectSolution (src/CORA.cpp:352-441): N x r -> N x d */
int cora_b200_project_solution(cora_b200_t *h, int r, const double *Y, double *Y_out);

/* One stage record of the staircase (one TNT + one certification). */
/* How certify_solution (src/CORA_problem.cpp:1030-1103) reached its verdict. */
enum {
  CORA_B200_CERT_NONE = 0,
  CORA_B200_CERT_SV_RATIO = 1,     /* sigma_max/sigma_min(Y) > 1e6 short-circuit (:1039-1049): certified                 */
  CORA_B200_CERT_PSD = 2,          /* Cholesky of S + eta I succeeded (src/CORA_utils.cpp:33-57): certified              */
  CORA_B200_CERT_EIGENPAIR = 3,    /* S + eta I not PD; eigen-search found x' S x < -eta/2: not certified, direction in x */
  CORA_B200_CERT_INCONCLUSIVE = 4  /* no factorisation of S + eta I available for this graph and the eigen-search found
                                      no negative curvature: NOT certified, and no descent direction either              */
};

typedef struct cora_b200_stage {
  int32_t rank;
  int32_t status;
  int32_t num_outer;
  int32_t certified;
  int64_t cg_iterations;
  double f;
  double gradfx_norm;
  double theta;
  double eta;
  double tnt_seconds;
  double cert_seconds;
  int32_t cert_branch;  /* CORA_B200_CERT_*: which test decided `certified` */
  int32_t reserved;
} cora_b200_stage;

typedef struct cora_b200_solve_result {
  double f;             /* final (rank-d refined) cost                         */
  double lifted_f;      /* cost at the certified lifted rank                   */
  int32_t final_rank;   /* = d after rounding                                  */
  int32_t lifted_rank;
  int32_t certified;    /* certificate at the lifted rank                      */
  int32_t num_stages;
  int64_t total_cg_iterations;
  double seconds;       /* solve-to-certificate wall-clock                     */
  int32_t stage_capacity;
  int32_t refined_certified; /* certificate of the refined rank-d solution (the reference's last certify_solution call) */
  cora_b200_stage *stages; /* caller-allocated, stage_capacity entries or NULL */
} cora_b200_solve_result;

/* solveCORA (src/CORA.cpp:26-243): staircase from rank r0 up to max_rank, rounding
 * and refinement.  X0 is N x r0; X_out is N x d. */
int cora_b200_solve(cora_b200_t *h, int r0, const double *X0, int max_rank,
                    const cora_b200_tnt_params *p, int verbose, double *X_out,
                    cora_b200_solve_result *res);

/* ---- multi-GPU helper: every rank passes its own result; on return
 * winner_rank is the arg-min of f over certified ranks (over all ranks when none
 * is certified) and X_inout (N x r_max, column-major, zero padded) holds the
 * winner's iterate on every rank.  `nccl_comm` is an ncclComm_t. */
int cora_b200_gather_best(void *nccl_comm, cora_b200_t *h, int world_size, int my_rank,
                          int r_max, double f, int certified, double *X_inout,
                          int *winner_rank, double *winner_f);
/* Same selection; the iterates stay on the devices: every rank's RESIDENT iterate (rank r, as left by
 * cora_b200_solve / cora_b200_tnt_resident) is replaced by the winner's with one ncclBroadcast, no host staging. */
int cora_b200_gather_best_resident(void *nccl_comm, cora_b200_t *h, int world_size, int my_rank, double f,
                                   int certified, int *winner_rank, double *winner_f);

/* the selection rule of gather_best as a pure host function (CPU-testable): arg-min of f over the
 * certified ranks, over all ranks when none is certified; ties go to the lowest rank */
int cora_b200_select_best(int world_size, const double *f, const int *certified, int *winner);
/* NCCL communicator plumbing (libnccl is dlopen-ed at first use): rank 0 creates the 128-byte
 * unique id and ships it to the other ranks by any out-of-band channel; every rank then calls
 * cora_b200_nccl_init with its CUDA device. */
int cora_b200_nccl_unique_id(void *id128);
int cora_b200_nccl_init(void **comm, int device, int world_size, int rank, const void *id128);
int cora_b200_nccl_destroy(void *comm);

/* ---- host-side assembly of the data matrix: Problem::fillDataMatrix and friends
 * (src/CORA_problem.cpp:115-377, 625-712) from the flattened measurement stacks, in
 * O(#factors log #factors) (the reference's add*Measurement is O(M^2)).
 * Translation indices address [n poses | l landmarks]; rot_R is row-major d x d per
 * factor.  Two-phase: call with rowptr == NULL to assemble and obtain *nnz, then again
 * with buffers (rowptr N+1, col/val *nnz entries) to copy the CSR out.  CPU only. */
int cora_b200_assemble(int d, int n_poses, int n_landmarks, int64_t E, const int64_t *rp_i,
                       const int64_t *rp_j, const double *rp_t, const double *rp_tau, int64_t Ep,
                       const int64_t *rot_i, const int64_t *rot_j, const double *rot_R,
                       const double *rot_kappa, int64_t m, const int64_t *rg_a, const int64_t *rg_b,
                       const double *rg_r, const double *rg_w, int64_t *nnz, int32_t *rowptr,
                       int32_t *col, double *val);

/* ---- either side of the solver (host): the odometry initialisation of the reference's experiments
 * (getOdomInitialization, examples/paper_experiments.cpp:426-534) from the measurement stacks of
 * cora_b200_assemble, and export of a rounded N x d solution (src/CORA_utils.cpp:204-350: getRotation's
 * determinant / orthogonality checks included -> ERUNTIME).  X is column-major in the reference row order.
 * reference_sign != 0: range rows initialised with (second - first) as paper_experiments.cpp:500-506 does; 0: the
 * sign the data matrix implies.  Every random draw (start pose of further chains, landmarks, SO(rank) factor)
 * comes from `seed` (the reference's are unseeded). */
int cora_b200_odometry_initialization(int d, int n_poses, int n_landmarks, int64_t E, const int64_t *rp_i,
                                      const int64_t *rp_j, const double *rp_t, int64_t Ep, const int64_t *rot_i,
                                      const int64_t *rot_j, const double *rot_R, int64_t m, const int64_t *rg_a,
                                      const int64_t *rg_b, int rank, uint64_t seed, int reference_sign,
                                      double *X_out /* N x rank */);
/* format 0: TUM "time x y z qx qy qz qw" (saveSolnToTum), 1: g2o VERTEX_SE3:QUAT / VERTEX_SE2 (saveSolnToG20);
 * poses first_pose .. first_pose + count - 1 in index order, time = position in that list. */
int cora_b200_save_solution(const char *path, int format, int d, int n_poses, int n_ranges, int n_trans,
                            const double *X /* N x d */, int64_t first_pose, int64_t count);

/* ---- PyFG text -> measurement stacks (host): parsePyfgTextToProblem (src/pyfg_text_parser.cpp:112-321)
 * with the data model of src/CORA_problem.cpp:24-113 and the precisions of
 * include/CORA/Measurements.h:79-152.  `path_or_text` is a file name, or the file contents when
 * from_text != 0.  Errors as the reference: unknown keyword / malformed line -> ERUNTIME, duplicate
 * variable or measurement, unknown symbol -> EINVAL.  The stacks (caller-allocated with the sizes
 * cora_b200_pyfg_sizes reports; rp_t is E x d, rot_R is Ep x d x d row-major) are the input of
 * cora_b200_assemble. */
typedef struct cora_b200_pyfg cora_b200_pyfg_t;
int cora_b200_pyfg_parse(const char *path_or_text, int from_text, cora_b200_pyfg_t **out);
int cora_b200_pyfg_sizes(const cora_b200_pyfg_t *g, int *d, int *n_poses, int *n_landmarks, int64_t *E,
                         int64_t *Ep, int64_t *m);
int cora_b200_pyfg_arrays(const cora_b200_pyfg_t *g, int64_t *rp_i, int64_t *rp_j, double *rp_t,
                          double *rp_tau, int64_t *rot_i, int64_t *rot_j, double *rot_R,
                          double *rot_kappa, int64_t *rg_a, int64_t *rg_b, double *rg_r, double *rg_w);
int cora_b200_pyfg_free(cora_b200_pyfg_t *g);

/* ---- test hooks (no compute): the internal device layout, rebuilt into CSR on
 * the host so the CPU test-suite can check the permutation / block-ELL / spill
 * split without a GPU.  Buffers are caller-allocated (nnz entries). */
int cora_b200_layout_roundtrip(int d, int n_poses, int n_ranges, int n_trans,
                               const int32_t *rowptr, const int32_t *col, const double *val,
                               int64_t nnz, int32_t *out_rowptr, int32_t *out_col,
                               double *out_val, int64_t *stats /* 8 entries */);
/* Same, through the STRIP layout (per-warp records of cora_b200/csrc/stream_layout.hpp) that the
 * rank-specialised streaming kernels consume. */
int cora_b200_strip_layout_roundtrip(int d, int n_poses, int n_ranges, int n_trans,
                                     const int32_t *rowptr, const int32_t *col, const double *val,
                                     int64_t nnz, int32_t *out_rowptr, int32_t *out_col,
                                     double *out_val, int64_t *stats /* 8 entries */);

/* Test hook (CPU only, no GPU): chain factorisation of (Q + shift I) [last row pinned when
 * pin_last] and M^-1 V executed on the host through the same per-chunk routines the device
 * kernels call.  V/out may be NULL (factor only); *pos_def receives the PD verdict. */
int cora_b200_debug_chain_host(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                               const int32_t *col, const double *val, int64_t nnz, double shift,
                               int pin_last, int r, const double *V, double *out, int *pos_def);

/* Test hook (CPU only): which factorisation the pose system of this matrix gets and its size.
 * stats[0] = 1: odometry chain (levels of chunks); 0: general sparse block Cholesky (loop closures, several
 * robots), then [1] pose couplings, [2] off-diagonal blocks of L, [3] elimination-tree height, [4] clusters,
 * [5] cluster levels (grid barriers per sweep of a triangular solve), [6] largest column, [7] poses, [8] blocks
 * of the per-cluster inverses, [9] most row blocks one cluster reads, [10] longest row of L in blocks, [11] rows
 * longer than 64 blocks. */
int cora_b200_debug_factor_stats(int d, int n_poses, int n_ranges, int n_trans, const int32_t *rowptr,
                                 const int32_t *col, const double *val, int64_t nnz, int64_t *stats /* 12 */);

#ifdef __cplusplus
}
#endif
#endif /* CORA_B200_H_ */
