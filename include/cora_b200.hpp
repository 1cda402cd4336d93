// cora_b200.hpp -- header-only C++17 host side above the C-ABI (cora_b200.h): the reference's
// operator interface for the staircase inner loop with the same names, argument meaning and error
// behaviour, so that a caller of CORA::Problem / CORA::solveCORA (include/CORA/CORA_problem.h:340-397,
// include/CORA/CORA.h:22-38 of MarineRoboticsGroup/cora @ 015dc43) switches by changing a namespace.
//
// The reference's dense type is Eigen::MatrixXd (column-major f64); Eigen is not a dependency of
// this repository, so `Matrix` below is a minimal column-major container with the same memory
// layout -- `Eigen::Map<Eigen::MatrixXd>(M.data(), M.rows(), M.cols())` views it in place, and
// INTEGRATION.md shows the one-line adaptor in the other direction.  The data matrix is handed over
// as the three CSR arrays of the reference's Eigen::SparseMatrix<double, RowMajor>
// (valuePtr / innerIndexPtr / outerIndexPtr after makeCompressed()).
//
// Error behaviour (include/CORA/CORA_types.h:15-39): CORA_B200_EINVAL -> std::invalid_argument
// (also used where the reference throws MatrixShapeException), CORA_B200_ENOTIMPL ->
// cora_b200::NotImplementedException, everything else -> std::runtime_error.  There is no CPU
// fallback: without a CUDA device every call that needs one throws std::runtime_error.
#ifndef CORA_B200_HPP_
#define CORA_B200_HPP_

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cora_b200.h"

namespace cora_b200 {

typedef double Scalar;

// Column-major dense matrix, the layout of Eigen::MatrixXd (include/CORA/CORA_types.h:47-48).
class Matrix {
 public:
  Matrix() = default;
  Matrix(std::ptrdiff_t rows, std::ptrdiff_t cols) : rows_(rows), cols_(cols), v_((size_t)rows * cols, 0.0) {}
  Matrix(std::ptrdiff_t rows, std::ptrdiff_t cols, const double *src) : rows_(rows), cols_(cols), v_(src, src + (size_t)rows * cols) {}
  std::ptrdiff_t rows() const { return rows_; }
  std::ptrdiff_t cols() const { return cols_; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return v_[(size_t)j * rows_ + i]; }
  double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return v_[(size_t)j * rows_ + i]; }

 private:
  std::ptrdiff_t rows_ = 0, cols_ = 0;
  std::vector<double> v_;
};
typedef std::vector<double> Vector;

struct NotImplementedException : public std::logic_error {  // include/CORA/CORA_types.h:15-19
  explicit NotImplementedException(const std::string &m) : std::logic_error(m) {}
};

// enum class Preconditioner (include/CORA/CORA_types.h:76-77)
enum class Preconditioner { None = CORA_B200_PRECON_NONE, Jacobi = CORA_B200_PRECON_JACOBI,
                            BlockCholesky = CORA_B200_PRECON_BLOCK_CHOLESKY,
                            RegularizedCholesky = CORA_B200_PRECON_REG_CHOLESKY };
// enum class Formulation (include/CORA/CORA_types.h:50-55)
enum class Formulation { Explicit = CORA_B200_FORMULATION_EXPLICIT, Implicit = CORA_B200_FORMULATION_IMPLICIT };

// CORA::CertResults (include/CORA/CORA_types.h:58-64)
struct CertResults {
  bool is_certified = false;
  Scalar theta = 0;
  Vector x;
  Matrix all_eigvecs;
  size_t num_iters = 0;
};

// Optimization::Riemannian::TNTResult<Matrix, Scalar> (TNT.h:168-194 and its bases)
enum class TNTStatus { Gradient, PreconditionedGradient, RelativeDecrease, Stepsize, TrustRegion,
                       IterationLimit, ElapsedTime, UserFunction };
struct CoraTntResult {
  Matrix x;
  Scalar f = 0, gradfx_norm = 0, preconditioned_grad_f_x_norm = 0, elapsed_time = 0;
  TNTStatus status = TNTStatus::IterationLimit;
  std::vector<Scalar> objective_values, gradient_norms, preconditioned_gradient_norms, trust_region_radius,
      time, update_step_norms, update_step_M_norms, gain_ratios;
  std::vector<size_t> inner_iterations;
};
typedef std::pair<CoraTntResult, std::vector<Matrix>> CoraResult;  // include/CORA/CORA.h:20

inline void check(int code) {
  if (code == CORA_B200_OK) return;
  const std::string msg = cora_b200_last_error();
  if (code == CORA_B200_EINVAL) throw std::invalid_argument(msg);
  if (code == CORA_B200_ENOTIMPL) throw NotImplementedException(msg);
  throw std::runtime_error(msg);
}

// The device side of CORA::Problem after updateProblemData(): the data matrix laid out on one GPU.
// Method names and semantics are those of include/CORA/CORA_problem.h:327-397.
class Problem {
 public:
  // dim d, n poses, m range measurements, n + l translations; CSR of the data matrix in the
  // reference row order (src/CORA_problem.cpp:964-1021).  What Problem::updateProblemData()
  // (src/CORA_problem.cpp:500-510) has produced when the solver starts.
  Problem(int dim, int num_poses, int num_ranges, int num_translations, const int32_t *outer_index,
          const int32_t *inner_index, const double *values, int64_t nnz, int relaxation_rank,
          Preconditioner preconditioner = Preconditioner::RegularizedCholesky, int device = 0, void *stream = nullptr)
      : dim_(dim), n_(num_poses), m_(num_ranges), nt_(num_translations), rank_(relaxation_rank),
        preconditioner_(preconditioner) {
    check(cora_b200_create(&h_, device, stream, dim, num_poses, num_ranges, num_translations, outer_index,
                           inner_index, values, nnz, (int)preconditioner, 0.0));
  }
  ~Problem() { cora_b200_destroy(h_); }
  Problem(const Problem &) = delete;
  Problem &operator=(const Problem &) = delete;

  int dim() const { return dim_; }
  int numPoses() const { return n_; }
  int numRangeMeasurements() const { return m_; }
  int numTranslationalStates() const { return nt_; }
  std::ptrdiff_t getDataMatrixSize() const { return (std::ptrdiff_t)dim_ * n_ + m_ + nt_; }  // :940-942
  int getRelaxationRank() const { return rank_; }
  void setRank(int r) { rank_ = r; }           // CORA_problem.h:331-334
  void incrementRank() { ++rank_; }            // :327-330
  void setPreconditioner(Preconditioner p) {   // :335-337 + updatePreconditioner
    check(cora_b200_set_preconditioner(h_, (int)p, 0.0));
    preconditioner_ = p;
  }
  void setFormulation(Formulation f) {  // CORA_problem.h:338 (+ fillImplicitFormulationMatrices, :714-741)
    check(cora_b200_set_formulation(h_, (int)f));
    formulation_ = f;
  }
  Formulation getFormulation() const { return formulation_; }  // CORA_problem.h:290
  std::ptrdiff_t rotAndRangeMatrixSize() const { return (std::ptrdiff_t)dim_ * n_ + m_; }
  std::ptrdiff_t getExpectedVariableSize() const {  // src/CORA_problem.cpp:944-954
    return formulation_ == Formulation::Explicit ? getDataMatrixSize() : rotAndRangeMatrixSize();
  }
  Matrix getTranslationExplicitSolution(const Matrix &Y) const {  // src/CORA_problem.cpp:1168-1197
    shape(Y, "Y");
    Matrix X(getDataMatrixSize(), Y.cols());
    check(cora_b200_translation_explicit_solution(h_, (int)Y.cols(), Y.data(), X.data()));
    return X;
  }

  Scalar evaluateObjective(const Matrix &Y) const {
    shape(Y, "Y");
    Scalar f = 0;
    check(cora_b200_objective(h_, (int)Y.cols(), Y.data(), &f));
    return f;
  }
  Matrix Euclidean_gradient(const Matrix &Y) const {
    shape(Y, "Y");
    Matrix G(Y.rows(), Y.cols());
    check(cora_b200_egrad(h_, (int)Y.cols(), Y.data(), G.data()));
    return G;
  }
  Matrix Riemannian_gradient(const Matrix &Y) const {
    shape(Y, "Y");
    Matrix G(Y.rows(), Y.cols());
    check(cora_b200_rgrad(h_, (int)Y.cols(), Y.data(), nullptr, G.data()));
    return G;
  }
  Matrix Riemannian_gradient(const Matrix &Y, const Matrix &NablaF_Y) const {
    shape(Y, "Y"); same(Y, NablaF_Y, "NablaF_Y");
    Matrix G(Y.rows(), Y.cols());
    check(cora_b200_rgrad(h_, (int)Y.cols(), Y.data(), NablaF_Y.data(), G.data()));
    return G;
  }
  Matrix Riemannian_Hessian_vector_product(const Matrix &Y, const Matrix &NablaF_Y, const Matrix &Ydot) const {
    shape(Y, "Y"); same(Y, NablaF_Y, "NablaF_Y"); same(Y, Ydot, "Ydot");
    Matrix H(Y.rows(), Y.cols());
    check(cora_b200_hessvec(h_, (int)Y.cols(), Y.data(), NablaF_Y.data(), Ydot.data(), H.data()));
    return H;
  }
  Matrix tangent_space_projection(const Matrix &Y, const Matrix &Ydot) const {
    shape(Y, "Y"); same(Y, Ydot, "Ydot");
    Matrix out(Y.rows(), Y.cols());
    check(cora_b200_tangent_proj(h_, (int)Y.cols(), Y.data(), Ydot.data(), out.data()));
    return out;
  }
  Matrix precondition(const Matrix &V) const {
    shape(V, "V");
    Matrix out(V.rows(), V.cols());
    check(cora_b200_precondition(h_, (int)V.cols(), V.data(), out.data()));
    return out;
  }
  Matrix projectToManifold(const Matrix &A) const {
    shape(A, "A");
    Matrix out(A.rows(), A.cols());
    check(cora_b200_project(h_, (int)A.cols(), A.data(), out.data()));
    return out;
  }
  Matrix retract(const Matrix &Y, const Matrix &V) const {
    shape(Y, "Y"); same(Y, V, "V");
    Matrix out(Y.rows(), Y.cols());
    check(cora_b200_retract(h_, (int)Y.cols(), Y.data(), V.data(), out.data()));
    return out;
  }
  // LambdaBlocks = (d x dn block row, m range multipliers), CORA_problem.h:355, :385
  std::pair<Matrix, Vector> compute_Lambda_blocks(const Matrix &Y) const {
    shape(Y, "Y");
    Matrix st(dim_, (std::ptrdiff_t)dim_ * n_);
    Vector ob((size_t)(m_ > 0 ? m_ : 1));
    check(cora_b200_lambda_blocks(h_, (int)Y.cols(), Y.data(), st.data(), ob.data()));
    ob.resize((size_t)m_);
    return {std::move(st), std::move(ob)};
  }
  // certify_solution, CORA_problem.h:371-375 (max_fill_factor / drop_tol belong to the reference's
  // ILDL preconditioner, which the Lanczos search here does not use)
  CertResults certify_solution(const Matrix &Y, Scalar eta, size_t nx, const Matrix &eigvec_bootstrap,
                               size_t max_LOBPCG_iters = 500, Scalar = 3, Scalar = 1e-3) const {
    shape(Y, "Y");
    const std::ptrdiff_t N = getExpectedVariableSize();  // (implicit: the truncated direction, :1085-1100)
    const int cap = (int)std::max<size_t>(nx, (size_t)Y.cols() + 2);
    CertResults out;
    out.x.assign((size_t)N, 0.0);
    Matrix ev(N, cap);
    int cert = 0, ncols = 0;
    int64_t iters = 0;
    const bool has_boot = eigvec_bootstrap.rows() == N && eigvec_bootstrap.cols() > 0;
    check(cora_b200_certify(h_, (int)Y.cols(), Y.data(), eta, (int)nx, has_boot ? eigvec_bootstrap.data() : nullptr,
                            has_boot ? (int)eigvec_bootstrap.cols() : 0, (int)max_LOBPCG_iters, &cert, &out.theta,
                            out.x.data(), ev.data(), cap, &ncols, &iters));
    out.is_certified = cert != 0;
    out.num_iters = (size_t)iters;
    out.all_eigvecs = Matrix(N, ncols, ev.data());
    return out;
  }

  cora_b200_t *handle() const { return h_; }

 private:
  void shape(const Matrix &Y, const char *name) const {  // checkMatrixShape, CORA_types.h:23-39
    if (Y.rows() != getExpectedVariableSize() || Y.cols() < 1)
      throw std::invalid_argument(std::string(name) + " has the wrong shape: expected " +
                                  std::to_string(getExpectedVariableSize()) + " rows, got " + std::to_string(Y.rows()) +
                                  " x " + std::to_string(Y.cols()));
  }
  static void same(const Matrix &A, const Matrix &B, const char *name) {
    if (A.rows() != B.rows() || A.cols() != B.cols())
      throw std::invalid_argument(std::string(name) + " must have the shape of Y");
  }
  cora_b200_t *h_ = nullptr;
  int dim_, n_, m_, nt_, rank_;
  Preconditioner preconditioner_;
  Formulation formulation_ = Formulation::Explicit;
};

namespace detail {
inline CoraTntResult unpack(const cora_b200_tnt_result &r, std::vector<std::vector<double>> &tr,
                            std::vector<int32_t> &inner, Matrix &&x) {
  CoraTntResult out;
  out.x = std::move(x);
  out.f = r.f; out.gradfx_norm = r.gradfx_norm; out.preconditioned_grad_f_x_norm = r.preconditioned_gradfx_norm;
  out.elapsed_time = r.elapsed_time;
  out.status = (TNTStatus)r.status;
  const size_t k = (size_t)r.num_outer;
  auto cut = [](std::vector<double> &v, size_t n) { v.resize(std::min(v.size(), n)); return v; };
  out.objective_values = cut(tr[0], k + 1); out.gradient_norms = cut(tr[1], k + 1);
  out.preconditioned_gradient_norms = cut(tr[2], k + 1); out.trust_region_radius = cut(tr[3], k + 1);
  out.time = cut(tr[4], k + 1); out.update_step_norms = cut(tr[5], k); out.update_step_M_norms = cut(tr[6], k);
  out.gain_ratios = cut(tr[7], k);
  for (size_t i = 0; i < k && i < inner.size(); ++i) out.inner_iterations.push_back((size_t)inner[i]);
  return out;
}
}  // namespace detail

// One call of Optimization::Riemannian::TNT with solveCORA's closures and parameters
// (src/CORA.cpp:52-122,139) on the device.  `params` == nullptr: the values of src/CORA.cpp:95-109.
inline CoraTntResult TNT(Problem &problem, const Matrix &x0, const cora_b200_tnt_params *params = nullptr) {
  cora_b200_tnt_params p;
  if (params) p = *params; else check(cora_b200_tnt_default_params(&p));
  const int cap = p.max_iterations + 2;
  std::vector<std::vector<double>> tr(8, std::vector<double>((size_t)cap, 0.0));
  std::vector<int32_t> inner((size_t)cap, 0);
  cora_b200_tnt_result r{};
  r.trace_capacity = cap;
  r.objective_values = tr[0].data(); r.gradient_norms = tr[1].data(); r.preconditioned_gradient_norms = tr[2].data();
  r.trust_region_radius = tr[3].data(); r.time = tr[4].data(); r.update_step_norms = tr[5].data();
  r.update_step_M_norms = tr[6].data(); r.gain_ratios = tr[7].data(); r.inner_iterations = inner.data();
  Matrix x(x0.rows(), x0.cols());
  check(cora_b200_tnt(problem.handle(), (int)x0.cols(), x0.data(), &p, x.data(), &r));
  return detail::unpack(r, tr, inner, std::move(x));
}

// saddleEscape, include/CORA/CORA.h:33-35 (src/CORA.cpp:245-350): Y is N x r, returns N x (r+1)
inline Matrix saddleEscape(const Problem &problem, const Matrix &Y, Scalar theta, const Vector &v,
                           Scalar gradient_tolerance, Scalar preconditioned_gradient_tolerance) {
  if ((std::ptrdiff_t)v.size() != problem.getExpectedVariableSize()) throw std::invalid_argument("v has the wrong size");
  Matrix out(Y.rows(), Y.cols() + 1);
  check(cora_b200_saddle_escape(problem.handle(), (int)Y.cols() + 1, Y.data(), theta, v.data(), gradient_tolerance,
                                preconditioned_gradient_tolerance, out.data()));
  return out;
}

// projectSolution, include/CORA/CORA.h:37-38 (src/CORA.cpp:352-441): N x r -> N x d
inline Matrix projectSolution(const Problem &problem, const Matrix &Y, bool = false) {
  Matrix out(Y.rows(), problem.dim());
  check(cora_b200_project_solution(problem.handle(), (int)Y.cols(), Y.data(), out.data()));
  return out;
}

// solveCORA, include/CORA/CORA.h:22-24 (src/CORA.cpp:26-243).  As in the reference the problem's
// rank is mutated (incrementRank / setRank(d), src/CORA.cpp:193,207); the iterate log holds the
// final iterate only (log_iterates feeds the reference's visualiser, which is out of scope).
inline CoraResult solveCORA(Problem &problem, const Matrix &x0, int max_relaxation_rank = 20, bool verbose = false,
                            bool /*log_iterates*/ = false, bool show_iterates = false) {
  if (x0.rows() != problem.getExpectedVariableSize())
    throw std::invalid_argument("x0 has the wrong number of rows");  // src/CORA.cpp:30-40
  cora_b200_tnt_params p;
  check(cora_b200_tnt_default_params(&p));
  p.verbose = show_iterates ? 1 : 0;
  cora_b200_solve_result res{};
  std::vector<cora_b200_stage> stages((size_t)2 * (max_relaxation_rank + 2));
  res.stage_capacity = (int)stages.size();
  res.stages = stages.data();
  Matrix x(x0.rows(), problem.dim());
  check(cora_b200_solve(problem.handle(), (int)x0.cols(), x0.data(), max_relaxation_rank, &p, verbose ? 1 : 0, x.data(),
                        &res));
  problem.setRank(problem.dim());
  CoraTntResult out;
  out.f = res.f;
  out.x = x;
  if (res.num_stages > 0) {
    const cora_b200_stage &s = stages[(size_t)std::min(res.num_stages, res.stage_capacity) - 1];
    out.gradfx_norm = s.gradfx_norm;
    out.status = (TNTStatus)s.status;
  }
  out.elapsed_time = res.seconds;
  return {out, std::vector<Matrix>{x}};
}

}  // namespace cora_b200
#endif  // CORA_B200_HPP_
