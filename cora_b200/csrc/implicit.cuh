// implicit.cuh -- Formulation::Implicit, the translation-marginalised form of the problem
// (src/CORA_problem.cpp:714-757 fillImplicitFormulationMatrices / dataMatrixProduct, :878-885 precondition,
// :1085-1100 certificate truncation, :1168-1197 getTranslationExplicitSolution; used by the paper experiments,
// examples/config.json:8).
//
// The variable is Y = [rotations; ranges] (d n + m rows); the reference multiplies with
//     Qmain Y - T_red chol(L_red)^-1 T_red^T Y ,   Qmain = Q[0:dn+m, 0:dn+m], T_red = Q[0:dn+m, dn+m:N-1],
//     L_red = Q33 without its last row and column.
// Here iterates keep all N internal rows with ZERO translation rows, and the product is evaluated as the rows of
// Q [Y; t*(Y)],  t*(Y) = -L_red^-1 T_red^T Y (last translation 0): the top rows are Qmain Y + T_red t* -- the
// reference's expression -- and the translation rows vanish (T^T Y + L t* = 0; the pinned last row too, because
// the rows of [T^T L] sum to zero: Q annihilates a common shift of all translations).  So the fused product
// kernels (gradient, Hessian, CG epilogues) run unchanged on the completed copy:
//     1. F = Y with zero translation rows              (k_translation_rows, mode 0)
//     2. G0 = Q F                                      (translation rows of G0 = T^T Y)
//     3. Z = Ltrans^-1 G0                              (factor of Q33 alone, last translation pinned: the chain /
//                                                       general pose-graph Cholesky in translation-only mode)
//     4. translation rows of F = -Z                    (mode 1)
// L_red is the translation Laplacian of the same pose graph + landmark border, so the solvers of chain_chol.cuh /
// gen_chol.hpp apply.
#pragma once
#include "chain_factor_dev.cuh"

namespace cora_b200 {

inline void implicit_ensure(H *h, int r) {
  if (h->formulation != CORA_B200_FORMULATION_IMPLICIT) throw Error(CORA_B200_EINVAL, "handle is not in the implicit formulation");
  const size_t need = (size_t)h->DL.numTiles * h->DL.TR * std::max(r, h->ws_r);
  for (auto &b : h->d_imp)
    if (b.n < need) {
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      b.alloc(need);
    }
  if (!h->ltrans) {
    bool pd = false;
    h->ltrans = build_chain_chol(h, h->d_bval.p, h->d_sdiag.p, 0.0, /*pin_last=*/true, &pd, /*want_solve=*/true,
                                 /*trans_only=*/true);
    if (!pd) {
      destroy_chain_chol(h->ltrans);
      h->ltrans = nullptr;
      throw Error(CORA_B200_ERUNTIME, "implicit formulation: the reduced translation Laplacian is not positive definite "
                                      "(disconnected translation graph?)");
    }
  }
}

inline void translation_rows(H *h, int mode, const double *x, const double *z, double *out, int r, const CgCtrl *ctrl) {
  const long long nE = (long long)h->DL.N * r;
  k_translation_rows<<<flat_grid(h, nE), kThreads, 0, h->stream>>>(mode, h->DL.nPoseRows, h->DL.D1, h->DL.l, r, x, z, out,
                                                                  nE, ctrl);
  check_launch(h);
}

inline void zero_translation_rows(H *h, double *V, int r, const CgCtrl *ctrl) {
  translation_rows(h, 2, nullptr, nullptr, V, r, ctrl);
}

// [Y; t*(Y)] in scratch (valid until the next completion)
inline const double *implicit_complete(H *h, const double *X, int r, CgCtrl *ctrl) {
  implicit_ensure(h, r);
  double *F = h->d_imp[0].p, *G0 = h->d_imp[1].p, *Z = h->d_imp[2].p;
  translation_rows(h, 0, X, nullptr, F, r, ctrl);
  launch_qprod_raw(h, QM_SPMM, F, nullptr, nullptr, G0, nullptr, r, POST_STORE, SC_TMP + 4, ctrl);
  chain_solve(h, h->ltrans, G0, Z, r, ctrl);
  translation_rows(h, 1, nullptr, Z, F, r, ctrl);
  return F;
}

inline void set_formulation(H *h, int formulation) {
  if (formulation != CORA_B200_FORMULATION_EXPLICIT && formulation != CORA_B200_FORMULATION_IMPLICIT)
    throw Error(CORA_B200_EINVAL, "Unknown formulation");  // src/CORA_problem.cpp:755
  if (formulation == CORA_B200_FORMULATION_IMPLICIT && h->DL.n + h->DL.l < 2)
    throw Error(CORA_B200_EINVAL, "implicit formulation needs at least two translational states");
  h->formulation = formulation;
  h->resident_r = 0;
  if (formulation == CORA_B200_FORMULATION_IMPLICIT) implicit_ensure(h, std::max(h->ws_r, 1));
}

}  // namespace cora_b200
