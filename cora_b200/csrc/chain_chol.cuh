// chain_chol.cuh -- structured Cholesky of (Q + lambda I) / (S + eta I) for chain + landmark
// graphs (RegularizedCholesky preconditioner and the PSD test of the certificate).
#pragma once
#include "ops.cuh"

namespace cora_b200 {

struct ChainChol {};

inline void destroy_chain_chol(ChainChol *c) { delete c; }

inline ChainChol *build_chain_chol(H *, const double *, const double *, double, bool, bool *) {
  throw Error(CORA_B200_ENOTIMPL, "RegularizedCholesky (chain Cholesky) not implemented yet");
}
inline void chain_solve(H *, ChainChol *, const double *, double *, int, const CgCtrl *) {
  throw Error(CORA_B200_ENOTIMPL, "RegularizedCholesky (chain Cholesky) not implemented yet");
}
inline double estimate_spectral_norm(H *) {
  throw Error(CORA_B200_ENOTIMPL, "spectral norm estimate not implemented yet");
}

}  // namespace cora_b200
