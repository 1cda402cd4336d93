// certify.cuh -- certification, saddle escape, rounding, staircase (placeholder).
#pragma once
#include "solver.cuh"

namespace cora_b200 {
inline void certify_host(H *, int, const double *, double, int, const double *, int, int, int *, double *,
                         double *, double *, int, int *, int64_t *) {
  throw Error(CORA_B200_ENOTIMPL, "certify not implemented yet");
}
inline void saddle_escape_host(H *, int, const double *, double, const double *, double, double, double *) {
  throw Error(CORA_B200_ENOTIMPL, "saddle escape not implemented yet");
}
inline void project_solution_host(H *, int, const double *, double *) {
  throw Error(CORA_B200_ENOTIMPL, "project solution not implemented yet");
}
inline void solve_staircase(H *, int, const double *, int, const cora_b200_tnt_params &, int, double *,
                            cora_b200_solve_result *) {
  throw Error(CORA_B200_ENOTIMPL, "staircase not implemented yet");
}
}  // namespace cora_b200
