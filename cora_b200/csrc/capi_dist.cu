// capi_dist.cu -- multi-GPU helper (NCCL): gather the best certified solution.
#include "ops.cuh"

using namespace cora_b200;

extern "C" int cora_b200_gather_best(void *, cora_b200_t *, int, int, int, double, int, double *, int *, double *) {
  set_last_error("gather_best not implemented yet");
  return CORA_B200_ENOTIMPL;
}
