// unity.cu -- single translation unit of libcora_b200.so (kernels live in headers).
#include "capi_core.cu"
#include "capi_solver.cu"
#include "capi_dist.cu"
