// unity.cu -- main translation unit of libcora_b200.so (host logic + the small kernels live in headers);
// the persistent kernels are instantiated in their own objects (pk_instance.cu).
#include "capi_core.cu"
#include "capi_solver.cu"
#include "capi_dist.cu"
#include "pk_registry.cuh"
