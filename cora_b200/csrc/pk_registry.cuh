// pk_registry.cuh -- lookup of the persistent kernel instantiations (pk_instance.cu) by (d, rank).
// CORA_PK_LIST is the X-macro list of compiled pairs, passed by cora_b200/build.py: X(3,5) X(3,0) ...
#pragma once

#ifdef CORA_PK_LIST_HEADER
#include "pk_list.gen.h"
#endif
#ifndef CORA_PK_LIST
#error "CORA_PK_LIST must list the compiled (d, rank) pairs"
#endif

namespace cora_b200 {
#define X(D, R) void *pk_tnt_##D##_##R(); void *pk_spmm_##D##_##R();
CORA_PK_LIST
#undef X

void *persistent_tnt_kernel(int d, int R) {
#define X(D, RR) if (d == D && R == RR) return pk_tnt_##D##_##RR();
  CORA_PK_LIST
#undef X
  return nullptr;
}
void *persistent_spmm_kernel(int d, int R) {
#define X(D, RR) if (d == D && R == RR) return pk_spmm_##D##_##RR();
  CORA_PK_LIST
#undef X
  return nullptr;
}
}  // namespace cora_b200
