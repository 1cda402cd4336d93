// pk_instance.cu -- ONE instantiation of the persistent kernels, k_tnt_persistent<PK_D, PK_R> and
// k_spmm_persistent<PK_D, PK_R>; compiled once per (d, rank) pair of cora_b200/build.py:PK_LIST into its own
// object (the kernels are large: separate translation units compile in parallel).  PK_R == 0 is the
// any-rank tile-pipeline kernel, PK_R > 0 the rank-specialised streaming kernel (stream.cuh).
#include "persistent_kernel.cuh"

#ifndef PK_D
#error "compile with -DPK_D=<2|3> -DPK_R=<rank>"
#endif

#define PK_CAT2(a, b, c, d) a##b##c##d
#define PK_CAT(a, b, c, d) PK_CAT2(a, b, c, d)

namespace cora_b200 {
void *PK_CAT(pk_tnt_, PK_D, _, PK_R)() { return (void *)k_tnt_persistent<PK_D, PK_R>; }
void *PK_CAT(pk_spmm_, PK_D, _, PK_R)() { return (void *)k_spmm_persistent<PK_D, PK_R>; }
}  // namespace cora_b200
