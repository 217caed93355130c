// reduce_i64.cu -- axis-reduction kernels instantiated for int64_t (templates: reduce_kernels.cuh, reduce_plan.cuh).
#include "reduce_plan.cuh"

FEO_REDUCE_INSTANTIATE(int64_t, i64)
