// reduce_i32.cu -- axis-reduction kernels instantiated for int32_t (templates: reduce_kernels.cuh, reduce_plan.cuh).
#include "reduce_plan.cuh"

FEO_REDUCE_INSTANTIATE(int32_t, i32)
