// reduce_u32.cu -- axis-reduction kernels instantiated for uint32_t (templates: reduce_kernels.cuh, reduce_plan.cuh).
#include "reduce_plan.cuh"

FEO_REDUCE_INSTANTIATE(uint32_t, u32)
