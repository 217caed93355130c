// reduce_f32.cu -- axis-reduction kernels instantiated for float (templates: reduce_kernels.cuh, reduce_plan.cuh).
#include "reduce_plan.cuh"

FEO_REDUCE_INSTANTIATE(float, f32)
