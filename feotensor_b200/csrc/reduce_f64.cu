// reduce_f64.cu -- axis-reduction kernels instantiated for double (templates: reduce_kernels.cuh, reduce_plan.cuh).
#include "reduce_plan.cuh"

FEO_REDUCE_INSTANTIATE(double, f64)
