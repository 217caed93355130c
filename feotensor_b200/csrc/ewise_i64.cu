// ewise_i64.cu -- K1 elementwise / broadcast kernels instantiated for int64_t (templates: ewise_impl.cuh).
#include "ewise_impl.cuh"

FEO_EWISE_INSTANTIATE(int64_t, i64)
