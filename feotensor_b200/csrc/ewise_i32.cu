// ewise_i32.cu -- K1 elementwise / broadcast kernels instantiated for int32_t (templates: ewise_impl.cuh).
#include "ewise_impl.cuh"

FEO_EWISE_INSTANTIATE(int32_t, i32)
