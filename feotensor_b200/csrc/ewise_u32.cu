// ewise_u32.cu -- K1 elementwise / broadcast kernels instantiated for uint32_t (templates: ewise_impl.cuh).
#include "ewise_impl.cuh"

FEO_EWISE_INSTANTIATE(uint32_t, u32)
