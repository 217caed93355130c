// ewise_f32.cu -- K1 elementwise / broadcast kernels instantiated for float (templates: ewise_impl.cuh).
#include "ewise_impl.cuh"

FEO_EWISE_INSTANTIATE(float, f32)
