// ewise_f64.cu -- K1 elementwise / broadcast kernels instantiated for double (templates: ewise_impl.cuh).
#include "ewise_impl.cuh"

FEO_EWISE_INSTANTIATE(double, f64)
