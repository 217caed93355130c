// matmul_tf32x3.cu -- K7: DynamicMatrix::matmul (matrix.rs:76-90) for f32 on the tcgen05 tensor cores
// with split-TF32 (3xTF32).  Placeholder until the tcgen05/TMEM/TMA kernel lands: reports the
// algorithm as unavailable so FEO_MATMUL_AUTO resolves to the SIMT kernel.
#include "feo_common.cuh"

bool feo_matmul_tf32x3_supported(size_t, size_t, size_t) { return false; }

int feo_impl_matmul_tf32x3(feo_ctx *ctx, void *, const void *, const void *, size_t, size_t, size_t) {
    return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "TF32X3 matmul is not built into this library yet");
}
