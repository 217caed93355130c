// dotlike.cu -- K5: DynamicVector::vecmul (dot, vector.rs:71-78), batched dot, DynamicMatrix::vecmul
// (matrix.rs:92-103) and DynamicVector::matmul (vector.rs:80-91).  All three are sum-reductions of an
// elementwise product and reuse the reduction kernels with a multiplying loader:
//   dot / batched dot   rows geometry  in[batch][n] * in2[batch][n]      -> out[batch]
//   matrix . vector     rows geometry  A[M][K] * v[K] (v re-read, cached) -> out[M]
//   vector . matrix     cols geometry  v[K] (one scalar per row) * A[K][N] -> out[N]
// HBM-bound: algorithmic bytes = sizeof(T) * (elements read once + outputs).
#include "reduce_kernels.cuh"

using namespace feo_red;

namespace {

template <typename T>
int dotlike_typed(feo_ctx *ctx, int loader, T *out, const T *a, const T *b, size_t outer, size_t inner) {
    Canon c{true, 1, 1, 1, 1};
    switch (loader) {
    case LD_MUL:  // a, b: [outer][inner]
        c.K1 = (long long)outer; c.R2 = (long long)inner;
        return run_canonical<T, K_SUM, LD_MUL>(ctx, c, out, a, b, nullptr, FIN_NONE, (T)1);
    case LD_MATVEC:  // a = A[outer=M][inner=K], b = v[K]
        c.K1 = (long long)outer; c.R2 = (long long)inner;
        return run_canonical<T, K_SUM, LD_MATVEC>(ctx, c, out, a, b, nullptr, FIN_NONE, (T)1);
    case LD_VECMAT:  // a = A[outer=K][inner=N], b = v[K]
        c.R1 = (long long)outer; c.K2 = (long long)inner;
        return run_canonical<T, K_SUM, LD_VECMAT>(ctx, c, out, a, b, nullptr, FIN_NONE, (T)1);
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown loader %d", loader);
    }
}

}  // namespace

int feo_impl_dotlike(feo_ctx *ctx, int loader, int dtype, void *out, const void *a, const void *b, size_t outer, size_t inner) {
    FEO_DISPATCH(ctx, dtype, T, return dotlike_typed<T>(ctx, loader, (T *)out, (const T *)a, (const T *)b, outer, inner));
    return FEO_OK;
}
