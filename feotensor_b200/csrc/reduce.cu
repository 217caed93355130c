// reduce.cu -- host planner for Tensor::{sum,mean,var,max,min} (tensor.rs:70-307): collapses the
// (shape, axes) pair to the canonical geometry of reduce_kernels.cuh, picks the kernel and the
// split count, and handles the reference's special cases:
//   * axes empty or nothing kept -> full reduction to shape [1] (tensor.rs:84-87);
//   * mean/var divide by `n` built in type T (tensor.rs:112-130) -- closed form feo_divisor<T>;
//   * integer var is the reference's own two-pass integer scheme (truncated mean, wrapping squares);
//   * float var is one pass (Welford/Chan).
#include "reduce_kernels.cuh"

using namespace feo_red;

#define FEO_DECL_TYPED(SFX) int feo_reduce_typed_##SFX(feo_ctx *, int, void *, const void *, const void *, int, const size_t *, int, const size_t *, int);
FEO_DECL_TYPED(f32) FEO_DECL_TYPED(f64) FEO_DECL_TYPED(i32) FEO_DECL_TYPED(i64) FEO_DECL_TYPED(u32)

namespace {

// Chan merge of equal-count per-shard partials [part][mean x nout | M2 x nout], in part order.
template <typename T>
__global__ void __launch_bounds__(kThreads) var_merge_kernel(T *__restrict__ out, const T *__restrict__ partials, long long nparts,
                                                             long long nout, T count, T divisor) {
    const long long o = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (o >= nout) return;
    Acc<T, K_WELFORD> acc;
    acc.init((T)0);
    for (long long p = 0; p < nparts; ++p) acc.merge_raw(count, partials[p * 2 * nout + o], partials[p * 2 * nout + nout + o]);
    out[o] = acc.m2 / divisor;
}

}  // namespace

int feo_impl_reduce(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, const void *aux, int rank,
                    const size_t *shape, int naxes, const size_t *axes, int mode) {
    switch (dtype) {
    case FEO_F32: return feo_reduce_typed_f32(ctx, kind, out, in, aux, rank, shape, naxes, axes, mode);
    case FEO_F64: return feo_reduce_typed_f64(ctx, kind, out, in, aux, rank, shape, naxes, axes, mode);
    case FEO_I32: return feo_reduce_typed_i32(ctx, kind, out, in, aux, rank, shape, naxes, axes, mode);
    case FEO_I64: return feo_reduce_typed_i64(ctx, kind, out, in, aux, rank, shape, naxes, axes, mode);
    case FEO_U32: return feo_reduce_typed_u32(ctx, kind, out, in, aux, rank, shape, naxes, axes, mode);
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    }
}

// Integers: partials are [part][S1 x nout | S2 x nout] (wrapping sum and sum of squares, Acc<T, K_SUMSQ>).
template <typename T>
__global__ void __launch_bounds__(kThreads) var_merge_int_kernel(T *__restrict__ out, const T *__restrict__ partials, long long nparts,
                                                                 long long nout, T divisor, int *flag) {
    const long long o = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (o >= nout) return;
    Acc<T, K_SUMSQ> acc;
    acc.init((T)0);
    for (long long p = 0; p < nparts; ++p) { acc.v = w_add(acc.v, partials[p * 2 * nout + o]); acc.q = w_add(acc.q, partials[p * 2 * nout + nout + o]); }
    finalize(acc, out, o, nout, FIN_DIV, divisor, flag);
}

int feo_impl_var_merge(feo_ctx *ctx, int dtype, void *out, const void *partials, size_t nparts, size_t nout,
                       size_t count_per_part, const void *divisor_host) {
    const unsigned blocks = (unsigned)feo_ceil_div(nout, (size_t)kThreads);
    if (dtype == FEO_F32) {
        var_merge_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float *)out, (const float *)partials, (long long)nparts,
                                                                       (long long)nout, (float)count_per_part, *(const float *)divisor_host);
    } else if (dtype == FEO_F64) {
        var_merge_kernel<double><<<blocks, kThreads, 0, ctx->stream>>>((double *)out, (const double *)partials, (long long)nparts,
                                                                        (long long)nout, (double)count_per_part, *(const double *)divisor_host);
    } else if (dtype == FEO_I32) {
        var_merge_int_kernel<int32_t><<<blocks, kThreads, 0, ctx->stream>>>((int32_t *)out, (const int32_t *)partials, (long long)nparts,
                                                                           (long long)nout, *(const int32_t *)divisor_host, ctx->arith_flag);
    } else if (dtype == FEO_I64) {
        var_merge_int_kernel<int64_t><<<blocks, kThreads, 0, ctx->stream>>>((int64_t *)out, (const int64_t *)partials, (long long)nparts,
                                                                           (long long)nout, *(const int64_t *)divisor_host, ctx->arith_flag);
    } else {
        var_merge_int_kernel<uint32_t><<<blocks, kThreads, 0, ctx->stream>>>((uint32_t *)out, (const uint32_t *)partials, (long long)nparts,
                                                                            (long long)nout, *(const uint32_t *)divisor_host, ctx->arith_flag);
    }
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}
