// reduce.cu -- host planner for Tensor::{sum,mean,var,max,min} (tensor.rs:70-307): collapses the
// (shape, axes) pair to the canonical geometry of reduce_kernels.cuh, picks the kernel and the
// split count, and handles the reference's special cases:
//   * axes empty or nothing kept -> full reduction to shape [1] (tensor.rs:84-87);
//   * mean/var divide by `n` built in type T (tensor.rs:112-130) -- closed form feo_divisor<T>;
//   * integer var is the reference's own two-pass integer scheme (truncated mean, wrapping squares);
//   * float var is one pass (Welford/Chan).
#include <string.h>

#include "reduce_kernels.cuh"

using namespace feo_red;

namespace {

struct Plan {
    int m;  // collapsed groups
    long long ext[FEO_MAX_RANK];
    bool red[FEO_MAX_RANK];
    long long nout;
    // un-collapsed dims for the generic kernel
    GenericGeom gen;
};

Plan make_plan(int rank, const size_t *shape, int naxes, const size_t *axes) {
    Plan p;
    memset(&p, 0, sizeof p);
    bool reduced[FEO_MAX_RANK];
    bool any_kept = false;
    for (int d = 0; d < rank; ++d) {
        reduced[d] = (naxes == 0);  // axes == [] reduces everything (tensor.rs:84)
        for (int i = 0; i < naxes; ++i) reduced[d] = reduced[d] || (axes[i] == (size_t)d);
        any_kept = any_kept || !reduced[d];
    }
    (void)any_kept;
    long long stride[FEO_MAX_RANK];
    long long s = 1;
    for (int d = rank - 1; d >= 0; --d) { stride[d] = s; s *= (long long)shape[d]; }
    p.nout = 1;
    p.gen.nout = 1;
    p.gen.nred = 1;
    for (int d = 0; d < rank; ++d) {
        const long long e = (long long)shape[d];
        if (!reduced[d]) p.nout *= e;
        if (e == 1) continue;
        if (reduced[d]) { p.gen.rext[p.gen.nr] = e; p.gen.rstr[p.gen.nr] = stride[d]; p.gen.nr++; p.gen.nred *= e; }
        else { p.gen.kext[p.gen.nk] = e; p.gen.kstr[p.gen.nk] = stride[d]; p.gen.nk++; p.gen.nout *= e; }
        if (p.m > 0 && p.red[p.m - 1] == reduced[d]) p.ext[p.m - 1] *= e;
        else { p.ext[p.m] = e; p.red[p.m] = reduced[d]; p.m++; }
    }
    return p;
}

template <typename T, int KIND>
int run_plan(feo_ctx *ctx, const Plan &p, T *out, const T *in, const T *aux, int fin, T divisor) {
    const Canon c = canonicalise(p.m, p.ext, p.red);
    if (c.ok) return run_canonical<T, KIND, LD_PLAIN>(ctx, c, out, in, nullptr, aux, fin, divisor);
    GenericGeom G = p.gen;
    G.fin = fin;
    long long blocks = G.nout;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    reduce_generic_kernel<T, KIND><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, in, aux, G, divisor, ctx->arith_flag);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

template <typename T>
int reduce_typed(feo_ctx *ctx, int kind, T *out, const T *in, const T *aux, int rank, const size_t *shape, int naxes,
                 const size_t *axes, int mode) {
    const Plan p = make_plan(rank, shape, naxes, axes);
    const T n = feo_divisor<T>(rank, shape, naxes, axes);
    if (mode == 2) return run_plan<T, K_SQDEV>(ctx, p, out, in, aux, FIN_NONE, n);
    const bool partial = (mode == 1);
    switch (kind) {
    case FEO_SUM: return run_plan<T, K_SUM>(ctx, p, out, in, nullptr, FIN_NONE, n);
    case FEO_MEAN: return run_plan<T, K_SUM>(ctx, p, out, in, nullptr, partial ? FIN_NONE : FIN_DIV, n);
    case FEO_MAX: return run_plan<T, K_MAX>(ctx, p, out, in, nullptr, FIN_NONE, n);
    case FEO_MIN: return run_plan<T, K_MIN>(ctx, p, out, in, nullptr, FIN_NONE, n);
    case FEO_VAR:
        if constexpr (std::is_floating_point<T>::value) {
            return run_plan<T, K_WELFORD>(ctx, p, out, in, nullptr, partial ? FIN_MEAN_M2 : FIN_VAR, n);
        } else {
            // tensor.rs:188,196-202 in integers: mean = sum / n (truncating), then sum (x - mean)^2 / n
            T *mean = nullptr;
            FEO_CUDA(ctx, cudaMallocAsync((void **)&mean, (size_t)p.nout * sizeof(T), ctx->stream));
            int rc = run_plan<T, K_SUM>(ctx, p, mean, in, nullptr, FIN_DIV, n);
            if (rc == FEO_OK) rc = run_plan<T, K_SQDEV>(ctx, p, out, in, mean, FIN_DIV, n);
            cudaFreeAsync(mean, ctx->stream);
            return rc;
        }
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown reduction %d", kind);
    }
}

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
    FEO_DISPATCH(ctx, dtype, T, return reduce_typed<T>(ctx, kind, (T *)out, (const T *)in, (const T *)aux, rank, shape, naxes, axes, mode));
    return FEO_OK;
}

int feo_impl_var_merge(feo_ctx *ctx, int dtype, void *out, const void *partials, size_t nparts, size_t nout,
                       size_t count_per_part, const void *divisor_host) {
    const unsigned blocks = (unsigned)feo_ceil_div(nout, (size_t)kThreads);
    if (dtype == FEO_F32) {
        var_merge_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float *)out, (const float *)partials, (long long)nparts,
                                                                       (long long)nout, (float)count_per_part, *(const float *)divisor_host);
    } else {
        var_merge_kernel<double><<<blocks, kThreads, 0, ctx->stream>>>((double *)out, (const double *)partials, (long long)nparts,
                                                                        (long long)nout, (double)count_per_part, *(const double *)divisor_host);
    }
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}
