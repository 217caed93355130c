// reduce_plan.cuh -- host planner templates for the axis reductions (see reduce.cu); instantiated once per
// element type in reduce_<dtype>.cu so the translation units build in parallel.
#pragma once
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
            if (!partial) {  // short reduced innermost axis: both passes in registers, one kernel
                const Canon c = canonicalise(p.m, p.ext, p.red);
                if (c.ok) {
                    const int rc1 = run_canonical<T, K_INTVAR, LD_PLAIN>(ctx, c, out, in, nullptr, nullptr, FIN_DIV, n);
                    if (rc1 != -1) return rc1;
                }
            }
            T *mean = nullptr;
            FEO_CUDA(ctx, feo_alloc_async(ctx, (void **)&mean, (size_t)p.nout * sizeof(T)));
            int rc = run_plan<T, K_SUM>(ctx, p, mean, in, nullptr, FIN_DIV, n);
            if (rc == FEO_OK) rc = run_plan<T, K_SQDEV>(ctx, p, out, in, mean, FIN_DIV, n);
            cudaFreeAsync(mean, ctx->stream);
            return rc;
        }
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown reduction %d", kind);
    }
}

}  // namespace

#define FEO_REDUCE_INSTANTIATE(T, SFX)                                                                                      \
    int feo_reduce_typed_##SFX(feo_ctx *ctx, int kind, void *out, const void *in, const void *aux, int rank,                 \
                               const size_t *shape, int naxes, const size_t *axes, int mode) {                               \
        return reduce_typed<T>(ctx, kind, (T *)out, (const T *)in, (const T *)aux, rank, shape, naxes, axes, mode);          \
    }
