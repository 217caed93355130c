// reduce_plan.cuh -- host planner templates for the axis reductions (see reduce.cu); instantiated once per
// element type in reduce_<dtype>.cu so the translation units build in parallel.
#pragma once
#include <string.h>

#include <utility>

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

// More alternating kept / reduced groups than the canonical kernels take (e.g. [16]^7 over axes 1, 3, 5): reduce the
// innermost canonical suffix first, with every group in front of it treated as kept; the result has two or three groups
// fewer and is 1 / (the reduced extents) of the size, so the later passes cost next to nothing.  Sum / max / min only:
// their partial results are plain values (mean divides in the last pass).  The one-kernel generic fallback walks each
// output with div / mod indexing and uncoalesced 4-byte loads (measured 0.55-0.9 TB/s).
template <typename T, int KIND>
int run_multipass(feo_ctx *ctx, const Plan &p, T *out, const T *in, int fin, T divisor) {
    long long ext[FEO_MAX_RANK];
    bool red[FEO_MAX_RANK];
    int m = p.m;
    for (int i = 0; i < m; ++i) { ext[i] = p.ext[i]; red[i] = p.red[i]; }
    const T *cur = in;
    T *owned = nullptr;  // intermediate produced by the previous pass
    int rc = FEO_OK;
    for (;;) {
        const Canon c = canonicalise(m, ext, red);
        if (c.ok) {
            rc = run_canonical<T, KIND, LD_PLAIN>(ctx, c, out, cur, nullptr, nullptr, fin, divisor);
            break;
        }
        // Two ways to peel a pass off.  "Outermost first": reduce the first reduced group with everything inside it treated as
        // one long kept row -- always the streaming cols kernel at full speed, shrinks the data by that group's extent.
        // "Suffix": reduce the innermost canonical [k r k (r)] with everything in front treated as kept -- shrinks by r1 * r2
        // but works on inner slabs that may be tiny ([16]^7: 16 x 16 floats).  Take the first unless it shrinks too little.
        const int take = red[m - 1] ? 4 : 3;
        const long long shrink_suffix = ext[m - take + 1] * (take == 4 ? ext[m - 1] : 1);
        const int lead = red[0] ? 0 : 1;  // index of the outermost reduced group
        const bool outer_first = ext[lead] >= 8 || ext[lead] >= shrink_suffix;
        Canon s{true, 1, 1, 1, 1};
        if (outer_first) {
            for (int i = 0; i < lead; ++i) s.K1 *= ext[i];
            s.R1 = ext[lead];
            for (int i = lead + 1; i < m; ++i) s.K2 *= ext[i];
        } else {
            for (int i = 0; i <= m - take; ++i) s.K1 *= ext[i];
            s.R1 = ext[m - take + 1];
            s.K2 = ext[m - take + 2];
            if (take == 4) s.R2 = ext[m - 1];
        }
        T *tmp = nullptr;
        if (feo_alloc_async(ctx, (void **)&tmp, (size_t)(s.K1 * s.K2) * sizeof(T)) != cudaSuccess) {
            rc = feo_fail(ctx, FEO_ERR_CUDA, "out of memory for a %lld-element intermediate", s.K1 * s.K2);
            break;
        }
        rc = run_canonical<T, KIND, LD_PLAIN>(ctx, s, tmp, cur, nullptr, nullptr, FIN_NONE, (T)1);
        if (owned) cudaFreeAsync(owned, ctx->stream);
        owned = tmp;
        cur = tmp;
        if (rc != FEO_OK) break;
        if (outer_first) {
            // group `lead` is gone; its kept neighbours (lead - 1, if any, and lead + 1) become one group
            if (lead == 0) {
                for (int i = 1; i < m; ++i) { ext[i - 1] = ext[i]; red[i - 1] = red[i]; }
                m -= 1;
            } else {
                ext[lead - 1] *= ext[lead + 1];
                for (int i = lead + 2; i < m; ++i) { ext[i - 2] = ext[i]; red[i - 2] = red[i]; }
                m -= 2;
            }
        } else {
            ext[m - take] *= s.K2;  // the suffix's leading kept group absorbs k2
            m = m - take + 1;
        }
    }
    if (owned) cudaFreeAsync(owned, ctx->stream);
    return rc;
}

// Float variance over more groups than the canonical kernels take: the innermost canonical suffix is reduced to Welford
// partials (mean | M2, equal counts) by the fast kernels -- one pass over the input -- and the partials, 1 / (r1 * r2) of
// the size, are merged over the remaining reduced groups by reduce_generic_pairs_kernel.
template <typename T>
int run_var_twostage(feo_ctx *ctx, const Plan &p, T *out, const T *in, int fin, T divisor) {
    const int m = p.m;
    const int take = p.red[m - 1] ? 4 : 3;
    Canon s{true, 1, 1, 1, 1};
    for (int i = 0; i <= m - take; ++i) s.K1 *= p.ext[i];
    s.R1 = p.ext[m - take + 1];
    s.K2 = p.ext[m - take + 2];
    if (take == 4) s.R2 = p.ext[m - 1];
    const long long n1 = s.K1 * s.K2;
    if (s.R1 * s.R2 < 32) return -1;  // the first pass would shrink the data too little to pay for a second one
    T *tmp = nullptr;
    FEO_CUDA(ctx, feo_alloc_async(ctx, (void **)&tmp, (size_t)n1 * 2 * sizeof(T)));
    int rc = run_canonical<T, K_WELFORD, LD_PLAIN>(ctx, s, tmp, in, nullptr, nullptr, FIN_MEAN_M2, (T)1);
    if (rc == FEO_OK) {
        // the intermediate is dense over groups 0 .. m - take (the last of them kept, extended by k2)
        GenericGeom G;
        memset(&G, 0, sizeof G);
        G.nout = 1; G.nred = 1; G.fin = fin;
        long long stride = 1;
        for (int i = m - take; i >= 0; --i) {
            const long long e = (i == m - take) ? p.ext[i] * s.K2 : p.ext[i];
            if (p.red[i]) { G.rext[G.nr] = e; G.rstr[G.nr] = stride; G.nr++; G.nred *= e; }
            else { G.kext[G.nk] = e; G.kstr[G.nk] = stride; G.nk++; G.nout *= e; }
            stride *= e;
        }
        // make_plan lists dims outermost first; this loop ran innermost first
        for (int i = 0; i < G.nk / 2; ++i) { std::swap(G.kext[i], G.kext[G.nk - 1 - i]); std::swap(G.kstr[i], G.kstr[G.nk - 1 - i]); }
        for (int i = 0; i < G.nr / 2; ++i) { std::swap(G.rext[i], G.rext[G.nr - 1 - i]); std::swap(G.rstr[i], G.rstr[G.nr - 1 - i]); }
        long long blocks = feo_ceil_div((size_t)G.nout, (size_t)(kThreads / 32));
        const long long cap = (long long)ctx->sm_count * 32;
        if (blocks > cap) blocks = cap;
        reduce_generic_pairs_kernel<T><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, tmp, tmp + n1, G, (T)(s.R1 * s.R2), divisor, ctx->arith_flag);
        ctx->launches++;
        ctx->window.clear();
        if (cudaGetLastError() != cudaSuccess) rc = feo_fail(ctx, FEO_ERR_CUDA, "reduce_generic_pairs_kernel launch failed");
    }
    cudaFreeAsync(tmp, ctx->stream);
    return rc;
}

// Float variance, any number of groups, one reduced group per pass from the outside in: the first pass is the streaming
// cols kernel with Welford accumulators (everything inside the outermost reduced group is one long kept row) and leaves
// equal-count (mean | M2) partials; each further reduced group is merged away by welford_pairs_cols_kernel on data that is
// already 1 / (first extent) of the input.  Used when the canonical suffix would shrink the data too little for
// run_var_twostage ([16]^7: 0.55 -> multi-TB/s).
template <typename T>
int run_var_multipass(feo_ctx *ctx, const Plan &p, T *out, const T *in, int fin, T divisor) {
    long long ext[FEO_MAX_RANK];
    bool red[FEO_MAX_RANK];
    int m = p.m;
    for (int i = 0; i < m; ++i) { ext[i] = p.ext[i]; red[i] = p.red[i]; }
    auto reduced_left = [&]() { int n = 0; for (int i = 0; i < m; ++i) n += red[i] ? 1 : 0; return n; };
    auto drop_lead = [&](int lead) {  // group `lead` is gone; its kept neighbours become one group
        if (lead == 0) {
            for (int i = 1; i < m; ++i) { ext[i - 1] = ext[i]; red[i - 1] = red[i]; }
            m -= 1;
        } else if (lead == m - 1) {
            m -= 1;
        } else {
            ext[lead - 1] *= ext[lead + 1];
            for (int i = lead + 2; i < m; ++i) { ext[i - 2] = ext[i]; red[i - 2] = red[i]; }
            m -= 2;
        }
    };
    T *cur = nullptr;       // current partials: means | M2, `ncur` each
    long long ncur = 0;
    long long count = 1;
    int rc = FEO_OK;
    bool first = true;
    while (rc == FEO_OK) {
        const int lead = red[0] ? 0 : 1;
        long long K1 = 1, K2 = 1;
        for (int i = 0; i < lead; ++i) K1 *= ext[i];
        const long long R1 = ext[lead];
        for (int i = lead + 1; i < m; ++i) K2 *= ext[i];
        const bool last = reduced_left() == 1;
        const long long n1 = K1 * K2;
        T *dst = out;
        if (!last && feo_alloc_async(ctx, (void **)&dst, (size_t)n1 * 2 * sizeof(T)) != cudaSuccess) {
            rc = feo_fail(ctx, FEO_ERR_CUDA, "out of memory for a %lld-element intermediate", 2 * n1);
            break;
        }
        const int f = last ? fin : FIN_MEAN_M2;
        if (first) {
            const Canon s{true, K1, R1, K2, 1};
            rc = run_canonical<T, K_WELFORD, LD_PLAIN>(ctx, s, dst, in, nullptr, nullptr, f, divisor);
        } else {
            const long long blocks = feo_ceil_div((size_t)n1, (size_t)kThreads);
            welford_pairs_cols_kernel<T><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(dst, cur, cur + ncur, K1, R1, K2, (T)count, divisor, f);
            ctx->launches++;
            ctx->window.clear();
            if (cudaGetLastError() != cudaSuccess) rc = feo_fail(ctx, FEO_ERR_CUDA, "welford_pairs_cols_kernel launch failed");
        }
        if (cur) cudaFreeAsync(cur, ctx->stream);
        cur = last ? nullptr : dst;
        ncur = n1;
        count *= R1;
        first = false;
        if (last) break;
        drop_lead(lead);
    }
    if (cur) cudaFreeAsync(cur, ctx->stream);
    return rc;
}

template <typename T, int KIND>
int run_plan(feo_ctx *ctx, const Plan &p, T *out, const T *in, const T *aux, int fin, T divisor) {
    const Canon c = canonicalise(p.m, p.ext, p.red);
    if (c.ok) return run_canonical<T, KIND, LD_PLAIN>(ctx, c, out, in, nullptr, aux, fin, divisor);
    if constexpr (KIND == K_SUM || KIND == K_MAX || KIND == K_MIN) {
        if (ctx->knobs[FEO_KNOB_REDUCE_SPLIT] != -1) return run_multipass<T, KIND>(ctx, p, out, in, fin, divisor);
    }
    if constexpr (KIND == K_WELFORD) {
        if (ctx->knobs[FEO_KNOB_REDUCE_SPLIT] != -1) {
            // outermost reduced group wide enough: its streaming pass shrinks the data 8x or more and the rest is cheap
            if (p.ext[p.red[0] ? 0 : 1] >= 8) return run_var_multipass<T>(ctx, p, out, in, fin, divisor);
            const int rc = run_var_twostage<T>(ctx, p, out, in, fin, divisor);
            if (rc != -1) return rc;
            return run_var_multipass<T>(ctx, p, out, in, fin, divisor);
        }
    }
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
            if (partial) return run_plan<T, K_SUMSQ>(ctx, p, out, in, nullptr, FIN_MEAN_M2, n);  // out = S1 x nout | S2 x nout
            if (ctx->knobs[FEO_KNOB_REDUCE_FUSE] != 3) {
                // one pass for every other geometry: wrapping sum and sum of squares (see Acc<T, K_SUMSQ>)
                return run_plan<T, K_SUMSQ>(ctx, p, out, in, nullptr, FIN_DIV, n);
            }
            T *mean = nullptr;  // the literal two passes (knob REDUCE_FUSE = 3; kept as the cross-check of the identity)
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
