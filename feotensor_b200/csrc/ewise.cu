// ewise.cu -- K1: elementwise Add/Sub/Mul/Div kernels (tensor.rs:336-489) for sm_100a.
//
//  * ewise_flat_kernel       same-shape and scalar-rhs operators: 128-bit loads/stores, 4 vectors in
//                            flight per thread, streaming cache hints (each byte is touched once).
//  * bcast_tile_kernel       broadcasting (stride-0) operands: every thread owns one 128/256-bit column
//                            slice and a PT x QT register tile over the two dimensions along which one
//                            operand is invariant, so an output vector costs (PT+QT)/(PT*QT) operand
//                            loads instead of 2 -- the L2 would otherwise carry 3x the DRAM traffic.
//                            Also serves the outer product (tensor.rs:311-322) as [na,nb] with
//                            strides a=(1,0), b=(0,1).
//  * bcast_generic_kernel    any strides / unaligned shapes: one element per thread, index decomposed
//                            by div/mod over the collapsed dims (metadata in the kernel parameter
//                            space, i.e. constant memory).
// HBM-bound: algorithmic bytes = sizeof(T) * (size(out) + un-broadcast size(a) + size(b)).
// The kernel templates live in ewise_impl.cuh and are instantiated per element type in ewise_<dtype>.cu;
// this file holds the dtype dispatch and the small constructor / generator kernels.
#include "feo_common.cuh"

#define FEO_DECL_TYPED(SFX)                                                                                                   \
    int feo_ewise_typed_##SFX(feo_ctx *, int, void *, const void *, const void *, const void *, size_t);                       \
    int feo_bcast_typed_##SFX(feo_ctx *, int, void *, const void *, const void *, int, const size_t *, const size_t *, const size_t *);
FEO_DECL_TYPED(f32) FEO_DECL_TYPED(f64) FEO_DECL_TYPED(i32) FEO_DECL_TYPED(i64) FEO_DECL_TYPED(u32)

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) fill_kernel(T *__restrict__ out, T value, size_t n) {
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) out[i] = value;
}
template <typename T>
__global__ void __launch_bounds__(kThreads) eye_kernel(T *__restrict__ out, size_t rows, size_t cols) {
    const size_t n = rows * cols;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
        out[i] = (i / cols == i % cols) ? (T)1 : (T)0;
}

// Counter-based generator: value = f(seed, global index).  Mirrors feo_oracle_fill_uniform bit for bit
// (separately rounded multiply and add, no FMA).
__host__ __device__ __forceinline__ unsigned long long feo_mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
template <typename T>
__global__ void __launch_bounds__(kThreads) fill_uniform_kernel(T *__restrict__ out, size_t n, unsigned long long seed,
                                                                unsigned long long first, T lo, T span, unsigned long long irange) {
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const unsigned long long h = feo_mix64(seed * 0xD1342543DE82EF95ULL + (first + i));
        if constexpr (std::is_same<T, float>::value) {
            const float u = __fmul_rn((float)(h >> 40), 1.0f / 16777216.0f);
            out[i] = __fadd_rn(lo, __fmul_rn(span, u));
        } else if constexpr (std::is_same<T, double>::value) {
            const double u = __dmul_rn((double)(h >> 11), 1.0 / 9007199254740992.0);
            out[i] = __dadd_rn(lo, __dmul_rn(span, u));
        } else {
            out[i] = w_add(lo, (T)(h % irange));
        }
    }
}

// Tensor::pow (tensor.rs:325-333).  mode: 0 generic powf, 1 x*x, 2 x, 3 sqrt, 4 one, 5 1/x.
template <typename T>
__global__ void __launch_bounds__(kThreads) pow_kernel(T *__restrict__ out, const T *__restrict__ a, T power, int mode, size_t n) {
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const T x = a[i];
        T r;
        switch (mode) {
        case 1: r = x * x; break;
        case 2: r = x; break;
        case 3: r = sqrt(x); break;
        case 4: r = (T)1; break;
        case 5: r = (T)1 / x; break;
        default: r = pow(x, power); break;
        }
        out[i] = r;
    }
}

}  // namespace

int feo_impl_pow(feo_ctx *ctx, int dtype, void *out, const void *a, const void *power_host, size_t n) {
    size_t blocks = feo_ceil_div(n, (size_t)kThreads);
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    auto mode_of = [](double p) { return p == 2.0 ? 1 : p == 1.0 ? 2 : p == 0.5 ? 3 : p == 0.0 ? 4 : p == -1.0 ? 5 : 0; };
    if (dtype == FEO_F32) {
        const float p = *(const float *)power_host;
        pow_kernel<float><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((float *)out, (const float *)a, p, mode_of(p), n);
    } else {
        const double p = *(const double *)power_host;
        pow_kernel<double><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((double *)out, (const double *)a, p, mode_of(p), n);
    }
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

int feo_impl_fill_uniform(feo_ctx *ctx, int dtype, void *out, size_t n, uint64_t seed, uint64_t first_index, double lo, double hi) {
    size_t blocks = feo_ceil_div(n, (size_t)kThreads);
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    FEO_DISPATCH(ctx, dtype, T, {
        T tlo = (T)lo, span = (T)0;
        unsigned long long irange = 1;
        if constexpr (std::is_floating_point<T>::value) span = (T)((T)hi - (T)lo);
        else irange = (unsigned long long)((long long)hi - (long long)lo);
        fill_uniform_kernel<T><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((T *)out, n, seed, first_index, tlo, span, irange);
        FEO_LAUNCH_CHECK(ctx);
    });
    return FEO_OK;
}

int feo_impl_ewise(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, const void *scalar_host, size_t n) {
    switch (dtype) {
    case FEO_F32: return feo_ewise_typed_f32(ctx, op, out, a, b, scalar_host, n);
    case FEO_F64: return feo_ewise_typed_f64(ctx, op, out, a, b, scalar_host, n);
    case FEO_I32: return feo_ewise_typed_i32(ctx, op, out, a, b, scalar_host, n);
    case FEO_I64: return feo_ewise_typed_i64(ctx, op, out, a, b, scalar_host, n);
    case FEO_U32: return feo_ewise_typed_u32(ctx, op, out, a, b, scalar_host, n);
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    }
}

int feo_impl_broadcast(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, int rank,
                       const size_t *shape, const size_t *sa, const size_t *sb) {
    switch (dtype) {
    case FEO_F32: return feo_bcast_typed_f32(ctx, op, out, a, b, rank, shape, sa, sb);
    case FEO_F64: return feo_bcast_typed_f64(ctx, op, out, a, b, rank, shape, sa, sb);
    case FEO_I32: return feo_bcast_typed_i32(ctx, op, out, a, b, rank, shape, sa, sb);
    case FEO_I64: return feo_bcast_typed_i64(ctx, op, out, a, b, rank, shape, sa, sb);
    case FEO_U32: return feo_bcast_typed_u32(ctx, op, out, a, b, rank, shape, sa, sb);
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    }
}

int feo_impl_fill(feo_ctx *ctx, int dtype, void *out, const void *value, size_t n) {
    if (n == 0) return FEO_OK;
    size_t blocks = feo_ceil_div(n, (size_t)kThreads);
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    FEO_DISPATCH(ctx, dtype, T, {
        fill_kernel<T><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((T *)out, *(const T *)value, n);
        FEO_LAUNCH_CHECK(ctx);
    });
    return FEO_OK;
}

int feo_impl_eye(feo_ctx *ctx, int dtype, void *out, size_t rows, size_t cols) {
    size_t blocks = feo_ceil_div(rows * cols, (size_t)kThreads);
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    FEO_DISPATCH(ctx, dtype, T, {
        eye_kernel<T><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((T *)out, rows, cols);
        FEO_LAUNCH_CHECK(ctx);
    });
    return FEO_OK;
}
