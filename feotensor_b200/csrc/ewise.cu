// ewise.cu -- K1: elementwise Add/Sub/Mul/Div kernels (tensor.rs:336-489) for sm_100a.
//
//  * ewise_flat_kernel       same-shape and scalar-rhs operators: 128-bit loads/stores, 4 vectors in
//                            flight per thread, streaming cache hints (each byte is touched once).
//  * bcast_tile_kernel       broadcasting (stride-0) operands: every thread owns one 128-bit column
//                            slice and a PT x QT register tile over the two dimensions along which one
//                            operand is invariant, so an output vector costs (PT+QT)/(PT*QT) operand
//                            loads instead of 2 -- the L2 would otherwise carry 3x the DRAM traffic.
//                            Also serves the outer product (tensor.rs:311-322) as [na,nb] with
//                            strides a=(1,0), b=(0,1).
//  * bcast_generic_kernel    any strides / unaligned shapes: one element per thread, index decomposed
//                            by div/mod over the collapsed dims (metadata in the kernel parameter
//                            space, i.e. constant memory).
// HBM-bound: algorithmic bytes = sizeof(T) * (size(out) + un-broadcast size(a) + size(b)).
#include "feo_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kFlatUnroll = 4;

template <typename T, int OP, bool SCALAR>
__global__ void __launch_bounds__(kThreads) ewise_flat_kernel(T *__restrict__ out, const T *__restrict__ a,
                                                              const T *__restrict__ b, T s, size_t nvec, int *flag) {
    constexpr int VN = VecN<T>::value;
    using V = AVec<T, VN>;
    const size_t base = (size_t)blockIdx.x * (kThreads * kFlatUnroll) + threadIdx.x;
    V va[kFlatUnroll], vb[kFlatUnroll];
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) {
        const size_t i = base + (size_t)u * kThreads;
        if (i < nvec) {
            va[u] = ld_stream(reinterpret_cast<const V *>(a) + i);
            if constexpr (!SCALAR) vb[u] = ld_stream(reinterpret_cast<const V *>(b) + i);
        }
    }
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) {
        const size_t i = base + (size_t)u * kThreads;
        if (i < nvec) {
            V r;
#pragma unroll
            for (int k = 0; k < VN; ++k) r.v[k] = apply_op<OP>(va[u].v[k], SCALAR ? s : vb[u].v[k], flag);
            st_stream(reinterpret_cast<V *>(out) + i, r);
        }
    }
}

// scalar path: unaligned pointers or the < VN element tail
template <typename T, int OP, bool SCALAR>
__global__ void __launch_bounds__(kThreads) ewise_scalar_path_kernel(T *__restrict__ out, const T *__restrict__ a,
                                                                     const T *__restrict__ b, T s, size_t start, size_t n,
                                                                     int *flag) {
    for (size_t i = start + (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
        out[i] = apply_op<OP>(a[i], SCALAR ? s : b[i], flag);
}

template <typename T, int OP, bool SCALAR>
int launch_flat(feo_ctx *ctx, T *out, const T *a, const T *b, T s, size_t n) {
    constexpr int VN = VecN<T>::value;
    const bool aligned = feo_aligned16(out) && feo_aligned16(a) && (SCALAR || feo_aligned16(b));
    size_t nvec = aligned ? n / VN : 0;
    if (nvec) {
        const size_t blocks = feo_ceil_div(nvec, (size_t)kThreads * kFlatUnroll);
        ewise_flat_kernel<T, OP, SCALAR><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, a, b, s, nvec, ctx->arith_flag);
        FEO_LAUNCH_CHECK(ctx);
    }
    const size_t done = nvec * VN;
    if (done < n) {
        const size_t rem = n - done;
        size_t blocks = feo_ceil_div(rem, (size_t)kThreads);
        if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
        ewise_scalar_path_kernel<T, OP, SCALAR><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, a, b, s, done, n, ctx->arith_flag);
        FEO_LAUNCH_CHECK(ctx);
    }
    return FEO_OK;
}

template <typename T, bool SCALAR>
int launch_flat_op(feo_ctx *ctx, int op, T *out, const T *a, const T *b, T s, size_t n) {
    switch (op) {
    case FEO_ADD: return launch_flat<T, FEO_ADD, SCALAR>(ctx, out, a, b, s, n);
    case FEO_SUB: return launch_flat<T, FEO_SUB, SCALAR>(ctx, out, a, b, s, n);
    case FEO_MUL: return launch_flat<T, FEO_MUL, SCALAR>(ctx, out, a, b, s, n);
    default: return launch_flat<T, FEO_DIV, SCALAR>(ctx, out, a, b, s, n);
    }
}

// ---- broadcasting ----------------------------------------------------------------------------------
constexpr int kMaxRest = FEO_MAX_RANK;

struct BcastParams {
    // inner (contiguous in out) dimension, in 128-bit vectors; operand inner stride is 1 or 0 (splat)
    long long ni_vec;
    int a_inner, b_inner;
    // P: dimension walked by the PT loop; Q: dimension walked by the QT loop (extent 1 if absent)
    long long np, nq;
    long long sa_p, sb_p, so_p, sa_q, sb_q, so_q;  // element strides
    // remaining dims, decomposed per block
    int nrest;
    long long rest_ext[kMaxRest], rest_sa[kMaxRest], rest_sb[kMaxRest], rest_so[kMaxRest];
    // block geometry
    int tx, ty;  // threads along inner vectors / along q tiles (tx * ty == kThreads)
    long long i_blocks, q_blocks, p_tiles;
};

template <typename T> __device__ __forceinline__ AVec<T, VecN<T>::value> load_operand(const T *p, int inner) {
    constexpr int VN = VecN<T>::value;
    using V = AVec<T, VN>;
    if (inner) return ld_keep(reinterpret_cast<const V *>(p));
    V r;
    const T s = __ldg(p);
#pragma unroll
    for (int k = 0; k < VN; ++k) r.v[k] = s;
    return r;
}

// A_QINV: a does not move along q (stride 0) -> PT loads of a instead of PT*QT.  B_PINV likewise for b.
template <typename T, int OP, int PT, int QT, bool A_QINV, bool B_PINV>
__global__ void __launch_bounds__(kThreads) bcast_tile_kernel(T *__restrict__ out, const T *__restrict__ a,
                                                              const T *__restrict__ b, const __grid_constant__ BcastParams P,
                                                              int *flag) {
    constexpr int VN = VecN<T>::value;
    using V = AVec<T, VN>;
    long long blk = blockIdx.x;
    const long long bi = blk % P.i_blocks; blk /= P.i_blocks;
    const long long bq = blk % P.q_blocks; blk /= P.q_blocks;
    const long long bp = blk % P.p_tiles;  blk /= P.p_tiles;
    long long oa = 0, ob = 0, oo = 0;
#pragma unroll 1
    for (int d = P.nrest - 1; d >= 0; --d) {
        const long long c = blk % P.rest_ext[d];
        blk /= P.rest_ext[d];
        oa += c * P.rest_sa[d]; ob += c * P.rest_sb[d]; oo += c * P.rest_so[d];
    }
    const int tx = threadIdx.x % P.tx, ty = threadIdx.x / P.tx;
    const long long iv = bi * P.tx + tx;
    if (iv >= P.ni_vec) return;
    const long long p0 = bp * PT;
    const long long q0 = (bq * P.ty + ty) * QT;
    if (q0 >= P.nq) return;
    oa += iv * VN * P.a_inner; ob += iv * VN * P.b_inner; oo += iv * VN;

    V va[A_QINV ? PT : PT * QT], vb[B_PINV ? QT : PT * QT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
        if (p0 + p < P.np) {
#pragma unroll
            for (int q = 0; q < (A_QINV ? 1 : QT); ++q)
                if (A_QINV || q0 + q < P.nq)
                    va[A_QINV ? p : p * QT + q] = load_operand<T>(a + oa + (p0 + p) * P.sa_p + (q0 + q) * P.sa_q, P.a_inner);
        }
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        if (q0 + q < P.nq) {
#pragma unroll
            for (int p = 0; p < (B_PINV ? 1 : PT); ++p)
                if (B_PINV || p0 + p < P.np)
                    vb[B_PINV ? q : p * QT + q] = load_operand<T>(b + ob + (p0 + p) * P.sb_p + (q0 + q) * P.sb_q, P.b_inner);
        }
    }
#pragma unroll
    for (int p = 0; p < PT; ++p) {
        if (p0 + p < P.np) {
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                if (q0 + q < P.nq) {
                    const V &x = va[A_QINV ? p : p * QT + q];
                    const V &y = vb[B_PINV ? q : p * QT + q];
                    V r;
#pragma unroll
                    for (int k = 0; k < VN; ++k) r.v[k] = apply_op<OP>(x.v[k], y.v[k], flag);
                    st_stream(reinterpret_cast<V *>(out + oo + (p0 + p) * P.so_p + (q0 + q) * P.so_q), r);
                }
            }
        }
    }
}

struct GenericParams {
    int rank;
    long long ext[FEO_MAX_RANK], sa[FEO_MAX_RANK], sb[FEO_MAX_RANK];
    long long n;
};

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads) bcast_generic_kernel(T *__restrict__ out, const T *__restrict__ a,
                                                                 const T *__restrict__ b, const __grid_constant__ GenericParams P,
                                                                 int *flag) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < P.n; i += (long long)gridDim.x * kThreads) {
        long long rem = i, oa = 0, ob = 0;
#pragma unroll 1
        for (int d = P.rank - 1; d >= 0; --d) {
            const long long c = rem % P.ext[d];
            rem /= P.ext[d];
            oa += c * P.sa[d]; ob += c * P.sb[d];
        }
        out[i] = apply_op<OP>(a[oa], b[ob], flag);
    }
}

struct Collapsed {
    int rank;
    long long ext[FEO_MAX_RANK], sa[FEO_MAX_RANK], sb[FEO_MAX_RANK];
};

// Drop extent-1 dims, then merge neighbours that are contiguous for out, a and b alike.
Collapsed collapse(int rank, const size_t *shape, const size_t *sa, const size_t *sb) {
    Collapsed c;
    c.rank = 0;
    for (int d = 0; d < rank; ++d) {
        if (shape[d] == 1) continue;
        const long long e = (long long)shape[d], xa = (long long)sa[d], xb = (long long)sb[d];
        if (c.rank > 0) {
            const int l = c.rank - 1;
            if (c.sa[l] == xa * e && c.sb[l] == xb * e) {  // previous dim steps exactly over this one
                c.ext[l] *= e; c.sa[l] = xa; c.sb[l] = xb;
                continue;
            }
        }
        c.ext[c.rank] = e; c.sa[c.rank] = xa; c.sb[c.rank] = xb; c.rank++;
    }
    if (c.rank == 0) { c.rank = 1; c.ext[0] = 1; c.sa[0] = 0; c.sb[0] = 0; }
    return c;
}

template <typename T, int OP, int PT, int QT, bool A_QINV, bool B_PINV>
int launch_tile(feo_ctx *ctx, T *out, const T *a, const T *b, const BcastParams &P) {
    long long blocks = P.i_blocks * P.q_blocks * P.p_tiles;
    for (int d = 0; d < P.nrest; ++d) blocks *= P.rest_ext[d];
    if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "broadcast grid too large");
    bcast_tile_kernel<T, OP, PT, QT, A_QINV, B_PINV><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, a, b, P, ctx->arith_flag);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

template <typename T, int OP>
int launch_generic(feo_ctx *ctx, T *out, const T *a, const T *b, const Collapsed &c) {
    GenericParams G;
    G.rank = c.rank;
    G.n = 1;
    for (int d = 0; d < c.rank; ++d) { G.ext[d] = c.ext[d]; G.sa[d] = c.sa[d]; G.sb[d] = c.sb[d]; G.n *= c.ext[d]; }
    long long blocks = (G.n + kThreads - 1) / kThreads;
    const long long cap = (long long)ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    bcast_generic_kernel<T, OP><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, a, b, G, ctx->arith_flag);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

template <typename T, int OP>
int broadcast_typed(feo_ctx *ctx, T *out, const T *a, const T *b, int rank, const size_t *shape, const size_t *sa,
                    const size_t *sb) {
    constexpr int VN = VecN<T>::value;
    const Collapsed c = collapse(rank, shape, sa, sb);
    const int knob = ctx->knobs[FEO_KNOB_BCAST_TILE];
    const int in = c.rank - 1;

    // same-shape after collapsing: the flat kernel is the right tool
    if (c.rank == 1 && c.sa[0] == 1 && c.sb[0] == 1 && knob != 4)
        return launch_flat<T, OP, false>(ctx, out, a, b, (T)0, (size_t)c.ext[0]);

    bool vec_ok = knob != 4 && (c.sa[in] == 0 || c.sa[in] == 1) && (c.sb[in] == 0 || c.sb[in] == 1) &&
                  (c.ext[in] % VN == 0) && feo_aligned16(out) && (c.sa[in] == 0 || feo_aligned16(a)) &&
                  (c.sb[in] == 0 || feo_aligned16(b));
    for (int d = 0; d < in && vec_ok; ++d) {
        if (c.sa[in] == 1 && c.sa[d] % VN != 0) vec_ok = false;
        if (c.sb[in] == 1 && c.sb[d] % VN != 0) vec_ok = false;
    }
    if (!vec_ok) return launch_generic<T, OP>(ctx, out, a, b, c);

    // strides of the dense output
    long long so[FEO_MAX_RANK];
    {
        long long s = 1;
        for (int d = c.rank - 1; d >= 0; --d) { so[d] = s; s *= c.ext[d]; }
    }
    // P = largest outer dim where b is invariant and a moves; Q = largest where a is invariant and b moves
    int pd = -1, qd = -1;
    for (int d = 0; d < in; ++d) {
        if (c.sb[d] == 0 && c.sa[d] != 0 && (pd < 0 || c.ext[d] > c.ext[pd])) pd = d;
        if (c.sa[d] == 0 && c.sb[d] != 0 && (qd < 0 || c.ext[d] > c.ext[qd])) qd = d;
    }
    BcastParams P;
    memset(&P, 0, sizeof P);
    P.ni_vec = c.ext[in] / VN;
    P.a_inner = (int)c.sa[in];
    P.b_inner = (int)c.sb[in];
    P.np = pd >= 0 ? c.ext[pd] : 1; P.sa_p = pd >= 0 ? c.sa[pd] : 0; P.sb_p = pd >= 0 ? c.sb[pd] : 0; P.so_p = pd >= 0 ? so[pd] : 0;
    P.nq = qd >= 0 ? c.ext[qd] : 1; P.sa_q = qd >= 0 ? c.sa[qd] : 0; P.sb_q = qd >= 0 ? c.sb[qd] : 0; P.so_q = qd >= 0 ? so[qd] : 0;
    for (int d = 0; d < in; ++d) {
        if (d == pd || d == qd) continue;
        P.rest_ext[P.nrest] = c.ext[d]; P.rest_sa[P.nrest] = c.sa[d]; P.rest_sb[P.nrest] = c.sb[d]; P.rest_so[P.nrest] = so[d];
        P.nrest++;
    }
    int tx = kThreads;
    while (tx > 1 && tx / 2 >= P.ni_vec) tx /= 2;  // smallest power of two >= ni_vec, capped at kThreads
    P.tx = tx;
    P.ty = kThreads / tx;
    P.i_blocks = (P.ni_vec + tx - 1) / tx;

    auto finish = [&](int PT, int QT) {
        P.p_tiles = (P.np + PT - 1) / PT;
        const long long qt = (P.nq + QT - 1) / QT;
        P.q_blocks = (qt + P.ty - 1) / P.ty;
    };
    const bool both = pd >= 0 && qd >= 0;
    if (both && knob != 3) {
        if (knob == 2 || P.np < 8) { finish(4, 4); return launch_tile<T, OP, 4, 4, true, true>(ctx, out, a, b, P); }
        finish(8, 4);
        return launch_tile<T, OP, 8, 4, true, true>(ctx, out, a, b, P);
    }
    if (pd >= 0 && knob != 3) {  // only b is invariant somewhere: hold b, walk p
        P.nq = 1; P.sa_q = P.sb_q = P.so_q = 0;
        if (qd >= 0) { /* unreachable: both handled above */ }
        finish(8, 1);
        return launch_tile<T, OP, 8, 1, false, true>(ctx, out, a, b, P);
    }
    if (qd >= 0 && knob != 3) {  // only a is invariant somewhere: hold a, walk q
        finish(1, 8);
        return launch_tile<T, OP, 1, 8, true, false>(ctx, out, a, b, P);
    }
    // no invariance to exploit (or tiling disabled): 1x1 vector tiles; fold P/Q back into the rest dims
    if (pd >= 0) { P.rest_ext[P.nrest] = P.np; P.rest_sa[P.nrest] = P.sa_p; P.rest_sb[P.nrest] = P.sb_p; P.rest_so[P.nrest] = P.so_p; P.nrest++; }
    if (qd >= 0) { P.rest_ext[P.nrest] = P.nq; P.rest_sa[P.nrest] = P.sa_q; P.rest_sb[P.nrest] = P.sb_q; P.rest_so[P.nrest] = P.so_q; P.nrest++; }
    P.np = P.nq = 1; P.sa_p = P.sb_p = P.so_p = P.sa_q = P.sb_q = P.so_q = 0;
    P.ty = 1;  // q extent is 1: spare rows of threads would idle, keep the block one row wide
    P.tx = kThreads;
    P.i_blocks = (P.ni_vec + P.tx - 1) / P.tx;
    finish(1, 1);
    return launch_tile<T, OP, 1, 1, true, true>(ctx, out, a, b, P);
}

template <typename T>
int broadcast_op(feo_ctx *ctx, int op, T *out, const T *a, const T *b, int rank, const size_t *shape, const size_t *sa,
                 const size_t *sb) {
    switch (op) {
    case FEO_ADD: return broadcast_typed<T, FEO_ADD>(ctx, out, a, b, rank, shape, sa, sb);
    case FEO_SUB: return broadcast_typed<T, FEO_SUB>(ctx, out, a, b, rank, shape, sa, sb);
    case FEO_MUL: return broadcast_typed<T, FEO_MUL>(ctx, out, a, b, rank, shape, sa, sb);
    default: return broadcast_typed<T, FEO_DIV>(ctx, out, a, b, rank, shape, sa, sb);
    }
}

// ---- constructors ------------------------------------------------------------------------------------
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

}  // namespace

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
    FEO_DISPATCH(ctx, dtype, T, {
        if (scalar_host) return launch_flat_op<T, true>(ctx, op, (T *)out, (const T *)a, nullptr, *(const T *)scalar_host, n);
        return launch_flat_op<T, false>(ctx, op, (T *)out, (const T *)a, (const T *)b, (T)0, n);
    });
    return FEO_OK;
}

int feo_impl_broadcast(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, int rank,
                       const size_t *shape, const size_t *sa, const size_t *sb) {
    FEO_DISPATCH(ctx, dtype, T, return broadcast_op<T>(ctx, op, (T *)out, (const T *)a, (const T *)b, rank, shape, sa, sb));
    return FEO_OK;
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
