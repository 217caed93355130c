// ewise_impl.cuh -- K1 kernel templates (see ewise.cu for the overview); instantiated once per element type
// in ewise_<dtype>.cu so the translation units build in parallel.
#pragma once
#include <string.h>

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
    int keep_operands;  // 1: load operands with the L2 evict_last policy
    // > 0: the CTAs of q-block bq also prefetch into L2 the rows of the P-invariant operand (b) that q-block bq + prefetch_q
    // will load -- a large b is cold for every launch (the write stream of the previous one evicted it), and a CTA that has
    // to wait for DRAM before its first store is ~10 % longer (measured: tools/slab_boundary.py)
    int prefetch_q;
};

template <typename T, int VB> __device__ __forceinline__ AVec<T, VecW<T, VB>::value> load_operand(const T *p, int inner, unsigned long long pol = 0) {
    constexpr int VN = VecW<T, VB>::value;
    using V = AVec<T, VN>;
    if (inner) {
        if constexpr (VB == 16) { if (pol) return ld_keep_last(reinterpret_cast<const V *>(p), pol); }
        return ld_keep(reinterpret_cast<const V *>(p));
    }
    V r;
    const T s = __ldg(p);
#pragma unroll
    for (int k = 0; k < VN; ++k) r.v[k] = s;
    return r;
}

// VB: bytes per thread access (16 = LDG/STG.128, 32 = the sm_100 256-bit forms).
// A_QINV: a does not move along q (stride 0) -> PT loads of a instead of PT*QT.  B_PINV likewise for b.
// Block order: inner blocks fastest, then the P tiles, then Q -- so the CTAs that share the same rows of the
// P-invariant operand (b) are co-resident and its second read is an L2 hit, not a DRAM read.
template <typename T, int OP, int VB, int PT, int QT, bool A_QINV, bool B_PINV>
__global__ void __launch_bounds__(kThreads) bcast_tile_kernel(T *__restrict__ out, const T *__restrict__ a,
                                                              const T *__restrict__ b, const __grid_constant__ BcastParams P,
                                                              int *flag) {
    constexpr int VN = VecW<T, VB>::value;
    using V = AVec<T, VN>;
    // programmatic dependent launch: let the next (independent) launch start filling SMs as this grid's tail drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    long long blk = blockIdx.x;
    const long long bi = blk % P.i_blocks; blk /= P.i_blocks;
    const long long bp = blk % P.p_tiles;  blk /= P.p_tiles;
    const long long bq = blk % P.q_blocks; blk /= P.q_blocks;
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
    // operands are re-read (by other CTAs now, by the next slab launch later) while the output streams through L2 once:
    // operands evict_last, output evict_first
    const unsigned long long pol = P.keep_operands ? l2_policy_evict_last() : 0ull;
    if constexpr (B_PINV) {
        // one lane per 128-byte line, first P tile only: fire-and-forget L2 prefetch of the b rows a later q-block needs
        if (P.prefetch_q > 0 && bp == 0 && P.b_inner && (tx & (128 / VB - 1)) == 0) {
            const long long qf = q0 + (long long)P.prefetch_q * P.ty * QT;
#pragma unroll
            for (int q = 0; q < QT; ++q)
                if (qf + q < P.nq) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(b + ob + (qf + q) * P.sb_q));
        }
    }

    V va[A_QINV ? PT : PT * QT], vb[B_PINV ? QT : PT * QT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
        if (p0 + p < P.np) {
#pragma unroll
            for (int q = 0; q < (A_QINV ? 1 : QT); ++q)
                if (A_QINV || q0 + q < P.nq)
                    va[A_QINV ? p : p * QT + q] = load_operand<T, VB>(a + oa + (p0 + p) * P.sa_p + (q0 + q) * P.sa_q, P.a_inner, pol);
        }
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        if (q0 + q < P.nq) {
#pragma unroll
            for (int p = 0; p < (B_PINV ? 1 : PT); ++p)
                if (B_PINV || p0 + p < P.np)
                    vb[B_PINV ? q : p * QT + q] = load_operand<T, VB>(b + ob + (p0 + p) * P.sb_p + (q0 + q) * P.sb_q, P.b_inner, pol);
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

// ---- hold + stream: the tiled kernel with a software pipeline ---------------------------------------------
// One operand ("held", h) is invariant along a loop dimension L; each thread keeps PT rows of it (rows along a tile
// dimension T) in registers for its 128/256-bit column slice, then walks a chunk of L: the other operand ("streamed",
// s) is prefetched QT rows ahead while the PT x QT results of the current rows are stored.  Loads per output vector
// fall to ~1/PT and their latency hides behind the stores.  SWAP: h is the right-hand operand (a op b == s op h).
struct StreamParams {
    long long ni_vec;
    int h_inner, s_inner;
    long long nt, nl;
    long long sh_t, ss_t, so_t, ss_l, so_l;
    int nrest;
    long long rest_ext[kMaxRest], rest_sh[kMaxRest], rest_ss[kMaxRest], rest_so[kMaxRest];
    int tx, ty;
    long long i_blocks, t_tiles, l_blocks, lchunk;
};

template <typename T, int OP, bool SWAP, int VB, int PT, int QT, int D, bool S_TINV>
__global__ void __launch_bounds__(kThreads) bcast_stream_kernel(T *__restrict__ out, const T *__restrict__ h,
                                                                const T *__restrict__ s, const __grid_constant__ StreamParams P,
                                                                int *flag) {
    constexpr int VN = VecW<T, VB>::value;
    constexpr int NS = S_TINV ? QT : PT * QT;
    using V = AVec<T, VN>;
    long long blk = blockIdx.x;
    const long long bi = blk % P.i_blocks; blk /= P.i_blocks;
    const long long bt = blk % P.t_tiles;  blk /= P.t_tiles;
    const long long bl = blk % P.l_blocks; blk /= P.l_blocks;
    long long oh = 0, os = 0, oo = 0;
#pragma unroll 1
    for (int d = P.nrest - 1; d >= 0; --d) {
        const long long c = blk % P.rest_ext[d];
        blk /= P.rest_ext[d];
        oh += c * P.rest_sh[d]; os += c * P.rest_ss[d]; oo += c * P.rest_so[d];
    }
    const int txi = threadIdx.x % P.tx, tyi = threadIdx.x / P.tx;
    const long long iv = bi * P.tx + txi;
    if (iv >= P.ni_vec) return;
    const long long t0 = bt * PT;
    const long long l_lo = (bl * P.ty + tyi) * P.lchunk;
    if (l_lo >= P.nl) return;
    const long long l_hi = (l_lo + P.lchunk < P.nl) ? l_lo + P.lchunk : P.nl;
    oh += iv * VN * P.h_inner; os += iv * VN * P.s_inner; oo += iv * VN;

    V vh[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p)
        if (t0 + p < P.nt) vh[p] = load_operand<T, VB>(h + oh + (t0 + p) * P.sh_t, P.h_inner);

    auto load_s = [&](V(&buf)[NS], long long l) {
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            if (l + q < l_hi) {
#pragma unroll
                for (int p = 0; p < (S_TINV ? 1 : PT); ++p)
                    if (S_TINV || t0 + p < P.nt)
                        buf[S_TINV ? q : p * QT + q] = load_operand<T, VB>(s + os + (l + q) * P.ss_l + (S_TINV ? 0 : (t0 + p) * P.ss_t), P.s_inner);
            }
        }
    };
    // ring of D prefetched steps: the load for step i + D is issued right after step i's stores
    V ring[D][NS];
#pragma unroll
    for (int d = 0; d < D; ++d) load_s(ring[d], l_lo + (long long)d * QT);
    for (long long l = l_lo; l < l_hi; l += (long long)D * QT) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const long long ll = l + (long long)d * QT;
            if (ll < l_hi) {
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    if (t0 + p < P.nt) {
#pragma unroll
                        for (int q = 0; q < QT; ++q) {
                            if (ll + q < l_hi) {
                                const V &x = vh[p];
                                const V &y = ring[d][S_TINV ? q : p * QT + q];
                                V r;
#pragma unroll
                                for (int k = 0; k < VN; ++k) r.v[k] = SWAP ? apply_op<OP>(y.v[k], x.v[k], flag) : apply_op<OP>(x.v[k], y.v[k], flag);
                                st_stream(reinterpret_cast<V *>(out + oo + (t0 + p) * P.so_t + (ll + q) * P.so_l), r);
                            }
                        }
                    }
                }
                load_s(ring[d], ll + (long long)D * QT);
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

template <typename T, int OP, int VB, int PT, int QT, bool A_QINV, bool B_PINV>
int launch_tile(feo_ctx *ctx, T *out, const T *a, const T *b, BcastParams &P) {
    P.p_tiles = (P.np + PT - 1) / PT;
    const long long qt = (P.nq + QT - 1) / QT;
    P.q_blocks = (qt + P.ty - 1) / P.ty;
    long long blocks = P.i_blocks * P.q_blocks * P.p_tiles;
    for (int d = 0; d < P.nrest; ++d) blocks *= P.rest_ext[d];
    if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "broadcast grid too large");
    // Back-to-back launches whose operands do not overlap (e.g. the 1 GiB slabs of one big broadcast) carry the
    // programmatic-stream-serialization attribute: their CTAs start as the previous grid's tail drains.
    const bool pdl = feo_pdl_admit(ctx);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (ctx->l2_persist_bytes && ctx->l2_win_ptr && ctx->l2_win_bytes) {
        // knob BCAST_L2KEEP = 2: the operand later launches read again (b of the config-2 slab loop) goes to the persisting
        // carve-out, everything else streams
        size_t bytes = ctx->l2_win_bytes;
        if (bytes > ctx->l2_window_max) bytes = ctx->l2_window_max;
        attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[na].val.accessPolicyWindow.base_ptr = const_cast<void *>(ctx->l2_win_ptr);
        attr[na].val.accessPolicyWindow.num_bytes = bytes;
        attr[na].val.accessPolicyWindow.hitRatio = bytes <= ctx->l2_persist_bytes ? 1.0f : (float)((double)ctx->l2_persist_bytes / (double)bytes);
        attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    int *flag = ctx->arith_flag;
    const T *ca = a, *cb = b;
    FEO_CUDA(ctx, cudaLaunchKernelEx(&cfg, bcast_tile_kernel<T, OP, VB, PT, QT, A_QINV, B_PINV>, out, ca, cb, P, flag));
    ctx->launches++;
    return FEO_OK;
}

template <typename T, int OP, bool SWAP, int VB, int PT, int QT, int D, bool S_TINV>
int launch_stream(feo_ctx *ctx, T *out, const T *h, const T *s, StreamParams &P) {
    P.t_tiles = (P.nt + PT - 1) / PT;
    long long other = P.i_blocks * P.t_tiles;
    for (int d = 0; d < P.nrest; ++d) other *= P.rest_ext[d];
    // chunk of the loop dimension per thread: enough CTAs for ~8 waves, but at least two pipeline steps per thread
    long long lchunk = ctx->knobs[FEO_KNOB_BCAST_LCHUNK];
    if (lchunk <= 0) {
        const long long target = (long long)ctx->sm_count * 24;
        const long long want_chunks = (target + other - 1) / other;
        lchunk = (P.nl + want_chunks - 1) / want_chunks;
        if (lchunk > 32) lchunk = 32;  // measured: ~32 rows per thread is the sweet spot (outer product 7.5 TB/s)
        if (lchunk < 2 * D * QT) lchunk = 2 * D * QT;
    }
    lchunk = (lchunk + QT - 1) / QT * QT;
    P.lchunk = lchunk;
    const long long chunks = (P.nl + lchunk - 1) / lchunk;
    P.l_blocks = (chunks + P.ty - 1) / P.ty;
    const long long blocks = other * P.l_blocks;
    if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "broadcast grid too large");
    bcast_stream_kernel<T, OP, SWAP, VB, PT, QT, D, S_TINV><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, h, s, P, ctx->arith_flag);
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

// ---- any shape, any strides: work in the FLAT index space of the dense output -----------------------------------------
// The vector planner needs the inner extent to be a multiple of the vector; [n, 3] * [3] or an outer product with 4099
// columns fell to one element per thread with 64-bit div/mod.  The output is dense whatever the operands look like, so a
// thread owns 4 consecutive flat outputs (one 16-byte store); operand offsets are decoded once per vector with 32-bit
// magic-number division, the other three elements step along the inner dim and re-decode only across a row end.
// Operand loads go through L1 (small broadcast operands live there; a dense operand's four scalar loads coalesce).
struct FastDiv {
    unsigned int d, mul, shift;
};
inline FastDiv make_fastdiv(unsigned int d) {
    FastDiv f;
    f.d = d;
    unsigned int sh = 0;
    while (sh < 32 && (1ull << sh) < d) ++sh;
    f.shift = sh;
    f.mul = (unsigned int)((((1ull << 32) * ((1ull << sh) - d)) / d) + 1);
    return f;
}
__device__ __forceinline__ unsigned int fd_div(unsigned int n, const FastDiv &f) { return (__umulhi(n, f.mul) + n) >> f.shift; }

template <int RANK> struct FlatParams {
    FastDiv ext[RANK];
    unsigned int sa[RANK], sb[RANK];
    unsigned int n;
};

template <int RANK> __device__ __forceinline__ void flat_decode(unsigned int i, const FlatParams<RANK> &P, unsigned int &oa, unsigned int &ob,
                                                               unsigned int &inner) {
    oa = 0; ob = 0;
    unsigned int rem = i;
#pragma unroll
    for (int d = RANK - 1; d >= 1; --d) {
        const unsigned int q = fd_div(rem, P.ext[d]);
        const unsigned int c = rem - q * P.ext[d].d;
        if (d == RANK - 1) inner = c;
        oa += c * P.sa[d]; ob += c * P.sb[d];
        rem = q;
    }
    if (RANK == 1) inner = rem;
    oa += rem * P.sa[0]; ob += rem * P.sb[0];
}

// U flat indices decoded together: each dim's divisor and strides are fetched once and applied to all U.  ncu on the
// one-at-a-time version ([16]^7): the parameter fetches (26 LDC.64 per vector, re-issued every loop trip) kept the ADU pipe
// 94 % busy with mio_throttle / short_scoreboard on top -- the constant bank, not HBM, was the bound (2.5 TB/s).
template <int RANK, int U> __device__ __forceinline__ void flat_decode_multi(const unsigned int (&i)[U], const FlatParams<RANK> &P,
                                                                            unsigned int (&oa)[U], unsigned int (&ob)[U], unsigned int (&inner)[U]) {
    unsigned int rem[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { oa[u] = 0; ob[u] = 0; rem[u] = i[u]; }
#pragma unroll
    for (int d = RANK - 1; d >= 1; --d) {
        const FastDiv f = P.ext[d];
        const unsigned int sa = P.sa[d], sb = P.sb[d];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int q = fd_div(rem[u], f);
            const unsigned int c = rem[u] - q * f.d;
            if (d == RANK - 1) inner[u] = c;
            oa[u] += c * sa; ob[u] += c * sb;
            rem[u] = q;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (RANK == 1) inner[u] = rem[u];
        oa[u] += rem[u] * P.sa[0]; ob[u] += rem[u] * P.sb[0];
    }
}

template <typename T, int OP, int RANK>
__global__ void __launch_bounds__(kThreads) bcast_flat_kernel(T *__restrict__ out, const T *__restrict__ a, const T *__restrict__ b,
                                                              const __grid_constant__ FlatParams<RANK> P, int *flag) {
    constexpr int VN = VecN<T>::value;
    constexpr int U = 2;  // vectors per loop trip (4 is slower: 3.0 vs 3.7 TB/s on [16]^7, registers)
    using V = AVec<T, VN>;
    const unsigned int nvec = P.n / VN;
    const unsigned int ext_in = P.ext[RANK - 1].d, sa_in = P.sa[RANK - 1], sb_in = P.sb[RANK - 1];
    const unsigned int stride = gridDim.x * kThreads;
    for (unsigned int v0 = blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += U * stride) {
        unsigned int i0[U], oa[U], ob[U], inner[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned int v = v0 + u * stride;
            i0[u] = (v < nvec ? v : v0) * VN;  // a trip's missing tail vectors redo the first one (not stored)
        }
        flat_decode_multi<RANK, U>(i0, P, oa, ob, inner);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v0 + u * stride >= nvec) break;
            T x[VN], y[VN];
            if (inner[u] + VN <= ext_in) {  // the whole vector sits in one row
                if (sa_in == 1 && (reinterpret_cast<uintptr_t>(a + oa[u]) & 15u) == 0) {
                    const V t = *reinterpret_cast<const V *>(a + oa[u]);
#pragma unroll
                    for (int k = 0; k < VN; ++k) x[k] = t.v[k];
                } else if (sa_in == 0) {
                    const T t = __ldg(a + oa[u]);
#pragma unroll
                    for (int k = 0; k < VN; ++k) x[k] = t;
                } else {
#pragma unroll
                    for (int k = 0; k < VN; ++k) x[k] = __ldg(a + oa[u] + k * sa_in);
                }
                if (sb_in == 1 && (reinterpret_cast<uintptr_t>(b + ob[u]) & 15u) == 0) {
                    const V t = *reinterpret_cast<const V *>(b + ob[u]);
#pragma unroll
                    for (int k = 0; k < VN; ++k) y[k] = t.v[k];
                } else if (sb_in == 0) {
                    const T t = __ldg(b + ob[u]);
#pragma unroll
                    for (int k = 0; k < VN; ++k) y[k] = t;
                } else {
#pragma unroll
                    for (int k = 0; k < VN; ++k) y[k] = __ldg(b + ob[u] + k * sb_in);
                }
            } else {  // the vector straddles a row end (always, when the inner extent is below VN): decode its elements together
                unsigned int ie[VN], ka[VN], kb[VN], dummy[VN];
#pragma unroll
                for (int k = 0; k < VN; ++k) ie[k] = i0[u] + k;
                flat_decode_multi<RANK, VN>(ie, P, ka, kb, dummy);
#pragma unroll
                for (int k = 0; k < VN; ++k) { x[k] = __ldg(a + ka[k]); y[k] = __ldg(b + kb[k]); }
            }
            V r;
#pragma unroll
            for (int k = 0; k < VN; ++k) r.v[k] = apply_op<OP>(x[k], y[k], flag);
            st_stream(reinterpret_cast<V *>(out + i0[u]), r);
        }
    }
    // the last n % VN elements
    if (blockIdx.x == 0 && threadIdx.x < P.n - nvec * VN) {
        const unsigned int i = nvec * VN + threadIdx.x;
        unsigned int oa, ob, inner;
        flat_decode<RANK>(i, P, oa, ob, inner);
        out[i] = apply_op<OP>(__ldg(a + oa), __ldg(b + ob), flag);
    }
}

// Returns -1 when the shape does not fit the 32-bit flat path (>= 2^31 elements or offsets, unaligned output).
template <typename T, int OP, int RANK>
int launch_flat_bcast_rank(feo_ctx *ctx, T *out, const T *a, const T *b, const Collapsed &c) {
    FlatParams<RANK> P;
    long long n = 1;
    for (int d = 0; d < RANK; ++d) {
        P.ext[d] = make_fastdiv((unsigned int)c.ext[d]);
        P.sa[d] = (unsigned int)c.sa[d]; P.sb[d] = (unsigned int)c.sb[d];
        n *= c.ext[d];
    }
    P.n = (unsigned int)n;
    constexpr int VN = VecN<T>::value;
    long long blocks = (n / VN + kThreads - 1) / kThreads;
    const long long cap = (long long)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    bcast_flat_kernel<T, OP, RANK><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, a, b, P, ctx->arith_flag);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}
template <typename T, int OP>
int launch_flat_bcast(feo_ctx *ctx, T *out, const T *a, const T *b, const Collapsed &c) {
    if (!feo_aligned16(out)) return -1;
    const long long lim = (1ll << 31) - 8;
    long long n = 1, ea = 0, eb = 0;
    for (int d = 0; d < c.rank; ++d) { n *= c.ext[d]; ea += (c.ext[d] - 1) * c.sa[d]; eb += (c.ext[d] - 1) * c.sb[d]; }
    if (n >= lim || ea >= lim || eb >= lim) return -1;
    switch (c.rank) {
    case 1: return launch_flat_bcast_rank<T, OP, 1>(ctx, out, a, b, c);
    case 2: return launch_flat_bcast_rank<T, OP, 2>(ctx, out, a, b, c);
    case 3: return launch_flat_bcast_rank<T, OP, 3>(ctx, out, a, b, c);
    case 4: return launch_flat_bcast_rank<T, OP, 4>(ctx, out, a, b, c);
    default: break;
    }
    // 5 .. 8 collapsed dims: pad with leading extent-1 dims up to 6 or 8
    Collapsed p;
    p.rank = c.rank <= 6 ? 6 : 8;
    const int pad = p.rank - c.rank;
    for (int d = 0; d < p.rank; ++d) {
        p.ext[d] = d < pad ? 1 : c.ext[d - pad];
        p.sa[d] = d < pad ? 0 : c.sa[d - pad];
        p.sb[d] = d < pad ? 0 : c.sb[d - pad];
    }
    if (p.rank == 6) return launch_flat_bcast_rank<T, OP, 6>(ctx, out, a, b, p);
    return launch_flat_bcast_rank<T, OP, 8>(ctx, out, a, b, p);
}

// Vector path with VB-byte accesses; returns -1 when the shape / alignment does not allow it.
template <typename T, int OP, int VB>
int broadcast_vec(feo_ctx *ctx, T *out, const T *a, const T *b, const Collapsed &c, int knob, bool may_defer = false) {
    constexpr int VN = VecW<T, VB>::value;
    const int in = c.rank - 1;
    auto aligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & (uintptr_t)(VB - 1)) == 0; };
    bool vec_ok = (c.sa[in] == 0 || c.sa[in] == 1) && (c.sb[in] == 0 || c.sb[in] == 1) && (c.ext[in] % VN == 0) && aligned(out) &&
                  (c.sa[in] == 0 || aligned(a)) && (c.sb[in] == 0 || aligned(b));
    for (int d = 0; d < in && vec_ok; ++d) {
        if (c.sa[in] == 1 && c.sa[d] % VN != 0) vec_ok = false;
        if (c.sb[in] == 1 && c.sb[d] % VN != 0) vec_ok = false;
    }
    if (!vec_ok) return -1;

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
    P.keep_operands = (ctx->knobs[FEO_KNOB_BCAST_L2KEEP] == 0);
    {
        // L2 prefetch distance for the P-invariant operand, in q-blocks (knob BCAST_PREFETCH: 0 = off, n = distance).  Off by
        // default: measured on the config-2 slab loop (profiles/r2_slab_loop.md) 128 q-blocks ahead gives +1 % (6508 -> 6577
        // GB/s, inside the run-to-run spread), 512 ahead -3 % (the write stream evicts the lines before they are used).
        const int kp = ctx->knobs[FEO_KNOB_BCAST_PREFETCH];
        P.prefetch_q = (kp > 0 && pd >= 0 && qd >= 0) ? kp : 0;
    }
    // the larger of the operands that do not move along some outer dim is the one worth keeping in L2 between launches
    ctx->l2_win_ptr = nullptr;
    ctx->l2_win_bytes = 0;
    if (ctx->l2_persist_bytes) {
        long long span_a = 1, span_b = 1;
        for (int d = 0; d < c.rank; ++d) { span_a += (c.ext[d] - 1) * c.sa[d]; span_b += (c.ext[d] - 1) * c.sb[d]; }
        const bool pick_b = (pd >= 0) && (qd < 0 || span_b >= span_a);
        if (pd >= 0 || qd >= 0) {
            ctx->l2_win_ptr = pick_b ? (const void *)b : (const void *)a;
            ctx->l2_win_bytes = (size_t)(pick_b ? span_b : span_a) * sizeof(T);
        }
    }

    const bool both = pd >= 0 && qd >= 0;
    // Many short alternating axes ([16]^7 with a and b on alternate axes): the P x Q x inner tile is a few KiB and every CTA
    // decodes the other dims with 64-bit div / mod -- the flat-index kernel is 2.6x faster there (0.96 -> 2.5 TB/s).
    if (both && may_defer && knob == 0 && P.nrest >= 2 && P.ni_vec * P.np * P.nq < 4096) return -2;
    if (both && knob != 3) {
        // Both operands are invariant somewhere (config 2: a along q, b along p): PT x QT register tile.  Measured on
        // B200 (profiles/): 8x4 un-pipelined tiles beat the pipelined hold+stream form here (6.3 vs 5.8 TB/s per 1 GiB
        // slab, 7.15 vs 6.5 TB/s on a 32 GiB call) -- many short CTAs overlap their load and store phases better.
        if (knob == 2 || P.np < 8) return launch_tile<T, OP, VB, 4, 4, true, true>(ctx, out, a, b, P);
        return launch_tile<T, OP, VB, 8, 4, true, true>(ctx, out, a, b, P);
    }
    // One operand invariant along one dim.  If the OTHER operand is a full-size stream (matrix (op) row vector: as many bytes
    // read as written), holding the small operand saves next to nothing and the prefetch ring only serialises: the plain
    // 4-rows-in-flight path below is faster (f32: 5.65 -> 6.9 TB/s).  The outer product (the streamed operand is a splat
    // scalar per row) stays on hold + stream (7.06 vs 4.25 TB/s).
    const bool streamed_full = (qd >= 0) ? (P.b_inner == 1) : (P.a_inner == 1);
    const bool plain_stream = sizeof(T) == 4 && streamed_full && knob != 6;
    if ((pd >= 0 || qd >= 0) && knob != 3 && knob != 5 && !plain_stream) {
        // One operand is invariant along one dimension (outer product, matrix (op) row vector): hold its vector in
        // registers and stream the other operand along that dimension with a two-step prefetch.
        StreamParams S;
        memset(&S, 0, sizeof S);
        S.ni_vec = P.ni_vec; S.tx = P.tx; S.ty = P.ty; S.i_blocks = P.i_blocks;
        S.nrest = P.nrest;
        S.nt = 1;
        const bool hold_a = (qd >= 0);  // a is invariant along q: hold a, stream b along q
        for (int d = 0; d < P.nrest; ++d) {
            S.rest_ext[d] = P.rest_ext[d]; S.rest_so[d] = P.rest_so[d];
            S.rest_sh[d] = hold_a ? P.rest_sa[d] : P.rest_sb[d];
            S.rest_ss[d] = hold_a ? P.rest_sb[d] : P.rest_sa[d];
        }
        if (hold_a) {
            S.h_inner = P.a_inner; S.s_inner = P.b_inner;
            S.nl = P.nq; S.ss_l = P.sb_q; S.so_l = P.so_q;
            return launch_stream<T, OP, false, VB, 1, 4, 2, true>(ctx, out, a, b, S);
        }
        // b is invariant along p: hold b, stream a along p, operands swapped
        S.h_inner = P.b_inner; S.s_inner = P.a_inner;
        S.nl = P.np; S.ss_l = P.sa_p; S.so_l = P.so_p;
        return launch_stream<T, OP, true, VB, 1, 4, 2, true>(ctx, out, b, a, S);
    }
    // no invariance to exploit (or tiling disabled): fold P/Q back into the rest dims
    if (pd >= 0) { P.rest_ext[P.nrest] = P.np; P.rest_sa[P.nrest] = P.sa_p; P.rest_sb[P.nrest] = P.sb_p; P.rest_so[P.nrest] = P.so_p; P.nrest++; }
    if (qd >= 0) { P.rest_ext[P.nrest] = P.nq; P.rest_sa[P.nrest] = P.sa_q; P.rest_sb[P.nrest] = P.sb_q; P.rest_so[P.nrest] = P.so_q; P.nrest++; }
    P.np = P.nq = 1; P.sa_p = P.sb_p = P.so_p = P.sa_q = P.sb_q = P.so_q = 0;
    if (knob != 3 && P.nrest > 0) {
        // e.g. matrix (op) column vector (b's inner stride is 0: a splat): nothing is re-used, but a thread that moves one
        // vector has one load in flight.  Walk the largest outer dim as the q loop instead: 4 rows per thread in flight, and
        // the block's spare thread rows (short inner extents) spread over it too.
        int big = 0;
        for (int d = 1; d < P.nrest; ++d)
            if (P.rest_ext[d] > P.rest_ext[big]) big = d;
        P.nq = P.rest_ext[big]; P.sa_q = P.rest_sa[big]; P.sb_q = P.rest_sb[big]; P.so_q = P.rest_so[big];
        for (int d = big; d + 1 < P.nrest; ++d) {
            P.rest_ext[d] = P.rest_ext[d + 1]; P.rest_sa[d] = P.rest_sa[d + 1]; P.rest_sb[d] = P.rest_sb[d + 1]; P.rest_so[d] = P.rest_so[d + 1];
        }
        P.nrest--;
        return launch_tile<T, OP, VB, 1, 4, false, true>(ctx, out, a, b, P);
    }
    P.ty = 1;  // q extent is 1: spare rows of threads would idle, keep the block one row wide
    P.tx = kThreads;
    P.i_blocks = (P.ni_vec + P.tx - 1) / P.tx;
    return launch_tile<T, OP, VB, 1, 1, true, true>(ctx, out, a, b, P);
}

template <typename T, int OP>
int broadcast_typed(feo_ctx *ctx, T *out, const T *a, const T *b, int rank, const size_t *shape, const size_t *sa,
                    const size_t *sb) {
    const Collapsed c = collapse(rank, shape, sa, sb);
    const int knob = ctx->knobs[FEO_KNOB_BCAST_TILE];
    // same-shape after collapsing: the flat kernel is the right tool
    if (c.rank == 1 && c.sa[0] == 1 && c.sb[0] == 1 && knob != 4)
        return launch_flat<T, OP, false>(ctx, out, a, b, (T)0, (size_t)c.ext[0]);
    {   // byte ranges of this launch for the overlap hazard check (feo_pdl_admit)
        long long no = 1, ea = 0, eb = 0;
        for (int d = 0; d < c.rank; ++d) { no *= c.ext[d]; ea += (c.ext[d] - 1) * c.sa[d]; eb += (c.ext[d] - 1) * c.sb[d]; }
        ctx->pending.out = {(uintptr_t)out, (uintptr_t)(out + no)};
        ctx->pending.in[0] = {(uintptr_t)a, (uintptr_t)(a + ea + 1)};
        ctx->pending.in[1] = {(uintptr_t)b, (uintptr_t)(b + eb + 1)};
        ctx->pending_valid = true;
    }
    if (knob != 4) {
        int rc = -1;
        if (knob != 7) rc = broadcast_vec<T, OP, 16>(ctx, out, a, b, c, knob, true);
        if (rc == -2) {  // the vector planner prefers the flat kernel; fall back to its own tiles if that cannot take the shape
            rc = launch_flat_bcast<T, OP>(ctx, out, a, b, c);
            if (rc == -1) rc = broadcast_vec<T, OP, 16>(ctx, out, a, b, c, knob, false);
        }
        if (rc != -1) return rc;
        rc = launch_flat_bcast<T, OP>(ctx, out, a, b, c);
        if (rc != -1) return rc;
    }
    return launch_generic<T, OP>(ctx, out, a, b, c);
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

}  // namespace

#define FEO_EWISE_INSTANTIATE(T, SFX)                                                                                        \
    int feo_ewise_typed_##SFX(feo_ctx *ctx, int op, void *out, const void *a, const void *b, const void *scalar_host, size_t n) { \
        if (scalar_host) return launch_flat_op<T, true>(ctx, op, (T *)out, (const T *)a, nullptr, *(const T *)scalar_host, n);  \
        return launch_flat_op<T, false>(ctx, op, (T *)out, (const T *)a, (const T *)b, (T)0, n);                              \
    }                                                                                                                        \
    int feo_bcast_typed_##SFX(feo_ctx *ctx, int op, void *out, const void *a, const void *b, int rank, const size_t *shape,  \
                              const size_t *sa, const size_t *sb) {                                                          \
        return broadcast_op<T>(ctx, op, (T *)out, (const T *)a, (const T *)b, rank, shape, sa, sb);                          \
    }
