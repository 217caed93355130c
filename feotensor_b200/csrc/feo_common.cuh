// feo_common.cuh -- shared host/device plumbing of libfeotensor_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <limits>
#include <string>
#include <vector>
#include <type_traits>

#include "../../include/feotensor_b200.h"

// ---- context ------------------------------------------------------------------------------------
enum feo_knob {
    FEO_KNOB_BCAST_TILE = 0,    // 0 = auto, 2 = 4x4 tiles, 3 = 1x1 vector tiles, 4 = scalar fallback, 5 = never hold+stream, 6 = always (see broadcast_vec), 7 = flat-index kernel
    FEO_KNOB_REDUCE_SPLIT = 1,  // 0 = auto, > 0 forced number of splits S, -1 = one-kernel generic fallback instead of multi-pass
    FEO_KNOB_REDUCE_BLOCKS_PER_SM = 2,  // target resident blocks per SM used by the split planner (0 = auto)
    FEO_KNOB_MATMUL_TF32_KC = 3,  // k-blocks (of 32) per TMEM accumulation chunk in the 3xTF32 matmul (0 = default 2)
    FEO_KNOB_BCAST_LCHUNK = 4,  // rows of the loop dimension per thread in the hold+stream broadcast kernel (0 = auto)
    FEO_KNOB_PDL = 5,           // programmatic dependent launch of independent back-to-back broadcast kernels: 0 = on, 1 = off
    FEO_KNOB_BCAST_L2KEEP = 6,  // broadcast operands in L2: 0 = evict_last load policy, 1 = nothing, 2 = persisting access-policy window on the
                                // operand that is re-read across launches (sets the device's persisting-L2 carve-out; 0 / 1 reset it)
    FEO_KNOB_REDUCE_FUSE = 7,   // split reductions: last-arriving CTA combines the partials in the same launch: 0 = auto (small inputs), 1 = off, 2 = always
    FEO_KNOB_GATHER_OVERLAP = 8,  // sharded-B matmul, how B's panels travel: 0 = auto (copy engines piecewise when every panel is whole k-blocks, else
                                  // the matmul kernel pulls them itself), 1 = gather kernel in stream order, 2 = gather kernel on a side stream,
                                  // 3 = copy engines, 4 = in-kernel pull (TMA bulk copies)
    FEO_KNOB_MATMUL_SIMT = 9,   // SIMT matmul, aligned 4-byte types: 0 = auto (cp.async pipeline), 1 = register double buffering, 2..13 = pipeline variants (matmul_simt.cu)
    FEO_KNOB_PEER_TIMEOUT_S = 10,  // seconds a kernel waits for a peer before it raises the status flag and goes on (0 = for ever); default: FEO_PEER_TIMEOUT_S or 600
    FEO_KNOB_MATMUL_TF32_PAIR = 11,  // 3xTF32 matmul: 0 = CTA-pair kernel (tcgen05 cta_group::2, persistent), 1 = one 128 x 256 tile per CTA
    FEO_KNOB_MATMUL_TF32_BN = 12,    // CTA-pair kernel: tile width (multiple of 32, <= 256); 0 = fewest waves x width
    FEO_KNOB_MATMUL_TF32_BRAW = 13,  // CTA-pair kernel: 0 = B read as raw f32 tiles and split hi/lo inside the kernel (needs N % 4 == 0), 1 = split + transpose pre-pass
    FEO_KNOB_BCAST_PREFETCH = 14,    // tiled broadcast kernel: L2 prefetch of the P-invariant operand, q-blocks ahead (0 = off)
    FEO_KNOB_PULL_SHAPE = 15,        // copy-engine route of the sharded-B matmul: 100 * streams (1..4) + k-blocks per piece (power of two); 0 = 2 streams x 16
    FEO_KNOB_COUNT = 16
};

// Address ranges of a launch, for the hazard check that lets independent back-to-back kernels overlap their tails.
struct feo_range { uintptr_t lo = 0, hi = 0; };
struct feo_launch_rec { feo_range out, in[2]; };

struct feo_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // stream created by feo_init
    cudaStream_t side_stream = nullptr;  // high-priority helper stream (created on first use, see feo_side_stream)
    cudaStream_t pull_streams[3] = {nullptr, nullptr, nullptr};  // further helper streams: the copy-engine pull deals its pieces over side_stream and these
    cudaEvent_t pull_joins[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_open = nullptr;
    cudaMemPool_t pool = nullptr;       // this context's own stream-ordered pool (never trimmed at synchronisation)
    void *workspace = nullptr;          // scratch for two-stage reductions
    size_t workspace_bytes = 0;
    int *arith_flag = nullptr;          // device int, sticky
    int *host_flags = nullptr;          // pinned [2]: where feo_sync / feo_download read arith_flag and status_flag into
    unsigned int *host_ones = nullptr;  // pinned array of ones (readiness flags published by H2D copies, see matmul_tf32x3.cu)
    size_t host_ones_len = 0;
    int *status_flag = nullptr;         // device int, sticky: feo_peer::kStatusPeerTimeout (arith_flag + 1)
    unsigned long long peer_timeout_ns = 600ull * 1000000000ull;
    unsigned int *counters = nullptr;   // zeroed arrival counters for the fused reduction combine
    size_t counters_len = 0;
    unsigned int *xtail_ticket = nullptr;  // zeroed CTA counter of the exchange-at-the-tail reductions (the entry behind `counters`)
    void *xtail_req = nullptr;          // feo_peer::XTailReq of the sharded reduction being issued (peer_sync.cuh), else null
    uint64_t launches = 0;
    size_t l2_persist_bytes = 0;        // persisting-L2 carve-out set by knob BCAST_L2KEEP = 2 (0: none)
    size_t l2_window_max = 0;           // cudaDevAttrMaxAccessPolicyWindowSize
    const void *l2_win_ptr = nullptr;   // operand range of the launch being prepared that deserves the window
    size_t l2_win_bytes = 0;
    int knobs[FEO_KNOB_COUNT] = {0};
    std::vector<feo_launch_rec> window;  // launches since the last fully serialising operation (see feo_pdl_admit)
    feo_launch_rec pending;              // ranges of the launch being prepared
    bool pending_valid = false;
    std::string last_error;
};

int feo_fail(feo_ctx *ctx, int status, const char *fmt, ...);
// Every entry point makes the context's device current first: a process may hold contexts on several devices (and the host
// framework may have switched devices since the last call).  cudaSetDevice on the already-current device is a cheap no-op.
#define FEO_ENTER(c)                                   \
    do {                                               \
        feo_ctx *c_ = (c);                             \
        if (c_) cudaSetDevice(c_->device);             \
    } while (0)
// True if the launch described by ctx->pending may start before the kernels in the window have finished (no RAW / WAR /
// WAW overlap with any of them, window not full); records it.  False: launch normally (full stream order).
bool feo_pdl_admit(feo_ctx *ctx);
int feo_workspace(feo_ctx *ctx, size_t bytes, void **ptr);
// Stream-ordered allocation from the context's own pool.  A private pool means a block freed on another context's stream
// is never handed out here with a hidden cross-stream dependency -- which would deadlock kernels that wait for peers.
cudaError_t feo_alloc_async(feo_ctx *ctx, void **ptr, size_t bytes);
// Creates ctx->side_stream / ev_fork / ev_join on first use.
int feo_side_stream(feo_ctx *ctx);

#define FEO_CUDA(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return feo_fail((ctx), FEO_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                            __FILE__, __LINE__);                                                  \
    } while (0)

// A normally launched kernel starts after everything before it in the stream has finished: the overlap window closes.
#define FEO_LAUNCH_CHECK(ctx)          \
    do {                               \
        (ctx)->launches++;             \
        (ctx)->window.clear();         \
        (ctx)->pending_valid = false;  \
        FEO_CUDA(ctx, cudaGetLastError()); \
    } while (0)

// dtype dispatch: FEO_DISPATCH(dtype, T, { ... code using T ... })
#define FEO_DISPATCH(ctx, dtype, T, ...)                                         \
    switch (dtype) {                                                             \
    case FEO_F32: { using T = float; __VA_ARGS__; break; }                       \
    case FEO_F64: { using T = double; __VA_ARGS__; break; }                      \
    case FEO_I32: { using T = int32_t; __VA_ARGS__; break; }                     \
    case FEO_I64: { using T = int64_t; __VA_ARGS__; break; }                     \
    case FEO_U32: { using T = uint32_t; __VA_ARGS__; break; }                    \
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", (int)(dtype)); \
    }

// ---- element arithmetic -------------------------------------------------------------------------
// Integers wrap like Rust release builds (arithmetic done in the unsigned twin); floats are one
// IEEE operation per call (the library is built without fast-math; nothing here can contract).
template <typename T> struct UnsignedOf { using type = T; };
template <> struct UnsignedOf<int32_t> { using type = uint32_t; };
template <> struct UnsignedOf<int64_t> { using type = unsigned long long; };

template <typename T> __host__ __device__ __forceinline__ T w_add(T a, T b) {
    if constexpr (std::is_floating_point<T>::value) return a + b;
    else { using U = typename UnsignedOf<T>::type; return (T)((U)a + (U)b); }
}
template <typename T> __host__ __device__ __forceinline__ T w_sub(T a, T b) {
    if constexpr (std::is_floating_point<T>::value) return a - b;
    else { using U = typename UnsignedOf<T>::type; return (T)((U)a - (U)b); }
}
template <typename T> __host__ __device__ __forceinline__ T w_mul(T a, T b) {
    if constexpr (std::is_floating_point<T>::value) return a * b;
    else { using U = typename UnsignedOf<T>::type; return (T)((U)a * (U)b); }
}
// Division: floats IEEE; integers write 0 and raise the sticky flag on x/0 and MIN/-1 (Rust panics).
template <typename T> __device__ __forceinline__ T w_div(T a, T b, int *flag) {
    if constexpr (std::is_floating_point<T>::value) {
        return a / b;
    } else {
        bool bad = (b == (T)0);
        if constexpr (std::is_signed<T>::value) bad = bad || (a == std::numeric_limits<T>::min() && b == (T)-1);
        if (bad) { if (flag) atomicOr(flag, 1); return (T)0; }
        if constexpr (sizeof(T) == 8 && std::is_signed<T>::value) {
            // i64 tensors mostly hold small numbers, and the 64-bit signed division is a ~100-instruction routine (4.3 TB/s on
            // a same-shape div, ALU-bound).  Both operands inside the i32 range: the quotient is the 32-bit one, except
            // INT_MIN / -1, which only overflows in 32 bits -- that one takes the long route with everything else.
            const int x = (int)a, y = (int)b;
            if ((T)x == a && (T)y == b && !(x == std::numeric_limits<int>::min() && y == -1)) return (T)(x / y);
        }
        return a / b;
    }
}
template <int OP, typename T> __device__ __forceinline__ T apply_op(T a, T b, int *flag) {
    if constexpr (OP == FEO_ADD) return w_add(a, b);
    else if constexpr (OP == FEO_SUB) return w_sub(a, b);
    else if constexpr (OP == FEO_MUL) return w_mul(a, b);
    else return w_div(a, b, flag);
}

// max / min of two values as the reductions use them (tensor.rs:248, :300 replace on a strict compare).  Floats go through
// fmax / fmin (one FMNMX): the result is one of the operands, and a NaN operand never wins -- the reference's strict
// compares skip NaN elements too -- and never swallows the other operand of a tree node either.
template <typename T> __device__ __forceinline__ T feo_max2(T a, T b) {
    if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmax(a, b);
    else return (b > a) ? b : a;
}
template <typename T> __device__ __forceinline__ T feo_min2(T a, T b) {
    if constexpr (std::is_same<T, float>::value) return fminf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmin(a, b);
    else return (b < a) ? b : a;
}

// ---- 128-bit (and narrower) vector access ---------------------------------------------------------
template <typename T, int N> struct alignas(sizeof(T) * N) AVec { T v[N]; };
template <typename T> struct VecN { static constexpr int value = 16 / sizeof(T); };

// sm_100 adds 256-bit global accesses (LDG.E.256 / STG.E.256): one 32-byte sector per thread per instruction.
template <typename T, int BYTES> struct VecW { static constexpr int value = BYTES / (int)sizeof(T); };

// Streaming (read-once) load: bypass L1 allocation.  (Not `volatile`: the compiler may hoist the next batch's loads
// above the current batch's arithmetic; batches are combined as trees so the loads of one batch stay grouped.)
template <typename V> __device__ __forceinline__ V ld_stream(const V *p) {
    static_assert(sizeof(V) == 32 || sizeof(V) == 16 || sizeof(V) == 8 || sizeof(V) == 4, "vector size");
    V r;
    if constexpr (sizeof(V) == 32) {
        unsigned long long t[4];
        asm("ld.global.nc.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(t[0]), "=l"(t[1]), "=l"(t[2]), "=l"(t[3]) : "l"(p));
        __builtin_memcpy(&r, t, 32);
    } else if constexpr (sizeof(V) == 16) {
        uint4 t;
        asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
        __builtin_memcpy(&r, &t, 16);
    } else if constexpr (sizeof(V) == 8) {
        uint2 t;
        asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(p));
        __builtin_memcpy(&r, &t, 8);
    } else {
        uint32_t t;
        asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(t) : "l"(p));
        __builtin_memcpy(&r, &t, 4);
    }
    return r;
}
// Reused (keep in L1/L2) read-only load.
template <typename V> __device__ __forceinline__ V ld_keep(const V *p) {
    V r;
    if constexpr (sizeof(V) == 32) {
        unsigned long long t[4];
        asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(t[0]), "=l"(t[1]), "=l"(t[2]), "=l"(t[3]) : "l"(p));
        __builtin_memcpy(&r, t, 32);
    } else if constexpr (sizeof(V) == 16) {
        uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
        __builtin_memcpy(&r, &t, 16);
    } else if constexpr (sizeof(V) == 8) {
        uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        __builtin_memcpy(&r, &t, 8);
    } else {
        uint32_t t = __ldg(reinterpret_cast<const uint32_t *>(p));
        __builtin_memcpy(&r, &t, 4);
    }
    return r;
}
// Reused operand that should survive a large write stream in L2 (e.g. the broadcast operand re-read by every slab
// launch): 128-bit read-only load with an L2 evict_last cache policy.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <typename V> __device__ __forceinline__ V ld_keep_last(const V *p, unsigned long long pol) {
    static_assert(sizeof(V) == 16, "128-bit only");
    V r;
    uint4 t;
    asm("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(pol));
    __builtin_memcpy(&r, &t, 16);
    return r;
}
// Streaming store (evict-first): results are written once and not re-read by the kernel.
template <typename V> __device__ __forceinline__ void st_stream(V *p, const V &v) {
    if constexpr (sizeof(V) == 32) {
        unsigned long long t[4]; __builtin_memcpy(t, &v, 32);
        asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(t[0]), "l"(t[1]), "l"(t[2]), "l"(t[3]) : "memory");
    } else if constexpr (sizeof(V) == 16) {
        uint4 t; __builtin_memcpy(&t, &v, 16);
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
    } else if constexpr (sizeof(V) == 8) {
        uint2 t; __builtin_memcpy(&t, &v, 8);
        asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(t.x), "r"(t.y) : "memory");
    } else {
        uint32_t t; __builtin_memcpy(&t, &v, 4);
        asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(t) : "memory");
    }
}

static inline bool feo_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool feo_aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
static inline size_t feo_ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- entry points implemented per kernel family (called from feo_api.cu) ---------------------------
int feo_impl_ewise(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, const void *scalar_host, size_t n);
int feo_impl_broadcast(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, int rank,
                       const size_t *shape, const size_t *sa, const size_t *sb);
int feo_impl_fill(feo_ctx *ctx, int dtype, void *out, const void *value, size_t n);
int feo_impl_pow(feo_ctx *ctx, int dtype, void *out, const void *a, const void *power_host, size_t n);
int feo_impl_eye(feo_ctx *ctx, int dtype, void *out, size_t rows, size_t cols);
int feo_impl_fill_uniform(feo_ctx *ctx, int dtype, void *out, size_t n, uint64_t seed, uint64_t first_index, double lo, double hi);

// reduce `mode`: 0 final result, 1 partial (no division; VAR writes mean|M2), 2 squared deviations from aux
int feo_impl_reduce(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, const void *aux, int rank,
                    const size_t *shape, int naxes, const size_t *axes, int mode);
int feo_impl_var_merge(feo_ctx *ctx, int dtype, void *out, const void *partials, size_t nparts, size_t nout,
                       size_t count_per_part, const void *divisor_host);
// loader: 1 = a[i]*b[i] rows (dot / batched dot), 2 = A[i,j]*v[j] (matvec), 3 = v[i]*A[i,j] (vecmat)
int feo_impl_dotlike(feo_ctx *ctx, int loader, int dtype, void *out, const void *a, const void *b, size_t outer, size_t inner);
int feo_impl_matmul_simt(feo_ctx *ctx, int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N);
int feo_impl_matmul_tf32x3(feo_ctx *ctx, void *C, const void *A, const void *B, size_t M, size_t K, size_t N);
bool feo_matmul_tf32x3_supported(size_t M, size_t K, size_t N);

// Closed form of the reference's count-in-T (tensor.rs:112-121): floats saturate, integers wrap.
template <typename T> static inline T feo_count_in_type(size_t count) {
    if constexpr (std::is_same<T, float>::value) return (float)(count < (size_t(1) << 24) ? count : (size_t(1) << 24));
    else if constexpr (std::is_same<T, double>::value) return (double)(count < (size_t(1) << 53) ? count : (size_t(1) << 53));
    else return (T)count;  // modular conversion == repeated wrapping +1
}
template <typename T> static inline T feo_divisor(int rank, const size_t *shape, int naxes, const size_t *axes) {
    if (naxes > 0) {
        T n = (T)1;
        for (int i = 0; i < naxes; ++i) n = w_mul(n, feo_count_in_type<T>(shape[axes[i]]));
        return n;
    }
    size_t size = 1;
    for (int d = 0; d < rank; ++d) size *= shape[d];
    return feo_count_in_type<T>(size);
}
