// reduce_kernels.cuh -- K2/K3/K5: axis reductions (tensor.rs:70-307), dot (vector.rs:71-78),
// matrix.vector (matrix.rs:92-103) and vector.matrix (vector.rs:80-91) for sm_100a.
//
// Geometry.  A dense row-major tensor with some axes reduced collapses (extent-1 dims dropped,
// neighbours with the same kept/reduced status merged) to alternating groups.  Up to four groups
// fit the canonical form  in[k1][r1][k2][r2] -> out[k1][k2]  (any extent may be 1):
//   cols  kernel  r2 == 1        kept axis innermost: threads along k2 (128-bit), loop over r1
//   micro kernel  2 <= r2 <= 32  short reduced innermost axis: one thread per output, r2 contiguous
//                                elements (128-bit loads) per r1 step
//   rows  kernel  r2 > 32        long reduced innermost axis: a warp or a block per output, lanes
//                                stride the row with 128-bit loads, warp-shuffle + shared-memory tree
//   generic       anything else  (more than four groups): a block per output, div/mod indexing
// Too few outputs to fill 148 SMs -> the reduced range is split S ways, stage 1 writes accumulators
// to a workspace and a deterministic stage 2 merges them in split order (no atomics anywhere).
// One pass over the input for every kind: HBM-bound, algorithmic bytes = sizeof(T)*(N_in + N_out).
#pragma once

#include "feo_common.cuh"
#include "peer_sync.cuh"

namespace feo_red {

constexpr int kThreads = 256;
enum { K_SUM = 0, K_MAX = 1, K_MIN = 2, K_WELFORD = 3, K_SQDEV = 4, K_INTVAR = 5, K_SUMSQ = 6 };
enum { FIN_NONE = 0, FIN_DIV = 1, FIN_VAR = 2, FIN_MEAN_M2 = 3 };
enum { LD_PLAIN = 0, LD_MUL = 1, LD_MATVEC = 2, LD_VECMAT = 3 };

// ---- accumulators ---------------------------------------------------------------------------------
template <typename T, int KIND> struct Acc;

template <typename T> struct Acc<T, K_SUM> {
    T v;
    __device__ __forceinline__ void init(T) { v = (T)0; }
    __device__ __forceinline__ void push(T x) { v = w_add(v, x); }
    __device__ __forceinline__ void merge(const Acc &o) { v = w_add(v, o.v); }
    // Batches are combined as a balanced tree: short dependency chains, and every load of the batch is needed by the
    // first level, so the scheduler keeps the loads grouped ahead of the arithmetic (a serial chain made ptxas sink each
    // load next to its use: one load in flight per thread, -20 % on the rows kernel).
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        if constexpr ((N & (N - 1)) == 0 && N >= 2) {
            T t[N / 2];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) t[i] = w_add(x[i], x[i + N / 2]);
#pragma unroll
            for (int w = N / 4; w >= 1; w /= 2)
#pragma unroll
                for (int i = 0; i < w; ++i) t[i] = w_add(t[i], t[i + w]);
            push(t[0]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) push(x[i]);
        }
    }
};
template <typename T> struct Acc<T, K_SQDEV> {  // sum of (x - mean)^2 for a given mean (integer var, tensor.rs:196-198)
    T v, m;
    __device__ __forceinline__ void init(T mean) { v = (T)0; m = mean; }
    __device__ __forceinline__ void push(T x) { const T c = w_sub(x, m); v = w_add(v, w_mul(c, c)); }
    __device__ __forceinline__ void merge(const Acc &o) { v = w_add(v, o.v); }
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) push(x[i]);
    }
};
template <typename T> struct Acc<T, K_MAX> {
    T v;
    // Seed: the identity of max.  Floats start at -inf, so a slice of -inf reduces to -inf exactly like the reference, whose
    // init min(0, every x) (tensor.rs:234-237) is <= every element.  NaN never wins the strict compare (as in tensor.rs:248);
    // the reference's order-dependent NaN handling in its init fold is not reproduced (unspecified).
    __device__ __forceinline__ void init(T) {
        if constexpr (std::is_floating_point<T>::value) v = -std::numeric_limits<T>::infinity();
        else v = std::numeric_limits<T>::lowest();
    }
    __device__ __forceinline__ void push(T x) { v = feo_max2(v, x); }  // replace on a strict compare, tensor.rs:248
    __device__ __forceinline__ void merge(const Acc &o) { push(o.v); }
    static __device__ __forceinline__ T pick(T a, T b) { return feo_max2(a, b); }
    // Batches are combined as a balanced tree: short dependency chains, and every load of the batch is needed by the
    // first level, so the scheduler keeps the loads grouped ahead of the arithmetic (a serial chain made ptxas sink each
    // load next to its use: one load in flight per thread, -20 % on the rows kernel).
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        if constexpr ((N & (N - 1)) == 0 && N >= 2) {
            T t[N / 2];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) t[i] = pick(x[i], x[i + N / 2]);
#pragma unroll
            for (int w = N / 4; w >= 1; w /= 2)
#pragma unroll
                for (int i = 0; i < w; ++i) t[i] = pick(t[i], t[i + w]);
            push(t[0]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) push(x[i]);
        }
    }
};
template <typename T> struct Acc<T, K_MIN> {
    T v;
    __device__ __forceinline__ void init(T) {  // +inf for floats: see K_MAX
        if constexpr (std::is_floating_point<T>::value) v = std::numeric_limits<T>::infinity();
        else v = std::numeric_limits<T>::max();
    }
    __device__ __forceinline__ void push(T x) { v = feo_min2(v, x); }  // tensor.rs:300
    __device__ __forceinline__ void merge(const Acc &o) { push(o.v); }
    static __device__ __forceinline__ T pick(T a, T b) { return feo_min2(a, b); }
    // Batches are combined as a balanced tree: short dependency chains, and every load of the batch is needed by the
    // first level, so the scheduler keeps the loads grouped ahead of the arithmetic (a serial chain made ptxas sink each
    // load next to its use: one load in flight per thread, -20 % on the rows kernel).
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        if constexpr ((N & (N - 1)) == 0 && N >= 2) {
            T t[N / 2];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) t[i] = pick(x[i], x[i + N / 2]);
#pragma unroll
            for (int w = N / 4; w >= 1; w /= 2)
#pragma unroll
                for (int i = 0; i < w; ++i) t[i] = pick(t[i], t[i + w]);
            push(t[0]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) push(x[i]);
        }
    }
};
// Integer variance of ONE complete batch (the whole reduced range of an output sits in a thread's registers: the
// single-row micro geometry): the reference's two passes (tensor.rs:171-177, 188-202) fused -- truncated integer mean,
// wrapping squares.  Only push_batch is meaningful; the planner uses this kind nowhere else.
template <typename T> struct Acc<T, K_INTVAR> {
    T v, n;
    __device__ __forceinline__ void init(T divisor) { v = (T)0; n = divisor; }
    __device__ __forceinline__ void push(T) {}
    __device__ __forceinline__ void merge(const Acc &) {}
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) s = w_add(s, x[i]);
        const T mean = w_div(s, n, (int *)nullptr);
        T q = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) { const T c = w_sub(x[i], mean); q = w_add(q, w_mul(c, c)); }
        v = q;
    }
};
// Integer variance in ONE pass for any geometry.  The reference's two passes (tensor.rs:171-177, 188-202) are wrapping
// integer arithmetic, i.e. arithmetic in the ring Z / 2^k, where  sum (x - m)^2 = sum x^2 - 2 m sum x + n m^2  holds exactly
// whatever overflows on the way.  So the wrapping sum and sum of squares determine the reference's result bit for bit:
// m = S1 / n (truncating, on the wrapped S1 like the reference's own mean), var = (S2 - 2 m S1 + n m^2) / n.
template <typename T> struct Acc<T, K_SUMSQ> {
    T v, q;
    __device__ __forceinline__ void init(T) { v = (T)0; q = (T)0; }
    __device__ __forceinline__ void push(T x) { v = w_add(v, x); q = w_add(q, w_mul(x, x)); }
    __device__ __forceinline__ void merge(const Acc &o) { v = w_add(v, o.v); q = w_add(q, o.q); }
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        T s = (T)0, t = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) { s = w_add(s, x[i]); t = w_add(t, w_mul(x[i], x[i])); }
        v = w_add(v, s); q = w_add(q, t);
    }
};
// Welford / Chan: (count, mean, M2).  Batches are combined two-pass in registers and merged with one
// division per batch, so the per-element cost stays ~4 flops while keeping Welford's stability.
template <typename T> struct Acc<T, K_WELFORD> {
    T n, mean, m2;
    __device__ __forceinline__ void init(T) { n = (T)0; mean = (T)0; m2 = (T)0; }
    __device__ __forceinline__ void push(T x) {
        n += (T)1;
        const T d = x - mean;
        mean += d / n;
        m2 += d * (x - mean);
    }
    __device__ __forceinline__ void merge_raw(T on, T omean, T om2) {
        if (on == (T)0) return;
        const T nn = n + on;
        const T d = omean - mean;
        // equal counts (every level of a shuffle / shared-memory tree, equal splits, equal shards): on / nn is exactly 0.5, so
        // the division -- a dozen instructions, and the bulk of a merge -- is skipped without changing a bit
        T f;
        if (on == n) f = (T)0.5;
        else f = on / nn;
        mean += d * f;
        m2 += om2 + d * d * n * f;
        n = nn;
    }
    __device__ __forceinline__ void merge(const Acc &o) { merge_raw(o.n, o.mean, o.m2); }
    // the first cnt of N values (the others are 0), two-pass in registers like push_batch
    template <int N> __device__ __forceinline__ void push_masked(const T (&x)[N], int cnt) {
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += x[i];
        const T bm = s / (T)cnt;
        T q = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) { const T c = x[i] - bm; q += (i < cnt) ? c * c : (T)0; }
        if (n == (T)0) { n = (T)cnt; mean = bm; m2 = q; }
        else merge_raw((T)cnt, bm, q);
    }
    template <int N> __device__ __forceinline__ void push_batch(const T (&x)[N]) {
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += x[i];
        const T bm = s * ((T)1 / (T)N);
        T q = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) { const T c = x[i] - bm; q += c * c; }
        if (n == (T)0) { n = (T)N; mean = bm; m2 = q; }  // first batch: nothing to merge, no division
        else merge_raw((T)N, bm, q);
    }
    // The same with the merge weight f = N / (n + N) handed in: accumulators that advance in lock step (the W components of a
    // column vector, the J columns of a thread) share one division per batch instead of doing one each.  Same bits.
    template <int N> __device__ __forceinline__ void push_batch_f(const T (&x)[N], T f) {
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += x[i];
        const T bm = s * ((T)1 / (T)N);
        T q = (T)0;
#pragma unroll
        for (int i = 0; i < N; ++i) { const T c = x[i] - bm; q += c * c; }
        if (n == (T)0) { n = (T)N; mean = bm; m2 = q; }
        else {
            const T d = bm - mean;
            mean += d * f;
            m2 += q + d * d * n * f;
            n += (T)N;
        }
    }
};

template <int KIND, typename T> __device__ __forceinline__ T aux_val(const T *aux, long long o) {
    if constexpr (KIND == K_SQDEV) return aux[o];
    else return (T)0;
}

template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }
template <> __device__ __forceinline__ int64_t shfl_down_t<int64_t>(int64_t v, int off) {
    return (int64_t)__shfl_down_sync(0xffffffffu, (long long)v, off);
}
template <typename T, int KIND> __device__ __forceinline__ Acc<T, KIND> shfl_down_acc(const Acc<T, KIND> &a, int off) {
    Acc<T, KIND> r;
    if constexpr (KIND == K_WELFORD) {
        r.n = shfl_down_t(a.n, off); r.mean = shfl_down_t(a.mean, off); r.m2 = shfl_down_t(a.m2, off);
    } else if constexpr (KIND == K_SQDEV) {
        r.v = shfl_down_t(a.v, off); r.m = a.m;
    } else if constexpr (KIND == K_INTVAR) {
        r.v = shfl_down_t(a.v, off); r.n = a.n;
    } else if constexpr (KIND == K_SUMSQ) {
        r.v = shfl_down_t(a.v, off); r.q = shfl_down_t(a.q, off);
    } else {
        r.v = shfl_down_t(a.v, off);
    }
    return r;
}
template <typename T, int KIND> __device__ __forceinline__ void warp_merge(Acc<T, KIND> &a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const Acc<T, KIND> o = shfl_down_acc(a, off); a.merge(o); }
}

template <typename T, int KIND>
__device__ __forceinline__ void finalize(const Acc<T, KIND> &acc, T *out, long long o, long long nout, int fin, T divisor, int *flag) {
    if constexpr (KIND == K_WELFORD) {
        if (fin == FIN_MEAN_M2) { out[o] = acc.mean; out[nout + o] = acc.m2; }
        else out[o] = acc.m2 / divisor;
    } else if constexpr (KIND == K_SUMSQ) {
        if (fin == FIN_MEAN_M2) { out[o] = acc.v; out[nout + o] = acc.q; return; }  // partial: S1 | S2
        const T m = w_div(acc.v, divisor, flag);
        T v = w_sub(acc.q, w_mul(w_mul((T)2, m), acc.v));
        v = w_add(v, w_mul(divisor, w_mul(m, m)));
        out[o] = w_div(v, divisor, flag);
    } else {
        T v = acc.v;
        if (fin == FIN_DIV) v = w_div(v, divisor, flag);
        out[o] = v;
    }
}
// L2-coherent read of a partial accumulator written by another CTA of the same launch.
template <typename T, int KIND> __device__ __forceinline__ Acc<T, KIND> load_partial(const void *partial, long long idx) {
    static_assert(sizeof(Acc<T, KIND>) % 4 == 0, "accumulators are made of 32-bit words");
    Acc<T, KIND> r;
    const unsigned int *src = reinterpret_cast<const unsigned int *>(reinterpret_cast<const Acc<T, KIND> *>(partial) + idx);
    unsigned int w[sizeof(Acc<T, KIND>) / 4];
#pragma unroll
    for (int i = 0; i < (int)(sizeof(Acc<T, KIND>) / 4); ++i) w[i] = __ldcg(src + i);
    __builtin_memcpy(&r, w, sizeof r);
    return r;
}

// The last arriver's merge, out of line so the rarely taken path does not inflate the streaming kernels' registers.
template <typename T, int KIND>
__device__ __noinline__ void combine_thread(T *out, const void *partial, long long o, long long nout, long long S, int fin, T divisor,
                                            int *flag, unsigned int *counters) {
    __threadfence();
    Acc<T, KIND> m = load_partial<T, KIND>(partial, o);
    constexpr int B = 8;  // partials fetched per round: independent loads in flight, merged in split order
    for (long long k = 1; k < S; k += B) {
        Acc<T, KIND> buf[B];
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (k + i < S) buf[i] = load_partial<T, KIND>(partial, (k + i) * nout + o);
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (k + i < S) m.merge(buf[i]);
    }
    finalize(m, out, o, nout, fin, divisor, flag);
    counters[o] = 0;  // ready for the next launch on this stream
}

// Result of one (output, split) pair, thread level (cols / micro kernels).  S == 1: final value.  S > 1: the accumulator
// goes to the workspace; with `counters` the LAST of the S writers of this output merges all S partials in split order
// (so the result does not depend on who is last) and finalises -- one launch instead of two, no atomics on data.
// The S > 1 path is out of line (by-value accumulator) so that it does not inflate the streaming kernels' registers.
template <typename T, int KIND>
__device__ __noinline__ void publish_partial(Acc<T, KIND> acc, T *out, void *partial, long long o, long long s, long long nout, long long S,
                                             int fin, T divisor, int *flag, unsigned int *counters) {
    reinterpret_cast<Acc<T, KIND> *>(partial)[s * nout + o] = acc;
    if (!counters) return;  // two-stage mode: reduce_stage2_kernel combines
    __threadfence();
    const unsigned int ticket = atomicAdd(&counters[o], 1u);
    if (ticket != (unsigned int)(S - 1)) return;
    combine_thread<T, KIND>(out, partial, o, nout, S, fin, divisor, flag, counters);
}
template <typename T, int KIND>
__device__ __forceinline__ void emit(const Acc<T, KIND> &acc, T *out, void *partial, long long o, long long s, long long nout,
                                     long long S, int fin, T divisor, int *flag, unsigned int *counters) {
    if (S <= 1) finalize(acc, out, o, nout, fin, divisor, flag);
    else publish_partial<T, KIND>(acc, out, partial, o, s, nout, S, fin, divisor, flag, counters);
}

// Warp-level merge of the S partials of output o (same order as reduce_stage2_kernel<WARP>): every lane of the calling
// warp takes a contiguous run of splits, then the fixed shuffle tree.
template <typename T, int KIND>
__device__ __noinline__ void warp_combine(T *out, const void *partial, long long o, long long nout, long long S, int fin, T divisor,
                                             int *flag, unsigned int *counters, int lane) {
    Acc<T, KIND> m;
    m.init((T)0);
    if constexpr (KIND == K_SQDEV) m.m = load_partial<T, KIND>(partial, o).m;
    const long long per = (S + 31) / 32;
    const long long lo = lane * per, hi = (lo + per < S) ? lo + per : S;
    constexpr int B = 4;
    for (long long k = lo; k < hi; k += B) {
        Acc<T, KIND> buf[B];
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (k + i < hi) buf[i] = load_partial<T, KIND>(partial, (k + i) * nout + o);
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (k + i < hi) m.merge(buf[i]);
    }
    warp_merge(m);
    if (lane == 0) { finalize(m, out, o, nout, fin, divisor, flag); counters[o] = 0; }
}

// ---- the exchange of a sharded reduction as a device function (csrc/peer.cu's exchange kernel, and the tail of the rows
// kernel below): element i of the result = epilogue(partial_0[i] (+) partial_1[i] (+) ...) in rank order; V elements per access.
enum { X_SUM = 0, X_MEAN = 1, X_VAR = 2, X_MAX = 3, X_MIN = 4 };
template <typename T, int XK, int V>
__device__ __forceinline__ void xchg_combine(const feo_peer::XchgParams &p, T *__restrict__ out, T divisor, T count, int *arith_flag,
                                             long long first, long long stride) {
    using feo_peer::kMaxPeers;
    using feo_peer::ld_peer;
    using Vec = AVec<T, V>;
    const long long nv = p.n / V;
    for (long long i = first; i < nv; i += stride) {
        Vec part[kMaxPeers], part2[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            if (r < p.world) {
                const Vec *src = reinterpret_cast<const Vec *>(p.win[r] + p.slot_off) + i;
                part[r] = ld_peer(src);
                if constexpr (XK == X_VAR) part2[r] = ld_peer(src + nv);  // M2 follows the n means
            }
        }
        Vec res;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if constexpr (XK == X_VAR && std::is_floating_point<T>::value) {
                Acc<T, K_WELFORD> acc;
                acc.init((T)0);
#pragma unroll
                for (int r = 0; r < kMaxPeers; ++r)
                    if (r < p.world) acc.merge_raw(count, part[r].v[j], part2[r].v[j]);
                res.v[j] = acc.m2 / divisor;
            } else if constexpr (XK == X_VAR) {  // integers: wrapping S1 | S2 partials, exact by the ring identity (K_SUMSQ)
                T s1 = (T)0, s2 = (T)0;
#pragma unroll
                for (int r = 0; r < kMaxPeers; ++r)
                    if (r < p.world) { s1 = w_add(s1, part[r].v[j]); s2 = w_add(s2, part2[r].v[j]); }
                const T m = w_div(s1, divisor, arith_flag);
                T v = w_sub(s2, w_mul(w_mul((T)2, m), s1));
                v = w_add(v, w_mul(divisor, w_mul(m, m)));
                res.v[j] = w_div(v, divisor, arith_flag);
            } else {
                T a = part[0].v[j];
#pragma unroll
                for (int r = 1; r < kMaxPeers; ++r)
                    if (r < p.world) {
                        if constexpr (XK == X_MAX) a = feo_max2(a, part[r].v[j]);
                        else if constexpr (XK == X_MIN) a = feo_min2(a, part[r].v[j]);
                        else a = w_add(a, part[r].v[j]);
                    }
                if constexpr (XK == X_MEAN) a = w_div(a, divisor, arith_flag);
                res.v[j] = a;
            }
        }
        reinterpret_cast<Vec *>(out)[i] = res;
    }
}

// Exchange at the tail of a reduction kernel (see feo_peer::XTailReq): plain data, so that Geom stays a kernel parameter.
struct XTail {
    int on, mean;
    feo_peer::XchgParams p;
    void *out;
    unsigned long long divisor_bits, count_bits;
    unsigned int *ticket;
};
template <typename T> __device__ __forceinline__ T xtail_value(unsigned long long bits) {
    T v;
    __builtin_memcpy(&v, &bits, sizeof(T));
    return v;
}
// Called by EVERY thread of EVERY CTA at the end of the kernel, after this CTA's outputs have been written: the CTA that
// takes the last ticket sees all of them (fence + atomic, the usual last-block pattern) and runs the whole exchange.
template <typename T, int KIND>
__device__ __forceinline__ void run_xtail(const XTail &xt, int *arith_flag) {
    if constexpr (KIND == K_SUM || KIND == K_MAX || KIND == K_MIN || KIND == K_WELFORD || KIND == K_SUMSQ) {
        __shared__ unsigned int s_ticket;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(xt.ticket, 1u);
        __syncthreads();
        if (s_ticket != gridDim.x - 1) return;
        if (threadIdx.x == 0) *xt.ticket = 0;  // ready for the next launch on this stream
        __threadfence();
        feo_peer::signal_and_wait(xt.p, true);
        T *out = static_cast<T *>(xt.out);
        const T divisor = xtail_value<T>(xt.divisor_bits), count = xtail_value<T>(xt.count_bits);
        if constexpr (KIND == K_SUM) {
            if (xt.mean) xchg_combine<T, X_MEAN, 1>(xt.p, out, divisor, count, arith_flag, threadIdx.x, blockDim.x);
            else xchg_combine<T, X_SUM, 1>(xt.p, out, divisor, count, arith_flag, threadIdx.x, blockDim.x);
        } else if constexpr (KIND == K_MAX) {
            xchg_combine<T, X_MAX, 1>(xt.p, out, divisor, count, arith_flag, threadIdx.x, blockDim.x);
        } else if constexpr (KIND == K_MIN) {
            xchg_combine<T, X_MIN, 1>(xt.p, out, divisor, count, arith_flag, threadIdx.x, blockDim.x);
        } else {
            xchg_combine<T, X_VAR, 1>(xt.p, out, divisor, count, arith_flag, threadIdx.x, blockDim.x);
        }
    }
}

struct Geom {
    long long K1, R1, K2, R2;  // canonical extents
    long long S, S1, S2;       // splits: S = S1 (over r1) * S2 (over r2 vectors)
    long long c1, c2;          // chunk sizes: r1 rows per split, r2 vectors per split
    long long nout;            // K1 * K2
    long long xblocks;         // blocks along the thread-mapped axis
    int fin;
    int lanes;                 // rows kernel: 32 (warp per output) or 256 (block per output)
    int lx;                    // rows kernel: lanes along the row (power of two <= lanes); lanes / lx rows at a time
    unsigned int *counters;    // per-output arrival counters for the fused combine (nullptr: two-stage)
    XTail xt;                  // xt.on: the last CTA of the (rows) kernel runs the exchange of the sharded reduction
};

// ---- cols: in[k1][r1][k2] -> out[k1][k2], threads along k2 ---------------------------------------------
// J column-vectors per thread (kThreads apart): J = 1 with 8 rows in flight is the streaming shape; J = 4 is used when
// there are fewer than 8 rows in all ([2, n] or [3, n] over axis 0: the loads in flight come from the columns instead).
template <typename T, int KIND, int LOADER, bool VEC, int J>
__global__ void __launch_bounds__(kThreads, (J > 1 && VEC && KIND == K_WELFORD) ? 2 : 3) reduce_cols_kernel(T *__restrict__ out, void *__restrict__ partial,
                                                               const T *__restrict__ in, const T *__restrict__ in2,
                                                               const T *__restrict__ aux, const __grid_constant__ Geom G,
                                                               T divisor, int *flag) {
    constexpr int W = VEC ? VecN<T>::value : 1;
    constexpr int U = (VEC || J > 1) ? 8 : 16;  // 4-byte loads: twice the rows in flight, half the Welford merges per element
    using V = AVec<T, W>;
    long long blk = blockIdx.x;
    const long long bx = blk % G.xblocks; blk /= G.xblocks;
    const long long s = blk % G.S;
    const long long k1 = blk / G.S;
    const long long kv = bx * (kThreads * J) + threadIdx.x;
    if (kv * W >= G.K2) return;
    const long long r_lo = s * G.c1;
    const long long r_hi = (r_lo + G.c1 < G.R1) ? r_lo + G.c1 : G.R1;
    const long long o0 = k1 * G.K2 + kv * W;
    const T *base = in + (k1 * G.R1) * G.K2 + kv * W;
    Acc<T, KIND> acc[J][W];
    bool live[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        live[j] = (kv + (long long)j * kThreads) * W < G.K2;
#pragma unroll
        for (int k = 0; k < W; ++k) acc[j][k].init(live[j] ? aux_val<KIND>(aux, o0 + (long long)j * kThreads * W + k) : (T)0);
    }
    long long r = r_lo;
    // J > 1 is only planned for fewer than 8 rows in all: the U-row loop would never run there, and compiling it in cost the
    // few-row kernels 0.8-1.3 KB of spills (80-register cap)
    for (; J == 1 && r + U <= r_hi; r += U) {
        V v[J][U];
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (live[j]) {
#pragma unroll
                for (int u = 0; u < U; ++u) v[j][u] = ld_stream(reinterpret_cast<const V *>(base + (r + u) * G.K2) + (long long)j * kThreads);
            }
        if constexpr (LOADER == LD_VECMAT) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const T w = __ldg(in2 + r + u);
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int k = 0; k < W; ++k) v[j][u].v[k] = w_mul(w, v[j][u].v[k]);
            }
        }
        T wf = (T)0;  // Welford: every accumulator of this thread holds the same count, so one merge weight serves them all
        if constexpr (KIND == K_WELFORD) {
            const T n0 = acc[0][0].n;
            if (n0 == (T)U) wf = (T)0.5;            // second batch: exactly one half, no division
            else if (n0 != (T)0) wf = (T)U / (n0 + (T)U);
        }
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (live[j]) {
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    T x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u] = v[j][u].v[k];
                    if constexpr (KIND == K_WELFORD) acc[j][k].push_batch_f(x, wf);
                    else acc[j][k].push_batch(x);
                }
            }
    }
    // fewer than U rows left (or, with few rows in all, the whole job): TU rows x J columns of loads go out together
    constexpr int TU = (J == 1 || !VEC) ? 4 : 2;
    for (; r < r_hi; r += TU) {
        const int nrem = (int)((r_hi - r < TU) ? r_hi - r : TU);
        V v[J][TU];
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (live[j]) {
#pragma unroll
                for (int u = 0; u < TU; ++u)
                    if (u < nrem) v[j][u] = ld_stream(reinterpret_cast<const V *>(base + (r + u) * G.K2) + (long long)j * kThreads);
            }
        T w[TU];
#pragma unroll
        for (int u = 0; u < TU; ++u) {
            w[u] = (T)1;
            if constexpr (LOADER == LD_VECMAT) { if (u < nrem) w[u] = __ldg(in2 + r + u); }
        }
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (live[j]) {
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    T x[TU];
#pragma unroll
                    for (int u = 0; u < TU; ++u) x[u] = (u < nrem) ? (LOADER == LD_VECMAT ? w_mul(w[u], v[j][u].v[k]) : v[j][u].v[k]) : (T)0;
                    if constexpr (KIND == K_WELFORD) {
                        acc[j][k].push_masked(x, nrem);  // one division per chunk, not one per element
                    } else {
#pragma unroll
                        for (int u = 0; u < TU; ++u)
                            if (u < nrem) acc[j][k].push(x[u]);
                    }
                }
            }
    }
    if (VEC && G.S <= 1 && G.fin != FIN_MEAN_M2 && KIND != K_SUMSQ) {
        // unsplit: one 128-bit store per column-vector (with few rows the output is as large as the input, and W scalar
        // stores per thread, 16 bytes apart across the warp, quarter-fill every sector they touch)
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (live[j]) {
                V r;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    if constexpr (KIND == K_WELFORD) r.v[k] = acc[j][k].m2 / divisor;
                    else r.v[k] = (G.fin == FIN_DIV) ? w_div(acc[j][k].v, divisor, flag) : acc[j][k].v;
                }
                *reinterpret_cast<V *>(out + o0 + (long long)j * kThreads * W) = r;
            }
        return;
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
        if (live[j]) {
#pragma unroll
            for (int k = 0; k < W; ++k)
                emit(acc[j][k], out, partial, o0 + (long long)j * kThreads * W + k, s, G.nout, G.S, G.fin, divisor, flag, G.counters);
        }
}

// ---- cols2d: in[k1][r1][k2] with a SHORT k2 row (at most 128 vectors: [n, 3] points, [.., 4096, 16] ...) ---------------
// The cols kernel would leave most of a block idle (k2 = 16: 4 of 256 threads).  Here a block owns one (k1, split): its
// threads form an lx x ly grid, lx = k2 / W lanes across the row (any value, not only powers of two), ly = 256 / lx rows
// walked concurrently with 8 sweeps in flight; so a warp still reads whole contiguous rows.  The ly partial rows are
// combined through shared memory in a fixed tree.
template <typename T, int KIND, int LOADER, bool VEC>
__global__ void __launch_bounds__(kThreads, 3) reduce_cols2d_kernel(T *__restrict__ out, void *__restrict__ partial,
                                                                 const T *__restrict__ in, const T *__restrict__ in2,
                                                                 const T *__restrict__ aux, const __grid_constant__ Geom G,
                                                                 T divisor, int *flag) {
    constexpr int W = VEC ? VecN<T>::value : 1;
    constexpr int U = VEC ? 8 : 16;  // 4-byte loads: 16 in flight per thread (ncu: 8 left the kernel latency-bound at 60 % of DRAM peak)
    using V = AVec<T, W>;
    __shared__ Acc<T, KIND> smem[W][kThreads];
    // lx x ly threads per k1; when the rows are few and short, several k1 share a block (G.c2 of them)
    const int lx = G.lx, ly = G.lanes, per = lx * ly, kz = (int)G.c2;
    const int tz = threadIdx.x / per, tin = threadIdx.x % per;
    const int tx = tin % lx, ty = tin / lx;
    const long long s = blockIdx.x % G.S, k1 = (blockIdx.x / G.S) * kz + tz;
    const bool active = tz < kz && k1 < G.K1;
    const long long r_lo = s * G.c1;
    const long long r_hi = (r_lo + G.c1 < G.R1) ? r_lo + G.c1 : G.R1;
    const long long o0 = k1 * G.K2 + (long long)tx * W;
    Acc<T, KIND> acc[W];
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k].init(active ? aux_val<KIND>(aux, o0 + k) : (T)0);
    if (active) {
        const T *base = in + (k1 * G.R1) * G.K2 + (long long)tx * W;
        long long r = r_lo + ty;
        for (; r + (long long)(U - 1) * ly < r_hi; r += (long long)U * ly) {
            V v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const V *>(base + (r + (long long)u * ly) * G.K2));
            if constexpr (LOADER == LD_VECMAT) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const T w = __ldg(in2 + r + (long long)u * ly);
#pragma unroll
                    for (int k = 0; k < W; ++k) v[u].v[k] = w_mul(w, v[u].v[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < W; ++k) {
                T x[U];
#pragma unroll
                for (int u = 0; u < U; ++u) x[u] = v[u].v[k];
                acc[k].push_batch(x);
            }
        }
        for (; r < r_hi; r += ly) {
            V v = ld_stream(reinterpret_cast<const V *>(base + r * G.K2));
            if constexpr (LOADER == LD_VECMAT) {
                const T w = __ldg(in2 + r);
#pragma unroll
                for (int k = 0; k < W; ++k) v.v[k] = w_mul(w, v.v[k]);
            }
#pragma unroll
            for (int k = 0; k < W; ++k) acc[k].push(v.v[k]);
        }
    }
    // Rows that sit in the same warp (lanes lx apart) meet through shuffles first: for short rows that is most of the tree
    // (k2 = 16: 3 of 6 levels) without a barrier each.  Needs whole warps per k1 group and power-of-two lx, ly.
    int rpw = 1;  // rows already merged per warp
    if ((lx & (lx - 1)) == 0 && lx <= 16 && (ly & (ly - 1)) == 0 && per >= 32) {
        rpw = 32 / lx;
        for (int off = 16; off >= lx; off >>= 1) {
#pragma unroll
            for (int k = 0; k < W; ++k) { const Acc<T, KIND> o = shfl_down_acc(acc[k], off); acc[k].merge(o); }
        }
    }
#pragma unroll
    for (int k = 0; k < W; ++k) smem[k][threadIdx.x] = acc[k];
    __syncthreads();
    int st = 1;
    while (st * 2 < ly) st *= 2;  // largest power of two below ly (ly >= 2)
    for (; st >= rpw; st >>= 1) {
        if (active && (ty & (rpw - 1)) == 0 && ty < st && ty + st < ly) {
#pragma unroll
            for (int k = 0; k < W; ++k) smem[k][threadIdx.x].merge(smem[k][threadIdx.x + st * lx]);
        }
        __syncthreads();
    }
    if (active && ty == 0) {
#pragma unroll
        for (int k = 0; k < W; ++k) emit(smem[k][threadIdx.x], out, partial, o0 + k, s, G.nout, G.S, G.fin, divisor, flag, G.counters);
    }
}

// ---- micro: in[k1][r1][k2][r2] with 2 <= r2 <= 32 ----------------------------------------------------------
// R2C > 0: compile-time r2.  A thread owns OUTS outputs (k2, k2 + 256, ...) and reads each one's r2 contiguous
// elements with 256-bit loads (one full 32-byte sector per access), OUTS x U of them in flight per step.
// R2C == 0: run-time r2, one output per thread, scalar loads.
// WIDE (r1 == 1: a single row per output): no row loop to unroll, so more outputs per thread keep ~256 B in flight.
template <int R2C, bool WIDE> struct MicroShape {
    static constexpr int OUTS = (R2C == 0) ? 4 : (WIDE ? ((32 / R2C) > 8 ? 8 : ((32 / R2C) < 2 ? 2 : (32 / R2C))) : (R2C <= 16 ? 4 : 2));
    static constexpr int U = (R2C == 0 || WIDE) ? 1 : ((64 / (OUTS * R2C)) > 8 ? 8 : ((64 / (OUTS * R2C)) < 1 ? 1 : (64 / (OUTS * R2C))));
};
template <typename T, int KIND, int R2C, bool WIDE>
__global__ void __launch_bounds__(kThreads) reduce_micro_kernel(T *__restrict__ out, void *__restrict__ partial,
                                                                const T *__restrict__ in, const T *__restrict__ aux,
                                                                const __grid_constant__ Geom G, T divisor, int *flag) {
    constexpr int OUTS = MicroShape<R2C, WIDE>::OUTS;
    long long blk = blockIdx.x;
    const long long bx = blk % G.xblocks; blk /= G.xblocks;
    const long long s = blk % G.S;
    const long long k1 = blk / G.S;
    const long long k2_0 = bx * (kThreads * OUTS) + threadIdx.x;
    if (k2_0 >= G.K2) return;
    const long long r_lo = s * G.c1;
    const long long r_hi = (r_lo + G.c1 < G.R1) ? r_lo + G.c1 : G.R1;
    const long long rstride = G.K2 * G.R2;
    const T *base = in + ((k1 * G.R1) * G.K2 + k2_0) * G.R2;
    Acc<T, KIND> acc[OUTS];
    bool live[OUTS];
#pragma unroll
    for (int j = 0; j < OUTS; ++j) {
        live[j] = (k2_0 + (long long)j * kThreads) < G.K2;
        acc[j].init(live[j] ? aux_val<KIND>(aux, k1 * G.K2 + k2_0 + (long long)j * kThreads) : (T)0);
    }
    if constexpr (R2C > 0) {
        constexpr int VN = VecW<T, 32>::value;  // 256-bit loads: every access is exactly one 32-byte sector or more
        constexpr int W = R2C < VN ? R2C : VN;
        constexpr int NV = R2C / W;
        constexpr int U = MicroShape<R2C, WIDE>::U;
        using V = AVec<T, W>;
        long long r = r_lo;
        if constexpr (WIDE) {  // r1 == 1: one row per output, straight-line code (the accumulator folds to one batch)
            V v[OUTS][NV];
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) v[j][i] = ld_stream(reinterpret_cast<const V *>(base + (long long)j * kThreads * G.R2) + i);
                }
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
                    T x[R2C];
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int k = 0; k < W; ++k) x[i * W + k] = v[j][i].v[k];
                    Acc<T, KIND> a1;
                    a1.init(KIND == K_INTVAR ? divisor : aux_val<KIND>(aux, k1 * G.K2 + k2_0 + (long long)j * kThreads));
                    a1.push_batch(x);
                    finalize(a1, out, k1 * G.K2 + k2_0 + (long long)j * kThreads, G.nout, G.fin, divisor, flag);  // r1 == 1: never split
                }
            return;
        } else {
        for (; r + U <= r_hi; r += U) {
            V v[OUTS][U][NV];
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int i = 0; i < NV; ++i)
                            v[j][u][i] = ld_stream(reinterpret_cast<const V *>(base + (long long)j * kThreads * G.R2 + (r + u) * rstride) + i);
                }
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        T x[R2C];
#pragma unroll
                        for (int i = 0; i < NV; ++i)
#pragma unroll
                            for (int k = 0; k < W; ++k) x[i * W + k] = v[j][u][i].v[k];
                        acc[j].push_batch(x);
                    }
                }
        }
        for (; r < r_hi; ++r) {
            V v[OUTS][NV];
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) v[j][i] = ld_stream(reinterpret_cast<const V *>(base + (long long)j * kThreads * G.R2 + r * rstride) + i);
                }
#pragma unroll
            for (int j = 0; j < OUTS; ++j)
                if (live[j]) {
                    T x[R2C];
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int k = 0; k < W; ++k) x[i * W + k] = v[j][i].v[k];
                    acc[j].push_batch(x);
                }
        }
        }
    } else {
        // run-time r2 (3, 5, 6 ... or an unaligned base): a thread still owns whole outputs; its 4-byte loads go through
        // L1, which turns the warp's stride-r2 accesses back into full sectors.  Groups of 8 loads are issued together.
        const int r2 = (int)G.R2;
        for (long long r = r_lo; r < r_hi; ++r) {
            const T *p = base + r * rstride;
            for (int j0 = 0; j0 < r2; j0 += 8) {
                T x[OUTS][8];
#pragma unroll
                for (int o = 0; o < OUTS; ++o)
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[o][j] = (live[o] && j0 + j < r2) ? __ldg(p + (long long)o * kThreads * G.R2 + j0 + j) : (T)0;
                const int cnt = (r2 - j0 < 8) ? r2 - j0 : 8;
#pragma unroll
                for (int o = 0; o < OUTS; ++o) {
                    if constexpr (KIND == K_WELFORD) {
                        acc[o].push_masked(x[o], cnt);  // masked lanes hold 0: one division per chunk instead of one per element
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < cnt) acc[o].push(x[o][j]);
                    }
                }
            }
        }
    }
    if constexpr (!(R2C > 0 && WIDE)) {  // the WIDE path has emitted and returned above
#pragma unroll
        for (int j = 0; j < OUTS; ++j)
            if (live[j]) emit(acc[j], out, partial, k1 * G.K2 + k2_0 + (long long)j * kThreads, s, G.nout, G.S, G.fin, divisor, flag, G.counters);
    }
}

// ---- microsmem: in[k][r2] -> out[k], run-time r2 <= 32 ([n, 3] points, [.., 6], [.., 31]), one row per output --------------
// A thread that owns whole outputs reads r2 four-byte words 4 * r2 bytes apart from its neighbours': 12 scalar loads per
// 48 bytes (measured 3.4 TB/s on [n, 3]).  Here the block copies its contiguous span of the input into shared memory with
// 128-bit loads (the span starts 16-byte aligned because a block owns a multiple of 32 outputs), then every thread reduces
// its outputs from shared memory (stride r2 words: conflict-free for odd r2).  Variance is the exact two-pass form.
constexpr int kMicroSmemBytes = 32768;
template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads) reduce_microsmem_kernel(T *__restrict__ out, const T *__restrict__ in, const T *__restrict__ aux,
                                                                   long long nout, int r2, int ob, int fin, T divisor, int *flag) {
    constexpr int VN = VecN<T>::value;
    using V = AVec<T, VN>;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    T *tile = reinterpret_cast<T *>(tile_raw);
    const long long o_first = (long long)blockIdx.x * ob;
    const int outs = (int)((nout - o_first < ob) ? nout - o_first : ob);
    const int elems = outs * r2;
    const T *src = in + o_first * r2;
    const int nvec = elems / VN;
    constexpr int U = 4;
    int v = threadIdx.x;
    for (; v + (U - 1) * kThreads < nvec; v += U * kThreads) {
        V t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = ld_stream(reinterpret_cast<const V *>(src) + v + u * kThreads);
#pragma unroll
        for (int u = 0; u < U; ++u) reinterpret_cast<V *>(tile)[v + u * kThreads] = t[u];
    }
    for (; v < nvec; v += kThreads) reinterpret_cast<V *>(tile)[v] = ld_stream(reinterpret_cast<const V *>(src) + v);
    for (int e = nvec * VN + threadIdx.x; e < elems; e += kThreads) tile[e] = __ldg(src + e);
    __syncthreads();
    for (int o = threadIdx.x; o < outs; o += kThreads) {
        const T *g = tile + o * r2;
        const long long oo = o_first + o;
        Acc<T, KIND> acc;
        if constexpr (KIND == K_WELFORD) {
            T sum = (T)0;
            for (int j = 0; j < r2; ++j) sum += g[j];
            const T mean = sum / (T)r2;
            T q = (T)0;
            for (int j = 0; j < r2; ++j) { const T c = g[j] - mean; q += c * c; }
            acc.n = (T)r2; acc.mean = mean; acc.m2 = q;
        } else if constexpr (KIND == K_INTVAR) {
            T sum = (T)0;
            for (int j = 0; j < r2; ++j) sum = w_add(sum, g[j]);
            const T mean = w_div(sum, divisor, (int *)nullptr);
            T q = (T)0;
            for (int j = 0; j < r2; ++j) { const T c = w_sub(g[j], mean); q = w_add(q, w_mul(c, c)); }
            acc.v = q; acc.n = divisor;
        } else {
            acc.init(aux_val<KIND>(aux, oo));
            for (int j = 0; j < r2; ++j) acc.push(g[j]);
        }
        finalize(acc, out, oo, nout, fin, divisor, flag);
    }
}

// ---- rows: in[k1][r1][k2][r2], long r2; `lanes` threads cooperate on one output ----------------------------
template <typename T, int KIND, int LOADER, bool VEC>
__global__ void __launch_bounds__(kThreads, 4) reduce_rows_kernel(T *__restrict__ out, void *__restrict__ partial,
                                                               const T *__restrict__ in, const T *__restrict__ in2,
                                                               const T *__restrict__ aux, const __grid_constant__ Geom G,
                                                               T divisor, int *flag) {
    constexpr int W = VEC ? VecN<T>::value : 1;
    constexpr int U = VEC ? 4 : 16;  // scalar path (odd / unaligned rows): 16 four-byte loads in flight per lane
    using V = AVec<T, W>;
    __shared__ Acc<T, KIND> smem[kThreads / 32];
    const int lanes = G.lanes;
    const int per_block = kThreads / lanes;  // outputs (x splits) handled by one block
    const int sub = threadIdx.x / lanes;     // which of them this thread works on
    const int lane = threadIdx.x % lanes;
    const long long unit = (long long)blockIdx.x * per_block + sub;  // = o * S + s
    const bool active = unit < G.nout * G.S;
    Acc<T, KIND> acc;
    long long o = 0, s = 0;
    if (active) {
        o = unit / G.S; s = unit % G.S;
        const long long k1 = o / G.K2, k2 = o % G.K2;
        const long long s1 = s / G.S2, s2 = s % G.S2;
        const long long r_lo = s1 * G.c1, r_hi = (r_lo + G.c1 < G.R1) ? r_lo + G.c1 : G.R1;
        const long long nvec = G.R2 / W;
        const long long c_lo = s2 * G.c2, c_hi = (c_lo + G.c2 < nvec) ? c_lo + G.c2 : nvec;
        const long long rstride = G.K2 * G.R2;
        const long long obase = ((k1 * G.R1) * G.K2 + k2) * G.R2;
        acc.init(aux_val<KIND>(aux, o));
        // lanes form an lx x ly grid: lx lanes stride the row, ly rows are walked concurrently (short rows
        // would otherwise leave most of the lanes idle)
        const int lx = G.lx, lxi = lane % lx, lyi = lane / lx, ly = lanes / lx;
        for (long long r = r_lo + lyi; r < r_hi; r += ly) {
            const T *row = in + obase + r * rstride;
            const T *row2 = nullptr;
            if constexpr (LOADER == LD_MUL) row2 = in2 + obase + r * rstride;
            if constexpr (LOADER == LD_MATVEC) row2 = in2;
            if constexpr (!VEC && LOADER == LD_PLAIN) {
                // odd or unaligned rows (e.g. 4099 columns): peel each row segment to a 16-byte boundary and read its body
                // with 128-bit loads like the aligned path; the <= 2 * (VN - 1) edge elements go through scalar loads
                constexpr int VNp = VecN<T>::value;
                constexpr int UP = 4;
                using VP = AVec<T, VNp>;
                const T *seg = row + c_lo;
                const long long len = c_hi - c_lo;
                const int mis = (int)((reinterpret_cast<uintptr_t>(seg) / sizeof(T)) % VNp);
                long long head = mis ? VNp - mis : 0;
                if (head > len) head = len;
                for (long long e = lxi; e < head; e += lx) acc.push(__ldg(seg + e));
                const VP *body = reinterpret_cast<const VP *>(seg + head);
                const long long nb = (len - head) / VNp;
                long long bv = lxi;
                for (; bv + (long long)(UP - 1) * lx < nb; bv += (long long)UP * lx) {
                    VP v[UP];
#pragma unroll
                    for (int u = 0; u < UP; ++u) v[u] = ld_stream(body + bv + (long long)u * lx);
                    T x[UP * VNp];
#pragma unroll
                    for (int u = 0; u < UP; ++u)
#pragma unroll
                        for (int k = 0; k < VNp; ++k) x[u * VNp + k] = v[u].v[k];
                    acc.push_batch(x);
                }
                for (; bv < nb; bv += lx) {
                    const VP v = ld_stream(body + bv);
#pragma unroll
                    for (int k = 0; k < VNp; ++k) acc.push(v.v[k]);
                }
                for (long long e = head + nb * VNp + lxi; e < len; e += lx) acc.push(__ldg(seg + e));
                continue;
            }
            long long cv = c_lo + lxi;
            for (; cv + (long long)(U - 1) * lx < c_hi; cv += (long long)U * lx) {
                V v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const V *>(row) + cv + (long long)u * lx);
                if constexpr (LOADER == LD_MUL || LOADER == LD_MATVEC) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const V w = (LOADER == LD_MUL) ? ld_stream(reinterpret_cast<const V *>(row2) + cv + (long long)u * lx)
                                                       : ld_keep(reinterpret_cast<const V *>(row2) + cv + (long long)u * lx);
#pragma unroll
                        for (int k = 0; k < W; ++k) v[u].v[k] = w_mul(v[u].v[k], w.v[k]);
                    }
                }
                T x[U * W];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < W; ++k) x[u * W + k] = v[u].v[k];
                acc.push_batch(x);
            }
            for (; cv < c_hi; cv += lx) {
                V v = ld_stream(reinterpret_cast<const V *>(row) + cv);
                if constexpr (LOADER == LD_MUL || LOADER == LD_MATVEC) {
                    const V w = ld_keep(reinterpret_cast<const V *>(row2) + cv);
#pragma unroll
                    for (int k = 0; k < W; ++k) v.v[k] = w_mul(v.v[k], w.v[k]);
                }
#pragma unroll
                for (int k = 0; k < W; ++k) acc.push(v.v[k]);
            }
        }
    } else {
        acc.init((T)0);
    }
    warp_merge(acc);
    const bool fused = (G.S > 1) && G.counters != nullptr;
    if (lanes == 32) {  // a warp per (output, split)
        unsigned int last = 0;
        if (active && lane == 0) {
            emit(acc, out, partial, o, s, G.nout, G.S, G.fin, divisor, flag, (unsigned int *)nullptr);
            if (fused) {
                __threadfence();
                last = (atomicAdd(&G.counters[o], 1u) == (unsigned int)(G.S - 1));
            }
        }
        if (fused) {
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                warp_combine<T, KIND>(out, partial, o, G.nout, G.S, G.fin, divisor, flag, G.counters, lane);
            }
        }
    } else {
        // block per (output, split): merge the eight warp results in warp order
        __shared__ unsigned int s_last;
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) smem[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int last = 0;
            if (active) {
                Acc<T, KIND> t = smem[0];
#pragma unroll
                for (int w = 1; w < kThreads / 32; ++w) t.merge(smem[w]);
                emit(t, out, partial, o, s, G.nout, G.S, G.fin, divisor, flag, (unsigned int *)nullptr);
                if (fused) {
                    __threadfence();
                    last = (atomicAdd(&G.counters[o], 1u) == (unsigned int)(G.S - 1));
                }
            }
            s_last = last;
        }
        if (fused) {
            __syncthreads();
            if (s_last && warp == 0) {
                __threadfence();
                warp_combine<T, KIND>(out, partial, o, G.nout, G.S, G.fin, divisor, flag, G.counters, threadIdx.x & 31);
            }
        }
    }
    // sharded reduction with the exchange at the tail: every output of this launch is final now (un-split, or merged by the
    // last arriver above); the last CTA to get here signals the peers, waits, pulls and combines
    if constexpr (LOADER == LD_PLAIN) {
        if (G.xt.on) run_xtail<T, KIND>(G.xt, flag);
    }
}

// ---- generic: any number of groups, a block per output ------------------------------------------------------
struct GenericGeom {
    int nk, nr;
    long long kext[FEO_MAX_RANK], kstr[FEO_MAX_RANK], rext[FEO_MAX_RANK], rstr[FEO_MAX_RANK];
    long long nout, nred;
    int fin;
};
template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads) reduce_generic_kernel(T *__restrict__ out, const T *__restrict__ in,
                                                                  const T *__restrict__ aux, const __grid_constant__ GenericGeom G,
                                                                  T divisor, int *flag) {
    __shared__ Acc<T, KIND> smem[kThreads / 32];
    for (long long o = blockIdx.x; o < G.nout; o += gridDim.x) {
        long long rem = o, base = 0;
#pragma unroll 1
        for (int d = G.nk - 1; d >= 0; --d) { base += (rem % G.kext[d]) * G.kstr[d]; rem /= G.kext[d]; }
        Acc<T, KIND> acc;
        acc.init(aux_val<KIND>(aux, o));
        for (long long i = threadIdx.x; i < G.nred; i += kThreads) {
            long long q = i, off = base;
#pragma unroll 1
            for (int d = G.nr - 1; d >= 0; --d) { off += (q % G.rext[d]) * G.rstr[d]; q /= G.rext[d]; }
            acc.push(__ldg(in + off));
        }
        warp_merge(acc);
        if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            Acc<T, KIND> t = smem[0];
#pragma unroll
            for (int w = 1; w < kThreads / 32; ++w) t.merge(smem[w]);
            finalize(t, out, o, G.nout, G.fin, divisor, flag);
        }
        __syncthreads();
    }
}

// Second stage of a many-group float variance: the input is an array of equal-count Welford partials (means in `mean`,
// M2 in `m2`, `count` elements each) laid out densely over the remaining groups; a warp merges the partials of one output
// in a fixed order (Chan).  The data is the small intermediate of the first, canonical pass, so simplicity wins here.
template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_generic_pairs_kernel(T *__restrict__ out, const T *__restrict__ mean,
                                                                        const T *__restrict__ m2, const __grid_constant__ GenericGeom G,
                                                                        T count, T divisor, int *flag) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kThreads / 32);
    for (long long o = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); o < G.nout; o += warps) {
        long long rem = o, base = 0;
#pragma unroll 1
        for (int d = G.nk - 1; d >= 0; --d) { base += (rem % G.kext[d]) * G.kstr[d]; rem /= G.kext[d]; }
        Acc<T, K_WELFORD> acc;
        acc.init((T)0);
        for (long long i = lane; i < G.nred; i += 32) {
            long long q = i, off = base;
#pragma unroll 1
            for (int d = G.nr - 1; d >= 0; --d) { off += (q % G.rext[d]) * G.rstr[d]; q /= G.rext[d]; }
            acc.merge_raw(count, __ldg(mean + off), __ldg(m2 + off));
        }
        warp_merge(acc);
        if (lane == 0) finalize(acc, out, o, G.nout, G.fin, divisor, flag);
    }
}

// Later passes of a many-group float variance (run_var_multipass): equal-count Welford partials laid out [k1][r1][k2]
// (means in `mean`, M2 in `m2`) are merged along r1.  The data is an intermediate a fraction of the input's size, so a
// thread per output with a serial walk is enough; consecutive threads read consecutive k2.
template <typename T>
__global__ void __launch_bounds__(kThreads) welford_pairs_cols_kernel(T *__restrict__ out, const T *__restrict__ mean, const T *__restrict__ m2,
                                                                      long long K1, long long R1, long long K2, T count, T divisor, int fin) {
    const long long nout = K1 * K2;
    const long long o = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (o >= nout) return;
    const long long k1 = o / K2, k2 = o - k1 * K2;
    Acc<T, K_WELFORD> acc;
    acc.init((T)0);
    const long long base = k1 * R1 * K2 + k2;
    for (long long r = 0; r < R1; ++r) acc.merge_raw(count, __ldg(mean + base + r * K2), __ldg(m2 + base + r * K2));
    finalize(acc, out, o, nout, fin, divisor, (int *)nullptr);
}

// ---- stage 2: merge S partial accumulators per output, in a fixed order (deterministic) -------------------------
// MODE 0: a thread per output (many outputs); 1: a warp per output; 2: a CTA per output (few outputs, thousands of splits:
// a full reduction of 1 GiB leaves ~4700 partials, which one warp merged in ~100 us).  Warp / CTA: every lane takes a
// contiguous run of splits, loads them in batches, then the shuffle tree (and, per CTA, the 8 warp results in order).
template <typename T, int KIND, int MODE>
__global__ void __launch_bounds__(kThreads) reduce_stage2_kernel(T *__restrict__ out, const void *__restrict__ partial,
                                                                 long long nout, long long S, int fin, T divisor, int *flag) {
    const Acc<T, KIND> *p = reinterpret_cast<const Acc<T, KIND> *>(partial);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // stage 1 in front of this launch is complete and visible (no-op without PDL)
    if constexpr (MODE >= 1) {
        constexpr int LANES = MODE == 1 ? 32 : kThreads;
        const long long o = MODE == 1 ? ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5 : (long long)blockIdx.x;
        const int lane = MODE == 1 ? (threadIdx.x & 31) : threadIdx.x;
        if (o >= nout) return;
        Acc<T, KIND> acc;
        acc.init((T)0);
        if constexpr (KIND == K_SQDEV) acc.m = p[o].m;
        const long long per = (S + LANES - 1) / LANES;
        const long long lo = lane * per, hi = (lo + per < S) ? lo + per : S;
        constexpr int B = 4;
        for (long long s = lo; s < hi; s += B) {
            Acc<T, KIND> buf[B];
#pragma unroll
            for (int i = 0; i < B; ++i)
                if (s + i < hi) buf[i] = p[(s + i) * nout + o];
#pragma unroll
            for (int i = 0; i < B; ++i)
                if (s + i < hi) acc.merge(buf[i]);
        }
        warp_merge(acc);
        if constexpr (MODE == 1) {
            if (lane == 0) finalize(acc, out, o, nout, fin, divisor, flag);
        } else {
            __shared__ Acc<T, KIND> smem[kThreads / 32];
            if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                Acc<T, KIND> t = smem[0];
#pragma unroll
                for (int w = 1; w < kThreads / 32; ++w) t.merge(smem[w]);
                finalize(t, out, o, nout, fin, divisor, flag);
            }
        }
    } else {
        // Many outputs: four threads per output, each a contiguous run of splits fetched in batches of 8 (independent loads
        // in flight), then a two-step shuffle tree.  (One thread per output with one load per merge was latency-bound: a
        // 1 GiB block [2048,16384,8] reduced over [0,2] leaves 74 Welford partials for each of 16384 outputs, and their
        // merge cost ~40 us of a 200 us call; neighbouring outputs still read neighbouring partials, so the loads coalesce.)
        constexpr int Q = 4;
        const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
        const long long o = t / Q;
        const int g = (int)(t % Q);
        const bool live = o < nout;  // whole quads are live or not (kThreads % Q == 0); dead lanes only take part in the shuffles
        Acc<T, KIND> acc;
        acc.init((T)0);
        if constexpr (KIND == K_SQDEV) { if (live) acc.m = p[o].m; }
        const long long per = (S + Q - 1) / Q;
        const long long lo = g * per, hi = (lo + per < S) ? lo + per : S;
        constexpr int B = 8;
        for (long long s = lo; live && s < hi; s += B) {
            Acc<T, KIND> buf[B];
#pragma unroll
            for (int i = 0; i < B; ++i)
                if (s + i < hi) buf[i] = p[(s + i) * nout + o];
#pragma unroll
            for (int i = 0; i < B; ++i)
                if (s + i < hi) acc.merge(buf[i]);
        }
#pragma unroll
        for (int off = Q / 2; off > 0; off >>= 1) { const Acc<T, KIND> other = shfl_down_acc(acc, off); acc.merge(other); }
        if (live && g == 0) finalize(acc, out, o, nout, fin, divisor, flag);
    }
}

// ---- host side --------------------------------------------------------------------------------------------
struct Canon {
    bool ok;  // fits the canonical 4-group form
    long long K1, R1, K2, R2;
};

// groups: extents with their reduced flag, alternating, outermost first
inline Canon canonicalise(int m, const long long *ext, const bool *red) {
    Canon c{true, 1, 1, 1, 1};
    if (m == 0) return c;
    if (red[m - 1]) {
        if (m > 4) { c.ok = false; return c; }
        c.R2 = ext[m - 1];
        if (m >= 2) c.K2 = ext[m - 2];
        if (m >= 3) c.R1 = ext[m - 3];
        if (m >= 4) c.K1 = ext[m - 4];
    } else {
        if (m > 3) { c.ok = false; return c; }
        c.K2 = ext[m - 1];
        if (m >= 2) c.R1 = ext[m - 2];
        if (m >= 3) c.K1 = ext[m - 3];
    }
    return c;
}

inline long long clampll(long long v, long long lo, long long hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Runs one canonical reduction.  `in2`: second operand for the dot-like loaders; `aux`: per-output mean (K_SQDEV).
template <typename T, int KIND, int LOADER>
int run_canonical(feo_ctx *ctx, const Canon &c, T *out, const T *in, const T *in2, const T *aux, int fin, T divisor) {
    constexpr int VN = VecN<T>::value;
    Geom G;
    memset(&G, 0, sizeof G);
    G.K1 = c.K1; G.R1 = c.R1; G.K2 = c.K2; G.R2 = c.R2;
    G.nout = c.K1 * c.K2;
    G.fin = fin;
    G.S = G.S1 = G.S2 = 1;
    G.lanes = 32;
    G.lx = 32;
    const long long target_blocks = (long long)ctx->sm_count * (ctx->knobs[FEO_KNOB_REDUCE_BLOCKS_PER_SM] > 0 ? ctx->knobs[FEO_KNOB_REDUCE_BLOCKS_PER_SM] : 8);
    const int forced = ctx->knobs[FEO_KNOB_REDUCE_SPLIT];
    const bool in_aligned = feo_aligned16(in) && (LOADER == LD_PLAIN || LOADER == LD_VECMAT || feo_aligned16(in2));
    enum { COLS, COLS2D, MICRO, ROWS } which;
    bool vec = false;
    int r2c = 0, cols_j = 1;
    // Small inputs (config 1: 8 MiB, L2-resident) are latency-bound: a split buys shorter blocks but pays the publish /
    // combine tail (1.5-6 us against ~3 us of streaming).  When a block's share of the input is small anyway, one
    // un-split launch is the fastest plan (graph-replayed, `tools/split_sweep.py`: [64,128,256] over [0,2] 5.96 -> 4.28 us,
    // over [0] 11.4 -> 5.3, over [1] 11.1 -> 5.1).
    const size_t in_bytes_all = (size_t)c.K1 * c.R1 * c.K2 * c.R2 * sizeof(T);
    const bool small_input = forced == 0 && in_bytes_all <= (size_t(32) << 20);
    auto unsplit_ok = [&](long long block_elems, long long limit_bytes) {
        return small_input && block_elems * (long long)sizeof(T) <= limit_bytes;
    };

    if (c.R2 == 1 && LOADER != LD_MUL && LOADER != LD_MATVEC && c.R1 >= 2 &&
        c.K2 / ((in_aligned && c.K2 % VN == 0) ? VN : 1) <= kThreads / 2) {
        which = COLS2D;  // short rows: lx lanes across k2, ly rows at a time
        vec = in_aligned && (c.K2 % VN == 0);
        const long long W = vec ? VN : 1;
        G.lx = (int)(c.K2 / W);
        G.lanes = kThreads / G.lx;  // ly
        if ((long long)G.lanes > c.R1) G.lanes = (int)c.R1;
        long long kz = kThreads / ((long long)G.lx * G.lanes);  // k1 per block
        if (kz > c.K1) kz = c.K1;
        G.c2 = kz;
        G.xblocks = feo_ceil_div((size_t)c.K1, (size_t)kz);  // blocks along k1
        const long long want = 4 * target_blocks;
        // enough blocks along k1 already: do not split (every split pays the shared-memory tree and a partial again)
        long long S = forced > 0 ? forced : (G.xblocks >= target_blocks ? 1 : feo_ceil_div((size_t)want, (size_t)G.xblocks));
        const long long min_rows = (long long)G.lanes * (vec ? 16 : 32);  // two full sweeps per thread
        S = clampll(S, 1, c.R1 / min_rows > 0 ? c.R1 / min_rows : 1);
        if (unsplit_ok(kz * c.R1 * c.K2, 256 << 10)) S = 1;
        G.c1 = feo_ceil_div((size_t)c.R1, (size_t)S);
        G.S = G.S1 = feo_ceil_div((size_t)c.R1, (size_t)G.c1);
    } else if (c.R2 == 1 && LOADER != LD_MUL && LOADER != LD_MATVEC) {
        which = COLS;
        vec = in_aligned && (c.K2 % VN == 0) && feo_aligned16(out);
        const long long W = vec ? VN : 1;
        // few rows: the columns must supply the loads in flight, 4 column-vectors per thread
        cols_j = (c.R1 < 8 && c.K2 / W >= 4 * kThreads) ? 4 : 1;
        // scalar columns ([3, n] with odd n: rows 1 and 2 start off the 16-byte grid): 4 columns x 3 rows of 4-byte loads are
        // only 48 bytes in flight per thread -- take 8 columns: sum 3.6 -> 4.7 TB/s (knob REDUCE_BLOCKS_PER_SM = -8 keeps 4 for comparison)
        // (not for float variance: its three-word accumulators push the 8-column form into spills, 2.4 -> 1.9 TB/s)
        if (cols_j == 4 && !vec && KIND != K_WELFORD && c.K2 >= 64 * kThreads && ctx->knobs[FEO_KNOB_REDUCE_BLOCKS_PER_SM] != -8) cols_j = 8;
        G.xblocks = feo_ceil_div((size_t)(c.K2 / W), (size_t)kThreads * cols_j);
        const long long base_blocks = G.xblocks * c.K1;
        long long S = forced > 0 ? forced : feo_ceil_div((size_t)target_blocks, (size_t)base_blocks);
        S = clampll(S, 1, c.R1 >= 32 ? c.R1 / 16 : 1);
        if (unsplit_ok(c.R1 * kThreads * cols_j * W, 256 << 10)) S = 1;
        G.c1 = feo_ceil_div((size_t)c.R1, (size_t)S);
        G.S = G.S1 = feo_ceil_div((size_t)c.R1, (size_t)G.c1);
    } else if (c.R2 <= 32 && LOADER == LD_PLAIN) {
        which = MICRO;
        const int cands[5] = {2, 4, 8, 16, 32};
        for (int i = 0; i < 5; ++i)
            if (c.R2 == cands[i] && feo_aligned32(in)) r2c = cands[i];
        const bool wide = (c.R1 == 1);
        int outs = r2c == 0 ? 4 : (r2c <= 16 ? 4 : 2);  // MicroShape<R2C, WIDE>::OUTS
        if (r2c && wide) { outs = 32 / r2c; outs = outs > 8 ? 8 : (outs < 2 ? 2 : outs); }
        G.xblocks = feo_ceil_div((size_t)c.K2, (size_t)kThreads * outs);
        const long long base_blocks = G.xblocks * c.K1;
        long long S = forced > 0 ? forced : feo_ceil_div((size_t)target_blocks, (size_t)base_blocks);
        S = clampll(S, 1, c.R1 >= 32 ? c.R1 / 16 : 1);
        if (unsplit_ok(c.R1 * c.R2 * kThreads * outs, 256 << 10)) S = 1;
        G.c1 = feo_ceil_div((size_t)c.R1, (size_t)S);
        G.S = G.S1 = feo_ceil_div((size_t)c.R1, (size_t)G.c1);
    } else {
        which = ROWS;
        vec = in_aligned && (c.R2 % VN == 0);
        const long long W = vec ? VN : 1;
        const long long nvec = c.R2 / W;
        const long long per_out = c.R1 * c.R2;
        // a block per output for long rows -- unless the outputs alone fill the machine with warps (65536 rows of 4099)
        const bool many_outputs = G.nout >= 16 * target_blocks && per_out < 65536;
        G.lanes = (per_out >= 4096 && !many_outputs) ? kThreads : 32;
        const long long outs_per_block = kThreads / G.lanes;
        const long long base_blocks = feo_ceil_div((size_t)G.nout, (size_t)outs_per_block);
        // several waves of small units: no ragged last wave.  Float variance: every block ends with a tree of Welford merges
        // (shuffles + shared memory), ~12 % of a block that streams only 227 KiB (1 GiB over 4736 blocks: 172 us against
        // 153 us for sum) -- so its blocks are four times longer
        const long long want = (KIND == K_WELFORD ? 1 : 4) * target_blocks;
        long long S = forced > 0 ? forced : feo_ceil_div((size_t)want, (size_t)base_blocks);
        const long long min_elems = 4096;  // keep at least 16 KiB (f32) of work per split
        S = clampll(S, 1, per_out / min_elems > 0 ? per_out / min_elems : 1);
        // (a block per output streams ~21 GB/s: 64 KiB is ~3 us, what the combine tail costs)
        if (G.lanes == kThreads && 2 * base_blocks >= ctx->sm_count && unsplit_ok(per_out, 64 << 10)) S = 1;
        // split rows first, then columns
        long long S1 = S < c.R1 ? S : c.R1;
        G.c1 = feo_ceil_div((size_t)c.R1, (size_t)S1);
        S1 = feo_ceil_div((size_t)c.R1, (size_t)G.c1);
        long long S2 = feo_ceil_div((size_t)S, (size_t)S1);
        long long c2 = feo_ceil_div((size_t)nvec, (size_t)S2);
        const long long sweep = vec ? 4 : 16;  // U of the rows kernel: loads per lane per sweep
        int lx = 1;
        while (lx < G.lanes && (long long)lx * sweep < c2) lx *= 2;
        const long long quantum = (long long)lx * sweep;
        c2 = feo_ceil_div((size_t)c2, (size_t)quantum) * quantum;
        S2 = feo_ceil_div((size_t)nvec, (size_t)c2);
        G.lx = lx;
        G.S1 = S1; G.S2 = S2; G.S = S1 * S2; G.c2 = c2;
    }

    const bool micro_smem = which == MICRO && r2c == 0 && c.R1 == 1 && feo_aligned16(in) && c.R2 >= 2 && forced <= 0;
    if constexpr (KIND == K_INTVAR) {  // only the single-batch geometries can fuse the two integer passes
        if (!(which == MICRO && c.R1 == 1 && (r2c > 0 || micro_smem))) return -1;
    }
    if (micro_smem) {
        // outputs per block: a multiple of 32 keeps block spans aligned; ~16 KiB of input per block so that 8 blocks fit an SM
        int ob = (int)((kMicroSmemBytes / 2 / sizeof(T)) / c.R2) / 32 * 32;
        if (ob < 32) ob = 32;
        if (ob > 1024) ob = 1024;
        const size_t smem = (size_t)ob * c.R2 * sizeof(T);
        const long long blocks = feo_ceil_div((size_t)G.nout, (size_t)ob);
        if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "reduction grid too large");
        reduce_microsmem_kernel<T, KIND><<<(unsigned)blocks, kThreads, smem, ctx->stream>>>(out, in, aux, G.nout, (int)c.R2, ob, fin, divisor, ctx->arith_flag);
        FEO_LAUNCH_CHECK(ctx);
        return FEO_OK;
    }
    void *partial = nullptr;
    G.counters = nullptr;
    if (G.S > 1) {
        int rc = feo_workspace(ctx, (size_t)G.S * (size_t)G.nout * sizeof(Acc<T, KIND>), &partial);
        if (rc != FEO_OK) return rc;
        // fused combine: the last writer of each output merges its S partials (needs one zeroed counter per output)
        // Worth it where a launch costs more than the last arriver's serial tail: small inputs (config 1: 13.6 -> 8.6 us);
        // on multi-GiB inputs the second launch is 0.2 % and the two-stage combine is a little faster.  Knob: 2 = always.
        const size_t in_bytes = (size_t)c.K1 * c.R1 * c.K2 * c.R2 * sizeof(T);
        const int fuse = ctx->knobs[FEO_KNOB_REDUCE_FUSE];
        // The kernels that publish per THREAD (cols, cols2d, micro: one fence + atomic per output and split, then a serial
        // merge by the last arriver) lose to the second launch even on the GPU clock ([64,128,256] over [0]: 11.4 vs 4.9 us,
        // over [0,1] with 128 splits: 31 vs 6.0), so only the rows kernel (one atomic per block, warp-parallel merge) fuses.
        const bool fuse_auto = which == ROWS && in_bytes <= (size_t(256) << 20);
        if (fuse != 1 && (fuse == 2 || fuse_auto) && (size_t)G.nout <= ctx->counters_len) G.counters = ctx->counters;
    }
    int *flag = ctx->arith_flag;
    if (which == COLS) {
        const long long blocks = G.xblocks * G.S * c.K1;
        if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "reduction grid too large");
        constexpr int L = (LOADER == LD_VECMAT) ? LD_VECMAT : LD_PLAIN;
        if (vec && cols_j == 1) reduce_cols_kernel<T, KIND, L, true, 1><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        else if (vec) reduce_cols_kernel<T, KIND, L, true, 4><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        else if (cols_j == 8) {
            if constexpr (KIND != K_WELFORD) reduce_cols_kernel<T, KIND, L, false, 8><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        }
        else if (cols_j == 1) reduce_cols_kernel<T, KIND, L, false, 1><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        else reduce_cols_kernel<T, KIND, L, false, 4><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
    } else if (which == COLS2D) {
        const long long blocks = G.S * G.xblocks;
        if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "reduction grid too large");
        constexpr int L = (LOADER == LD_VECMAT) ? LD_VECMAT : LD_PLAIN;
        if (vec) reduce_cols2d_kernel<T, KIND, L, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        else reduce_cols2d_kernel<T, KIND, L, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
    } else if (which == MICRO) {
        const long long blocks = G.xblocks * G.S * c.K1;
        if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "reduction grid too large");
#define FEO_MICRO(R)                                                                                                          \
    do {                                                                                                                      \
        if (c.R1 == 1) reduce_micro_kernel<T, KIND, R, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, aux, G, divisor, flag); \
        else reduce_micro_kernel<T, KIND, R, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, aux, G, divisor, flag);        \
    } while (0)
        switch (r2c) {
        case 2: FEO_MICRO(2); break;
        case 4: FEO_MICRO(4); break;
        case 8: FEO_MICRO(8); break;
        case 16: FEO_MICRO(16); break;
        case 32: FEO_MICRO(32); break;
        default: FEO_MICRO(0); break;
        }
#undef FEO_MICRO
    } else {
        const long long units = G.nout * G.S;
        const long long blocks = feo_ceil_div((size_t)units, (size_t)(kThreads / G.lanes));
        if (blocks > 0x7fffffffLL) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "reduction grid too large");
        if constexpr (LOADER == LD_PLAIN && (KIND == K_SUM || KIND == K_MAX || KIND == K_MIN || KIND == K_WELFORD || KIND == K_SUMSQ)) {
            // a sharded reduction whose local part is this one launch with few, final outputs: the kernel's last CTA runs
            // the exchange itself (feo_reduce_sharded then skips the exchange kernel)
            auto *rq = static_cast<feo_peer::XTailReq *>(ctx->xtail_req);
            const bool final_in_one = (G.S == 1) || (G.counters != nullptr);
            if (rq && !rq->consumed && (const void *)out == rq->slot && final_in_one && G.nout <= 8192 && ctx->xtail_ticket &&
                ctx->knobs[FEO_KNOB_REDUCE_FUSE] != 1) {
                G.xt.on = 1;
                G.xt.mean = rq->mean;
                G.xt.p = rq->p;
                G.xt.out = rq->out;
                G.xt.divisor_bits = rq->divisor_bits;
                G.xt.count_bits = rq->count_bits;
                G.xt.ticket = ctx->xtail_ticket;
                rq->consumed = true;
            }
        }
        constexpr int L = (LOADER == LD_MUL || LOADER == LD_MATVEC) ? LOADER : LD_PLAIN;
        if (vec) reduce_rows_kernel<T, KIND, L, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
        else reduce_rows_kernel<T, KIND, L, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(out, partial, in, in2, aux, G, divisor, flag);
    }
    FEO_LAUNCH_CHECK(ctx);
    if (G.S > 1 && !G.counters) {
        // programmatic dependent launch: stage 2 is set up while stage 1 still runs and starts the moment it has finished
        // (the kernel executes griddepcontrol.wait before it reads a partial)
        auto stage2 = [&](auto kern, long long blocks) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)blocks);
            cfg.blockDim = dim3(kThreads);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = ctx->knobs[FEO_KNOB_PDL] == 1 ? 0 : 1;
            return cudaLaunchKernelEx(&cfg, kern, out, (const void *)partial, (long long)G.nout, (long long)G.S, fin, divisor, flag);
        };
        if (G.nout <= 64 && G.S >= 256) {
            FEO_CUDA(ctx, stage2(reduce_stage2_kernel<T, KIND, 2>, G.nout));
        } else if (G.nout < 4096) {
            FEO_CUDA(ctx, stage2(reduce_stage2_kernel<T, KIND, 1>, (long long)feo_ceil_div((size_t)G.nout * 32, (size_t)kThreads)));
        } else {
            FEO_CUDA(ctx, stage2(reduce_stage2_kernel<T, KIND, 0>, (long long)feo_ceil_div((size_t)G.nout * 4, (size_t)kThreads)));  // four threads per output
        }
        FEO_LAUNCH_CHECK(ctx);
    }
    return FEO_OK;
}

}  // namespace feo_red
