// Cross-GPU epilogues over peer memory: the exchange step that follows a sharded reduction
// or dot, written as ONE kernel per rank that signals, waits and combines directly over NVLink instead of an NCCL call.
//
// Every rank owns a *window* of peer-mappable memory (cudaMalloc + CUDA IPC; the handles travel through whatever
// rendezvous the host has -- torch.distributed in this repo):
//
//     [ 8 flags, one 128-byte line each | slot 0 | slot 1 ]
//
// Exchange number e (1, 2, ...; every rank issues the same sequence of exchanges, like any collective):
//   1. the local kernels (feo_reduce_partial, feo_dot, ...) write this rank's partial straight into slot e & 1 of the
//      rank's OWN window -- no staging copy;
//   2. the exchange kernel stores e into flag[rank] of every peer's window (st.release.sys), polls its own flags until all
//      ranks have reached e (ld.acquire.sys; bounded by the context's peer timeout: on expiry the status flag is raised, see peer_sync.cuh),
//   3. then every thread loads its elements of all `world` partials (the remote ones are plain peer loads over
//      NVLink/NVSwitch), combines them in RANK ORDER -- so the result is deterministic and bit-identical on every rank --
//      and applies the epilogue of the reference operation (mean: / n; var: Chan merge, / n) before the one store.
// Two slots are enough: a rank can only write slot e & 1 again (for e + 2) after it has left exchange e + 1, which needs
// every peer's flag e + 1, which a peer raises only after it has finished reading in exchange e.
#include "feo_common.cuh"
#include "peer_sync.cuh"
#include "reduce_kernels.cuh"
#include <cstring>

using feo_red::Acc;
using feo_red::K_WELFORD;

using namespace feo_peer;

namespace {

using feo_red::X_SUM;
using feo_red::X_MEAN;
using feo_red::X_VAR;
using feo_red::X_MAX;
using feo_red::X_MIN;

// out[o] = epilogue(partial_0[o] (+) partial_1[o] (+) ... ) in rank order; V = elements per access.
template <typename T, int KIND, int V>
__global__ void __launch_bounds__(kXThreads) xchg_allreduce_kernel(const __grid_constant__ XchgParams p, T *__restrict__ out, T divisor,
                                                                   T count, int *arith_flag) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the local partial in front of this launch is complete and visible
    signal_and_wait(p);
    feo_red::xchg_combine<T, KIND, V>(p, out, divisor, count, arith_flag, (long long)blockIdx.x * kXThreads + threadIdx.x,
                                      (long long)gridDim.x * kXThreads);
}

// Barrier only (n = 0): used before / after kernels that read peers' buffers directly.
__global__ void __launch_bounds__(kXThreads) xchg_barrier_kernel(const __grid_constant__ XchgParams p) { signal_and_wait(p); }

// All-gather by pulling: out = part_0 | part_1 | ... (each `bytes_per_part` long), U-byte accesses.
struct GatherParams {
    const char *part[kMaxPeers];
    size_t bytes_per_part;
};
template <typename U>
__global__ void __launch_bounds__(kXThreads) xchg_gather_kernel(const __grid_constant__ XchgParams p, const __grid_constant__ GatherParams g,
                                                                U *__restrict__ out) {
    signal_and_wait(p);
    const size_t per = g.bytes_per_part / sizeof(U), total = per * p.world;
    for (size_t i = (size_t)blockIdx.x * kXThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kXThreads) {
        const size_t r = i / per;
        out[i] = ld_peer(reinterpret_cast<const U *>(g.part[r]) + (i - r * per));
    }
}

template <typename T, int KIND>
int launch_allreduce(feo_comm *comm, const XchgParams &p, void *out, size_t n, size_t count, const void *divisor_host) {
    feo_ctx *ctx = comm->ctx;
    const T divisor = divisor_host ? *static_cast<const T *>(divisor_host) : (T)1;
    // 16-byte accesses when the row of n elements (and, for var, the M2 row behind it) stays aligned
    const bool vec = (n * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    constexpr int V = 16 / (int)sizeof(T);
    const size_t units = vec ? n / V : n;
    const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(feo_ceil_div(units, (size_t)kXThreads), 1), (size_t)ctx->sm_count * 4);
    // programmatic dependent launch: the exchange kernel is set up while the local reduction in front of it still runs and
    // starts the moment that one finishes (it executes griddepcontrol.wait before touching the partial)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kXThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->knobs[FEO_KNOB_PDL] == 1 ? 0 : 1;
    T *o = (T *)out;
    const T cnt = (T)count;
    int *flag = ctx->arith_flag;
    if (vec) FEO_CUDA(ctx, cudaLaunchKernelEx(&cfg, xchg_allreduce_kernel<T, KIND, V>, p, o, divisor, cnt, flag));
    else FEO_CUDA(ctx, cudaLaunchKernelEx(&cfg, xchg_allreduce_kernel<T, KIND, 1>, p, o, divisor, cnt, flag));
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

template <typename T> int dispatch_kind(feo_comm *comm, int kind, const XchgParams &p, void *out, size_t n, size_t count, const void *div) {
    switch (kind) {
    case FEO_SUM: return launch_allreduce<T, X_SUM>(comm, p, out, n, count, div);
    case FEO_MEAN: return launch_allreduce<T, X_MEAN>(comm, p, out, n, count, div);
    case FEO_MAX: return launch_allreduce<T, X_MAX>(comm, p, out, n, count, div);
    case FEO_MIN: return launch_allreduce<T, X_MIN>(comm, p, out, n, count, div);
    case FEO_VAR: return launch_allreduce<T, X_VAR>(comm, p, out, n, count, div);
    default: return feo_fail(comm->ctx, FEO_ERR_UNSUPPORTED, "unknown reduction kind %d", kind);
    }
}

XchgParams next_exchange(feo_comm *comm, long long n) {
    XchgParams p{};
    for (int r = 0; r < comm->world; ++r) p.win[r] = static_cast<char *>(comm->windows[r]);
    p.rank = comm->rank;
    p.world = comm->world;
    p.epoch = ++comm->epoch;
    p.slot_off = kFlagBytes + (p.epoch & 1) * comm->slot_bytes;
    p.n = n;
    p.timeout_ns = comm->ctx->peer_timeout_ns;
    p.status = comm->ctx->status_flag;
    return p;
}

}  // namespace

extern "C" {

int feo_peer_alloc(feo_ctx *ctx, size_t bytes, void **dptr) {
    FEO_ENTER(ctx);
    if (!ctx || !dptr || bytes == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    FEO_CUDA(ctx, cudaSetDevice(ctx->device));
    FEO_CUDA(ctx, cudaMalloc(dptr, bytes));  // not the stream-ordered pool: legacy IPC handles need a plain allocation
    FEO_CUDA(ctx, cudaMemset(*dptr, 0, bytes));
    FEO_CUDA(ctx, cudaDeviceSynchronize());
    return FEO_OK;
}
int feo_peer_free(feo_ctx *ctx, void *dptr) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_ERR_SHAPE;
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FEO_CUDA(ctx, cudaFree(dptr));
    return FEO_OK;
}
int feo_peer_export(feo_ctx *ctx, void *dptr, unsigned char *handle) {
    FEO_ENTER(ctx);
    if (!ctx || !dptr || !handle) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == FEO_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    FEO_CUDA(ctx, cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, sizeof(h));
    return FEO_OK;
}
int feo_peer_open(feo_ctx *ctx, const unsigned char *handle, void **dptr) {
    FEO_ENTER(ctx);
    if (!ctx || !dptr || !handle) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    FEO_CUDA(ctx, cudaSetDevice(ctx->device));
    FEO_CUDA(ctx, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FEO_OK;
}
int feo_peer_close(feo_ctx *ctx, void *dptr) {
    FEO_ENTER(ctx);
    if (!ctx || !dptr) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FEO_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return FEO_OK;
}

int feo_comm_create(feo_ctx *ctx, int rank, int world, void *const *windows, size_t window_bytes, feo_comm **out) {
    FEO_ENTER(ctx);
    if (!ctx || !windows || !out) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return feo_fail(ctx, FEO_ERR_SHAPE, "rank %d of %d: at most %d peers (one node)", rank, world, kMaxPeers);
    if (window_bytes < kFlagBytes + 2 * 256) return feo_fail(ctx, FEO_ERR_SHAPE, "window of %zu bytes is too small", window_bytes);
    for (int r = 0; r < world; ++r)
        if (!windows[r] || reinterpret_cast<uintptr_t>(windows[r]) % 256 != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "window %d is null or unaligned", r);
    feo_comm *c = new feo_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    for (int r = 0; r < world; ++r) c->windows[r] = windows[r];
    c->window_bytes = window_bytes;
    c->slot_bytes = ((window_bytes - kFlagBytes) / 2) & ~(size_t)255;
    c->epoch = 0;
    *out = c;
    return FEO_OK;
}
int feo_comm_destroy(feo_comm *comm) {
    delete comm;
    return FEO_OK;
}
int feo_comm_slot(feo_comm *comm, void **dptr, size_t *bytes) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !dptr) return FEO_ERR_SHAPE;
    *dptr = static_cast<char *>(comm->windows[comm->rank]) + kFlagBytes + ((comm->epoch + 1) & 1) * comm->slot_bytes;
    if (bytes) *bytes = comm->slot_bytes;
    return FEO_OK;
}
int feo_comm_barrier(feo_comm *comm) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    const XchgParams p = next_exchange(comm, 0);
    xchg_barrier_kernel<<<1, kXThreads, 0, ctx->stream>>>(p);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}
int feo_comm_allreduce(feo_comm *comm, int kind, int dtype, void *out, size_t n, size_t count_per_rank, const void *divisor_host) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !out) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    if (n == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "empty exchange");
    if ((kind == FEO_MEAN || kind == FEO_VAR) && !divisor_host) return feo_fail(ctx, FEO_ERR_SHAPE, "mean / var need the divisor");
    const size_t need = n * feo_dtype_size(dtype) * (kind == FEO_VAR ? 2 : 1);
    if (need > comm->slot_bytes) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "partial of %zu bytes exceeds the %zu-byte slot", need, comm->slot_bytes);
    const XchgParams p = next_exchange(comm, (long long)n);
    int rc = FEO_OK;
    FEO_DISPATCH(ctx, dtype, T, rc = dispatch_kind<T>(comm, kind, p, out, n, count_per_rank, divisor_host));
    return rc;
}

// Tensor::{sum,mean,var,max,min} of a tensor sharded on its leading axis, the leading axis among the removed ones
// (tensor.rs:70-307 on the global tensor): local one-pass reduction straight into the window, then the exchange kernel.
// `shape` is the LOCAL block, `divisor_host` the reference's counted-in-T divisor of the GLOBAL shape (feo_reduce_divisor).
int feo_reduce_sharded(feo_comm *comm, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape, int naxes,
                       const size_t *axes, const void *divisor_host) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !out || !in || !shape) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    if (naxes < 1 || !axes || axes[0] != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "the leading (sharded) axis must be reduced; otherwise no exchange is needed");
    int out_rank = 0;
    size_t out_shape[FEO_MAX_RANK];
    if (feo_reduce_shape(rank, shape, naxes, axes, &out_rank, out_shape) != FEO_OK)
        return feo_fail(ctx, FEO_ERR_SHAPE, "axes must be strictly ascending and in range");
    const size_t nout = feo_shape_size(out_rank, out_shape);
    size_t count = 1;
    for (int i = 0; i < naxes; ++i) count *= shape[axes[i]];
    void *slot = nullptr;
    size_t slot_bytes = 0;
    feo_comm_slot(comm, &slot, &slot_bytes);
    if (nout * feo_dtype_size(dtype) * (kind == FEO_VAR ? 2 : 1) > slot_bytes)
        return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "partial exceeds the %zu-byte slot", slot_bytes);
    if ((kind == FEO_MEAN || kind == FEO_VAR) && !divisor_host) return feo_fail(ctx, FEO_ERR_SHAPE, "mean / var need the divisor");
    // var partials: floats (mean | M2) merged by Chan, integers (S1 | S2) summed and resolved by the ring identity -- one
    // pass over the data and one exchange either way.  The exchange is offered to the local reduction first: if its plan is a
    // single launch with few outputs, that kernel's last CTA signals, waits and combines itself (one launch, ~3 us less at
    // config-1 sizes); otherwise the exchange kernel follows with programmatic dependent launch.
    const XchgParams p = next_exchange(comm, (long long)nout);
    feo_peer::XTailReq rq{};
    rq.consumed = false;
    rq.mean = kind == FEO_MEAN;
    rq.p = p;
    rq.out = out;
    rq.slot = slot;
    FEO_DISPATCH(ctx, dtype, T, {
        const T d = divisor_host ? *static_cast<const T *>(divisor_host) : (T)1;
        const T c = (T)count;
        memcpy(&rq.divisor_bits, &d, sizeof(T));
        memcpy(&rq.count_bits, &c, sizeof(T));
    });
    ctx->xtail_req = &rq;
    int rc = feo_reduce_partial(ctx, kind, dtype, slot, in, rank, shape, naxes, axes);
    ctx->xtail_req = nullptr;
    if (rc != FEO_OK || rq.consumed) return rc;
    FEO_DISPATCH(ctx, dtype, T, rc = dispatch_kind<T>(comm, kind, p, out, nout, count, divisor_host));
    return rc;
}

// Vector::vecmul (vector.rs:71-78) on two identically sharded vectors: `batch` partial dots into the window, then the sum.
int feo_dot_sharded(feo_comm *comm, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n_local) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !out) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    void *slot = nullptr;
    size_t slot_bytes = 0;
    feo_comm_slot(comm, &slot, &slot_bytes);
    if (batch * feo_dtype_size(dtype) > slot_bytes) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "partial exceeds the %zu-byte slot", slot_bytes);
    int rc = feo_dot_batched(ctx, dtype, slot, a, b, batch, n_local);
    if (rc != FEO_OK) return rc;
    return feo_comm_allreduce(comm, FEO_SUM, dtype, out, batch, n_local, nullptr);
}

// All-gather of equally sized parts that live in peer-mapped memory (parts[r] = rank r's buffer as mapped HERE).
int feo_comm_gather(feo_comm *comm, void *out, const void *const *parts, size_t bytes_per_part) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !out || !parts) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    if (bytes_per_part == 0 || bytes_per_part % 4 != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "parts must be a positive multiple of 4 bytes");
    GatherParams g{};
    bool wide = bytes_per_part % 16 == 0 && feo_aligned16(out);
    for (int r = 0; r < comm->world; ++r) {
        if (!parts[r]) return feo_fail(ctx, FEO_ERR_SHAPE, "part %d is null", r);
        g.part[r] = static_cast<const char *>(parts[r]);
        wide = wide && feo_aligned16(parts[r]);
    }
    g.bytes_per_part = bytes_per_part;
    const XchgParams p = next_exchange(comm, 0);
    const size_t units = bytes_per_part * comm->world / (wide ? 16 : 4);
    const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(feo_ceil_div(units, (size_t)kXThreads * 4), 1), (size_t)ctx->sm_count * 8);
    if (wide) xchg_gather_kernel<uint4><<<blocks, kXThreads, 0, ctx->stream>>>(p, g, static_cast<uint4 *>(out));
    else xchg_gather_kernel<uint32_t><<<blocks, kXThreads, 0, ctx->stream>>>(p, g, static_cast<uint32_t *>(out));
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

// Matrix::matmul (matrix.rs:76-90) with A row-sharded (this rank's [M,K] block) and B row-sharded too: b_parts[r] is rank
// r's [K / world, N] panel in peer-mapped memory.  f32 on the tensor cores: the split / transpose pre-pass pulls the panels
// over NVLink itself (no gathered B is materialised); other dtypes / SIMT: pull-gather into a scratch B, then the kernel.
// A closing barrier keeps every rank's panel untouched until all ranks have read it.
int feo_matmul_sharded_b(feo_comm *comm, int algo, int dtype, void *C, const void *A, const void *const *b_parts, size_t M, size_t K,
                         size_t N) {
    FEO_ENTER(comm ? comm->ctx : nullptr);
    if (!comm || !C || !A || !b_parts) return FEO_ERR_SHAPE;
    feo_ctx *ctx = comm->ctx;
    if (M == 0 || K == 0 || N == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    if (K % comm->world != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "B's %zu rows do not split evenly over %d ranks", K, comm->world);
    const size_t esz = feo_dtype_size(dtype);
    if (esz == 0) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    for (int r = 0; r < comm->world; ++r)
        if (!b_parts[r]) return feo_fail(ctx, FEO_ERR_SHAPE, "B panel %d is null", r);
    int rc;
    const bool tensor = dtype == FEO_F32 && (algo == FEO_MATMUL_TF32X3 || (algo == FEO_MATMUL_AUTO && feo_matmul_tf32x3_supported(M, K, N)));
    if (algo == FEO_MATMUL_TF32X3 && dtype != FEO_F32) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "TF32X3 is f32-only");
    if (tensor) {
        BParts parts{};
        for (int r = 0; r < comm->world; ++r) parts.part[r] = static_cast<const float *>(b_parts[r]);
        parts.world = comm->world;
        parts.rows_per_part = (int)(K / comm->world);
        const XchgParams p = next_exchange(comm, 0);
        if ((rc = feo_impl_matmul_tf32x3_parts(ctx, C, A, parts, p, M, K, N)) != FEO_OK) return rc;
    } else {
        void *full = nullptr;
        FEO_CUDA(ctx, feo_alloc_async(ctx, &full, K * N * esz));
        rc = feo_comm_gather(comm, full, b_parts, K / comm->world * N * esz);
        if (rc == FEO_OK) rc = feo_matmul(ctx, algo, dtype, C, A, full, M, K, N);
        cudaFreeAsync(full, ctx->stream);
        if (rc != FEO_OK) return rc;
    }
    return feo_comm_barrier(comm);
}

}  // extern "C"
