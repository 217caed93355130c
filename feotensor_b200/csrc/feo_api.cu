// feo_api.cu -- the C ABI of libfeotensor_b200.so: context, stream-ordered memory, bookkeeping and
// argument validation in front of the sm_100a kernels.  See include/feotensor_b200.h for the
// reference function (file:line) each entry point replaces.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "feo_common.cuh"

int feo_fail(feo_ctx *ctx, int status, const char *fmt, ...) {
    FEO_ENTER(ctx);
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return status;
}

bool feo_pdl_admit(feo_ctx *ctx) {
    if (!ctx->pending_valid || ctx->knobs[FEO_KNOB_PDL] == 1) { ctx->window.clear(); return false; }
    ctx->pending_valid = false;
    const feo_launch_rec &r = ctx->pending;
    auto hit = [](const feo_range &x, const feo_range &y) { return x.lo < y.hi && y.lo < x.hi; };
    // A kernel may still have CTAs running when the kernel 2+ launches later starts, so the check covers the whole window,
    // and a full window forces one fully ordered launch.
    bool independent = ctx->window.size() < 16;
    for (size_t i = 0; independent && i < ctx->window.size(); ++i) {
        const feo_launch_rec &w = ctx->window[i];
        if (hit(r.out, w.out)) independent = false;
        for (int k = 0; k < 2 && independent; ++k)
            if (hit(r.out, w.in[k]) || hit(r.in[k], w.out)) independent = false;
    }
    if (!independent) ctx->window.clear();
    ctx->window.push_back(r);
    return independent;
}

cudaError_t feo_alloc_async(feo_ctx *ctx, void **ptr, size_t bytes) {
    if (ctx->pool) return cudaMallocFromPoolAsync(ptr, bytes, ctx->pool, ctx->stream);
    return cudaMallocAsync(ptr, bytes, ctx->stream);
}

int feo_side_stream(feo_ctx *ctx) {
    FEO_ENTER(ctx);
    if (ctx->side_stream) return FEO_OK;
    int lo = 0, hi = 0;
    FEO_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    FEO_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 3; ++i) {
        FEO_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->pull_streams[i], cudaStreamNonBlocking, hi));
        FEO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pull_joins[i], cudaEventDisableTiming));
    }
    FEO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    FEO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    FEO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_open, cudaEventDisableTiming));
    return FEO_OK;
}

int feo_workspace(feo_ctx *ctx, size_t bytes, void **ptr) {
    FEO_ENTER(ctx);
    if (bytes > ctx->workspace_bytes) {
        if (ctx->workspace) FEO_CUDA(ctx, cudaFreeAsync(ctx->workspace, ctx->stream));
        ctx->workspace = nullptr;
        ctx->workspace_bytes = 0;
        size_t want = bytes < (size_t(8) << 20) ? (size_t(8) << 20) : bytes;
        FEO_CUDA(ctx, feo_alloc_async(ctx, &ctx->workspace, want));
        ctx->workspace_bytes = want;
    }
    *ptr = ctx->workspace;
    return FEO_OK;
}

// The synchronisation points (feo_sync, feo_download) are where the sticky device flags reach the host: the reference panics
// at `a / b` itself (integer x/0, MIN/-1); an asynchronous backend reports it at the next point the host looks at results.
static int feo_sync_and_check(feo_ctx *ctx) {
    if (ctx->host_flags) FEO_CUDA(ctx, cudaMemcpyAsync(ctx->host_flags, ctx->arith_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!ctx->host_flags) return FEO_OK;
    if (ctx->host_flags[0]) {
        FEO_CUDA(ctx, cudaMemsetAsync(ctx->arith_flag, 0, sizeof(int), ctx->stream));
        return feo_fail(ctx, FEO_ERR_ARITH, "attempt to divide by zero, or to divide with overflow, in an integer kernel since the last "
                                            "synchronisation (the reference panics; the kernel wrote 0 there)");
    }
    if (ctx->host_flags[1]) {
        FEO_CUDA(ctx, cudaMemsetAsync(ctx->status_flag, 0, sizeof(int), ctx->stream));
        return feo_fail(ctx, FEO_ERR_CUDA, "a kernel gave up waiting for a peer rank after %.0f s: the results of that exchange are invalid",
                        ctx->peer_timeout_ns * 1e-9);
    }
    return FEO_OK;
}

extern "C" {

int feo_abi_version(void) { return FEO_ABI_VERSION; }

size_t feo_dtype_size(int dtype) {
    switch (dtype) {
    case FEO_F32: case FEO_I32: case FEO_U32: return 4;
    case FEO_F64: case FEO_I64: return 8;
    default: return 0;
    }
}

int feo_init(int device, feo_ctx **out) {
    if (!out) return FEO_ERR_SHAPE;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return FEO_ERR_CUDA;  // no CPU fallback
    feo_ctx *ctx = new feo_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FEO_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return FEO_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FEO_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    {
        // own pool; keep freed blocks cached across synchronisations (the default pool trims to 0 at every sync)
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            ctx->pool = nullptr;  // plain cudaMallocAsync from the device's default pool
            cudaGetLastError();
        }
    }
    if (cudaMalloc(&ctx->arith_flag, 2 * sizeof(int)) != cudaSuccess || cudaMemset(ctx->arith_flag, 0, 2 * sizeof(int)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return FEO_ERR_CUDA;
    }
    ctx->status_flag = ctx->arith_flag + 1;
    if (cudaMallocHost(&ctx->host_flags, 2 * sizeof(int)) != cudaSuccess) { ctx->host_flags = nullptr; cudaGetLastError(); }
    if (const char *env = getenv("FEO_PEER_TIMEOUT_S")) {
        const double sec = atof(env);
        ctx->peer_timeout_ns = sec > 0 ? (unsigned long long)(sec * 1e9) : 0;
    }
    ctx->counters_len = size_t(1) << 20;
    if (cudaMalloc(&ctx->counters, (ctx->counters_len + 1) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(ctx->counters, 0, (ctx->counters_len + 1) * sizeof(unsigned int)) != cudaSuccess) {
        ctx->counters = nullptr;
        ctx->counters_len = 0;  // fall back to the two-launch combine
    }
    ctx->xtail_ticket = ctx->counters ? ctx->counters + ctx->counters_len : nullptr;
    *out = ctx;
    return FEO_OK;
}

int feo_shutdown(feo_ctx *ctx) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->workspace) cudaFreeAsync(ctx->workspace, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->arith_flag) cudaFree(ctx->arith_flag);
    if (ctx->host_flags) cudaFreeHost(ctx->host_flags);
    if (ctx->host_ones) cudaFreeHost(ctx->host_ones);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    for (int i = 0; i < 3; ++i) {
        if (ctx->pull_streams[i]) { cudaStreamSynchronize(ctx->pull_streams[i]); cudaStreamDestroy(ctx->pull_streams[i]); }
        if (ctx->pull_joins[i]) cudaEventDestroy(ctx->pull_joins[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_open) cudaEventDestroy(ctx->ev_open);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return FEO_OK;
}

const char *feo_last_error(feo_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int feo_set_stream(feo_ctx *ctx, void *s) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_ERR_SHAPE;
    // the workspace was allocated stream-ordered on the old stream: settle it before switching
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return FEO_OK;
}
void *feo_get_stream(feo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int feo_sync(feo_ctx *ctx) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_ERR_SHAPE;
    return feo_sync_and_check(ctx);
}
uint64_t feo_launch_count(feo_ctx *ctx) { return ctx ? ctx->launches : 0; }
int feo_tune(feo_ctx *ctx, int knob, int value) {
    if (!ctx || knob < 0 || knob >= FEO_KNOB_COUNT) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown knob %d", knob);
    ctx->knobs[knob] = value;
    if (knob == FEO_KNOB_PEER_TIMEOUT_S) ctx->peer_timeout_ns = value > 0 ? (unsigned long long)value * 1000000000ull : 0;
    if (knob == FEO_KNOB_BCAST_L2KEEP) {
        // 2: reserve the largest persisting-L2 carve-out the device allows; broadcast launches then put an access-policy
        // window on the operand that later launches re-read.  Anything else: give the carve-out back.
        FEO_ENTER(ctx);
        if (value >= 2) {  // 2: the largest carve-out the device allows; > 2: that many MiB
            int max_persist = 0, max_window = 0;
            FEO_CUDA(ctx, cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
            FEO_CUDA(ctx, cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
            if (value > 2 && ((size_t)value << 20) < (size_t)max_persist) max_persist = (int)((size_t)value << 20);
            FEO_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
            ctx->l2_persist_bytes = (size_t)max_persist;
            ctx->l2_window_max = (size_t)max_window;
        } else if (ctx->l2_persist_bytes) {
            FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            FEO_CUDA(ctx, cudaCtxResetPersistingL2Cache());
            FEO_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
            ctx->l2_persist_bytes = 0;
        }
    }
    return FEO_OK;
}

int feo_malloc(feo_ctx *ctx, size_t bytes, void **dptr) {
    FEO_ENTER(ctx);
    if (!dptr) return feo_fail(ctx, FEO_ERR_SHAPE, "null output pointer");
    *dptr = nullptr;
    FEO_CUDA(ctx, feo_alloc_async(ctx, dptr, bytes ? bytes : 1));
    return FEO_OK;
}
int feo_free(feo_ctx *ctx, void *dptr) {
    FEO_ENTER(ctx);
    if (dptr) FEO_CUDA(ctx, cudaFreeAsync(dptr, ctx->stream));  // stream-ordered: takes effect after every earlier kernel
    return FEO_OK;
}
int feo_trim(feo_ctx *ctx) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_ERR_SHAPE;
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pool) FEO_CUDA(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
    return FEO_OK;
}
int feo_host_alloc(feo_ctx *ctx, size_t bytes, void **hptr) {
    FEO_ENTER(ctx);
    FEO_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 1));
    return FEO_OK;
}
int feo_host_free(feo_ctx *ctx, void *hptr) {
    FEO_ENTER(ctx);
    if (hptr) FEO_CUDA(ctx, cudaFreeHost(hptr));
    return FEO_OK;
}
int feo_upload(feo_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FEO_ENTER(ctx);
    if (bytes) FEO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FEO_OK;
}
int feo_download_async(feo_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FEO_ENTER(ctx);
    if (bytes) FEO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FEO_OK;
}
int feo_download(feo_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FEO_ENTER(ctx);
    int rc = feo_download_async(ctx, dst, src, bytes);
    if (rc != FEO_OK) return rc;
    return feo_sync_and_check(ctx);  // the data is copied either way; FEO_ERR_ARITH says it contains the zeros of a x/0
}
int feo_copy(feo_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FEO_ENTER(ctx);
    if (bytes) FEO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return FEO_OK;
}

// ---- bookkeeping ---------------------------------------------------------------------------------
int feo_shape_check(int rank, const size_t *shape) {
    if (rank < 0) return FEO_ERR_SHAPE;
    for (int d = 0; d < rank; ++d)
        if (shape[d] == 0) return FEO_ERR_SHAPE;  // "Dimension cannot be zero" (shape.rs:13-15)
    return FEO_OK;
}
size_t feo_shape_size(int rank, const size_t *shape) {
    size_t s = 1;
    for (int d = 0; d < rank; ++d) s *= shape[d];
    return s;
}
int feo_flatten(int coord_rank, const size_t *coord, int rank, const size_t *shape, size_t *index, int *detail) {
    if (coord_rank != rank) { if (detail) *detail = -1; return FEO_ERR_SHAPE; }
    for (int i = 0; i < rank; ++i)
        if (coord[i] >= shape[i]) { if (detail) *detail = i; return FEO_ERR_SHAPE; }
    size_t idx = 0, stride = 1;
    for (int k = rank - 1; k >= 0; --k) { idx += coord[k] * stride; stride *= shape[k]; }
    if (index) *index = idx;
    return FEO_OK;
}
static int axes_check(int rank, int naxes, const size_t *axes) {
    if (naxes < 0 || naxes > rank) return FEO_ERR_SHAPE;
    for (int i = 0; i < naxes; ++i) {
        if (axes[i] >= (size_t)rank) return FEO_ERR_SHAPE;
        if (i > 0 && axes[i] <= axes[i - 1]) return FEO_ERR_SHAPE;
    }
    return FEO_OK;
}
int feo_reduce_shape(int rank, const size_t *shape, int naxes, const size_t *axes, int *out_rank, size_t *out_shape) {
    if (feo_shape_check(rank, shape) != FEO_OK || axes_check(rank, naxes, axes) != FEO_OK) return FEO_ERR_SHAPE;
    int nk = 0;
    for (int d = 0; d < rank; ++d) {
        bool removed = false;
        for (int i = 0; i < naxes; ++i) removed |= (axes[i] == (size_t)d);
        if (!removed) out_shape[nk++] = shape[d];
    }
    if (naxes == 0 || nk == 0) { out_shape[0] = 1; nk = 1; }
    *out_rank = nk;
    return FEO_OK;
}
int feo_reduce_divisor(int dtype, int rank, const size_t *shape, int naxes, const size_t *axes, void *out) {
    if (axes_check(rank, naxes, axes) != FEO_OK) return FEO_ERR_SHAPE;
    FEO_DISPATCH(nullptr, dtype, T, *(T *)out = feo_divisor<T>(rank, shape, naxes, axes));
    return FEO_OK;
}

// ---- constructors --------------------------------------------------------------------------------
int feo_fill(feo_ctx *ctx, int dtype, void *out, const void *value, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !value) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    return feo_impl_fill(ctx, dtype, out, value, n);
}
int feo_eye(feo_ctx *ctx, int dtype, void *out, size_t rows, size_t cols) {
    FEO_ENTER(ctx);
    if (!ctx || !out) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (rows == 0 || cols == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    if (rows > cols) return feo_fail(ctx, FEO_ERR_SHAPE, "out of bounds for dimension 1");  // matrix.rs:38-40
    return feo_impl_eye(ctx, dtype, out, rows, cols);
}

int feo_fill_uniform(feo_ctx *ctx, int dtype, void *out, size_t n, uint64_t seed, uint64_t first_index, double lo, double hi) {
    FEO_ENTER(ctx);
    if (!ctx || !out) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (!(hi > lo)) return feo_fail(ctx, FEO_ERR_SHAPE, "fill_uniform needs hi > lo");
    if (n == 0) return FEO_OK;
    return feo_impl_fill_uniform(ctx, dtype, out, n, seed, first_index, lo, hi);
}

// ---- elementwise ---------------------------------------------------------------------------------
int feo_ewise(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !b) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (op < FEO_ADD || op > FEO_DIV) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown op %d", op);
    if (n == 0) return FEO_OK;
    return feo_impl_ewise(ctx, op, dtype, out, a, b, nullptr, n);
}
int feo_ewise_scalar(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *scalar_host, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !scalar_host) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (op < FEO_ADD || op > FEO_DIV) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown op %d", op);
    if (n == 0) return FEO_OK;
    if (op == FEO_DIV) {  // `tensor / 0` on an integer tensor panics in the reference at the first element (tensor.rs:465-475)
        bool zero = false;
        if (dtype == FEO_I32 || dtype == FEO_U32) zero = *static_cast<const uint32_t *>(scalar_host) == 0;
        else if (dtype == FEO_I64) zero = *static_cast<const int64_t *>(scalar_host) == 0;
        if (zero) return feo_fail(ctx, FEO_ERR_ARITH, "attempt to divide by zero");
    }
    return feo_impl_ewise(ctx, op, dtype, out, a, nullptr, scalar_host, n);
}
int feo_pow(feo_ctx *ctx, int dtype, void *out, const void *a, const void *power_host, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !power_host) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (dtype != FEO_F32 && dtype != FEO_F64) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "pow needs a Float element type (tensor.rs:325)");
    if (n == 0) return FEO_OK;
    return feo_impl_pow(ctx, dtype, out, a, power_host, n);
}
int feo_ewise_broadcast(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, int rank,
                        const size_t *shape, const size_t *sa, const size_t *sb) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !b) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (op < FEO_ADD || op > FEO_DIV) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown op %d", op);
    if (rank < 0 || rank > FEO_MAX_RANK) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "rank %d > %d", rank, FEO_MAX_RANK);
    if (feo_shape_check(rank, shape) != FEO_OK) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    return feo_impl_broadcast(ctx, op, dtype, out, a, b, rank, shape, sa, sb);
}
int feo_arith_flag(feo_ctx *ctx, int *flag, int clear) {
    FEO_ENTER(ctx);
    int v = 0;
    FEO_CUDA(ctx, cudaMemcpyAsync(&v, ctx->arith_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) *flag = v;
    if (clear) FEO_CUDA(ctx, cudaMemsetAsync(ctx->arith_flag, 0, sizeof(int), ctx->stream));
    return FEO_OK;
}

int feo_peer_status(feo_ctx *ctx, int *timed_out, int clear) {
    FEO_ENTER(ctx);
    if (!ctx) return FEO_ERR_SHAPE;
    int v = 0;
    FEO_CUDA(ctx, cudaMemcpyAsync(&v, ctx->status_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FEO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (timed_out) *timed_out = v & 1;
    if (clear) FEO_CUDA(ctx, cudaMemsetAsync(ctx->status_flag, 0, sizeof(int), ctx->stream));
    if (v & 1) return feo_fail(ctx, FEO_ERR_CUDA, "a kernel gave up waiting for a peer rank after %.0f s: the results of that exchange are invalid",
                               ctx->peer_timeout_ns * 1e-9);
    return FEO_OK;
}

// ---- reductions ----------------------------------------------------------------------------------
static int reduce_args(feo_ctx *ctx, int kind, void *out, const void *in, int rank, const size_t *shape, int naxes,
                       const size_t *axes) {
    if (!ctx || !out || !in) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (kind < FEO_SUM || kind > FEO_MIN) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown reduction %d", kind);
    if (rank < 1 || rank > FEO_MAX_RANK) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "rank %d not in 1..%d", rank, FEO_MAX_RANK);
    if (feo_shape_check(rank, shape) != FEO_OK) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    if (axes_check(rank, naxes, axes) != FEO_OK)
        return feo_fail(ctx, FEO_ERR_SHAPE, "axes must be strictly ascending, unique and < order (the reference panics)");
    return FEO_OK;
}
int feo_reduce(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape, int naxes,
               const size_t *axes) {
    FEO_ENTER(ctx);
    int rc = reduce_args(ctx, kind, out, in, rank, shape, naxes, axes);
    if (rc != FEO_OK) return rc;
    return feo_impl_reduce(ctx, kind, dtype, out, in, nullptr, rank, shape, naxes, axes, 0);
}
int feo_reduce_partial(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape,
                       int naxes, const size_t *axes) {
    FEO_ENTER(ctx);
    int rc = reduce_args(ctx, kind, out, in, rank, shape, naxes, axes);
    if (rc != FEO_OK) return rc;
    return feo_impl_reduce(ctx, kind, dtype, out, in, nullptr, rank, shape, naxes, axes, 1);
}
int feo_reduce_sqdev(feo_ctx *ctx, int dtype, void *out, const void *in, const void *mean, int rank, const size_t *shape,
                     int naxes, const size_t *axes) {
    FEO_ENTER(ctx);
    int rc = reduce_args(ctx, FEO_SUM, out, in, rank, shape, naxes, axes);
    if (rc != FEO_OK) return rc;
    if (!mean) return feo_fail(ctx, FEO_ERR_SHAPE, "null mean");
    return feo_impl_reduce(ctx, FEO_VAR, dtype, out, in, mean, rank, shape, naxes, axes, 2);
}
int feo_var_merge(feo_ctx *ctx, int dtype, void *out, const void *partials, size_t nparts, size_t nout, size_t count_per_part,
                  const void *divisor_host) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !partials || !divisor_host) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (feo_dtype_size(dtype) == 0) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    if (nparts == 0 || nout == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "empty merge");
    return feo_impl_var_merge(ctx, dtype, out, partials, nparts, nout, count_per_part, divisor_host);
}

// ---- products ------------------------------------------------------------------------------------
int feo_outer(feo_ctx *ctx, int dtype, void *out, const void *a, size_t na, const void *b, size_t nb) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !b) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (na == 0 || nb == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    const size_t shape[2] = {na, nb}, sa[2] = {1, 0}, sb[2] = {0, 1};
    return feo_impl_broadcast(ctx, FEO_MUL, dtype, out, a, b, 2, shape, sa, sb);
}
int feo_dot(feo_ctx *ctx, int dtype, void *out, const void *a, const void *b, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !b) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (n == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    return feo_impl_dotlike(ctx, 1, dtype, out, a, b, 1, n);
}
int feo_dot_batched(feo_ctx *ctx, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !a || !b) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (n == 0 || batch == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    return feo_impl_dotlike(ctx, 1, dtype, out, a, b, batch, n);
}
int feo_matvec(feo_ctx *ctx, int dtype, void *out, const void *A, const void *v, size_t M, size_t K) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !A || !v) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (M == 0 || K == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    return feo_impl_dotlike(ctx, 2, dtype, out, A, v, M, K);
}
int feo_vecmat(feo_ctx *ctx, int dtype, void *out, const void *v, const void *A, size_t K, size_t N) {
    FEO_ENTER(ctx);
    if (!ctx || !out || !A || !v) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (N == 0 || K == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    return feo_impl_dotlike(ctx, 3, dtype, out, A, v, K, N);
}
int feo_matmul(feo_ctx *ctx, int algo, int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N) {
    FEO_ENTER(ctx);
    if (!ctx || !C || !A || !B) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (M == 0 || K == 0 || N == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    if (algo == FEO_MATMUL_AUTO)
        algo = (dtype == FEO_F32 && feo_matmul_tf32x3_supported(M, K, N)) ? FEO_MATMUL_TF32X3 : FEO_MATMUL_SIMT;
    if (algo == FEO_MATMUL_TF32X3) {
        if (dtype != FEO_F32) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "TF32X3 is f32-only");
        if (!feo_matmul_tf32x3_supported(M, K, N))
            return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "TF32X3 is used for M*N*K >= 2^21 (smaller products run on the SIMT kernel)");
        return feo_impl_matmul_tf32x3(ctx, C, A, B, M, K, N);
    }
    if (algo != FEO_MATMUL_SIMT) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unknown matmul algo %d", algo);
    return feo_impl_matmul_simt(ctx, dtype, C, A, B, M, K, N);
}

}  // extern "C"
