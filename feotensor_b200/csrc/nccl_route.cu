// nccl_route.cu -- the NCCL route of the sharded path, inside the C ABI (north star: "reductions finish with an
// NCCL allreduce over NVLink, matmul ... uses an NCCL allgather of B only when B is not already replicated").
//
// A host that is not Python needs no torch.distributed for this: rank 0 asks for a unique id (feo_nccl_unique_id), moves
// the 128 bytes to the other ranks through whatever rendezvous it has, and every rank calls feo_nccl_comm_create -- or, for
// the one-process / G-devices model of SURVEY 8e, feo_nccl_comm_create_all (ncclCommInitAll) with grouped calls.
// libnccl.so.2 is opened at run time (dlopen; FEO_NCCL_LIB overrides the name), so the library has no link-time
// dependency on NCCL and a process that already carries an NCCL (e.g. PyTorch's) shares that copy.
//
// The peer-memory exchange kernels of peer.cu are the library's own (faster at these sizes) route; this one is the
// baseline they are measured against and the portable fallback when peer mapping is unavailable.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "feo_common.cuh"

// ---- the slice of NCCL's stable C ABI used here (nccl.h, unchanged since 2.x) ---------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;      // ncclSuccess = 0
typedef int ncclDataType_t;    // ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8
typedef int ncclRedOp_t;       // ncclSum = 0, ncclMax = 2, ncclMin = 3

struct feo_nccl {
    feo_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *why = "";
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char *name = getenv("FEO_NCCL_LIB");
    void *h = dlopen(name && *name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h && !(name && *name)) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = "libnccl.so.2 could not be opened (set FEO_NCCL_LIB)"; return nullptr; }
#define FEO_NCCL_SYM(field, sym)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));                   \
    if (!api.field) { api.why = "libnccl lacks " sym; dlclose(h); return nullptr; }
    FEO_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    FEO_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    FEO_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    FEO_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FEO_NCCL_SYM(AllReduce, "ncclAllReduce")
    FEO_NCCL_SYM(AllGather, "ncclAllGather")
    FEO_NCCL_SYM(GroupStart, "ncclGroupStart")
    FEO_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FEO_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FEO_NCCL_SYM
    api.handle = h;
    return &api;
}

#define FEO_NCCL(ctx, api, expr)                                                                                  \
    do {                                                                                                          \
        ncclResult_t r_ = (expr);                                                                                 \
        if (r_ != 0) return feo_fail((ctx), FEO_ERR_CUDA, "%s: %s (%s:%d)", #expr, (api)->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

int nccl_dtype(int dtype) {
    switch (dtype) {
    case FEO_I32: return 2;
    case FEO_U32: return 3;
    case FEO_I64: return 4;
    case FEO_F32: return 7;
    case FEO_F64: return 8;
    default: return -1;
    }
}

}  // namespace

extern "C" {

int feo_nccl_unique_id(unsigned char *id) {
    NcclApi *api = nccl_api();
    if (!api || !id) return FEO_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == FEO_NCCL_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != 0) return FEO_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return FEO_OK;
}

int feo_nccl_comm_create(feo_ctx *ctx, const unsigned char *id, int rank, int world, feo_nccl **out) {
    FEO_ENTER(ctx);
    if (!ctx || !id || !out) return feo_fail(ctx, FEO_ERR_SHAPE, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return feo_fail(ctx, FEO_ERR_SHAPE, "rank %d of %d", rank, world);
    NcclApi *api = nccl_api();
    if (!api) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "NCCL is not available: libnccl.so.2 could not be opened or is too old (set FEO_NCCL_LIB)");
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    feo_nccl *c = new feo_nccl();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; return feo_fail(ctx, FEO_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r)); }
    *out = c;
    return FEO_OK;
}

// One process, one host thread, ndev devices (SURVEY.md 8e process-model row): ncclCommInitAll over the contexts' devices.
// Calls on several of these communicators from one thread must sit between feo_nccl_group_start / feo_nccl_group_end.
int feo_nccl_comm_create_all(feo_ctx *const *ctxs, int ndev, feo_nccl **out) {
    if (!ctxs || !out || ndev < 1 || ndev > 64) return FEO_ERR_SHAPE;
    NcclApi *api = nccl_api();
    if (!api) return feo_fail(ctxs[0], FEO_ERR_UNSUPPORTED, "NCCL is not available (libnccl.so.2; set FEO_NCCL_LIB)");
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < ndev; ++i) {
        if (!ctxs[i]) return FEO_ERR_SHAPE;
        devs[i] = ctxs[i]->device;
    }
    ncclResult_t r = api->CommInitAll(comms, ndev, devs);
    if (r != 0) return feo_fail(ctxs[0], FEO_ERR_CUDA, "ncclCommInitAll: %s", api->GetErrorString(r));
    for (int i = 0; i < ndev; ++i) {
        feo_nccl *c = new feo_nccl();
        c->ctx = ctxs[i];
        c->comm = comms[i];
        c->rank = i;
        c->world = ndev;
        out[i] = c;
    }
    return FEO_OK;
}

int feo_nccl_comm_destroy(feo_nccl *c) {
    if (!c) return FEO_OK;
    FEO_ENTER(c->ctx);
    NcclApi *api = nccl_api();
    cudaStreamSynchronize(c->ctx->stream);
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
    return FEO_OK;
}

int feo_nccl_group_start(void) {
    NcclApi *api = nccl_api();
    return api && api->GroupStart() == 0 ? FEO_OK : FEO_ERR_CUDA;
}
int feo_nccl_group_end(void) {
    NcclApi *api = nccl_api();
    return api && api->GroupEnd() == 0 ? FEO_OK : FEO_ERR_CUDA;
}

// In-place all-reduce of n elements on the context's stream: kind = FEO_SUM / FEO_MAX / FEO_MIN.
int feo_nccl_allreduce(feo_nccl *c, int kind, int dtype, void *buf, size_t n) {
    if (!c || !buf) return FEO_ERR_SHAPE;
    feo_ctx *ctx = c->ctx;
    FEO_ENTER(ctx);
    NcclApi *api = nccl_api();
    const int dt = nccl_dtype(dtype);
    if (dt < 0) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    int op;
    switch (kind) {
    case FEO_SUM: op = 0; break;
    case FEO_MAX: op = 2; break;
    case FEO_MIN: op = 3; break;
    default: return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "all-reduce kind %d (SUM, MAX, MIN)", kind);
    }
    FEO_NCCL(ctx, api, api->AllReduce(buf, buf, n, dt, op, c->comm, ctx->stream));
    ctx->window.clear();
    return FEO_OK;
}

// out = part 0 | part 1 | ... ; every rank contributes bytes_per_rank bytes (a multiple of 4).
int feo_nccl_allgather(feo_nccl *c, void *out, const void *in, size_t bytes_per_rank) {
    if (!c || !out || !in) return FEO_ERR_SHAPE;
    feo_ctx *ctx = c->ctx;
    FEO_ENTER(ctx);
    NcclApi *api = nccl_api();
    if (bytes_per_rank == 0 || bytes_per_rank % 4 != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "parts must be a positive multiple of 4 bytes");
    FEO_NCCL(ctx, api, api->AllGather(in, out, bytes_per_rank / 4, /*ncclUint32*/ 3, c->comm, ctx->stream));
    ctx->window.clear();
    return FEO_OK;
}

// Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) of a tensor sharded on its leading axis, the leading axis removed:
// local one-pass partial, NCCL exchange, epilogue.  sum / mean / max / min: ncclAllReduce (mean divides afterwards by the
// GLOBAL counted-in-T divisor); var: ncclAllGather of the per-rank (mean | M2) -- integers (S1 | S2) -- partials and the
// rank-ordered merge of feo_var_merge (deterministic; exact for integers).  One process (or thread) per GPU: inside a
// feo_nccl group the three phases would be enqueued out of order, so grouped hosts call the phases themselves
// (feo_reduce_partial, feo_nccl_allreduce / feo_nccl_allgather, feo_ewise_scalar / feo_var_merge).
int feo_reduce_sharded_nccl(feo_nccl *c, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape, int naxes,
                            const size_t *axes, const void *divisor_host) {
    if (!c || !out || !in || !shape) return FEO_ERR_SHAPE;
    feo_ctx *ctx = c->ctx;
    FEO_ENTER(ctx);
    if (naxes < 1 || !axes || axes[0] != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "the leading (sharded) axis must be reduced; otherwise no exchange is needed");
    if ((kind == FEO_MEAN || kind == FEO_VAR) && !divisor_host) return feo_fail(ctx, FEO_ERR_SHAPE, "mean / var need the divisor");
    int out_rank = 0;
    size_t out_shape[FEO_MAX_RANK];
    if (feo_reduce_shape(rank, shape, naxes, axes, &out_rank, out_shape) != FEO_OK)
        return feo_fail(ctx, FEO_ERR_SHAPE, "axes must be strictly ascending and in range");
    const size_t nout = feo_shape_size(out_rank, out_shape), esz = feo_dtype_size(dtype);
    if (esz == 0) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    int rc;
    if (kind != FEO_VAR) {
        if ((rc = feo_reduce_partial(ctx, kind, dtype, out, in, rank, shape, naxes, axes)) != FEO_OK) return rc;
        if ((rc = feo_nccl_allreduce(c, kind == FEO_MEAN ? FEO_SUM : kind, dtype, out, nout)) != FEO_OK) return rc;
        if (kind == FEO_MEAN) rc = feo_ewise_scalar(ctx, FEO_DIV, dtype, out, out, divisor_host, nout);
        return rc;
    }
    size_t count = 1;
    for (int i = 0; i < naxes; ++i) count *= shape[axes[i]];
    void *scratch = nullptr;  // [world + 1][2 * nout]: the gathered partials, then this rank's own
    FEO_CUDA(ctx, feo_alloc_async(ctx, &scratch, (size_t)(c->world + 1) * 2 * nout * esz));
    char *mine = static_cast<char *>(scratch) + (size_t)c->world * 2 * nout * esz;
    rc = feo_reduce_partial(ctx, FEO_VAR, dtype, mine, in, rank, shape, naxes, axes);
    if (rc == FEO_OK) rc = feo_nccl_allgather(c, scratch, mine, 2 * nout * esz);
    if (rc == FEO_OK) rc = feo_var_merge(ctx, dtype, out, scratch, (size_t)c->world, nout, count, divisor_host);
    cudaFreeAsync(scratch, ctx->stream);
    return rc;
}

// DynamicVector::vecmul (vector.rs:71-78) on identically sharded operands [batch, n_local]: partial dots, ncclAllReduce.
int feo_dot_sharded_nccl(feo_nccl *c, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n_local) {
    if (!c || !out) return FEO_ERR_SHAPE;
    feo_ctx *ctx = c->ctx;
    FEO_ENTER(ctx);
    int rc = feo_dot_batched(ctx, dtype, out, a, b, batch, n_local);
    if (rc != FEO_OK) return rc;
    return feo_nccl_allreduce(c, FEO_SUM, dtype, out, batch);
}

// DynamicMatrix::matmul (matrix.rs:76-90), A row-sharded ([M,K] block, C likewise), B row-sharded: b_local is this rank's
// [K / world, N] panel.  ncclAllGather of B into scratch, then the local kernel (any dtype / algo).
int feo_matmul_sharded_b_nccl(feo_nccl *c, int algo, int dtype, void *C, const void *A, const void *b_local, size_t M, size_t K, size_t N) {
    if (!c || !C || !A || !b_local) return FEO_ERR_SHAPE;
    feo_ctx *ctx = c->ctx;
    FEO_ENTER(ctx);
    if (M == 0 || K == 0 || N == 0) return feo_fail(ctx, FEO_ERR_SHAPE, "Dimension cannot be zero");
    if (K % c->world != 0) return feo_fail(ctx, FEO_ERR_SHAPE, "B's %zu rows do not split evenly over %d ranks", K, c->world);
    const size_t esz = feo_dtype_size(dtype);
    if (esz == 0) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
    void *full = nullptr;
    FEO_CUDA(ctx, feo_alloc_async(ctx, &full, K * N * esz));
    int rc = feo_nccl_allgather(c, full, b_local, K / c->world * N * esz);
    if (rc == FEO_OK) rc = feo_matmul(ctx, algo, dtype, C, A, full, M, K, N);
    cudaFreeAsync(full, ctx->stream);
    return rc;
}

}  // extern "C"
