/* feotensor_b200.h -- C ABI of libfeotensor_b200.so, the B200-native compute backend for the
 * feotensor hot path (elementwise operators, axis reductions, outer product, dot, matmul).
 *
 * The reference (Rust crate feotensor) has NO plugin / FFI seam: its only boundary is the public
 * Rust API on concrete types whose storage field is private (tensor.rs:15-18).  This header is the
 * boundary a `DeviceStorage<T>` placed beside `DynamicStorage<T>` (storage.rs:7-10) binds with
 * `extern "C"`; every entry point names the reference function it replaces (file:line relative to
 * the reference's src/).  INTEGRATION.md shows the Rust-side binding.
 *
 * Conventions
 *  - plain pointers and sizes only; device pointers are raw CUDA device addresses (void*), dense
 *    row-major exactly like DynamicStorage::flatten (storage.rs:18-38);
 *  - every call returns a feo_status; the text of the last failure is feo_last_error(ctx);
 *    FEO_ERR_SHAPE covers the reference's ShapeError cases AND the preconditions the reference
 *    panics on (operator shape mismatch tensor.rs:366, inner-dimension mismatch matrix.rs:78,
 *    axes that are not strictly ascending / in range -- SURVEY.md 8a-quirks);
 *  - compute calls are asynchronous on the context's CUDA stream; feo_download and feo_sync are
 *    the synchronisation points; one host thread per context;
 *  - there is NO CPU fallback: without a CUDA device feo_init fails with FEO_ERR_CUDA.
 */
#ifndef FEOTENSOR_B200_H
#define FEOTENSOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FEO_API __attribute__((visibility("default")))
#else
#define FEO_API
#endif

#define FEO_MAX_RANK 8
#define FEO_ABI_VERSION 1

typedef struct feo_ctx feo_ctx;

typedef enum {
    FEO_OK = 0,
    FEO_ERR_SHAPE = 1,       /* ShapeError (error.rs:3-20) or a reference panic precondition */
    FEO_ERR_CUDA = 2,        /* CUDA runtime / driver failure, or no device */
    FEO_ERR_UNSUPPORTED = 3, /* dtype / algorithm / rank not supported by this build */
    FEO_ERR_ARITH = 4        /* integer x/0 or MIN/-1 (the reference panics): returned at once for a zero scalar divisor, else by
                                the next feo_sync / feo_download after the kernel that met it (which wrote 0 there) */
} feo_status;

/* Element types: any `T: Num + PartialOrd + Copy` primitive the reference admits (tensor.rs:21). */
typedef enum { FEO_F32 = 0, FEO_F64 = 1, FEO_I32 = 2, FEO_I64 = 3, FEO_U32 = 4 } feo_dtype;

/* std::ops::{Add, Sub, Mul, Div} (tensor.rs:336-489). */
typedef enum { FEO_ADD = 0, FEO_SUB = 1, FEO_MUL = 2, FEO_DIV = 3 } feo_op;

/* Tensor::{sum, mean, var, max, min} (tensor.rs:70-307). */
typedef enum { FEO_SUM = 0, FEO_MEAN = 1, FEO_VAR = 2, FEO_MAX = 3, FEO_MIN = 4 } feo_reduce_kind;

/* Matrix::matmul algorithms.  AUTO = TF32X3 for f32 when the shape allows it, else SIMT. */
typedef enum { FEO_MATMUL_AUTO = 0, FEO_MATMUL_SIMT = 1, FEO_MATMUL_TF32X3 = 2 } feo_matmul_algo;

/* ---- context, stream, memory ------------------------------------------------------------------ */
FEO_API int feo_abi_version(void);
FEO_API size_t feo_dtype_size(int dtype);
/* Creates a context on CUDA device `device` with its own non-blocking stream. */
FEO_API int feo_init(int device, feo_ctx **ctx);
FEO_API int feo_shutdown(feo_ctx *ctx);
FEO_API const char *feo_last_error(feo_ctx *ctx);
/* Adopt a caller-owned cudaStream_t (e.g. the host framework's current stream); NULL restores the
 * context's own stream. */
FEO_API int feo_set_stream(feo_ctx *ctx, void *cuda_stream);
FEO_API void *feo_get_stream(feo_ctx *ctx);
FEO_API int feo_sync(feo_ctx *ctx);
/* Number of kernels this context has launched since feo_init (monotonic). */
FEO_API uint64_t feo_launch_count(feo_ctx *ctx);
/* Tuning knobs for experiments; unknown knobs return FEO_ERR_UNSUPPORTED. */
FEO_API int feo_tune(feo_ctx *ctx, int knob, int value);

/* Stream-ordered device allocation: the storage behind DeviceStorage<T> (replaces the Vec<T> of
 * DynamicStorage, storage.rs:7-15). */
FEO_API int feo_malloc(feo_ctx *ctx, size_t bytes, void **dptr);
FEO_API int feo_free(feo_ctx *ctx, void *dptr);
/* feo_malloc / feo_free are stream-ordered on the context's own memory pool, which keeps freed blocks
 * for reuse; feo_trim synchronises the stream and returns the cached blocks to the driver. */
FEO_API int feo_trim(feo_ctx *ctx);
/* Pinned host staging buffers for uploads / downloads. */
FEO_API int feo_host_alloc(feo_ctx *ctx, size_t bytes, void **hptr);
FEO_API int feo_host_free(feo_ctx *ctx, void *hptr);
/* Tensor::new copies host data in (tensor.rs:22-30): asynchronous H2D on the context stream. */
FEO_API int feo_upload(feo_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
/* Reads results back (the reference hands out &T / iterators, storage.rs:40-100): D2H, then waits. */
FEO_API int feo_download(feo_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
/* D2H without the wait (pair with feo_sync). */
FEO_API int feo_download_async(feo_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
FEO_API int feo_copy(feo_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);

/* ---- bookkeeping (host-only, exact integer logic) --------------------------------------------- */
/* Shape::new (shape.rs:12-17): FEO_ERR_SHAPE if any dimension is zero; rank 0 is allowed. */
FEO_API int feo_shape_check(int rank, const size_t *shape);
/* Shape::size (shape.rs:19-21). */
FEO_API size_t feo_shape_size(int rank, const size_t *shape);
/* DynamicStorage::flatten (storage.rs:18-38).  Returns FEO_OK and *index, or FEO_ERR_SHAPE with
 * *detail = -1 for an order mismatch (:19-22) or the first out-of-bounds dimension (:24-30). */
FEO_API int feo_flatten(int coord_rank, const size_t *coord, int rank, const size_t *shape, size_t *index, int *detail);
/* Result shape of a reduction: kept dims ascending (tensor.rs:72-80); [1] when axes is empty or
 * nothing is kept (tensor.rs:84-87).  FEO_ERR_SHAPE unless axes are strictly ascending and in range. */
FEO_API int feo_reduce_shape(int rank, const size_t *shape, int naxes, const size_t *axes, int *out_rank, size_t *out_shape);
/* The divisor `n` of mean / var, built in type T by counting (tensor.rs:112-130, :136-155), in
 * closed form.  *out is one element of `dtype`. */
FEO_API int feo_reduce_divisor(int dtype, int rank, const size_t *shape, int naxes, const size_t *axes, void *out);

/* ---- constructors ----------------------------------------------------------------------------- */
/* Tensor::fill / zeros / ones (tensor.rs:32-44): `value` points at one host element of dtype. */
FEO_API int feo_fill(feo_ctx *ctx, int dtype, void *out, const void *value, size_t n);
/* Matrix::eye (matrix.rs:36-42): zeros, then ones on (i, i) for i < rows; FEO_ERR_SHAPE if rows > cols
 * (the reference panics through an out-of-bounds set). */
FEO_API int feo_eye(feo_ctx *ctx, int dtype, void *out, size_t rows, size_t cols);

/* Synthetic data for benchmarks and full-size parity checks (no reference counterpart): element i
 * gets a value that depends only on (seed, first_index + i) -- a counter-based generator, so sharded
 * and unsharded runs see identical data and the host can regenerate any element.  Floats: uniform in
 * [lo, hi); integers: uniform in [lo, hi) with hi > lo.  oracle/feo_oracle.c holds the host twin. */
FEO_API int feo_fill_uniform(feo_ctx *ctx, int dtype, void *out, size_t n, uint64_t seed, uint64_t first_index, double lo, double hi);

/* ---- elementwise operators -------------------------------------------------------------------- */
/* Add/Sub/Mul/Div<Tensor<T>> for Tensor<T> (tensor.rs:362-373, 405-416, 435-446, 478-489):
 * out[i] = a[i] op b[i], i < n.  The caller has checked shape equality (the reference asserts). */
FEO_API int feo_ewise(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, size_t n);
/* Add/Sub/Mul/Div<T> for Tensor<T> (tensor.rs:336-359, 392-402, 465-475): out[i] = a[i] op *scalar.
 * Integer division by a zero scalar returns FEO_ERR_ARITH without launching anything. */
FEO_API int feo_ewise_scalar(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *scalar_host, size_t n);
/* Tensor::pow (tensor.rs:325-333), floats only: out[i] = a[i].powf(*power).  Exponents 2, 1, 0.5, 0, -1 use the
 * correctly rounded multiply / sqrt / divide (what a correctly rounded powf returns); other exponents use CUDA's
 * powf / pow, i.e. tolerance-level parity with the reference's libm (a few ulp). */
FEO_API int feo_pow(feo_ctx *ctx, int dtype, void *out, const void *a, const void *power_host, size_t n);
/* Broadcasting elementwise (NEW surface): out is dense row-major of `shape`;
 * a_strides / b_strides are element strides of the operands viewed at `shape`, 0 = broadcast.
 * Semantics: materialise both operands, then the reference operator. */
FEO_API int feo_ewise_broadcast(feo_ctx *ctx, int op, int dtype, void *out, const void *a, const void *b, int rank,
                        const size_t *shape, const size_t *a_strides, const size_t *b_strides);
/* Sticky per-context flag set when an integer kernel met x/0 or MIN/-1 (the reference panics; the
 * kernel writes 0 there).  *flag receives it; clear != 0 resets it.  Synchronises the stream.
 * feo_sync and feo_download read (and clear) the same flag and return FEO_ERR_ARITH when it was set, so a
 * host that only ever synchronises through them sees the reference's panic at its next look at the results. */
FEO_API int feo_arith_flag(feo_ctx *ctx, int *flag, int clear);

/* ---- reductions ------------------------------------------------------------------------------- */
/* Tensor::sum / mean / var / max / min (tensor.rs:70-108, 110-132, 134-203, 205-255, 257-307).
 * `in` is dense row-major of `shape`; `out` has feo_reduce_shape elements.  Floats: tree order
 * (reassociation-bounded vs the reference's sequential order), var by Welford/Chan in one pass;
 * integers: bit-exact, var by the reference's own two-pass integer scheme. */
FEO_API int feo_reduce(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape,
               int naxes, const size_t *axes);
/* Building blocks for a tensor sharded on its leading axis:
 *  SUM/MEAN -> per-shard sums, MAX/MIN -> per-shard extrema (combine with an all-reduce);
 *  VAR -> out holds 2 * nout values (combine with feo_var_merge): floats nout means | nout M2;
 *         integers nout wrapping sums | nout wrapping sums of squares (exact: see feo_var_merge). */
FEO_API int feo_reduce_partial(feo_ctx *ctx, int kind, int dtype, void *out, const void *in, int rank,
                       const size_t *shape, int naxes, const size_t *axes);
/* Merges `nparts` VAR partials laid out [part][2 x nout], then divides by *divisor_host (one element
 * of dtype).  Floats: Chan merge of equal-count (mean | M2) partials, out[o] = M2_total[o] / n.
 * Integers: S1 = sum of the wrapping sums, S2 = sum of the wrapping sums of squares; with
 * m = S1 / n (truncating) the reference's two-pass result (tensor.rs:171-202) is exactly
 * (S2 - 2 m S1 + n m^2) / n in wrapping arithmetic (a ring identity modulo 2^k). */
FEO_API int feo_var_merge(feo_ctx *ctx, int dtype, void *out, const void *partials, size_t nparts, size_t nout,
                  size_t count_per_part, const void *divisor_host);
/* Second pass of the integer variance (tensor.rs:196-198): out[o] = sum over the removed axes of
 * (x - mean[o])^2, no division. */
FEO_API int feo_reduce_sqdev(feo_ctx *ctx, int dtype, void *out, const void *in, const void *mean, int rank,
                     const size_t *shape, int naxes, const size_t *axes);

/* ---- products --------------------------------------------------------------------------------- */
/* Tensor::prod (tensor.rs:311-322): out[i * nb + j] = a[i] * b[j]. */
FEO_API int feo_outer(feo_ctx *ctx, int dtype, void *out, const void *a, size_t na, const void *b, size_t nb);
/* DynamicVector::vecmul (vector.rs:71-78): *out = sum_i a[i] * b[i] (one device element). */
FEO_API int feo_dot(feo_ctx *ctx, int dtype, void *out, const void *a, const void *b, size_t n);
/* `batch` independent dots of length n over row-major [batch, n] operands (BASELINE config 5). */
FEO_API int feo_dot_batched(feo_ctx *ctx, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n);
/* DynamicMatrix::vecmul (matrix.rs:92-103): out[i] = sum_j A[i,j] * v[j];  A is [M, K]. */
FEO_API int feo_matvec(feo_ctx *ctx, int dtype, void *out, const void *A, const void *v, size_t M, size_t K);
/* DynamicVector::matmul (vector.rs:80-91): out[j] = sum_i v[i] * A[i,j];  A is [K, N]. */
FEO_API int feo_vecmat(feo_ctx *ctx, int dtype, void *out, const void *v, const void *A, size_t K, size_t N);
/* DynamicMatrix::matmul (matrix.rs:76-90): C[M,N] = A[M,K] * B[K,N], all row-major.  The caller has
 * checked A.cols == B.rows (the reference assert_eq!s, matrix.rs:78). */
FEO_API int feo_matmul(feo_ctx *ctx, int algo, int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N);

/* ---- cross-GPU epilogues over peer memory (one process per GPU, one node) -------- */
/* The exchange step after a sharded reduction / dot, as ONE kernel per rank that signals, waits and
 * combines over NVLink peer loads in rank order (deterministic, bit-identical on every rank) with
 * the operation's epilogue (mean: / n; var: Chan merge, / n) fused in -- instead of an NCCL call.
 * Every rank allocates a window, exports its 64-byte handle, opens the peers' handles (the host's
 * rendezvous carries them; torch.distributed in this repo) and builds a feo_comm from the `world`
 * window addresses.  All ranks must issue the same sequence of comm calls. */
#define FEO_MAX_PEERS 8
#define FEO_PEER_HANDLE_BYTES 64
typedef struct feo_comm feo_comm;
/* Zero-filled, IPC-exportable device memory (plain allocation, not the stream-ordered pool). */
FEO_API int feo_peer_alloc(feo_ctx *ctx, size_t bytes, void **dptr);
FEO_API int feo_peer_free(feo_ctx *ctx, void *dptr);
FEO_API int feo_peer_export(feo_ctx *ctx, void *dptr, unsigned char *handle /* [FEO_PEER_HANDLE_BYTES] */);
/* Maps another process's allocation (enables peer access when it lives on another GPU). */
FEO_API int feo_peer_open(feo_ctx *ctx, const unsigned char *handle, void **dptr);
FEO_API int feo_peer_close(feo_ctx *ctx, void *dptr);
/* windows[r] = address of rank r's window in THIS process (windows[rank] is the local one). */
FEO_API int feo_comm_create(feo_ctx *ctx, int rank, int world, void *const *windows, size_t window_bytes, feo_comm **comm);
FEO_API int feo_comm_destroy(feo_comm *comm);
/* Where this rank's partial for the NEXT exchange goes (inside its own window) and its capacity. */
FEO_API int feo_comm_slot(feo_comm *comm, void **dptr, size_t *bytes);
/* Device-side waits for a peer (exchange flags, the sharded matmul's k-block flags) are bounded by ONE per-context
 * timeout: environment variable FEO_PEER_TIMEOUT_S or feo_tune(ctx, 10, seconds); default 600 s, 0 = wait for ever.
 * A wait that expires does not trap: the kernel raises a sticky status flag and goes on, the context stays usable.
 * feo_peer_status synchronises the stream, stores the flag in *timed_out, clears it if asked, and returns FEO_ERR_CUDA
 * (with a message) when it was set -- the results of that exchange are invalid. */
FEO_API int feo_peer_status(feo_ctx *ctx, int *timed_out, int clear);
/* Stream-ordered barrier across the ranks (no data). */
FEO_API int feo_comm_barrier(feo_comm *comm);
/* Combines the `world` partials sitting in the current slots: SUM / MAX / MIN; MEAN = sum then
 * / *divisor_host; VAR (floats) = Chan merge of [n means | n M2] partials of count_per_rank
 * elements each, then / *divisor_host.  `out` (n elements, local) is identical on every rank. */
FEO_API int feo_comm_allreduce(feo_comm *comm, int kind, int dtype, void *out, size_t n, size_t count_per_rank,
                       const void *divisor_host);
/* Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) of a tensor sharded on its leading axis with
 * axis 0 among the removed axes: `shape` is the LOCAL block, *divisor_host the counted-in-T divisor
 * of the GLOBAL shape (feo_reduce_divisor).  Local one-pass reduction written straight into the
 * window, then the exchange -- by the reduction kernel's own last CTA when the local plan is a single
 * launch with few outputs, else by the exchange kernel (var: one pass and one exchange for floats and
 * integers alike). */
FEO_API int feo_reduce_sharded(feo_comm *comm, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape,
                       int naxes, const size_t *axes, const void *divisor_host);
/* All-gather by pulling over NVLink: out = part 0 | part 1 | ...; parts[r] is rank r's buffer (feo_peer_alloc'ed
 * there) as mapped in THIS process, each bytes_per_part long. */
FEO_API int feo_comm_gather(feo_comm *comm, void *out, const void *const *parts, size_t bytes_per_part);
/* DynamicMatrix::matmul (matrix.rs:76-90) with A row-sharded (this rank's [M,K] block, C likewise) and B row-sharded:
 * b_parts[r] = rank r's [K / world, N] panel in peer-mapped memory.  f32 TF32X3: the panels arrive WHILE the matmul
 * kernel runs and are consumed k-block by k-block in a rotated order (this rank's own panel first, read in place) --
 * pulled piecewise by the copy engines, by the matmul kernel itself (TMA bulk copies), or by a gather kernel on a side
 * stream; otherwise pull-gather into scratch, then the kernel.  Ends with a barrier, so a panel may
 * be overwritten once the call has completed on its owner's stream. */
FEO_API int feo_matmul_sharded_b(feo_comm *comm, int algo, int dtype, void *C, const void *A, const void *const *b_parts, size_t M,
                         size_t K, size_t N);
/* DynamicVector::vecmul (vector.rs:71-78) on identically sharded operands [batch, n_local]. */
FEO_API int feo_dot_sharded(feo_comm *comm, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n_local);

/* ---- the NCCL route (north star: NCCL allreduce after reductions, NCCL allgather of B) ------------- */
/* The same exchanges through NCCL, inside the ABI, so a host that is not Python needs no torch.distributed: rank 0 calls
 * feo_nccl_unique_id and moves the 128 bytes to the other ranks through its own rendezvous, every rank then calls
 * feo_nccl_comm_create (ncclCommInitRank) -- or one process calls feo_nccl_comm_create_all over its G contexts
 * (ncclCommInitAll, the process model of SURVEY 8e) and brackets calls on several communicators with
 * feo_nccl_group_start / feo_nccl_group_end.  libnccl.so.2 is dlopen'ed on first use (FEO_NCCL_LIB overrides the
 * name); without it these return FEO_ERR_UNSUPPORTED.  The peer-memory kernels above are the faster route at the
 * BASELINE sizes; this is the portable one and the baseline they are measured against. */
#define FEO_NCCL_UNIQUE_ID_BYTES 128
typedef struct feo_nccl feo_nccl;
FEO_API int feo_nccl_unique_id(unsigned char *id /* [FEO_NCCL_UNIQUE_ID_BYTES] */);
FEO_API int feo_nccl_comm_create(feo_ctx *ctx, const unsigned char *id, int rank, int world, feo_nccl **comm);
FEO_API int feo_nccl_comm_create_all(feo_ctx *const *ctxs, int ndev, feo_nccl **comms /* [ndev] */);
FEO_API int feo_nccl_comm_destroy(feo_nccl *comm);
FEO_API int feo_nccl_group_start(void);
FEO_API int feo_nccl_group_end(void);
/* In-place ncclAllReduce of n elements on the context's stream; kind = FEO_SUM, FEO_MAX or FEO_MIN. */
FEO_API int feo_nccl_allreduce(feo_nccl *comm, int kind, int dtype, void *buf, size_t n);
/* ncclAllGather: out = part 0 | part 1 | ...; bytes_per_rank is a multiple of 4. */
FEO_API int feo_nccl_allgather(feo_nccl *comm, void *out, const void *in, size_t bytes_per_rank);
/* Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) of a leading-axis shard, leading axis removed: local partial,
 * ncclAllReduce (var: ncclAllGather + the rank-ordered merge of feo_var_merge), epilogue.  Arguments as
 * feo_reduce_sharded.  One process (or thread) per GPU; grouped single-thread hosts call the phases themselves. */
FEO_API int feo_reduce_sharded_nccl(feo_nccl *comm, int kind, int dtype, void *out, const void *in, int rank, const size_t *shape,
                            int naxes, const size_t *axes, const void *divisor_host);
/* DynamicVector::vecmul (vector.rs:71-78) on identically sharded operands: partial dots, then ncclAllReduce. */
FEO_API int feo_dot_sharded_nccl(feo_nccl *comm, int dtype, void *out, const void *a, const void *b, size_t batch, size_t n_local);
/* DynamicMatrix::matmul (matrix.rs:76-90), A and B row-sharded: ncclAllGather of B (b_local = this rank's [K / world, N]
 * panel) into scratch, then the local kernel. */
FEO_API int feo_matmul_sharded_b_nccl(feo_nccl *comm, int algo, int dtype, void *C, const void *A, const void *b_local, size_t M,
                              size_t K, size_t N);

#ifdef __cplusplus
}
#endif
#endif /* FEOTENSOR_B200_H */
