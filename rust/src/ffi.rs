//! bindgen-style declarations of include/feotensor_b200.h: EVERY FEO_API symbol, generated from the header by
//! tools/gen_rust_ffi.py (tests/test_abi_and_host.py fails when this file and the header drift apart).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct feo_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct feo_comm {
    _private: [u8; 0],
}

#[repr(C)]
pub struct feo_nccl {
    _private: [u8; 0],
}
pub const FEO_NCCL_UNIQUE_ID_BYTES: usize = 128;
pub const FEO_MAX_PEERS: usize = 8;
pub const FEO_PEER_HANDLE_BYTES: usize = 64;

pub const FEO_OK: c_int = 0;
pub const FEO_ERR_SHAPE: c_int = 1;
pub const FEO_ERR_CUDA: c_int = 2;
pub const FEO_ERR_UNSUPPORTED: c_int = 3;
pub const FEO_ERR_ARITH: c_int = 4;

pub const FEO_F32: c_int = 0;
pub const FEO_F64: c_int = 1;
pub const FEO_I32: c_int = 2;
pub const FEO_I64: c_int = 3;
pub const FEO_U32: c_int = 4;

pub const FEO_ADD: c_int = 0;
pub const FEO_SUB: c_int = 1;
pub const FEO_MUL: c_int = 2;
pub const FEO_DIV: c_int = 3;

pub const FEO_SUM: c_int = 0;
pub const FEO_MEAN: c_int = 1;
pub const FEO_VAR: c_int = 2;
pub const FEO_MAX: c_int = 3;
pub const FEO_MIN: c_int = 4;

pub const FEO_MATMUL_AUTO: c_int = 0;
pub const FEO_MATMUL_SIMT: c_int = 1;
pub const FEO_MATMUL_TF32X3: c_int = 2;

extern "C" {
    // BEGIN GENERATED (tools/gen_rust_ffi.py)
    pub fn feo_abi_version() -> c_int;
    pub fn feo_dtype_size(dtype: c_int) -> usize;
    pub fn feo_init(device: c_int, ctx: *mut *mut feo_ctx) -> c_int;
    pub fn feo_shutdown(ctx: *mut feo_ctx) -> c_int;
    pub fn feo_last_error(ctx: *mut feo_ctx) -> *const c_char;
    pub fn feo_set_stream(ctx: *mut feo_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn feo_get_stream(ctx: *mut feo_ctx) -> *mut c_void;
    pub fn feo_sync(ctx: *mut feo_ctx) -> c_int;
    pub fn feo_launch_count(ctx: *mut feo_ctx) -> u64;
    pub fn feo_tune(ctx: *mut feo_ctx, knob: c_int, value: c_int) -> c_int;
    pub fn feo_malloc(ctx: *mut feo_ctx, bytes: usize, dptr: *mut *mut c_void) -> c_int;
    pub fn feo_free(ctx: *mut feo_ctx, dptr: *mut c_void) -> c_int;
    pub fn feo_trim(ctx: *mut feo_ctx) -> c_int;
    pub fn feo_host_alloc(ctx: *mut feo_ctx, bytes: usize, hptr: *mut *mut c_void) -> c_int;
    pub fn feo_host_free(ctx: *mut feo_ctx, hptr: *mut c_void) -> c_int;
    pub fn feo_upload(ctx: *mut feo_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: usize) -> c_int;
    pub fn feo_download(ctx: *mut feo_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    pub fn feo_download_async(ctx: *mut feo_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    pub fn feo_copy(ctx: *mut feo_ctx, dst_dev: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    pub fn feo_shape_check(rank: c_int, shape: *const usize) -> c_int;
    pub fn feo_shape_size(rank: c_int, shape: *const usize) -> usize;
    pub fn feo_flatten(coord_rank: c_int, coord: *const usize, rank: c_int, shape: *const usize, index: *mut usize, detail: *mut c_int) -> c_int;
    pub fn feo_reduce_shape(rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize, out_rank: *mut c_int, out_shape: *mut usize) -> c_int;
    pub fn feo_reduce_divisor(dtype: c_int, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize, out: *mut c_void) -> c_int;
    pub fn feo_fill(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, value: *const c_void, n: usize) -> c_int;
    pub fn feo_eye(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, rows: usize, cols: usize) -> c_int;
    pub fn feo_fill_uniform(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, n: usize, seed: u64, first_index: u64, lo: f64, hi: f64) -> c_int;
    pub fn feo_ewise(ctx: *mut feo_ctx, op: c_int, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, n: usize) -> c_int;
    pub fn feo_ewise_scalar(ctx: *mut feo_ctx, op: c_int, dtype: c_int, out: *mut c_void, a: *const c_void, scalar_host: *const c_void, n: usize) -> c_int;
    pub fn feo_pow(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, a: *const c_void, power_host: *const c_void, n: usize) -> c_int;
    pub fn feo_ewise_broadcast(ctx: *mut feo_ctx, op: c_int, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, rank: c_int, shape: *const usize, a_strides: *const usize, b_strides: *const usize) -> c_int;
    pub fn feo_arith_flag(ctx: *mut feo_ctx, flag: *mut c_int, clear: c_int) -> c_int;
    pub fn feo_reduce(ctx: *mut feo_ctx, kind: c_int, dtype: c_int, out: *mut c_void, input: *const c_void, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize) -> c_int;
    pub fn feo_reduce_partial(ctx: *mut feo_ctx, kind: c_int, dtype: c_int, out: *mut c_void, input: *const c_void, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize) -> c_int;
    pub fn feo_var_merge(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, partials: *const c_void, nparts: usize, nout: usize, count_per_part: usize, divisor_host: *const c_void) -> c_int;
    pub fn feo_reduce_sqdev(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, input: *const c_void, mean: *const c_void, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize) -> c_int;
    pub fn feo_outer(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, a: *const c_void, na: usize, b: *const c_void, nb: usize) -> c_int;
    pub fn feo_dot(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, n: usize) -> c_int;
    pub fn feo_dot_batched(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, batch: usize, n: usize) -> c_int;
    pub fn feo_matvec(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, a: *const c_void, v: *const c_void, m: usize, k: usize) -> c_int;
    pub fn feo_vecmat(ctx: *mut feo_ctx, dtype: c_int, out: *mut c_void, v: *const c_void, a: *const c_void, k: usize, n: usize) -> c_int;
    pub fn feo_matmul(ctx: *mut feo_ctx, algo: c_int, dtype: c_int, c: *mut c_void, a: *const c_void, b: *const c_void, m: usize, k: usize, n: usize) -> c_int;
    pub fn feo_peer_alloc(ctx: *mut feo_ctx, bytes: usize, dptr: *mut *mut c_void) -> c_int;
    pub fn feo_peer_free(ctx: *mut feo_ctx, dptr: *mut c_void) -> c_int;
    pub fn feo_peer_export(ctx: *mut feo_ctx, dptr: *mut c_void, handle: *mut u8) -> c_int;
    pub fn feo_peer_open(ctx: *mut feo_ctx, handle: *const u8, dptr: *mut *mut c_void) -> c_int;
    pub fn feo_peer_close(ctx: *mut feo_ctx, dptr: *mut c_void) -> c_int;
    pub fn feo_comm_create(ctx: *mut feo_ctx, rank: c_int, world: c_int, windows: *const *mut c_void, window_bytes: usize, comm: *mut *mut feo_comm) -> c_int;
    pub fn feo_comm_destroy(comm: *mut feo_comm) -> c_int;
    pub fn feo_comm_slot(comm: *mut feo_comm, dptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
    pub fn feo_peer_status(ctx: *mut feo_ctx, timed_out: *mut c_int, clear: c_int) -> c_int;
    pub fn feo_comm_barrier(comm: *mut feo_comm) -> c_int;
    pub fn feo_comm_allreduce(comm: *mut feo_comm, kind: c_int, dtype: c_int, out: *mut c_void, n: usize, count_per_rank: usize, divisor_host: *const c_void) -> c_int;
    pub fn feo_reduce_sharded(comm: *mut feo_comm, kind: c_int, dtype: c_int, out: *mut c_void, input: *const c_void, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize, divisor_host: *const c_void) -> c_int;
    pub fn feo_comm_gather(comm: *mut feo_comm, out: *mut c_void, parts: *const *const c_void, bytes_per_part: usize) -> c_int;
    pub fn feo_matmul_sharded_b(comm: *mut feo_comm, algo: c_int, dtype: c_int, c: *mut c_void, a: *const c_void, b_parts: *const *const c_void, m: usize, k: usize, n: usize) -> c_int;
    pub fn feo_dot_sharded(comm: *mut feo_comm, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, batch: usize, n_local: usize) -> c_int;
    pub fn feo_nccl_unique_id(id: *mut u8) -> c_int;
    pub fn feo_nccl_comm_create(ctx: *mut feo_ctx, id: *const u8, rank: c_int, world: c_int, comm: *mut *mut feo_nccl) -> c_int;
    pub fn feo_nccl_comm_create_all(ctxs: *const *mut feo_ctx, ndev: c_int, comms: *mut *mut feo_nccl) -> c_int;
    pub fn feo_nccl_comm_destroy(comm: *mut feo_nccl) -> c_int;
    pub fn feo_nccl_group_start() -> c_int;
    pub fn feo_nccl_group_end() -> c_int;
    pub fn feo_nccl_allreduce(comm: *mut feo_nccl, kind: c_int, dtype: c_int, buf: *mut c_void, n: usize) -> c_int;
    pub fn feo_nccl_allgather(comm: *mut feo_nccl, out: *mut c_void, input: *const c_void, bytes_per_rank: usize) -> c_int;
    pub fn feo_reduce_sharded_nccl(comm: *mut feo_nccl, kind: c_int, dtype: c_int, out: *mut c_void, input: *const c_void, rank: c_int, shape: *const usize, naxes: c_int, axes: *const usize, divisor_host: *const c_void) -> c_int;
    pub fn feo_dot_sharded_nccl(comm: *mut feo_nccl, dtype: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, batch: usize, n_local: usize) -> c_int;
    pub fn feo_matmul_sharded_b_nccl(comm: *mut feo_nccl, algo: c_int, dtype: c_int, c: *mut c_void, a: *const c_void, b_local: *const c_void, m: usize, k: usize, n: usize) -> c_int;
    // END GENERATED
}
