//! Cross-GPU epilogues over peer memory (one process per GPU, one node): the exchange step after a reduction over the
//! sharded leading axis, a dot, or a matmul whose B is row-sharded too -- run by the library's own kernels over NVLink
//! instead of an NCCL call (`csrc/peer.cu`).
//!
//! The crate does not prescribe a rendezvous: `PeerWindow::export()` yields 64 opaque bytes that the host program carries
//! to the other ranks however it likes (MPI, a file, a socket); `PeerComm::new` takes every rank's bytes.
use crate::context::Context;
use crate::device_storage::{dims_of, DeviceElem};
use crate::device_tensor::DeviceTensor;
use crate::ffi;
use feotensor::{Axes, Shape};
use std::os::raw::{c_int, c_void};
use std::rc::Rc;

pub type PeerHandle = [u8; ffi::FEO_PEER_HANDLE_BYTES];

/// Peer-mappable device memory owned by this rank (zero-filled, freed on drop).
pub struct PeerWindow {
    ctx: Rc<Context>,
    ptr: *mut c_void,
    bytes: usize,
}

impl PeerWindow {
    pub fn new(ctx: &Rc<Context>, bytes: usize) -> PeerWindow {
        let mut ptr = std::ptr::null_mut();
        ctx.check(unsafe { ffi::feo_peer_alloc(ctx.raw, bytes, &mut ptr) });
        PeerWindow { ctx: ctx.clone(), ptr, bytes }
    }
    pub fn export(&self) -> PeerHandle {
        let mut h = [0u8; ffi::FEO_PEER_HANDLE_BYTES];
        self.ctx.check(unsafe { ffi::feo_peer_export(self.ctx.raw, self.ptr, h.as_mut_ptr()) });
        h
    }
    pub fn as_ptr(&self) -> *mut c_void { self.ptr }
    pub fn len(&self) -> usize { self.bytes }
}
impl Drop for PeerWindow {
    fn drop(&mut self) {
        unsafe { ffi::feo_peer_free(self.ctx.raw, self.ptr) };
    }
}

/// The `world` windows of one node as seen from this rank, and the exchange kernels over them.
/// Every rank must issue the same sequence of calls, like any collective.
pub struct PeerComm {
    ctx: Rc<Context>,
    raw: *mut ffi::feo_comm,
    own: PeerWindow,
    opened: Vec<*mut c_void>,
    pub rank: usize,
    pub world: usize,
}

impl PeerComm {
    /// `handles[r]` = rank r's `PeerWindow::export()`; `own` is this rank's window (its handle sits at `handles[rank]`).
    /// The caller synchronises the ranks (all windows created and mapped) before the first collective call.
    pub fn new(ctx: &Rc<Context>, rank: usize, own: PeerWindow, handles: &[PeerHandle]) -> PeerComm {
        let world = handles.len();
        let mut windows: Vec<*mut c_void> = Vec::with_capacity(world);
        let mut opened = Vec::new();
        for (r, h) in handles.iter().enumerate() {
            if r == rank {
                windows.push(own.as_ptr());
            } else {
                let mut p = std::ptr::null_mut();
                ctx.check(unsafe { ffi::feo_peer_open(ctx.raw, h.as_ptr(), &mut p) });
                opened.push(p);
                windows.push(p);
            }
        }
        let mut raw = std::ptr::null_mut();
        ctx.check(unsafe { ffi::feo_comm_create(ctx.raw, rank as c_int, world as c_int, windows.as_ptr(), own.len(), &mut raw) });
        PeerComm { ctx: ctx.clone(), raw, own, opened, rank, world }
    }

    /// `Tensor::{sum,mean,var,max,min}` (tensor.rs:70-307) of the GLOBAL tensor whose block of leading rows is `local`;
    /// `axes` contains 0 (otherwise no exchange is needed: call the local method).  `global` is the global shape.
    /// The result is replicated: bit-identical on every rank.
    pub fn reduce<T: DeviceElem>(&self, kind: c_int, local: &DeviceTensor<T>, global: &Shape, axes: Axes) -> DeviceTensor<T> {
        let ldims = dims_of(local.shape());
        let gdims = dims_of(global);
        let (mut out_rank, mut out_dims) = (0 as c_int, vec![0usize; gdims.len().max(1)]);
        let rc = unsafe {
            ffi::feo_reduce_shape(gdims.len() as c_int, gdims.as_ptr(), axes.len() as c_int, axes.as_ptr(), &mut out_rank, out_dims.as_mut_ptr())
        };
        assert!(rc == ffi::FEO_OK, "axes must be strictly ascending, unique and < order");
        out_dims.truncate(out_rank as usize);
        // the reference counts the divisor in T over the GLOBAL shape (tensor.rs:112-130)
        let mut divisor = T::zero();
        unsafe {
            ffi::feo_reduce_divisor(T::DTYPE, gdims.len() as c_int, gdims.as_ptr(), axes.len() as c_int, axes.as_ptr(),
                                    (&mut divisor as *mut T).cast());
        }
        // axes == [] reduces everything (tensor.rs:84): per shard, all local axes
        let local_axes: Vec<usize> = if axes.is_empty() { (0..ldims.len()).collect() } else { axes.clone() };
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(out_dims).unwrap());
        self.ctx.check(unsafe {
            ffi::feo_reduce_sharded(self.raw, kind, T::DTYPE, out.data.ptr, local.data.ptr, ldims.len() as c_int, ldims.as_ptr(),
                                    local_axes.len() as c_int, local_axes.as_ptr(), (&divisor as *const T).cast())
        });
        out
    }

    /// `DynamicVector::vecmul` (vector.rs:71-78) on two identically sharded vectors: a replicated one-element result.
    pub fn dot<T: DeviceElem>(&self, a: &DeviceTensor<T>, b: &DeviceTensor<T>) -> DeviceTensor<T> {
        assert!(a.shape() == b.shape(), "assertion failed: self.shape() == rhs.shape()");
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(vec![1]).unwrap());
        self.ctx.check(unsafe { ffi::feo_dot_sharded(self.raw, T::DTYPE, out.data.ptr, a.data.ptr, b.data.ptr, 1, a.size()) });
        out
    }

    /// `DynamicMatrix::matmul` (matrix.rs:76-90) with A row-sharded (`a_local` = this rank's [M, K] rows) and B row-sharded:
    /// `b_panels[r]` = rank r's [K / world, N] panel, a `PeerWindow` opened here.  The panels are pulled over NVLink while
    /// the matmul kernel runs; returns this rank's [M, N] rows of the product.
    pub fn matmul_sharded_b<T: DeviceElem>(&self, a_local: &DeviceTensor<T>, b_panels: &[*const c_void], k: usize, n: usize) -> DeviceTensor<T> {
        let m = dims_of(a_local.shape())[0];
        assert_eq!(dims_of(a_local.shape())[1], k, "assertion `left == right` failed (inner dimensions)");
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(vec![m, n]).unwrap());
        self.ctx.check(unsafe {
            ffi::feo_matmul_sharded_b(self.raw, ffi::FEO_MATMUL_AUTO, T::DTYPE, out.data.ptr, a_local.data.ptr, b_panels.as_ptr(), m, k, n)
        });
        out
    }

    /// Stream-ordered barrier across the ranks.
    pub fn barrier(&self) {
        self.ctx.check(unsafe { ffi::feo_comm_barrier(self.raw) });
    }
}

impl Drop for PeerComm {
    /// Collective, like every call on the communicator: a slower rank may still be loading this rank's slot or storing its
    /// flag into this window, so the ranks meet in one more barrier kernel before anything is unmapped or freed.
    fn drop(&mut self) {
        unsafe { ffi::feo_comm_barrier(self.raw) };
        self.ctx.sync();
        unsafe {
            ffi::feo_comm_destroy(self.raw);
            for p in &self.opened {
                ffi::feo_peer_close(self.ctx.raw, *p);
            }
        }
        let _ = &self.own; // freed by its own Drop, after the peers' mappings of it are no longer used by this rank
    }
}

/// NCCL's 128-byte unique id: rank 0 creates it, the host's own rendezvous carries it to the other ranks.
pub type NcclId = [u8; ffi::FEO_NCCL_UNIQUE_ID_BYTES];

/// The NCCL route of the same exchanges (csrc/nccl_route.cu): `ncclAllReduce` after a reduction over the sharded leading
/// axis or a dot, `ncclAllGather` of a row-sharded B in front of `matmul` -- issued inside the C ABI (libnccl.so.2 is
/// dlopen'ed by the library), so this crate needs neither NCCL bindings nor a Python process.
pub struct NcclComm {
    ctx: Rc<Context>,
    raw: *mut ffi::feo_nccl,
    pub rank: usize,
    pub world: usize,
}

impl NcclComm {
    pub fn unique_id() -> NcclId {
        let mut id = [0u8; ffi::FEO_NCCL_UNIQUE_ID_BYTES];
        let rc = unsafe { ffi::feo_nccl_unique_id(id.as_mut_ptr()) };
        assert!(rc == ffi::FEO_OK, "feo_nccl_unique_id failed with status {rc}: is libnccl.so.2 on the loader path (FEO_NCCL_LIB)?");
        id
    }

    /// One process (or thread) per GPU: every rank passes the id rank 0 created (`ncclCommInitRank`).
    pub fn new(ctx: &Rc<Context>, id: &NcclId, rank: usize, world: usize) -> NcclComm {
        let mut raw = std::ptr::null_mut();
        ctx.check(unsafe { ffi::feo_nccl_comm_create(ctx.raw, id.as_ptr(), rank as c_int, world as c_int, &mut raw) });
        NcclComm { ctx: ctx.clone(), raw, rank, world }
    }

    /// `Tensor::{sum, mean, var, max, min}` (tensor.rs:70-307) of a leading-axis shard with the leading axis removed.
    pub fn reduce<T: DeviceElem>(&self, kind: c_int, local: &DeviceTensor<T>, global_dims: &[usize], axes: &Axes) -> DeviceTensor<T> {
        let ldims = dims_of(local.shape());
        let mut out_rank: c_int = 0;
        let mut out_dims = vec![0usize; global_dims.len().max(1)];
        let rc = unsafe {
            ffi::feo_reduce_shape(global_dims.len() as c_int, global_dims.as_ptr(), axes.len() as c_int, axes.as_ptr(), &mut out_rank, out_dims.as_mut_ptr())
        };
        assert!(rc == ffi::FEO_OK, "axes must be strictly ascending, unique and < order (the reference panics)");
        out_dims.truncate(out_rank as usize);
        let mut divisor = std::mem::MaybeUninit::<T>::uninit();
        unsafe {
            ffi::feo_reduce_divisor(T::DTYPE, global_dims.len() as c_int, global_dims.as_ptr(), axes.len() as c_int, axes.as_ptr(), divisor.as_mut_ptr().cast());
        }
        let local_axes: Vec<usize> = if axes.is_empty() { (0..ldims.len()).collect() } else { axes.clone() };
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(out_dims).unwrap());
        self.ctx.check(unsafe {
            ffi::feo_reduce_sharded_nccl(self.raw, kind, T::DTYPE, out.data.ptr, local.data.ptr, ldims.len() as c_int, ldims.as_ptr(),
                                         local_axes.len() as c_int, local_axes.as_ptr(), divisor.as_ptr().cast())
        });
        out
    }

    /// `DynamicVector::vecmul` (vector.rs:71-78) on two identically sharded vectors.
    pub fn dot<T: DeviceElem>(&self, a: &DeviceTensor<T>, b: &DeviceTensor<T>) -> DeviceTensor<T> {
        assert!(a.shape() == b.shape(), "assertion failed: self.shape() == rhs.shape()");
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(vec![1]).unwrap());
        self.ctx.check(unsafe { ffi::feo_dot_sharded_nccl(self.raw, T::DTYPE, out.data.ptr, a.data.ptr, b.data.ptr, 1, a.size()) });
        out
    }

    /// `DynamicMatrix::matmul` (matrix.rs:76-90), A and B row-sharded: all-gather of B, then the local kernel.
    pub fn matmul_sharded_b<T: DeviceElem>(&self, a_local: &DeviceTensor<T>, b_local: &DeviceTensor<T>, k: usize, n: usize) -> DeviceTensor<T> {
        let m = dims_of(a_local.shape())[0];
        assert_eq!(dims_of(a_local.shape())[1], k, "assertion `left == right` failed (inner dimensions)");
        let out = DeviceTensor::<T>::uninit(&self.ctx, &Shape::new(vec![m, n]).unwrap());
        self.ctx.check(unsafe {
            ffi::feo_matmul_sharded_b_nccl(self.raw, ffi::FEO_MATMUL_AUTO, T::DTYPE, out.data.ptr, a_local.data.ptr, b_local.data.ptr, m, k, n)
        });
        out
    }
}

impl Drop for NcclComm {
    fn drop(&mut self) {
        unsafe { ffi::feo_nccl_comm_destroy(self.raw) };
    }
}
