"""Leading-axis sharding of the hot path across the GPUs of one node.

One process per GPU (torchrun); `torch.distributed` is the plumbing (NCCL on GPUs, gloo in the CPU tests).
A ShardedTensor is one dense row-major block of leading rows per rank.  Which ops need an exchange:

  elementwise / scalar / broadcast / outer           none (b with leading extent 1 is replicated)
  reduce, leading axis kept                           none (result stays sharded)
  reduce, leading axis removed (or axes == [])        partial per shard, then
      sum / mean      all-reduce(SUM); mean divides by the GLOBAL counted-in-T divisor afterwards
      max / min       all-reduce(MAX / MIN)
      var (floats)    all-gather of per-shard (mean | M2), Chan merge in rank order (deterministic)
      var (ints)      all-gather of per-shard wrapping (sum | sum of squares); m = S1 / n, var = (S2 - 2 m S1 + n m^2) / n
                      in wrapping arithmetic = the reference's two integer passes, exactly
  dot                                                 all-reduce(SUM) of per-shard partial dots
  matmul (A row-sharded)                              none if B is replicated, else all-gather of B

With an `NcclComm` attached (`CudaEngine(nccl=NcclComm(ctx))`) the same NCCL exchanges are issued by the library itself
(`csrc/nccl_route.cu`: feo_reduce_sharded_nccl, feo_dot_sharded_nccl, feo_matmul_sharded_b_nccl) -- the route a host
that is not Python takes; torch.distributed then only carries NCCL's 128-byte unique id at start-up.

With a `PeerComm` attached to the engine (`CudaEngine(comm=PeerComm(...))`) the exchange steps above do not call
NCCL: the local reduction writes its partial straight into a peer-mapped window and ONE kernel per rank signals, waits
and combines the partials over NVLink in rank order with the epilogue (mean's division, var's Chan merge -- or, for
integers, the exact resolution of wrapping sum / sum-of-squares partials) fused in
(`csrc/peer.cu`).  torch.distributed then only carries the 64-byte IPC handles at start-up.

The local arithmetic goes through an *engine*: `CudaEngine` (the C ABI; the only engine this package ships) --
the CPU tests inject their own checker engine to exercise exactly this exchange logic under gloo.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _dist():
    import torch.distributed as dist
    return dist


class PeerComm:
    """The `world` peer-mapped windows of one node and the native exchange kernels over them (csrc/peer.cu).

    `windows=` lets a test hand in raw device addresses (several ranks inside one process on one GPU); the normal
    constructor allocates this rank's window, exchanges CUDA IPC handles through torch.distributed and opens the peers'."""

    def __init__(self, ctx, rank=None, world=None, group=None, window_bytes=4 << 20, windows=None):
        self.ctx, self.lib = ctx, ctx.lib
        self._opened, self._own = [], None
        if windows is None:
            dist = _dist()
            rank = dist.get_rank(group) if rank is None else rank
            world = dist.get_world_size(group) if world is None else world
            p = C.c_void_p()
            ctx.check(self.lib.feo_peer_alloc(ctx.handle, window_bytes, C.byref(p)))
            self._own = p.value
            handle = C.create_string_buffer(L.PEER_HANDLE_BYTES)
            ctx.check(self.lib.feo_peer_export(ctx.handle, p, handle))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            windows = []
            for r in range(world):
                if r == rank:
                    windows.append(self._own)
                    continue
                q = C.c_void_p()
                ctx.check(self.lib.feo_peer_open(ctx.handle, handles[r], C.byref(q)))
                self._opened.append(q.value)
                windows.append(q.value)
            dist.barrier(group=group)  # every window is zeroed and mapped before the first signal
        self.rank, self.world, self.window_bytes = int(rank), int(world), int(window_bytes)
        arr = (C.c_void_p * self.world)(*[C.c_void_p(w) for w in windows])
        h = C.c_void_p()
        ctx.check(self.lib.feo_comm_create(ctx.handle, self.rank, self.world, arr, self.window_bytes, C.byref(h)))
        self.handle = h
        self.slot_bytes = self.slot()[1]

    def slot(self):
        p, n = C.c_void_p(), C.c_size_t()
        self.ctx.check(self.lib.feo_comm_slot(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def barrier(self):
        self.ctx.check(self.lib.feo_comm_barrier(self.handle))

    def close(self, barrier=True):
        """Collective: every rank calls it.  A slower rank may still be inside the last exchange kernel, loading this
        rank's slot or storing its flag into this window, so the ranks meet in one more barrier kernel before anyone
        unmaps or frees anything (barrier=False: the caller has already made sure of that, e.g. several ranks driven by
        one host thread enqueue their barriers first and close afterwards)."""
        if self.handle:
            if barrier:
                self.lib.feo_comm_barrier(self.handle)
            self.ctx.sync()
            self.lib.feo_comm_destroy(self.handle)
            self.handle = None
            for q in self._opened:
                self.lib.feo_peer_close(self.ctx.handle, C.c_void_p(q))
            if self._own:
                self.lib.feo_peer_free(self.ctx.handle, C.c_void_p(self._own))
            self._opened, self._own = [], None


class NcclComm:
    """The library's own NCCL communicator (csrc/nccl_route.cu): the exchange steps as NCCL calls issued INSIDE the C ABI.

    torch.distributed only carries the 128-byte unique id from rank 0 to the others at start-up -- any rendezvous a host
    has would do; afterwards `feo_reduce_sharded_nccl` / `feo_dot_sharded_nccl` / `feo_matmul_sharded_b_nccl` are single
    C calls a Rust or C++ host makes just the same."""

    def __init__(self, ctx, rank=None, world=None, group=None):
        dist = _dist()
        self.ctx, self.lib = ctx, ctx.lib
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.world = dist.get_world_size(group) if world is None else int(world)
        uid = C.create_string_buffer(L.NCCL_UNIQUE_ID_BYTES)
        if self.rank == 0:
            ctx.check(self.lib.feo_nccl_unique_id(uid))
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        h = C.c_void_p()
        ctx.check(self.lib.feo_nccl_comm_create(ctx.handle, box[0], self.rank, self.world, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.lib.feo_nccl_comm_destroy(self.handle)
            self.handle = None


class PeerBuffer:
    """One device allocation per rank that every rank can read: `ptrs[r]` is rank r's block as mapped in this process.

    Holds operands that a fused kernel pulls over NVLink (the row panels of a sharded B, `feo_matmul_sharded_b`).
    `ptrs=` hands in raw addresses (single-process tests); otherwise the handles travel through torch.distributed."""

    def __init__(self, ctx, nbytes, rank=None, world=None, group=None, ptrs=None):
        self.ctx, self.lib, self.nbytes = ctx, ctx.lib, int(nbytes)
        self._opened, self._own = [], None
        if ptrs is None:
            dist = _dist()
            rank = dist.get_rank(group) if rank is None else rank
            world = dist.get_world_size(group) if world is None else world
            p = C.c_void_p()
            ctx.check(self.lib.feo_peer_alloc(ctx.handle, self.nbytes, C.byref(p)))
            self._own = p.value
            handle = C.create_string_buffer(L.PEER_HANDLE_BYTES)
            ctx.check(self.lib.feo_peer_export(ctx.handle, p, handle))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self._own)
                    continue
                q = C.c_void_p()
                ctx.check(self.lib.feo_peer_open(ctx.handle, handles[r], C.byref(q)))
                self._opened.append(q.value)
                ptrs.append(q.value)
        self.rank, self.world, self.ptrs = int(rank), int(world), [int(q) for q in ptrs]

    def tensor(self, dims, dtype):
        """This rank's block viewed as a dense row-major tensor (no copy)."""
        from .core import DeviceStorage, Shape, Tensor
        shp = Shape(list(dims))
        if shp.size() * L.np_dtype(dtype).itemsize > self.nbytes:
            raise ValueError("the buffer is smaller than the tensor")
        return Tensor._wrap(shp, DeviceStorage(self.ctx, shp.size(), dtype, ptr=self.ptrs[self.rank], owner=self))

    def close(self, comm=None):
        """Every rank closes its buffer only once NO rank can still be reading it: pass the PeerComm (a barrier kernel runs
        first), or put a host-side barrier (dist.barrier after a sync) in front of this call."""
        if comm is not None and comm.handle:
            self.lib.feo_comm_barrier(comm.handle)
        self.ctx.sync()
        for q in self._opened:
            self.lib.feo_peer_close(self.ctx.handle, C.c_void_p(q))
        if self._own:
            self.lib.feo_peer_free(self.ctx.handle, C.c_void_p(self._own))
        self._opened, self._own = [], None


class CudaEngine:
    """Local compute on this rank's GPU through libfeotensor_b200 (arrays are core.Tensor)."""

    def __init__(self, ctx=None, comm=None, nccl=None):
        import torch

        from .core import Context
        self.torch = torch
        self.ctx = ctx or Context.default()
        self.comm = comm  # PeerComm: exchanges run as native peer-memory kernels instead of NCCL calls
        self.nccl = nccl  # NcclComm: exchanges are NCCL calls issued inside the C ABI (no torch.distributed on the data path)
        # one stream for the library and for torch's collectives, so NCCL is ordered after the kernels
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    # -- array protocol
    def dims(self, x):
        return list(x.shape().dims)

    def dtype(self, x):
        return x.dtype

    def empty(self, dims, dtype):
        from .core import Shape, Tensor
        return Tensor._empty(Shape(list(dims)), dtype, self.ctx)

    def from_host(self, arr):
        from .core import Shape, Tensor
        return Tensor(Shape(list(arr.shape)), arr, ctx=self.ctx)

    def as_torch(self, x):
        """Zero-copy torch view of the device buffer (for the collectives)."""
        class _CAI:
            pass
        obj = _CAI()
        dt = np.dtype(x.dtype)
        if dt == np.uint32:
            dt = np.dtype(np.int32)  # torch collectives lack uint32; callers handle the order-preserving bias
        obj.__cuda_array_interface__ = {"shape": (x.size(),), "typestr": dt.str, "data": (x.data.ptr, False), "version": 3,
                                        "strides": None}
        t = self.torch.as_tensor(obj, device=f"cuda:{self.ctx.device}")
        t._feo_keepalive = x
        return t

    # -- local arithmetic
    def ewise(self, op, a, b):
        return a._tensor_op(op, b)

    def scalar(self, op, a, s):
        return a._scalar_op(op, s)

    def broadcast(self, op, a, b):
        return a.broadcast(op, b)

    def reduce(self, kind, x, axes):
        return x._reduce(kind, axes)

    def reduce_partial(self, kind, x, axes, nout):
        """SUM/MEAN -> sums, MAX/MIN -> extrema ([nout]); VAR -> [2, nout]: floats means | M2, integers wrapping S1 | S2."""
        ctx, dims = self.ctx, self.dims(x)
        out = self.empty([2, nout] if kind == "var" else [nout], x.dtype)
        ctx.check(ctx.lib.feo_reduce_partial(ctx.handle, L.KIND_CODES[kind], x.data.code, C.c_void_p(out.data.ptr),
                                             C.c_void_p(x.data.ptr), len(dims), L.sizes(dims), len(axes), L.sizes(axes)))
        return out

    def var_merge(self, partials, nparts, nout, count_per_part, divisor):
        ctx = self.ctx
        out = self.empty([nout], partials.dtype)
        d = np.array([divisor], dtype=partials.dtype)
        ctx.check(ctx.lib.feo_var_merge(ctx.handle, partials.data.code, C.c_void_p(out.data.ptr), C.c_void_p(partials.data.ptr),
                                        nparts, nout, count_per_part, d.ctypes.data_as(C.c_void_p)))
        return out

    def reduce_sharded(self, kind, x, local_axes, nout, divisor):
        """Local reduction + peer-memory exchange + epilogue in one stream-ordered native sequence (csrc/peer.cu)."""
        ctx, dims = self.ctx, self.dims(x)
        out = self.empty([nout], x.dtype)
        d = np.array([divisor], dtype=x.dtype)
        ctx.check(ctx.lib.feo_reduce_sharded(self.comm.handle, L.KIND_CODES[kind], x.data.code, C.c_void_p(out.data.ptr),
                                             C.c_void_p(x.data.ptr), len(dims), L.sizes(dims), len(local_axes), L.sizes(local_axes),
                                             d.ctypes.data_as(C.c_void_p)))
        return out

    def reduce_sharded_nccl(self, kind, x, local_axes, nout, divisor):
        """Local reduction + NCCL exchange + epilogue, all inside the C ABI (csrc/nccl_route.cu)."""
        ctx, dims = self.ctx, self.dims(x)
        out = self.empty([nout], x.dtype)
        d = np.array([divisor], dtype=x.dtype)
        ctx.check(ctx.lib.feo_reduce_sharded_nccl(self.nccl.handle, L.KIND_CODES[kind], x.data.code, C.c_void_p(out.data.ptr),
                                                  C.c_void_p(x.data.ptr), len(dims), L.sizes(dims), len(local_axes), L.sizes(local_axes),
                                                  d.ctypes.data_as(C.c_void_p)))
        return out

    def dot_sharded_nccl(self, a, b):
        ctx = self.ctx
        out = self.empty([1], a.dtype)
        ctx.check(ctx.lib.feo_dot_sharded_nccl(self.nccl.handle, a.data.code, C.c_void_p(out.data.ptr), C.c_void_p(a.data.ptr),
                                               C.c_void_p(b.data.ptr), 1, a.size()))
        return out

    def matmul_sharded_b_nccl(self, A, B_local, K, N, algo="auto"):
        ctx, M = self.ctx, self.dims(A)[0]
        out = self.empty([M, N], A.dtype)
        ctx.check(ctx.lib.feo_matmul_sharded_b_nccl(self.nccl.handle, L.ALGO_CODES[algo], A.data.code, C.c_void_p(out.data.ptr),
                                                    C.c_void_p(A.data.ptr), C.c_void_p(B_local.data.ptr), M, K, N))
        return out

    def dot_sharded(self, a, b):
        ctx = self.ctx
        out = self.empty([1], a.dtype)
        ctx.check(ctx.lib.feo_dot_sharded(self.comm.handle, a.data.code, C.c_void_p(out.data.ptr), C.c_void_p(a.data.ptr),
                                          C.c_void_p(b.data.ptr), 1, a.size()))
        return out

    def dot(self, a, b):
        from .core import Vector
        return Vector.from_tensor(a).vecmul(Vector.from_tensor(b))._tensor

    def outer(self, a, b):
        return a.prod(b)

    def matmul(self, A, B, algo="auto"):
        from .core import Matrix
        return Matrix.from_tensor(A).matmul(Matrix.from_tensor(B), algo=algo)._tensor

    def matmul_sharded_b(self, A, b_ptrs, K, N, algo="auto"):
        """A[M,K] (this rank's rows) times a B whose row panels sit in peer memory: they are pulled while the matmul kernel runs."""
        ctx, M = self.ctx, self.dims(A)[0]
        out = self.empty([M, N], A.dtype)
        arr = (C.c_void_p * len(b_ptrs))(*[C.c_void_p(q) for q in b_ptrs])
        ctx.check(ctx.lib.feo_matmul_sharded_b(self.comm.handle, L.ALGO_CODES[algo], A.data.code, C.c_void_p(out.data.ptr),
                                               C.c_void_p(A.data.ptr), arr, M, K, N))
        return out

    def reshape(self, x, dims):
        from .core import Shape, Tensor
        return Tensor._wrap(Shape(list(dims)), x.data)



def reduce_divisor(dtype, global_dims, axes):
    """The reference's counted-in-T divisor for the GLOBAL shape (host-only entry point of the C ABI)."""
    lib = L.load()
    out = np.zeros(1, dtype=L.np_dtype(dtype))
    rc = lib.feo_reduce_divisor(L.code_of(dtype), len(global_dims), L.sizes(global_dims), len(axes), L.sizes(axes),
                                out.ctypes.data_as(C.c_void_p))
    if rc != L.OK:
        raise ValueError("axes must be strictly ascending and in range")
    return out[0]


def reduce_dims(global_dims, axes):
    lib = L.load()
    out_rank, out_shape = C.c_int(0), (C.c_size_t * max(len(global_dims), 1))()
    rc = lib.feo_reduce_shape(len(global_dims), L.sizes(global_dims), len(axes), L.sizes(axes), C.byref(out_rank), out_shape)
    if rc != L.OK:
        raise ValueError("axes must be strictly ascending and in range")
    return [int(out_shape[i]) for i in range(out_rank.value)]


class ShardedTensor:
    """Block `rank` of `world` equal leading-axis blocks of a dense row-major tensor of `global_dims`."""

    def __init__(self, engine, local, global_dims, rank=None, world=None, group=None):
        dist = _dist()
        self.engine = engine
        self.local = local
        self.global_dims = [int(d) for d in global_dims]
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self._plans = {}  # (kind, axes) -> validated shapes / divisor, so repeated reductions skip the host bookkeeping
        if self.global_dims[0] % self.world != 0:
            raise ValueError(f"leading extent {self.global_dims[0]} is not divisible by {self.world} ranks")
        want = [self.global_dims[0] // self.world] + self.global_dims[1:]
        if engine.dims(local) != want:
            raise ValueError(f"local block has dims {engine.dims(local)}, expected {want}")

    @property
    def rows_per_rank(self):
        return self.global_dims[0] // self.world

    def _like(self, local, global_dims=None):
        return ShardedTensor(self.engine, local, global_dims or self.global_dims, self.rank, self.world, self.group)

    # ---- no-communication ops ---------------------------------------------------------------------------
    def ewise(self, op, rhs):
        """Same-shape operator (tensor.rs:362-489); rhs is a ShardedTensor of the same global shape or a scalar."""
        if isinstance(rhs, ShardedTensor):
            if rhs.global_dims != self.global_dims:
                raise ValueError("assertion failed: self.shape == rhs.shape")
            return self._like(self.engine.ewise(op, self.local, rhs.local))
        return self._like(self.engine.scalar(op, self.local, rhs))

    def broadcast(self, op, rhs_replicated):
        """Broadcast against a replicated operand whose leading extent is 1 (config 2: b[1,4096,4096])."""
        rd = self.engine.dims(rhs_replicated)
        if len(rd) != len(self.global_dims) or rd[0] != 1:
            raise ValueError("the replicated operand must have leading extent 1 and the same order")
        out_local = self.engine.broadcast(op, self.local, rhs_replicated)
        od = self.engine.dims(out_local)
        return self._like(out_local, [od[0] * self.world] + od[1:])

    def outer(self, rhs_replicated):
        """Tensor::prod with a replicated right operand (tensor.rs:311-322): the result is sharded like self."""
        out_local = self.engine.outer(self.local, rhs_replicated)
        return self._like(out_local, self.global_dims + self.engine.dims(rhs_replicated))

    def matmul(self, B, b_sharded=False, algo="auto"):
        """Matrix::matmul with A = self row-sharded (matrix.rs:76-90).  B replicated, or row-sharded -> all-gather."""
        eng, dist = self.engine, _dist()
        if len(self.global_dims) != 2:
            raise ValueError("Shape must have order of 2")
        if b_sharded:
            if not isinstance(B, ShardedTensor) or B.global_dims[0] != self.global_dims[1]:
                raise ValueError("assertion `left == right` failed (inner dimensions)")
            peer = getattr(B, "peer", None)
            if getattr(eng, "comm", None) is not None and peer is not None:
                out_local = eng.matmul_sharded_b(self.local, peer.ptrs, B.global_dims[0], B.global_dims[1], algo)
                return self._like(out_local, [self.global_dims[0], B.global_dims[1]])
            if getattr(eng, "nccl", None) is not None:
                out_local = eng.matmul_sharded_b_nccl(self.local, B.local, B.global_dims[0], B.global_dims[1], algo)
                return self._like(out_local, [self.global_dims[0], B.global_dims[1]])
            mine = eng.as_torch(B.local)
            full = eng.empty(B.global_dims, eng.dtype(B.local))
            dist.all_gather(list(eng.as_torch(full).view(self.world, -1).unbind(0)), mine, group=self.group)
            B = full
        elif eng.dims(B)[0] != self.global_dims[1]:
            raise ValueError("assertion `left == right` failed (inner dimensions)")
        out_local = eng.matmul(self.local, B, algo)
        return self._like(out_local, [self.global_dims[0], eng.dims(B)[1]])

    # ---- reductions ------------------------------------------------------------------------------------------
    def reduce(self, kind, axes):
        """Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) on the global tensor.

        Returns a ShardedTensor when the leading axis is kept, else a replicated array (identical on every rank)."""
        eng, dist = self.engine, _dist()
        axes = [int(a) for a in axes]
        plan = self._plans.get((kind, tuple(axes)))
        if plan is None:
            out_dims = reduce_dims(self.global_dims, axes)  # validates the axes like the reference would panic
            dtype = np.dtype(eng.dtype(self.local))
            plan = (out_dims, int(np.prod(out_dims)), dtype, axes if axes else list(range(len(self.global_dims))),
                    reduce_divisor(dtype, self.global_dims, axes))
            self._plans[(kind, tuple(axes))] = plan
        out_dims, nout, dtype, local_axes, n_div = plan  # axes == [] reduces everything (tensor.rs:84): per shard, all local axes
        if axes and 0 not in axes:
            out_local = eng.reduce(kind, self.local, axes)
            return self._like(out_local, out_dims)
        comm = getattr(eng, "comm", None)
        if comm is not None and kind in ("sum", "mean", "var", "max", "min") and \
                nout * dtype.itemsize * (2 if kind == "var" else 1) <= comm.slot_bytes:
            return eng.reshape(eng.reduce_sharded(kind, self.local, local_axes, nout, n_div), out_dims)
        if getattr(eng, "nccl", None) is not None:
            return eng.reshape(eng.reduce_sharded_nccl(kind, self.local, local_axes, nout, n_div), out_dims)

        def allreduce(t, op):
            if dtype == np.uint32 and op != "sum":  # order-preserving bias for the int32 view
                t.bitwise_xor_(-2 ** 31)
            dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op], group=self.group)
            if dtype == np.uint32 and op != "sum":
                t.bitwise_xor_(-2 ** 31)

        if kind in ("sum", "mean", "max", "min"):
            part = eng.reduce_partial(kind, self.local, local_axes, nout)
            allreduce(eng.as_torch(part), "sum" if kind in ("sum", "mean") else kind)
            if kind == "mean":
                part = eng.scalar("div", part, n_div)
            return eng.reshape(part, out_dims)
        if kind != "var":
            raise ValueError(f"unknown reduction {kind}")
        # var: one pass per shard, one exchange.  Floats: (mean | M2) partials, Chan merge in rank order; integers: wrapping
        # (sum | sum of squares) partials, resolved exactly by feo_var_merge (the reference's two passes are ring arithmetic).
        local_count = int(np.prod([eng.dims(self.local)[a] for a in local_axes]))
        part = eng.reduce_partial("var", self.local, local_axes, nout)  # [2, nout]
        gathered = part
        if self.world > 1:
            gathered = eng.empty([self.world, 2, nout], dtype)
            dist.all_gather(list(eng.as_torch(gathered).view(self.world, -1).unbind(0)), eng.as_torch(part), group=self.group)
        merged = eng.var_merge(gathered, self.world, nout, local_count, n_div)
        return eng.reshape(merged, out_dims)

    def dot(self, rhs):
        """Vector::vecmul (vector.rs:71-78) on two identically sharded vectors: a replicated [1] result."""
        eng = self.engine
        if rhs.global_dims != self.global_dims or len(self.global_dims) != 1:
            raise ValueError("assertion failed: self.shape() == rhs.shape()")
        if getattr(eng, "comm", None) is not None:
            return eng.dot_sharded(self.local, rhs.local)
        if getattr(eng, "nccl", None) is not None:
            return eng.dot_sharded_nccl(self.local, rhs.local)
        part = eng.dot(self.local, rhs.local)
        _dist().all_reduce(eng.as_torch(part), group=self.group)
        return part


def shard_rows(engine, full_host: np.ndarray, rank=None, world=None, group=None, peer=False):
    """Upload this rank's block of leading rows of a host array (test / example helper).

    peer=True places the block in a PeerBuffer (readable by every rank; `.peer` on the result)."""
    dist = _dist()
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    rows = full_host.shape[0] // world
    block = np.ascontiguousarray(full_host[rank * rows:(rank + 1) * rows])
    if not peer:
        return ShardedTensor(engine, engine.from_host(block), list(full_host.shape), rank, world, group)
    buf = PeerBuffer(engine.ctx, block.nbytes, rank, world, group)
    local = buf.tensor(block.shape, block.dtype)
    engine.ctx.upload(local.data.ptr, block.reshape(-1))
    st = ShardedTensor(engine, local, list(full_host.shape), rank, world, group)
    st.peer = buf
    return st
