"""Host-side mirror of feotensor's public API over the B200 C ABI.

Same names, argument meaning and error behaviour as the reference (Rust crate feotensor), so the
parity tests read like the reference's own tests:

    Shape / shape()            shape.rs:6-66        Coordinate / coord()     coordinate.rs:6-73
    IndexIterator              iter.rs:6-48         ShapeError               error.rs:3-20
    DeviceStorage              beside DynamicStorage, storage.rs:7-100
    Tensor / Matrix / Vector   tensor.rs, matrix.rs, vector.rs

Where the reference returns Err(ShapeError) this raises ShapeError; where the reference panics
(operator shape mismatch tensor.rs:366, inner-dimension mismatch matrix.rs:78, bad axes) this raises
ReferencePanic.  All arithmetic runs in the CUDA library; nothing here computes on the CPU and
nothing imports the oracle.  Python cannot move values, so operators do not consume their operands
(the reference's by-value `self`): device buffers are released when the object is collected.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


class ShapeError(Exception):
    """error.rs:3-20 -- Display is "ShapeError: {reason}"."""

    def __init__(self, reason: str):
        super().__init__(f"ShapeError: {reason}")
        self.reason = reason


class ReferencePanic(Exception):
    """Raised where the reference panics (assert!/assert_eq!/unwrap on an Err)."""


class BackendError(RuntimeError):
    """CUDA / library failure (FEO_ERR_CUDA, FEO_ERR_UNSUPPORTED)."""


# ---- context -----------------------------------------------------------------------------------------
class Context:
    """One feo_ctx: a device and a stream.  `Context.default()` is created on first use."""

    _default = None

    def __init__(self, device: int | None = None):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        handle = C.c_void_p()
        rc = self.lib.feo_init(int(device), C.byref(handle))
        if rc != L.OK:
            raise BackendError(
                f"feo_init(device={device}) failed with status {rc}: no usable CUDA device. "
                "feotensor_b200 is GPU-only (no CPU fallback).")
        self.handle = handle
        self.device = int(device)

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context()
        return cls._default

    def check(self, rc: int):
        if rc == L.OK:
            return
        msg = (self.lib.feo_last_error(self.handle) or b"").decode()
        if rc == L.ERR_SHAPE:
            raise ShapeError(msg)
        if rc == L.ERR_ARITH:  # integer x/0 or MIN/-1: the reference panics (at the operator; here at the next sync point)
            raise ReferencePanic(msg)
        raise BackendError(f"status {rc}: {msg}")

    # memory
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.feo_malloc(self.handle, nbytes, C.byref(p)))
        return p.value or 0

    def free(self, ptr: int):
        if ptr and self.handle:
            self.lib.feo_free(self.handle, C.c_void_p(ptr))

    def trim(self):
        """Synchronise and hand the pool's cached (freed) blocks back to the driver."""
        self.check(self.lib.feo_trim(self.handle))

    def upload(self, dst: int, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self.check(self.lib.feo_upload(self.handle, C.c_void_p(dst), host.ctypes.data_as(C.c_void_p), host.nbytes))
        # pageable source: the copy is staged before returning, so `host` may be released

    def download(self, src: int, count: int, dtype) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        self.check(self.lib.feo_download(self.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(src), out.nbytes))
        return out

    def sync(self):
        self.check(self.lib.feo_sync(self.handle))

    def set_stream(self, cuda_stream: int | None):
        self.check(self.lib.feo_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def launch_count(self) -> int:
        return int(self.lib.feo_launch_count(self.handle))

    def tune(self, knob: int, value: int):
        self.check(self.lib.feo_tune(self.handle, knob, value))

    def arith_flag(self, clear: bool = True) -> int:
        f = C.c_int(0)
        self.check(self.lib.feo_arith_flag(self.handle, C.byref(f), int(clear)))
        return f.value

    def close(self):
        if self.handle:
            self.lib.feo_shutdown(self.handle)
            self.handle = None


# ---- bookkeeping types ---------------------------------------------------------------------------------
class Shape:
    """shape.rs:6-56."""

    def __init__(self, dims):
        dims = [int(d) for d in dims]
        if any(d < 0 for d in dims):
            raise ShapeError("Dimension cannot be negative")
        if L.load().feo_shape_check(len(dims), L.sizes(dims)) != L.OK:
            raise ShapeError("Dimension cannot be zero")  # shape.rs:13-15
        self.dims = dims

    def size(self) -> int:
        return int(L.load().feo_shape_size(len(self.dims), L.sizes(self.dims)))

    def order(self) -> int:
        return len(self.dims)

    def stack(self, rhs: "Shape") -> "Shape":
        return Shape(self.dims + rhs.dims)

    def __getitem__(self, i):
        return self.dims[i]

    def __eq__(self, other):
        return isinstance(other, Shape) and self.dims == other.dims

    def __hash__(self):
        return hash(tuple(self.dims))

    def __str__(self):
        return "(" + ", ".join(str(d) for d in self.dims) + ")"

    __repr__ = __str__


def shape(*dims) -> Shape:
    """The shape! macro (shape.rs:58-66)."""
    return Shape(list(dims))


class Coordinate:
    """coordinate.rs:6-56."""

    def __init__(self, indices):
        indices = [int(i) for i in indices]
        if not indices:
            raise ShapeError("Coordinate cannot be empty")  # coordinate.rs:13-15
        self.indices = indices

    def order(self) -> int:
        return len(self.indices)

    def iter(self):
        return iter(self.indices)

    __iter__ = iter

    def insert(self, index: int, axis: int) -> "Coordinate":
        """coordinate.rs:27-33 -- returns a new coordinate with value `axis` inserted at `index`."""
        if index > len(self.indices):
            raise ReferencePanic("insertion index out of bounds")
        new = list(self.indices)
        new.insert(index, int(axis))
        return Coordinate(new)

    def __getitem__(self, i):
        return self.indices[i]

    def __setitem__(self, i, v):
        self.indices[i] = int(v)

    def __eq__(self, other):
        return isinstance(other, Coordinate) and self.indices == other.indices

    def __str__(self):
        return "(" + ", ".join(str(i) for i in self.indices) + ")"

    __repr__ = __str__


def coord(*indices, count: int | None = None) -> Coordinate:
    """The coord! macro (coordinate.rs:58-73): coord(1, 2, 3) or coord(1, count=3)."""
    if count is not None:
        return Coordinate([indices[0]] * count)
    return Coordinate(list(indices))


class IndexIterator:
    """iter.rs:6-48 -- row-major odometer; yields nothing for order-0 shapes."""

    def __init__(self, shp: Shape):
        self.shape = shp
        self.current = [0] * max(shp.order(), 1)
        self.done = False

    def __iter__(self):
        return self

    def __next__(self) -> Coordinate:
        if self.done or self.shape.order() == 0:
            raise StopIteration
        result = Coordinate(self.current)
        for i in range(self.shape.order() - 1, -1, -1):
            if self.current[i] + 1 < self.shape[i]:
                self.current[i] += 1
                break
            self.current[i] = 0
            if i == 0:
                self.done = True
        return result


def _as_shape(s) -> Shape:
    return s if isinstance(s, Shape) else Shape(list(s))


# ---- device storage ----------------------------------------------------------------------------------------
class DeviceStorage:
    """Device-resident dense row-major buffer: the type that sits beside DynamicStorage (storage.rs:7-100)."""

    def __init__(self, ctx: Context, count: int, dtype, ptr: int | None = None, owner=None):
        self.ctx = ctx
        self.dtype = L.np_dtype(dtype)
        self.code = L.code_of(self.dtype)
        self.count = int(count)
        self._owner = owner  # keeps a parent allocation alive for views
        self._owns = ptr is None
        self.ptr = ctx.malloc(self.count * self.dtype.itemsize) if ptr is None else int(ptr)

    @classmethod
    def from_host(cls, ctx: Context, data: np.ndarray) -> "DeviceStorage":
        data = np.ascontiguousarray(data)
        st = cls(ctx, data.size, data.dtype)
        ctx.upload(st.ptr, data.reshape(-1))
        return st

    def flatten(self, coordinate: Coordinate, shp: Shape) -> int:
        """storage.rs:18-38."""
        idx, detail = C.c_size_t(0), C.c_int(0)
        rc = self.ctx.lib.feo_flatten(coordinate.order(), L.sizes(coordinate.indices), shp.order(), L.sizes(shp.dims),
                                      C.byref(idx), C.byref(detail))
        if rc != L.OK:
            if detail.value < 0:
                raise ShapeError(f"incorrect order ({coordinate.order()} vs {shp.order()}).")
            raise ShapeError(f"out of bounds for dimension {detail.value}")
        return int(idx.value)

    def size(self) -> int:
        return self.count

    def to_numpy(self) -> np.ndarray:
        return self.ctx.download(self.ptr, self.count, self.dtype)

    def __getitem__(self, i):
        if isinstance(i, slice):
            lo, hi, step = i.indices(self.count)
            if step != 1:
                raise IndexError("step must be 1")
            return self.ctx.download(self.ptr + lo * self.dtype.itemsize, max(hi - lo, 0), self.dtype)
        if not 0 <= i < self.count:
            raise IndexError("index out of bounds")
        return self.ctx.download(self.ptr + i * self.dtype.itemsize, 1, self.dtype)[0]

    def __setitem__(self, i, value):
        if not 0 <= i < self.count:
            raise IndexError("index out of bounds")
        self.ctx.upload(self.ptr + i * self.dtype.itemsize, np.array([value], dtype=self.dtype))

    def __iter__(self):
        return iter(self.to_numpy())

    def __eq__(self, other):
        if isinstance(other, DeviceStorage):
            return self.dtype == other.dtype and np.array_equal(self.to_numpy(), other.to_numpy())
        return NotImplemented

    def __del__(self):
        try:
            if self._owns and self.ptr:
                self.ctx.free(self.ptr)
                self.ptr = 0
        except Exception:
            pass


def _host_array(data, dtype):
    if isinstance(data, np.ndarray):
        if dtype is not None:
            return np.ascontiguousarray(data, dtype=L.np_dtype(dtype))
        if data.dtype not in L.DTYPE_CODES:
            raise TypeError(f"unsupported dtype {data.dtype}")
        return np.ascontiguousarray(data)
    return np.array(data, dtype=L.np_dtype(dtype if dtype is not None else "f64"))  # Rust float literals are f64


# ---- Tensor ---------------------------------------------------------------------------------------------------
class Tensor:
    """DynamicTensor<T> (tensor.rs:15-19) with device-resident storage."""

    def __init__(self, shp, data, dtype=None, ctx: Context | None = None):
        """Tensor::new (tensor.rs:22-30): copies `data` in; ShapeError on a length mismatch."""
        ctx = ctx or Context.default()
        shp = _as_shape(shp)
        host = _host_array(data, dtype).reshape(-1)
        if host.size != shp.size():
            raise ShapeError("Data length does not match shape size")
        self._shape = shp
        self.data = DeviceStorage.from_host(ctx, host)

    new = classmethod(lambda cls, shp, data, dtype=None, ctx=None: cls(shp, data, dtype, ctx))

    @classmethod
    def _wrap(cls, shp: Shape, storage: DeviceStorage):
        t = cls.__new__(cls)
        t._shape = shp
        t.data = storage
        return t

    @classmethod
    def _empty(cls, shp: Shape, dtype, ctx: Context):
        return cls._wrap(shp, DeviceStorage(ctx, shp.size(), dtype))

    @classmethod
    def fill(cls, shp, value, dtype="f64", ctx: Context | None = None):
        """tensor.rs:32-38."""
        ctx = ctx or Context.default()
        shp = _as_shape(shp)
        t = Tensor._empty(shp, dtype, ctx)
        v = np.array([value], dtype=t.dtype)
        ctx.check(ctx.lib.feo_fill(ctx.handle, t.data.code, C.c_void_p(t.data.ptr), v.ctypes.data_as(C.c_void_p), shp.size()))
        return t

    @classmethod
    def zeros(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 0, dtype, ctx)

    @classmethod
    def ones(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 1, dtype, ctx)

    @classmethod
    def uniform(cls, shp, seed: int, lo=-1.0, hi=1.0, dtype="f32", first_index: int = 0, ctx: Context | None = None):
        """Synthetic data generated on the device: element i = f(seed, first_index + i) (feo_fill_uniform)."""
        ctx = ctx or Context.default()
        shp = _as_shape(shp)
        t = Tensor._empty(shp, dtype, ctx)
        ctx.check(ctx.lib.feo_fill_uniform(ctx.handle, t.data.code, C.c_void_p(t.data.ptr), shp.size(), int(seed),
                                           int(first_index), float(lo), float(hi)))
        return t

    # properties
    @property
    def ctx(self) -> Context:
        return self.data.ctx

    @property
    def dtype(self):
        return self.data.dtype

    def shape(self) -> Shape:
        return self._shape

    def size(self) -> int:
        return self._shape.size()

    def get(self, coordinate: Coordinate):
        """tensor.rs:54-56."""
        return self.data[self.data.flatten(coordinate, self._shape)]

    def set(self, coordinate: Coordinate, value):
        """tensor.rs:63-67."""
        self.data[self.data.flatten(coordinate, self._shape)] = value

    def to_numpy(self) -> np.ndarray:
        return self.data.to_numpy().reshape(self._shape.dims if self._shape.order() else ())

    # reductions (tensor.rs:70-307)
    def _reduce(self, kind: str, axes):
        axes = [int(a) for a in axes]
        ctx, lib = self.ctx, self.ctx.lib
        dims = self._shape.dims
        out_rank, out_shape = C.c_int(0), (C.c_size_t * max(len(dims), 1))()
        rc = lib.feo_reduce_shape(len(dims), L.sizes(dims), len(axes), L.sizes(axes), C.byref(out_rank), out_shape)
        if rc != L.OK:
            raise ReferencePanic(f"axes {axes} are not strictly ascending and < order {len(dims)} (the reference panics)")
        oshape = Shape([out_shape[i] for i in range(out_rank.value)])
        out = Tensor._empty(oshape, self.dtype, ctx)
        ctx.check(lib.feo_reduce(ctx.handle, L.KIND_CODES[kind], self.data.code, C.c_void_p(out.data.ptr),
                                 C.c_void_p(self.data.ptr), len(dims), L.sizes(dims), len(axes), L.sizes(axes)))
        return out

    def sum(self, axes):
        return self._reduce("sum", axes)

    def mean(self, axes):
        return self._reduce("mean", axes)

    def var(self, axes):
        return self._reduce("var", axes)

    def max(self, axes):
        return self._reduce("max", axes)

    def min(self, axes):
        return self._reduce("min", axes)

    def prod(self, other: "Tensor") -> "Tensor":
        """Tensor / outer product (tensor.rs:311-322): shape = self.shape ++ other.shape."""
        other = _tensor_of(other)
        _same_dtype(self, other)
        ctx = self.ctx
        out = Tensor._empty(self._shape.stack(other._shape), self.dtype, ctx)
        ctx.check(ctx.lib.feo_outer(ctx.handle, self.data.code, C.c_void_p(out.data.ptr), C.c_void_p(self.data.ptr),
                                    self.size(), C.c_void_p(other.data.ptr), other.size()))
        return out

    def pow(self, power) -> "Tensor":
        """Elementwise powf, Float element types only (tensor.rs:325-333)."""
        if self.dtype.kind != "f":
            raise TypeError("pow needs a Float element type (tensor.rs:325)")
        ctx = self.ctx
        out = Tensor._empty(self._shape, self.dtype, ctx)
        p = np.array([power], dtype=self.dtype)
        ctx.check(ctx.lib.feo_pow(ctx.handle, self.data.code, C.c_void_p(out.data.ptr), C.c_void_p(self.data.ptr),
                                  p.ctypes.data_as(C.c_void_p), self.size()))
        return out

    def display(self) -> str:
        """Nested-bracket rendering (tensor.rs:507-551); host-side formatting of a downloaded copy."""
        def item(x):
            if self.dtype.kind != "f":
                return str(int(x))
            if np.isnan(x):
                return "NaN"
            if np.isinf(x):
                return "inf" if x > 0 else "-inf"
            return np.format_float_positional(x, unique=True, trim="-")  # Rust's `{}`: shortest round-trip, no exponent

        def fmt(data, dims, level):
            if len(dims) == 1:
                return "[" + ", ".join(item(x) for x in data) + "]"
            sub = int(np.prod(dims[1:]))
            parts = []
            for i in range(dims[0]):
                lead = "" if i == 0 else ",\n" + "\n" * (len(dims) - 2) + " " * level
                parts.append(lead + fmt(data[i * sub:(i + 1) * sub], dims[1:], level + 1))
            return "[" + "".join(parts) + "]"

        return fmt(self.data.to_numpy(), list(self._shape.dims), 1)

    # elementwise operators (tensor.rs:336-505)
    def _scalar_op(self, op: str, rhs):
        ctx = self.ctx
        out = Tensor._empty(self._shape, self.dtype, ctx)
        s = np.array([rhs], dtype=self.dtype)
        ctx.check(ctx.lib.feo_ewise_scalar(ctx.handle, L.OP_CODES[op], self.data.code, C.c_void_p(out.data.ptr),
                                           C.c_void_p(self.data.ptr), s.ctypes.data_as(C.c_void_p), self.size()))
        return out

    def _tensor_op(self, op: str, rhs: "Tensor"):
        if self._shape != rhs._shape:
            raise ReferencePanic(f"assertion failed: self.shape == rhs.shape ({self._shape} vs {rhs._shape})")
        _same_dtype(self, rhs)
        ctx = self.ctx
        out = Tensor._empty(self._shape, self.dtype, ctx)
        ctx.check(ctx.lib.feo_ewise(ctx.handle, L.OP_CODES[op], self.data.code, C.c_void_p(out.data.ptr),
                                    C.c_void_p(self.data.ptr), C.c_void_p(rhs.data.ptr), self.size()))
        return out

    def _binary(self, op: str, rhs):
        if isinstance(rhs, Vector):  # Tensor (op) Vector -> Vector (tensor.rs:375-381, 418-424, 448-454, 491-497)
            if op == "div" and self._shape.order() != 1:
                raise ReferencePanic("called `Result::unwrap()` on an `Err` value: ShapeError: Shape must have order of 1")
            return Vector.from_tensor(self._tensor_op(op, rhs._tensor))
        if isinstance(rhs, Matrix):
            if op == "div" and self._shape.order() != 2:
                raise ReferencePanic("called `Result::unwrap()` on an `Err` value: ShapeError: Shape must have order of 2")
            return Matrix.from_tensor(self._tensor_op(op, rhs._tensor))
        if isinstance(rhs, Tensor):
            return self._tensor_op(op, rhs)
        return self._scalar_op(op, rhs)

    def __add__(self, rhs):
        return self._binary("add", rhs)

    def __sub__(self, rhs):
        return self._binary("sub", rhs)

    def __mul__(self, rhs):
        return self._binary("mul", rhs)

    def __truediv__(self, rhs):
        return self._binary("div", rhs)

    # broadcasting: NEW surface (the reference asserts equal shapes)
    def broadcast(self, op: str, rhs: "Tensor") -> "Tensor":
        """out = self (op) rhs with numpy-style broadcasting of size-1 dims; both operands must have the same order."""
        rhs = _tensor_of(rhs)
        _same_dtype(self, rhs)
        a, b = self._shape.dims, rhs._shape.dims
        if len(a) != len(b):
            raise ShapeError(f"incorrect order ({len(a)} vs {len(b)}).")
        out_dims = []
        for x, y in zip(a, b):
            if x != y and x != 1 and y != 1:
                raise ShapeError(f"shapes {self._shape} and {rhs._shape} are not broadcastable")
            out_dims.append(x if x > y else y)
        sa, sb = _bcast_strides(a, out_dims), _bcast_strides(b, out_dims)
        ctx = self.ctx
        out = Tensor._empty(Shape(out_dims), self.dtype, ctx)
        ctx.check(ctx.lib.feo_ewise_broadcast(ctx.handle, L.OP_CODES[op], self.data.code, C.c_void_p(out.data.ptr),
                                              C.c_void_p(self.data.ptr), C.c_void_p(rhs.data.ptr), len(out_dims),
                                              L.sizes(out_dims), L.sizes(sa), L.sizes(sb)))
        return out


def _bcast_strides(dims, out_dims):
    st, s = [0] * len(dims), 1
    for d in range(len(dims) - 1, -1, -1):
        st[d] = s if (dims[d] == out_dims[d] and dims[d] != 1) else 0
        s *= dims[d]
    return st


def _tensor_of(x) -> Tensor:
    if isinstance(x, (Matrix, Vector)):
        return x._tensor
    return x


def _same_dtype(a: Tensor, b: Tensor):
    if a.dtype != b.dtype:
        raise TypeError(f"mismatched element types {a.dtype} vs {b.dtype} (the reference is generic over one T)")


# ---- newtype wrappers ----------------------------------------------------------------------------------------------
class _Wrapper:
    """Shared plumbing of the Matrix / Vector newtypes: Deref<Target = DynamicTensor> (matrix.rs:216-228)."""

    _ORDER = 0

    @classmethod
    def from_tensor(cls, tensor: Tensor):
        if tensor.shape().order() != cls._ORDER:
            raise ShapeError(f"Shape must have order of {cls._ORDER}")  # matrix.rs:26-28, vector.rs:25-27
        return cls._from_tensor_unchecked(tensor)

    @classmethod
    def _from_tensor_unchecked(cls, tensor: Tensor):
        w = cls.__new__(cls)
        w._tensor = tensor
        return w

    # Deref
    def shape(self):
        return self._tensor.shape()

    def size(self):
        return self._tensor.size()

    def get(self, c):
        return self._tensor.get(c)

    def set(self, c, v):
        return self._tensor.set(c, v)

    def prod(self, other):
        return self._tensor.prod(other)

    def pow(self, power):
        """matrix.rs:106-110 / vector.rs:94-98."""
        return type(self).from_tensor(self._tensor.pow(power))

    def display(self):
        return self._tensor.display()

    def to_numpy(self):
        return self._tensor.to_numpy()

    @property
    def dtype(self):
        return self._tensor.dtype

    @property
    def data(self):
        return self._tensor.data

    @property
    def ctx(self):
        return self._tensor.ctx

    def _binary(self, op, rhs):
        # Matrix/Vector (op) {scalar, same newtype, Tensor} -> same newtype (matrix.rs:113-214, vector.rs:101-202)
        if isinstance(rhs, _Wrapper):
            if type(rhs) is not type(self):
                raise TypeError("no such operator impl in the reference")
            res = self._tensor._tensor_op(op, rhs._tensor)
        elif isinstance(rhs, Tensor):
            res = self._tensor._tensor_op(op, rhs)
        else:
            res = self._tensor._scalar_op(op, rhs)
        return type(self).from_tensor(res)

    def __add__(self, rhs):
        return self._binary("add", rhs)

    def __sub__(self, rhs):
        return self._binary("sub", rhs)

    def __mul__(self, rhs):
        return self._binary("mul", rhs)

    def __truediv__(self, rhs):
        return self._binary("div", rhs)


class Matrix(_Wrapper):
    """DynamicMatrix<T> (matrix.rs:13-16)."""

    _ORDER = 2

    def __init__(self, shp, data, dtype=None, ctx=None):
        """matrix.rs:19-23 -- note: does not check order == 2 (only from_tensor does)."""
        self._tensor = Tensor(shp, data, dtype, ctx)

    new = classmethod(lambda cls, shp, data, dtype=None, ctx=None: cls(shp, data, dtype, ctx))

    @classmethod
    def fill(cls, shp, value, dtype="f64", ctx=None):
        return cls._from_tensor_unchecked(Tensor.fill(shp, value, dtype, ctx))

    @classmethod
    def zeros(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 0, dtype, ctx)

    @classmethod
    def ones(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 1, dtype, ctx)

    @classmethod
    def eye(cls, shp, dtype="f64", ctx=None):
        """matrix.rs:36-42."""
        ctx = ctx or Context.default()
        shp = _as_shape(shp)
        if shp.order() < 2:
            raise ReferencePanic("index out of bounds")
        if shp.order() != 2:
            raise ReferencePanic("called `Result::unwrap()` on an `Err` value: ShapeError: incorrect order")
        if shp[0] > shp[1]:
            raise ReferencePanic("called `Result::unwrap()` on an `Err` value: ShapeError: out of bounds for dimension 1")
        t = Tensor._empty(shp, dtype, ctx)
        ctx.check(ctx.lib.feo_eye(ctx.handle, t.data.code, C.c_void_p(t.data.ptr), shp[0], shp[1]))
        return cls._from_tensor_unchecked(t)

    def sum(self, axes):
        return Vector.from_tensor_or_panic(self._tensor.sum(axes))

    def mean(self, axes):
        return Vector.from_tensor_or_panic(self._tensor.mean(axes))

    def var(self, axes):
        return Vector.from_tensor_or_panic(self._tensor.var(axes))

    def max(self, axes):
        return Vector.from_tensor_or_panic(self._tensor.max(axes))

    def min(self, axes):
        return Vector.from_tensor_or_panic(self._tensor.min(axes))

    def matmul(self, rhs: "Matrix", algo: str = "auto") -> "Matrix":
        """matrix.rs:76-90."""
        if self.shape()[1] != rhs.shape()[0]:
            raise ReferencePanic(f"assertion `left == right` failed\n  left: {self.shape()[1]}\n right: {rhs.shape()[0]}")
        _same_dtype(self._tensor, rhs._tensor)
        ctx = self.ctx
        M, K, N = self.shape()[0], self.shape()[1], rhs.shape()[1]
        out = Tensor._empty(Shape([M, N]), self.dtype, ctx)
        ctx.check(ctx.lib.feo_matmul(ctx.handle, L.ALGO_CODES[algo], self.data.code, C.c_void_p(out.data.ptr),
                                     C.c_void_p(self.data.ptr), C.c_void_p(rhs.data.ptr), M, K, N))
        return Matrix.from_tensor(out)

    def vecmul(self, rhs: "Vector") -> "Vector":
        """matrix.rs:92-103."""
        if self.shape()[1] != rhs.shape()[0]:
            raise ReferencePanic(f"assertion `left == right` failed\n  left: {self.shape()[1]}\n right: {rhs.shape()[0]}")
        _same_dtype(self._tensor, rhs._tensor)
        ctx = self.ctx
        M, K = self.shape()[0], self.shape()[1]
        out = Tensor._empty(Shape([M]), self.dtype, ctx)
        ctx.check(ctx.lib.feo_matvec(ctx.handle, self.data.code, C.c_void_p(out.data.ptr), C.c_void_p(self.data.ptr),
                                     C.c_void_p(rhs.data.ptr), M, K))
        return Vector.from_tensor(out)

    def __getitem__(self, c: Coordinate):
        """Index<Coordinate> (matrix.rs:230-236): unwrap -> panic on a bad coordinate."""
        try:
            return self._tensor.get(c)
        except ShapeError as e:
            raise ReferencePanic(f"called `Result::unwrap()` on an `Err` value: {e}") from None

    def __setitem__(self, c: Coordinate, v):
        try:
            self._tensor.set(c, v)
        except ShapeError as e:
            raise ReferencePanic(f"called `Result::unwrap()` on an `Err` value: {e}") from None


class Vector(_Wrapper):
    """DynamicVector<T> (vector.rs:12-15)."""

    _ORDER = 1

    def __init__(self, data, dtype=None, ctx=None):
        """vector.rs:18-22."""
        host = _host_array(data, dtype).reshape(-1)
        if host.size == 0:
            raise ReferencePanic("called `Result::unwrap()` on an `Err` value: ShapeError: Dimension cannot be zero")
        self._tensor = Tensor(Shape([host.size]), host, None, ctx)

    new = classmethod(lambda cls, data, dtype=None, ctx=None: cls(data, dtype, ctx))

    @classmethod
    def from_tensor_or_panic(cls, tensor: Tensor):
        try:
            return cls.from_tensor(tensor)
        except ShapeError as e:
            raise ReferencePanic(f"called `Result::unwrap()` on an `Err` value: {e}") from None

    @classmethod
    def fill(cls, shp, value, dtype="f64", ctx=None):
        """vector.rs:31-37."""
        shp = _as_shape(shp)
        if shp.order() != 1:
            raise ShapeError("Shape must have order of 1")
        return cls._from_tensor_unchecked(Tensor.fill(shp, value, dtype, ctx))

    @classmethod
    def zeros(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 0, dtype, ctx)

    @classmethod
    def ones(cls, shp, dtype="f64", ctx=None):
        return cls.fill(shp, 1, dtype, ctx)

    # whole-vector reductions always pass axes = [] (vector.rs:45-68)
    def sum(self):
        return Vector.from_tensor(self._tensor.sum([]))

    def mean(self):
        return Vector.from_tensor(self._tensor.mean([]))

    def var(self):
        return Vector.from_tensor(self._tensor.var([]))

    def max(self):
        return Vector.from_tensor(self._tensor.max([]))

    def min(self):
        return Vector.from_tensor(self._tensor.min([]))

    def vecmul(self, rhs: "Vector") -> "Vector":
        """Dot product (vector.rs:71-78): a one-element vector."""
        if self.shape() != rhs.shape():
            raise ReferencePanic("assertion failed: self.shape() == rhs.shape()")
        _same_dtype(self._tensor, rhs._tensor)
        ctx = self.ctx
        out = Tensor._empty(Shape([1]), self.dtype, ctx)
        ctx.check(ctx.lib.feo_dot(ctx.handle, self.data.code, C.c_void_p(out.data.ptr), C.c_void_p(self.data.ptr),
                                  C.c_void_p(rhs.data.ptr), self.size()))
        return Vector.from_tensor(out)

    dot = vecmul  # the name BASELINE.json uses

    def matmul(self, rhs: Matrix) -> "Vector":
        """vector . matrix (vector.rs:80-91)."""
        if self.shape()[0] != rhs.shape()[0]:
            raise ReferencePanic(f"assertion `left == right` failed\n  left: {self.shape()[0]}\n right: {rhs.shape()[0]}")
        _same_dtype(self._tensor, rhs._tensor)
        ctx = self.ctx
        K, N = rhs.shape()[0], rhs.shape()[1]
        out = Tensor._empty(Shape([N]), self.dtype, ctx)
        ctx.check(ctx.lib.feo_vecmat(ctx.handle, self.data.code, C.c_void_p(out.data.ptr), C.c_void_p(self.data.ptr),
                                     C.c_void_p(rhs.data.ptr), K, N))
        return Vector.from_tensor(out)

    def __getitem__(self, i: int):
        """Index<usize> (vector.rs:218-224)."""
        try:
            return self._tensor.get(Coordinate([i]))
        except ShapeError as e:
            raise ReferencePanic(f"called `Result::unwrap()` on an `Err` value: {e}") from None

    def __setitem__(self, i: int, v):
        try:
            self._tensor.set(Coordinate([i]), v)
        except ShapeError as e:
            raise ReferencePanic(f"called `Result::unwrap()` on an `Err` value: {e}") from None


def dot_batched(a: Matrix, b: Matrix) -> Vector:
    """`batch` independent dots over the rows of two [batch, n] matrices (BASELINE config 5)."""
    if a.shape() != b.shape():
        raise ReferencePanic("assertion failed: a.shape() == b.shape()")
    _same_dtype(a._tensor, b._tensor)
    ctx = a.ctx
    out = Tensor._empty(Shape([a.shape()[0]]), a.dtype, ctx)
    ctx.check(ctx.lib.feo_dot_batched(ctx.handle, a.data.code, C.c_void_p(out.data.ptr), C.c_void_p(a.data.ptr),
                                      C.c_void_p(b.data.ptr), a.shape()[0], a.shape()[1]))
    return Vector.from_tensor(out)
