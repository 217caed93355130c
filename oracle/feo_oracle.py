"""ctypes front-end of the CPU oracle (oracle/feo_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package (feotensor_b200) never does.

The oracle restates the reference's algorithms (Rust crate feotensor) in C, in the reference's own
loop and accumulation order; see the header of feo_oracle.c for the citations and the parity pin.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfeo_oracle.so")

DTYPES = {"f32": 0, "f64": 1, "i32": 2, "i64": 3, "u32": 4}
NP_DTYPES = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64, "u32": np.uint32}
OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3}
KINDS = {"sum": 0, "mean": 1, "var": 2, "max": 3, "min": 4}

OK, ERR_SHAPE, ERR_UNSUPPORTED, ERR_ARITH, ERR_NOMEM = 0, 1, 3, 4, 5


class OracleError(Exception):
    def __init__(self, code, what):
        super().__init__(f"oracle {what}: status {code}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (idempotent).  Returns the path of the shared library."""
    src = [os.path.join(_HERE, "feo_oracle.c"), os.path.join(_HERE, "oracle_impl.inc")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.feo_oracle_shape_size.restype = C.c_size_t
        _lib.feo_oracle_index_iterator.restype = C.c_size_t
        _lib.feo_oracle_dtype_size.restype = C.c_size_t
    return _lib


def _code(np_dtype) -> int:
    for k, v in NP_DTYPES.items():
        if np.dtype(np_dtype) == np.dtype(v):
            return DTYPES[k]
    raise OracleError(ERR_UNSUPPORTED, f"dtype {np_dtype}")


def _sz(seq):
    arr = (C.c_size_t * max(len(seq), 1))(*[int(s) for s in seq])
    return arr


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _chk(rc, what):
    if rc != OK:
        raise OracleError(rc, what)


def _flat(a):
    a = np.ascontiguousarray(a)
    return a.reshape(-1), a.shape


# ---- bookkeeping ---------------------------------------------------------------------------------
def shape_check(shape) -> bool:
    """shape.rs:12-17 -- True if Shape::new(shape) is Ok."""
    return lib().feo_oracle_shape_check(len(shape), _sz(shape)) == OK


def shape_size(shape) -> int:
    return int(lib().feo_oracle_shape_size(len(shape), _sz(shape)))


def flatten(coord, shape):
    """storage.rs:18-38.  Returns (status, index): 0 ok, 1 order mismatch, 2+i out of bounds in dim i."""
    idx = C.c_size_t(0)
    rc = lib().feo_oracle_flatten(_sz(coord), len(coord), _sz(shape), len(shape), C.byref(idx))
    return rc, int(idx.value)


def index_iterator(shape):
    """iter.rs:27-47 -- list of coordinates in row-major order."""
    order = len(shape)
    n = shape_size(shape) if order else 0
    buf = (C.c_size_t * max(n * max(order, 1), 1))()
    cnt = lib().feo_oracle_index_iterator(order, _sz(shape), buf)
    return [tuple(buf[i * order + d] for d in range(order)) for i in range(cnt)]


def reduce_shape(shape, axes):
    out = (C.c_size_t * max(len(shape), 1))()
    n = lib().feo_oracle_reduce_shape(len(shape), _sz(shape), len(axes), _sz(axes), out)
    return tuple(int(out[i]) for i in range(n))


def count_in_type(dtype: str, count: int):
    out = np.zeros(1, NP_DTYPES[dtype])
    _chk(lib().feo_oracle_count_in_type(DTYPES[dtype], C.c_size_t(count), _p(out)), "count_in_type")
    return out[0]


def reduce_divisor(dtype: str, shape, axes):
    out = np.zeros(1, NP_DTYPES[dtype])
    _chk(lib().feo_oracle_reduce_divisor(DTYPES[dtype], len(shape), _sz(shape), len(axes), _sz(axes), _p(out)),
         "reduce_divisor")
    return out[0]


# ---- elementwise ---------------------------------------------------------------------------------
def ewise(op: str, a, b, faithful: bool = False):
    """tensor.rs:362-489 (same-shape operators).  Shapes must be equal (the reference asserts)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise OracleError(ERR_SHAPE, "ewise shape/dtype mismatch")
    out = np.empty_like(a)
    fn = lib().feo_oracle_ewise_faithful if faithful else lib().feo_oracle_ewise
    _chk(fn(_code(a.dtype), OPS[op], _p(out), _p(a), _p(b), C.c_size_t(a.size)), "ewise")
    return out


def ewise_scalar(op: str, a, s):
    """tensor.rs:336-359, 392-402, 465-475."""
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    sv = np.array([s], dtype=a.dtype)
    _chk(lib().feo_oracle_ewise_scalar(_code(a.dtype), OPS[op], _p(out), _p(a), _p(sv), C.c_size_t(a.size)),
         "ewise_scalar")
    return out


def broadcast_strides(shape, out_shape):
    """Element strides of a dense row-major operand of `shape` viewed at `out_shape` (0 = broadcast)."""
    if len(shape) != len(out_shape):
        raise OracleError(ERR_SHAPE, "broadcast rank mismatch")
    st, s = [0] * len(shape), 1
    for d in range(len(shape) - 1, -1, -1):
        if shape[d] == out_shape[d]:
            st[d] = s if shape[d] != 1 else 0
        elif shape[d] == 1:
            st[d] = 0
        else:
            raise OracleError(ERR_SHAPE, "not broadcastable")
        s *= shape[d]
    return st


def ewise_broadcast(op: str, a, b):
    """NEW surface (no reference behaviour): materialise both to the broadcast shape, then the operator."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.ndim != b.ndim or a.dtype != b.dtype:
        raise OracleError(ERR_SHAPE, "broadcast rank/dtype mismatch")
    out_shape = tuple(max(x, y) for x, y in zip(a.shape, b.shape))
    sa, sb = broadcast_strides(a.shape, out_shape), broadcast_strides(b.shape, out_shape)
    out = np.empty(out_shape, a.dtype)
    _chk(lib().feo_oracle_ewise_broadcast(_code(a.dtype), OPS[op], _p(out), _p(a), _p(b), a.ndim,
                                          _sz(out_shape), _sz(sa), _sz(sb)), "ewise_broadcast")
    return out


def materialize(a, out_shape):
    a = np.ascontiguousarray(a)
    st = broadcast_strides(a.shape, out_shape)
    out = np.empty(out_shape, a.dtype)
    _chk(lib().feo_oracle_materialize(_code(a.dtype), _p(out), _p(a), a.ndim, _sz(out_shape), _sz(st)), "materialize")
    return out


# ---- reductions ----------------------------------------------------------------------------------
def reduce(kind: str, x, axes, faithful: bool = False, recompute_mean_per_target: bool = False):
    """tensor.rs:70-307.  Returns an array of the reference's result shape ([1] on the scalar path)."""
    x = np.ascontiguousarray(x)
    axes = list(axes)
    shape = x.shape
    if lib().feo_oracle_axes_check(len(shape), len(axes), _sz(axes)) != OK:
        raise OracleError(ERR_SHAPE, "axes must be strictly ascending and in range")
    oshape = reduce_shape(shape, axes)
    out = np.empty(oshape, x.dtype)
    if faithful:
        rc = lib().feo_oracle_reduce_faithful(_code(x.dtype), KINDS[kind], _p(out), _p(x), len(shape), _sz(shape),
                                              len(axes), _sz(axes), int(recompute_mean_per_target))
    else:
        rc = lib().feo_oracle_reduce(_code(x.dtype), KINDS[kind], _p(out), _p(x), len(shape), _sz(shape),
                                     len(axes), _sz(axes))
    _chk(rc, f"reduce {kind}")
    return out


# ---- products ------------------------------------------------------------------------------------
def prod(a, b):
    """tensor.rs:311-322 -- shape = a.shape ++ b.shape."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = np.empty(a.shape + b.shape, a.dtype)
    _chk(lib().feo_oracle_prod(_code(a.dtype), _p(out), _p(a), C.c_size_t(a.size), _p(b), C.c_size_t(b.size)), "prod")
    return out


def dot(a, b):
    """vector.rs:71-78 -- returns a 1-element array (shape [1])."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.ndim != 1:
        raise OracleError(ERR_SHAPE, "dot shape mismatch")
    out = np.empty(1, a.dtype)
    _chk(lib().feo_oracle_dot(_code(a.dtype), _p(out), _p(a), _p(b), C.c_size_t(a.size)), "dot")
    return out


def dot_batched(a, b):
    """One reference dot per row (closest reference analogue: matrix.rs:92-103, one dot per row)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return np.concatenate([dot(a[i], b[i]) for i in range(a.shape[0])])


def matvec(A, v):
    """matrix.rs:92-103."""
    A = np.ascontiguousarray(A)
    v = np.ascontiguousarray(v)
    if A.ndim != 2 or v.ndim != 1 or A.shape[1] != v.shape[0]:
        raise OracleError(ERR_SHAPE, "matvec shape mismatch")
    out = np.empty(A.shape[0], A.dtype)
    _chk(lib().feo_oracle_matvec(_code(A.dtype), _p(out), _p(A), _p(v), C.c_size_t(A.shape[0]),
                                 C.c_size_t(A.shape[1])), "matvec")
    return out


def vecmat(v, A):
    """vector.rs:80-91."""
    A = np.ascontiguousarray(A)
    v = np.ascontiguousarray(v)
    if A.ndim != 2 or v.ndim != 1 or A.shape[0] != v.shape[0]:
        raise OracleError(ERR_SHAPE, "vecmat shape mismatch")
    out = np.empty(A.shape[1], A.dtype)
    _chk(lib().feo_oracle_vecmat(_code(A.dtype), _p(out), _p(v), _p(A), C.c_size_t(A.shape[0]),
                                 C.c_size_t(A.shape[1])), "vecmat")
    return out


def matmul(A, B, faithful: bool = False):
    """matrix.rs:76-90."""
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    if A.ndim != 2 or B.ndim != 2 or A.shape[1] != B.shape[0]:
        raise OracleError(ERR_SHAPE, "matmul shape mismatch")
    out = np.empty((A.shape[0], B.shape[1]), A.dtype)
    fn = lib().feo_oracle_matmul_faithful if faithful else lib().feo_oracle_matmul
    _chk(fn(_code(A.dtype), _p(out), _p(A), _p(B), C.c_size_t(A.shape[0]), C.c_size_t(A.shape[1]),
            C.c_size_t(B.shape[1])), "matmul")
    return out


def matmul_entries(A, B, rows, cols):
    """Selected C[row, col] entries with the reference's k-order (matrix.rs:82-85)."""
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    cols = np.ascontiguousarray(cols, dtype=np.uint64)
    out = np.empty(rows.size, A.dtype)
    _chk(lib().feo_oracle_matmul_entries(_code(A.dtype), _p(out), _p(A), _p(B), C.c_size_t(A.shape[0]),
                                         C.c_size_t(A.shape[1]), C.c_size_t(B.shape[1]), _p(rows), _p(cols),
                                         C.c_size_t(rows.size)), "matmul_entries")
    return out


def pow(a, power):
    """tensor.rs:325-333 (Float only)."""
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    pv = np.array([power], dtype=a.dtype)
    _chk(lib().feo_oracle_pow(_code(a.dtype), _p(out), _p(a), _p(pv), C.c_size_t(a.size)), "pow")
    return out


# ---- synthetic data (twin of the library's feo_fill_uniform; not part of the reference) --------------
def fill_uniform(dtype: str, n: int, seed: int, first_index: int = 0, lo: float = -1.0, hi: float = 1.0, stride: int = 1):
    """Elements first_index, first_index + stride, ... of the counter-based stream `seed`."""
    out = np.empty(n, NP_DTYPES[dtype])
    _chk(lib().feo_oracle_fill_uniform(DTYPES[dtype], _p(out), C.c_size_t(n), C.c_uint64(seed), C.c_uint64(first_index),
                                       C.c_size_t(stride), C.c_double(lo), C.c_double(hi)), "fill_uniform")
    return out
