"""ctypes binding of libfeotensor_b200.so (include/feotensor_b200.h).

The library is the product: if it is missing, or no CUDA device is present, everything here fails
loudly -- there is no CPU path and nothing in this package imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfeotensor_b200.so")

OK, ERR_SHAPE, ERR_CUDA, ERR_UNSUPPORTED, ERR_ARITH = 0, 1, 2, 3, 4
F32, F64, I32, I64, U32 = 0, 1, 2, 3, 4
ADD, SUB, MUL, DIV = 0, 1, 2, 3
SUM, MEAN, VAR, MAX, MIN = 0, 1, 2, 3, 4
MATMUL_AUTO, MATMUL_SIMT, MATMUL_TF32X3 = 0, 1, 2
KNOB_BCAST_TILE, KNOB_REDUCE_SPLIT, KNOB_REDUCE_BLOCKS_PER_SM, KNOB_MATMUL_TF32_KC, KNOB_BCAST_LCHUNK, KNOB_PDL, KNOB_BCAST_L2KEEP, KNOB_REDUCE_FUSE, KNOB_GATHER_OVERLAP, KNOB_MATMUL_SIMT, KNOB_PEER_TIMEOUT_S, KNOB_MATMUL_TF32_PAIR, KNOB_MATMUL_TF32_BN, KNOB_MATMUL_TF32_BRAW, KNOB_BCAST_PREFETCH, KNOB_PULL_SHAPE = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15
MAX_RANK = 8

DTYPE_CODES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
               np.dtype(np.int64): I64, np.dtype(np.uint32): U32}
DTYPE_NAMES = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64, "u32": np.uint32}
OP_CODES = {"add": ADD, "sub": SUB, "mul": MUL, "div": DIV}
KIND_CODES = {"sum": SUM, "mean": MEAN, "var": VAR, "max": MAX, "min": MIN}
ALGO_CODES = {"auto": MATMUL_AUTO, "simt": MATMUL_SIMT, "tf32x3": MATMUL_TF32X3}

_vp, _sz, _int = C.c_void_p, C.c_size_t, C.c_int
_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); must list every FEO_API symbol of include/feotensor_b200.h
PEER_HANDLE_BYTES, MAX_PEERS, NCCL_UNIQUE_ID_BYTES = 64, 8, 128

SIGNATURES = {
    "feo_abi_version": (_int, []),
    "feo_dtype_size": (_sz, [_int]),
    "feo_init": (_int, [_int, C.POINTER(_vp)]),
    "feo_shutdown": (_int, [_vp]),
    "feo_last_error": (C.c_char_p, [_vp]),
    "feo_set_stream": (_int, [_vp, _vp]),
    "feo_get_stream": (_vp, [_vp]),
    "feo_sync": (_int, [_vp]),
    "feo_launch_count": (C.c_uint64, [_vp]),
    "feo_tune": (_int, [_vp, _int, _int]),
    "feo_malloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "feo_free": (_int, [_vp, _vp]),
    "feo_trim": (_int, [_vp]),
    "feo_host_alloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "feo_host_free": (_int, [_vp, _vp]),
    "feo_upload": (_int, [_vp, _vp, _vp, _sz]),
    "feo_download": (_int, [_vp, _vp, _vp, _sz]),
    "feo_download_async": (_int, [_vp, _vp, _vp, _sz]),
    "feo_copy": (_int, [_vp, _vp, _vp, _sz]),
    "feo_shape_check": (_int, [_int, _szp]),
    "feo_shape_size": (_sz, [_int, _szp]),
    "feo_flatten": (_int, [_int, _szp, _int, _szp, _szp, C.POINTER(_int)]),
    "feo_reduce_shape": (_int, [_int, _szp, _int, _szp, C.POINTER(_int), _szp]),
    "feo_reduce_divisor": (_int, [_int, _int, _szp, _int, _szp, _vp]),
    "feo_fill": (_int, [_vp, _int, _vp, _vp, _sz]),
    "feo_eye": (_int, [_vp, _int, _vp, _sz, _sz]),
    "feo_fill_uniform": (_int, [_vp, _int, _vp, _sz, C.c_uint64, C.c_uint64, C.c_double, C.c_double]),
    "feo_ewise": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz]),
    "feo_ewise_scalar": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz]),
    "feo_pow": (_int, [_vp, _int, _vp, _vp, _vp, _sz]),
    "feo_ewise_broadcast": (_int, [_vp, _int, _int, _vp, _vp, _vp, _int, _szp, _szp, _szp]),
    "feo_arith_flag": (_int, [_vp, C.POINTER(_int), _int]),
    "feo_reduce": (_int, [_vp, _int, _int, _vp, _vp, _int, _szp, _int, _szp]),
    "feo_reduce_partial": (_int, [_vp, _int, _int, _vp, _vp, _int, _szp, _int, _szp]),
    "feo_var_merge": (_int, [_vp, _int, _vp, _vp, _sz, _sz, _sz, _vp]),
    "feo_reduce_sqdev": (_int, [_vp, _int, _vp, _vp, _vp, _int, _szp, _int, _szp]),
    "feo_outer": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _sz]),
    "feo_dot": (_int, [_vp, _int, _vp, _vp, _vp, _sz]),
    "feo_dot_batched": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz]),
    "feo_matvec": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz]),
    "feo_vecmat": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz]),
    "feo_matmul": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _sz, _sz]),
    "feo_peer_alloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "feo_peer_free": (_int, [_vp, _vp]),
    "feo_peer_export": (_int, [_vp, _vp, C.c_char_p]),
    "feo_peer_open": (_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "feo_peer_close": (_int, [_vp, _vp]),
    "feo_comm_create": (_int, [_vp, _int, _int, C.POINTER(_vp), _sz, C.POINTER(_vp)]),
    "feo_comm_destroy": (_int, [_vp]),
    "feo_comm_slot": (_int, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "feo_peer_status": (_int, [_vp, C.POINTER(_int), _int]),
    "feo_comm_barrier": (_int, [_vp]),
    "feo_comm_allreduce": (_int, [_vp, _int, _int, _vp, _sz, _sz, _vp]),
    "feo_reduce_sharded": (_int, [_vp, _int, _int, _vp, _vp, _int, _szp, _int, _szp, _vp]),
    "feo_comm_gather": (_int, [_vp, _vp, C.POINTER(_vp), _sz]),
    "feo_matmul_sharded_b": (_int, [_vp, _int, _int, _vp, _vp, C.POINTER(_vp), _sz, _sz, _sz]),
    "feo_dot_sharded": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz]),
    "feo_nccl_unique_id": (_int, [C.c_char_p]),
    "feo_nccl_comm_create": (_int, [_vp, C.c_char_p, _int, _int, C.POINTER(_vp)]),
    "feo_nccl_comm_create_all": (_int, [C.POINTER(_vp), _int, C.POINTER(_vp)]),
    "feo_nccl_comm_destroy": (_int, [_vp]),
    "feo_nccl_group_start": (_int, []),
    "feo_nccl_group_end": (_int, []),
    "feo_nccl_allreduce": (_int, [_vp, _int, _int, _vp, _sz]),
    "feo_nccl_allgather": (_int, [_vp, _vp, _vp, _sz]),
    "feo_reduce_sharded_nccl": (_int, [_vp, _int, _int, _vp, _vp, _int, _szp, _int, _szp, _vp]),
    "feo_dot_sharded_nccl": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz]),
    "feo_matmul_sharded_b_nccl": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _sz, _sz]),
}

_lib = None


def load():
    """dlopen the C-ABI library and set every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C feotensor_b200/csrc`).  feotensor_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def sizes(seq):
    seq = [int(s) for s in seq]
    return (C.c_size_t * max(len(seq), 1))(*seq)


def code_of(dtype) -> int:
    dt = np.dtype(DTYPE_NAMES.get(dtype, dtype) if isinstance(dtype, str) else dtype)
    if dt not in DTYPE_CODES:
        raise TypeError(f"unsupported dtype {dt}; supported: f32 f64 i32 i64 u32")
    return DTYPE_CODES[dt]


def np_dtype(dtype):
    return np.dtype(DTYPE_NAMES.get(dtype, dtype) if isinstance(dtype, str) else dtype)
