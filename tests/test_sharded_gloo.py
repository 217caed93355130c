"""World-size-2 gloo test of the leading-axis sharding logic (feotensor_b200/sharded.py) on CPU.

The exchange logic (which op needs which collective, partial layouts, merge order, the GLOBAL counting
divisor) is product code and is exercised for real over gloo.  The *local arithmetic* is injected: a
test-only engine built on the oracle stands in for the CUDA engine, which cannot run here.  Every sharded
result is compared with the oracle applied to the unsharded tensor: bit-exact for integers and max/min,
tolerance for float sums / var.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """Test double for sharded.CudaEngine: numpy arrays, oracle arithmetic."""

    def __init__(self):
        import torch
        from oracle import feo_oracle
        self.torch, self.o = torch, feo_oracle

    def dims(self, x):
        return list(x.shape)

    def dtype(self, x):
        return x.dtype

    def empty(self, dims, dtype):
        return np.empty(dims, dtype=dtype)

    def from_host(self, arr):
        return np.ascontiguousarray(arr).copy()

    def as_torch(self, x):
        flat = x.reshape(-1)
        assert flat.base is not None or flat is x  # a view: the collective must write through to x
        t = self.torch.from_numpy(flat.view(np.int32) if x.dtype == np.uint32 else flat)
        t._feo_keepalive = x
        return t

    def ewise(self, op, a, b):
        return self.o.ewise(op, a, b)

    def scalar(self, op, a, s):
        return self.o.ewise_scalar(op, a, s)

    def broadcast(self, op, a, b):
        return self.o.ewise_broadcast(op, a, b)

    def reduce(self, kind, x, axes):
        return self.o.reduce(kind, x, axes)

    def reduce_partial(self, kind, x, axes, nout):
        if kind == "var":
            ax = tuple(axes)
            if x.dtype.kind != "f":  # integers: wrapping sum | wrapping sum of squares
                with np.errstate(over="ignore"):
                    return np.stack([x.sum(axis=ax, dtype=x.dtype).reshape(-1), (x * x).sum(axis=ax, dtype=x.dtype).reshape(-1)])
            mean = x.astype(np.float64).mean(axis=ax).reshape(-1)
            m2 = ((x.astype(np.float64) - x.astype(np.float64).mean(axis=ax, keepdims=True)) ** 2).sum(axis=ax).reshape(-1)
            return np.stack([mean, m2]).astype(x.dtype)
        return self.o.reduce("sum" if kind == "mean" else kind, x, axes).reshape(-1).copy()

    def var_merge(self, partials, nparts, nout, count, divisor):
        if partials.dtype.kind != "f":
            # exact in the ring Z / 2^k: m = S1 / n (truncating), var = (S2 - 2 m S1 + n m^2) / n; python ints, then wrap
            bits, signed = partials.dtype.itemsize * 8, partials.dtype.kind == "i"

            def wrap(v):
                v &= (1 << bits) - 1
                return v - (1 << bits) if signed and v >= 1 << (bits - 1) else v

            def tdiv(a, b):
                q = abs(a) // abs(b)
                return q if (a >= 0) == (b >= 0) else -q
            p = partials.reshape(nparts, 2, nout)
            n = int(divisor)
            out = np.empty(nout, partials.dtype)
            for o in range(nout):
                s1 = wrap(sum(int(v) for v in p[:, 0, o]))
                s2 = wrap(sum(int(v) for v in p[:, 1, o]))
                m = tdiv(s1, n)
                out[o] = wrap(tdiv(wrap(s2 - 2 * m * s1 + n * m * m), n))
            return out
        p = partials.reshape(nparts, 2, nout).astype(np.float64)
        n, mean, m2 = 0.0, np.zeros(nout), np.zeros(nout)
        for g in range(nparts):  # Chan merge in rank order
            nn = n + count
            d = p[g, 0] - mean
            mean = mean + d * (count / nn)
            m2 = m2 + p[g, 1] + d * d * n * (count / nn)
            n = nn
        return (m2 / float(divisor)).astype(partials.dtype)

    def dot(self, a, b):
        return self.o.dot(a, b)

    def outer(self, a, b):
        return self.o.prod(a, b)

    def matmul(self, A, B, algo="auto"):
        return self.o.matmul(A, B)

    def reshape(self, x, dims):
        return x.reshape(dims)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from feotensor_b200 import sharded
        from oracle import feo_oracle as o
        eng = OracleEngine()
        rng = np.random.default_rng(1234)  # same stream on every rank: identical global tensors

        def gen(dt, shape):
            if dt in ("f32", "f64"):
                return rng.uniform(-1, 1, shape).astype(o.NP_DTYPES[dt])
            return rng.integers(0 if dt == "u32" else -1000, 2000 if dt == "u32" else 1000, shape).astype(o.NP_DTYPES[dt])

        rows = lambda full: full[rank * (full.shape[0] // world):(rank + 1) * (full.shape[0] // world)]
        for dt in ("f32", "f64", "i32", "i64", "u32"):
            x, y = gen(dt, (8, 5, 6)), gen(dt, (8, 5, 6))
            sx, sy = sharded.shard_rows(eng, x), sharded.shard_rows(eng, y)
            # no-communication ops: the local block equals the block of the global result
            np.testing.assert_array_equal(sx.ewise("add", sy).local, rows(o.ewise("add", x, y)))
            np.testing.assert_array_equal(sx.ewise("mul", 3).local, rows(o.ewise_scalar("mul", x, 3)))
            a, b = gen(dt, (8, 1, 6)), gen(dt, (1, 5, 6))
            r = sharded.shard_rows(eng, a).broadcast("mul", b)
            assert r.global_dims == [8, 5, 6]
            np.testing.assert_array_equal(r.local, rows(o.ewise_broadcast("mul", a, b)))
            u, v = gen(dt, (8,)), gen(dt, (7,))
            r = sharded.shard_rows(eng, u).outer(v)
            assert r.global_dims == [8, 7]
            np.testing.assert_array_equal(r.local, rows(o.prod(u, v)))
            # reductions
            for axes in ([1], [2], [1, 2], [0], [0, 2], [0, 1], [0, 1, 2], []):
                for kind in ("sum", "mean", "var", "max", "min"):
                    got = sx.reduce(kind, axes)
                    want = o.reduce(kind, x, axes)
                    msg = f"{dt} {kind} {axes}"
                    if isinstance(got, sharded.ShardedTensor):
                        assert axes and 0 not in axes, msg
                        assert got.global_dims == list(want.shape), msg
                        got, want = got.local, rows(want)
                    else:
                        assert (not axes) or 0 in axes, msg
                        assert list(got.shape) == list(want.shape), msg
                    if dt in ("i32", "i64", "u32") or kind in ("max", "min"):
                        np.testing.assert_array_equal(got, want, err_msg=msg)
                    else:
                        np.testing.assert_allclose(got, want, rtol=2e-4 if dt == "f32" else 1e-11, atol=1e-6 if dt == "f32" else 1e-13, err_msg=msg)
            # dot: all-reduce of partial dots
            p, q = gen(dt, (64,)), gen(dt, (64,))
            got = sharded.shard_rows(eng, p).dot(sharded.shard_rows(eng, q))
            if dt in ("i32", "i64", "u32"):
                np.testing.assert_array_equal(got, o.dot(p, q))
            else:
                np.testing.assert_allclose(got, o.dot(p, q), rtol=1e-4 if dt == "f32" else 1e-12)
            # matmul: A row-sharded; B replicated or row-sharded (all-gather)
            A, B = gen(dt, (8, 6)), gen(dt, (6, 4))
            want = o.matmul(A, B)
            for b_sharded in (False, True):
                Bop = sharded.shard_rows(eng, B) if b_sharded else B
                r = sharded.shard_rows(eng, A).matmul(Bop, b_sharded=b_sharded)
                assert r.global_dims == [8, 4]
                if dt in ("i32", "i64", "u32"):
                    np.testing.assert_array_equal(r.local, rows(want))
                else:
                    np.testing.assert_allclose(r.local, rows(want), rtol=1e-5)
        # the divisor is the GLOBAL counted-in-T one (tensor.rs:112-130), not the shard's
        assert sharded.reduce_divisor("f32", [2 ** 25, 2], [0]) == np.float32(2 ** 24)
        # bad axes are refused like the reference panics; uneven shards are refused
        with pytest.raises(ValueError):
            sx.reduce("sum", [2, 0])
        with pytest.raises(ValueError):
            sharded.ShardedTensor(eng, np.zeros((3, 2), np.float32), [7, 2])
    finally:
        dist.destroy_process_group()


def test_sharded_ops_world2_gloo():
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
