#!/usr/bin/env python
"""Several ranks of the peer-memory exchange (csrc/peer.cu) inside ONE process on ONE GPU.

Each rank has its own feo_ctx (own stream) and window; the windows are plain device addresses, so no IPC is involved.
The host enqueues rank 0, 1, ... in turn: a rank's exchange kernel spins (bounded) until the later ranks' kernels, running
concurrently on their streams, raise their flags.  Checks the sharded reductions and dot against the CPU oracle, that
every rank ends with bit-identical replicated results, and that back-to-back exchanges reuse the two slots safely.

    python tests/multigpu/peer_loopback.py [--world 4]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

# Lazy module loading synchronises the context the first time a kernel is launched; with one rank's exchange kernel already
# spinning for the others that would deadlock this single-process arrangement (separate processes are not affected).
os.environ["CUDA_MODULE_LOADING"] = "EAGER"
# Each rank uses two streams (main + the side stream of the overlapped matmul pre-pass); with more streams than hardware queues
# (8 by default) a rank's pre-pass can be queued behind another rank's waiting kernel.  Again a single-process artefact.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=4)
    ap.add_argument("--only-matmul", action="store_true")
    args = ap.parse_args()
    import feotensor_b200 as feo
    from feotensor_b200 import sharded
    from oracle import feo_oracle as o  # checker only

    world, wbytes = args.world, 1 << 20
    ctxs = [feo.Context(0) for _ in range(world)]
    wins = []
    for c in ctxs:
        p = C.c_void_p()
        c.check(c.lib.feo_peer_alloc(c.handle, wbytes, C.byref(p)))
        wins.append(p.value)
    engines = []
    for r, c in enumerate(ctxs):
        eng = sharded.CudaEngine(c, comm=sharded.PeerComm(c, rank=r, world=world, windows=wins, window_bytes=wbytes))
        c.set_stream(None)  # every rank on its own stream: they must run concurrently
        engines.append(eng)
    rng = np.random.default_rng(5)
    checks = 0

    def gen(dt, shape):
        if dt in ("f32", "f64"):
            return rng.uniform(-1, 1, shape).astype(o.NP_DTYPES[dt])
        return rng.integers(0 if dt == "u32" else -1000, 2000 if dt == "u32" else 1000, shape).astype(o.NP_DTYPES[dt])

    def shard(x, r):
        rows = x.shape[0] // world
        return sharded.ShardedTensor(engines[r], engines[r].from_host(np.ascontiguousarray(x[r * rows:(r + 1) * rows])), list(x.shape), r, world)

    # sharded-B matmul: every rank's B panel is a plain device buffer the others read directly.  (Runs before the reductions:
    # the 197 KB-shared-memory tensor-core kernel after small-carve-out kernels makes the SMs change their L1/shared split,
    # which did not happen while another rank's kernel was spinning -- an artefact of ranks sharing one GPU and process.)
    for dt, algo, (M, K, N) in (("i32", "simt", (8 * world, 16 * world, 24)), ("f64", "simt", (8 * world, 16 * world, 24)),
                                ("f32", "tf32x3", (64 * world, 96 * world, 200)), ("f32", "tf32x3", (32 * world, 33 * world, 257)),
                                ("f32", "tf32x3", (96 * world, 40 * world, 2084))):
        A, B = gen(dt, (M, K)), gen(dt, (K, N))
        kr = K // world
        ptrs = []
        for c in ctxs:
            p_ = C.c_void_p()
            c.check(c.lib.feo_peer_alloc(c.handle, kr * N * B.itemsize, C.byref(p_)))
            ptrs.append(p_.value)
        outs = []
        for r in range(world):
            buf = sharded.PeerBuffer(ctxs[r], kr * N * B.itemsize, rank=r, world=world, ptrs=ptrs)
            ctxs[r].upload(ptrs[r], np.ascontiguousarray(B[r * kr:(r + 1) * kr]).reshape(-1))
            bs = sharded.ShardedTensor(engines[r], buf.tensor((kr, N), B.dtype), [K, N], r, world)
            bs.peer = buf
            outs.append((shard(A, r), bs))
        res = [a.matmul(b, b_sharded=True, algo=algo) for a, b in outs]
        got = np.concatenate([r_.local.to_numpy() for r_ in res], axis=0)
        if dt == "i32":
            np.testing.assert_array_equal(got, o.matmul(A, B))
        elif algo == "simt":
            np.testing.assert_allclose(got, o.matmul(A, B), rtol=1e-11, atol=1e-12)
        else:
            exact = A.astype(np.float64) @ B.astype(np.float64)
            bound = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
            assert np.all(np.abs(got - exact) <= np.maximum(2 * np.abs(o.matmul(A, B) - exact), 2.0 ** -20 * bound)), f"tf32x3 {M}x{K}x{N}"
        for c, q in zip(ctxs, ptrs):
            c.check(c.lib.feo_peer_free(c.handle, C.c_void_p(q)))
        checks += 1
    for dt in (() if args.only_matmul else ("f32", "f64", "i32", "i64", "u32")):
        for shape in ((4 * world, 40, 24), (2 * world, 2999, 5)):
            x = gen(dt, shape)
            shards = [shard(x, r) for r in range(world)]
            for axes in ([0], [0, 2], [0, 1, 2], []):
                kinds = ("sum", "mean", "var", "max", "min")
                # enqueue everything on every rank first (5 exchanges back to back, 6 with the integer var), then read back
                outs = [[s.reduce(k, axes) for k in kinds] for s in shards]
                for ki, kind in enumerate(kinds):
                    got = [outs[r][ki].to_numpy() for r in range(world)]
                    want = o.reduce(kind, x, axes)
                    msg = f"{dt} {shape} {kind} {axes}"
                    for r in range(1, world):
                        np.testing.assert_array_equal(got[r], got[0], err_msg=msg + " (ranks differ)")
                    if dt in ("i32", "i64", "u32") or kind in ("max", "min"):
                        np.testing.assert_array_equal(got[0], want, err_msg=msg)
                    else:
                        tol = dict(rtol=3e-4, atol=2e-5) if dt == "f32" else dict(rtol=1e-11, atol=1e-12)
                        np.testing.assert_allclose(got[0], want, err_msg=msg, **tol)
                    checks += 1
        p, q = gen(dt, (world * 5000,)), gen(dt, (world * 5000,))
        dots = [shard(p, r).dot(shard(q, r)) for r in range(world)]
        got = [d.to_numpy() for d in dots]
        for r in range(1, world):
            np.testing.assert_array_equal(got[r], got[0])
        if dt in ("i32", "i64", "u32"):
            np.testing.assert_array_equal(got[0], o.dot(p, q))
        else:
            np.testing.assert_allclose(got[0], o.dot(p, q), rtol=1e-3 if dt == "f32" else 1e-11, atol=1e-3 if dt == "f32" else 1e-10)
        checks += 1
    launches = sum(c.launch_count() for c in ctxs)
    for e in engines:  # one host thread drives every rank: enqueue all the closing barriers first, then tear down
        e.comm.barrier()
    for e in engines:
        e.comm.close(barrier=False)
    print(f"peer_loopback ok: world={world}, {checks} checks, {launches} kernel launches")


if __name__ == "__main__":
    main()
