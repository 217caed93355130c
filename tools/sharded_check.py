#!/usr/bin/env python
"""Multi-GPU parity of the sharded path (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py

Every rank builds the same global tensors, uploads its block of leading rows, runs the sharded op through
feotensor_b200.sharded.CudaEngine (C ABI kernels + NCCL), and compares with the same op on the whole tensor
computed on its own GPU and with the CPU oracle: ints / max / min bit-exact, float sums within tolerance.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import feotensor_b200 as feo
    from feotensor_b200 import sharded
    from oracle import feo_oracle as o  # checker only

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = feo.Context(local)
    eng = sharded.CudaEngine(ctx)
    rng = np.random.default_rng(99)
    checks = 0

    def gen(dt, shape):
        if dt in ("f32", "f64"):
            return rng.uniform(-1, 1, shape).astype(o.NP_DTYPES[dt])
        return rng.integers(0 if dt == "u32" else -1000, 2000 if dt == "u32" else 1000, shape).astype(o.NP_DTYPES[dt])

    def rows(full):
        r = full.shape[0] // world
        return full[rank * r:(rank + 1) * r]

    R = 8 * world
    for dt in ("f32", "f64", "i32", "i64", "u32"):
        x = gen(dt, (R, 40, 24))
        sx = sharded.shard_rows(eng, x)
        full = feo.Tensor(feo.Shape(list(x.shape)), x, ctx=ctx)
        a, b = gen(dt, (R, 1, 64)), gen(dt, (1, 9, 64))
        got = sharded.shard_rows(eng, a).broadcast("mul", feo.Tensor(feo.shape(1, 9, 64), b, ctx=ctx))
        np.testing.assert_array_equal(got.local.to_numpy(), rows(o.ewise_broadcast("mul", a, b)))
        checks += 1
        for axes in ([1], [1, 2], [0], [0, 2], [0, 1, 2], []):
            for kind in ("sum", "mean", "var", "max", "min"):
                got = sx.reduce(kind, axes)
                single = getattr(full, kind)(axes).to_numpy()
                want = o.reduce(kind, x, axes)
                if isinstance(got, sharded.ShardedTensor):
                    g, single, want = got.local.to_numpy(), rows(single), rows(want)
                else:
                    g = got.to_numpy()
                msg = f"{dt} {kind} {axes}"
                if dt in ("i32", "i64", "u32") or kind in ("max", "min"):
                    np.testing.assert_array_equal(g, want, err_msg=msg)
                    np.testing.assert_array_equal(g, single, err_msg=msg)
                else:
                    tol = dict(rtol=3e-4, atol=2e-5) if dt == "f32" else dict(rtol=1e-11, atol=1e-12)
                    np.testing.assert_allclose(g, want, err_msg=msg, **tol)
                    np.testing.assert_allclose(g, single, err_msg=msg, **tol)
                checks += 1
        p, q = gen(dt, (R * 1000,)), gen(dt, (R * 1000,))
        got = sharded.shard_rows(eng, p).dot(sharded.shard_rows(eng, q)).to_numpy()
        if dt in ("i32", "i64", "u32"):
            np.testing.assert_array_equal(got, o.dot(p, q))
        else:
            np.testing.assert_allclose(got, o.dot(p, q), rtol=1e-3 if dt == "f32" else 1e-11, atol=1e-3 if dt == "f32" else 1e-10)
        A, B = gen(dt, (16 * world, 48)), gen(dt, (48 * world // world, 32))
        want = o.matmul(A, B)
        r1 = sharded.shard_rows(eng, A).matmul(feo.Tensor(feo.shape(*B.shape), B, ctx=ctx), algo="simt")
        Bs = gen(dt, (16 * world, 32))
        A2 = gen(dt, (16 * world, 16 * world))
        r2 = sharded.shard_rows(eng, A2).matmul(sharded.shard_rows(eng, Bs), b_sharded=True, algo="simt")
        for r, w in ((r1, want), (r2, o.matmul(A2, Bs))):
            if dt in ("i32", "i64", "u32"):
                np.testing.assert_array_equal(r.local.to_numpy(), rows(w))
            else:
                np.testing.assert_allclose(r.local.to_numpy(), rows(w), rtol=1e-4, atol=1e-5)
            checks += 1
    ctx.sync()
    dist.barrier()
    if rank == 0:
        print(f"sharded_check ok: world={world}, {checks} checks per rank, {ctx.launch_count()} kernel launches on rank 0", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
