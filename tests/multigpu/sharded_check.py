#!/usr/bin/env python
"""Multi-GPU parity of the sharded path (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/sharded_check.py

Every rank builds the same global tensors, uploads its block of leading rows, runs the sharded op through
feotensor_b200.sharded.CudaEngine (C ABI kernels + NCCL), and compares with the same op on the whole tensor
computed on its own GPU and with the CPU oracle: ints / max / min bit-exact, float sums within tolerance.

The checks run three times: with torch.distributed's NCCL as the exchange, with the library's own NCCL route
(csrc/nccl_route.cu, `NcclComm`: the exchange issued inside the C ABI), and with the native peer-memory exchange kernels
(csrc/peer.cu, `PeerComm`), whose replicated results must also be bit-identical across the ranks.
`--same-device` puts every rank on cuda:0 with gloo as the rendezvous (a one-GPU box can then exercise the IPC windows
and the exchange kernels; the ranks' kernels time-slice, so it is a correctness run only); NCCL is skipped there.
"""
import argparse
import zlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import feotensor_b200 as feo
    from feotensor_b200 import sharded
    from oracle import feo_oracle as o  # checker only

    ap = argparse.ArgumentParser()
    ap.add_argument("--same-device", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if args.same_device:
        local = 0
    torch.cuda.set_device(local)
    if args.same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = feo.Context(local)
    comm = sharded.PeerComm(ctx)
    total = 0
    ncomm = None if args.same_device else sharded.NcclComm(ctx)  # the library's own NCCL communicator (feo_nccl_*)
    for label, eng in (("nccl", None if args.same_device else sharded.CudaEngine(ctx)),
                       ("nccl-abi", None if args.same_device else sharded.CudaEngine(ctx, nccl=ncomm)),
                       ("peer", sharded.CudaEngine(ctx, comm=comm))):
        if eng is not None:
            total += run_checks(feo, sharded, o, dist, ctx, eng, rank, world, label, with_gather=not args.same_device)
    ctx.sync()
    dist.barrier()
    comm.close()
    if ncomm is not None:
        ncomm.close()
    if rank == 0:
        print(f"sharded_check ok: world={world}, {total} checks per rank, {ctx.launch_count()} kernel launches on rank 0", flush=True)
    dist.destroy_process_group()


def run_checks(feo, sharded, o, dist, ctx, eng, rank, world, label, with_gather=True):
    import feotensor_b200._lib as L
    rng = np.random.default_rng(99)
    checks = 0

    def same_on_every_rank(arr, msg):
        crcs = [None] * world
        dist.all_gather_object(crcs, zlib.crc32(np.ascontiguousarray(arr).tobytes()))
        assert len(set(crcs)) == 1, f"{label} {msg}: replicated result differs across ranks {crcs}"

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
                    same_on_every_rank(g, f"{dt} {kind} {axes}")
                msg = f"{label} {dt} {kind} {axes}"
                if dt in ("i32", "i64", "u32") or kind in ("max", "min"):
                    np.testing.assert_array_equal(g, want, err_msg=msg)
                    np.testing.assert_array_equal(g, single, err_msg=msg)
                else:
                    tol = dict(rtol=3e-4, atol=2e-5) if dt == "f32" else dict(rtol=1e-11, atol=1e-12)
                    np.testing.assert_allclose(g, want, err_msg=msg, **tol)
                    np.testing.assert_allclose(g, single, err_msg=msg, **tol)
                checks += 1
        if dt in ("f32", "f64"):
            # slices that are entirely -inf / +inf across every rank: the exchange's max / min must return the reference's
            # values (seeded with the infinities, not with +-FLT_MAX)
            xi = x.copy()
            xi[:, 3, :] = -np.inf
            xi[:, 5, 7] = np.inf
            sxi = sharded.shard_rows(eng, xi)
            for axes in ([0], [0, 2], [0, 1, 2], [1]):
                for kind in ("max", "min"):
                    got = sxi.reduce(kind, axes)
                    g = got.local.to_numpy() if isinstance(got, sharded.ShardedTensor) else got.to_numpy()
                    w = o.reduce(kind, xi, axes)
                    np.testing.assert_array_equal(g, rows(w) if isinstance(got, sharded.ShardedTensor) else w, err_msg=f"{label} {dt} {kind} {axes} inf")
                    checks += 1
        # Tensor::prod (tensor.rs:311-322) with the left operand split and the right one replicated: no exchange, one IEEE
        # multiply per element -> bit-exact for every dtype
        u, v = gen(dt, (R * 6,)), gen(dt, (77,))
        got = sharded.shard_rows(eng, u).outer(feo.Tensor(feo.shape(77), v, ctx=ctx))
        assert got.global_dims == [R * 6, 77]
        np.testing.assert_array_equal(got.local.to_numpy(), rows(o.prod(u, v)), err_msg=f"{label} {dt} outer")
        u2 = gen(dt, (R, 3, 5))
        got = sharded.shard_rows(eng, u2).outer(feo.Tensor(feo.shape(4, 2), v[:8].reshape(4, 2).copy(), ctx=ctx))
        assert got.global_dims == [R, 3, 5, 4, 2]
        np.testing.assert_array_equal(got.local.to_numpy(), rows(o.prod(u2, v[:8].reshape(4, 2))), err_msg=f"{label} {dt} outer 3d x 2d")
        checks += 2
        p, q = gen(dt, (R * 1000,)), gen(dt, (R * 1000,))
        got = sharded.shard_rows(eng, p).dot(sharded.shard_rows(eng, q)).to_numpy()
        same_on_every_rank(got, f"{dt} dot")
        if dt in ("i32", "i64", "u32"):
            np.testing.assert_array_equal(got, o.dot(p, q))
        else:
            np.testing.assert_allclose(got, o.dot(p, q), rtol=1e-3 if dt == "f32" else 1e-11, atol=1e-3 if dt == "f32" else 1e-10)
        A, B = gen(dt, (16 * world, 48)), gen(dt, (48 * world // world, 32))
        want = o.matmul(A, B)
        r1 = sharded.shard_rows(eng, A).matmul(feo.Tensor(feo.shape(*B.shape), B, ctx=ctx), algo="simt")
        Bs = gen(dt, (16 * world, 32))
        A2 = gen(dt, (16 * world, 16 * world))
        pairs = [(r1, want)]
        if with_gather or label == "peer":
            # B row-sharded: NCCL all-gather, or (peer engine) panels pulled over NVLink by the kernels themselves
            Bsh = sharded.shard_rows(eng, Bs, peer=(label == "peer"))
            pairs.append((sharded.shard_rows(eng, A2).matmul(Bsh, b_sharded=True, algo="simt"), o.matmul(A2, Bs)))
        for r, w in pairs:
            if dt in ("i32", "i64", "u32"):
                np.testing.assert_array_equal(r.local.to_numpy(), rows(w))
            else:
                np.testing.assert_allclose(r.local.to_numpy(), rows(w), rtol=1e-4, atol=1e-5)
            checks += 1
    if label == "peer":
        # f32 on the tensor cores, B's panels pulled from the peers while the matmul kernel runs: as raw rows when N % 4 == 0
        # (split inside the kernel), fused into the split / transpose pre-pass otherwise (ragged N and K included)
        # Every way the panels can travel (knob GATHER_OVERLAP): 0 = auto, 3 = the copy engines pull them piecewise, 4 = the matmul
        # kernel pulls them itself with TMA bulk copies, 2 / 1 = the gather kernel beside / in front of the matmul kernel -- same bits.
        for (M, K, N) in ((64 * world, 96 * world, 200), (128 * world, 32 * world, 257), (192 * world, 40 * world, 2084), (256 * world, 512 * world, 4352)):
            A, B = gen("f32", (M, K)), gen("f32", (K, N))
            Bsh = sharded.shard_rows(eng, B, peer=True)
            exact = rows(A.astype(np.float64) @ B.astype(np.float64))
            bound = rows(np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64))
            ref_err = 2 * np.abs(rows(o.matmul(A, B)) - exact) if M * K * N <= 1 << 28 else 0.0
            first = None
            try:
                for knob in (0, 3, 4, 2, 1):
                    ctx.tune(L.KNOB_GATHER_OVERLAP, knob)
                    got = sharded.shard_rows(eng, A).matmul(Bsh, b_sharded=True, algo="tf32x3").local.to_numpy()
                    err = np.abs(got - exact)
                    assert np.all(err <= np.maximum(ref_err, 2.0 ** -20 * bound)), f"tf32x3 sharded-B {M}x{K}x{N} knob {knob}"
                    if first is None:
                        first = got
                    np.testing.assert_array_equal(got, first, err_msg=f"tf32x3 sharded-B {M}x{K}x{N}: knob {knob} differs from knob 0")
                    checks += 1
            finally:
                ctx.tune(L.KNOB_GATHER_OVERLAP, 0)
            dist.barrier()
            Bsh.peer.close()
    return checks


if __name__ == "__main__":
    main()
