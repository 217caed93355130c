#!/usr/bin/env python
"""Exchange step of the sharded reductions: NCCL all-reduce / all-gather vs the native peer-memory kernels (csrc/peer.cu).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_bench.py

Per case: the whole sharded op (local reduction + exchange + epilogue) timed with CUDA events on the launching stream,
`--iters` back-to-back calls after warm-up, max over ranks; rank 0 prints one JSON line per case.
"""
import argparse
import json
import os
import sys

import ctypes as C

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--big", action="store_true", help="add the config-3 sized cases (1 GiB per rank)")
    ap.add_argument("--pull-shapes", type=lambda v: [int(x) for x in v.split(",") if x], default=[], help="extra copy-engine shapes to time: 100 * streams + k-blocks per piece, comma separated")
    ap.add_argument("--only-matmul", action="store_true", help="skip the reductions and the dot: only the config-4 sharded-B matmul (implies --big)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import feotensor_b200 as feo
    from feotensor_b200 import _lib as L
    from feotensor_b200 import sharded

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = feo.Context(local)
    comm = sharded.PeerComm(ctx)
    engines = {"nccl": sharded.CudaEngine(ctx), "peer": sharded.CudaEngine(ctx, comm=comm)}

    def nvlink_kib():
        """(Tx, Rx) KiB summed over this GPU's NVLinks, from the driver's counters (`nvidia-smi nvlink -gt d`); None if unavailable."""
        try:
            import re
            import subprocess
            txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True, timeout=20).stdout
            tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
            rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
            return (tx, rx) if (tx or rx) else None
        except Exception:
            return None

    def counted(fn):
        """timed(fn) plus the NVLink bytes this rank moved per call (driver counters around warm-up + timed calls)."""
        c0 = nvlink_kib()
        us = timed(fn)
        torch.cuda.synchronize()
        c1 = nvlink_kib()
        calls = 5 + args.iters
        per = None if (c0 is None or c1 is None) else {"tx_MiB_per_call": round((c1[0] - c0[0]) / 1024 / calls, 2), "rx_MiB_per_call": round((c1[1] - c0[1]) / 1024 / calls, 2)}
        return us, per

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / args.iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cases = [("c1_sum_[0,2]", "f32", [64, 128, 256], "sum", [0, 2]),
             ("c1_var_[0,2]", "f32", [64, 128, 256], "var", [0, 2]),
             ("c1_var_i32_[0,2]", "i32", [64, 128, 256], "var", [0, 2]),
             ("mid_sum_[0]", "f32", [8 * world, 16384, 8], "sum", [0]),
             ("mid_max_all", "f32", [8 * world, 16384, 8], "max", [0, 1, 2])]
    if args.only_matmul:
        args.big, cases = True, []
    if args.big and not args.only_matmul:
        cases += [("c3_sum_[0]", "f32", [2048 * world, 16384, 8], "sum", [0]),
                  ("c3_var_[0]", "f32", [2048 * world, 16384, 8], "var", [0]),
                  ("c3_sum_[0,2]", "f32", [2048 * world, 16384, 8], "sum", [0, 2]),
                  ("c3_var_all", "f32", [2048 * world, 16384, 8], "var", [0, 1, 2])]
    for name, dt, gdims, kind, axes in cases:
        ldims = [gdims[0] // world] + gdims[1:]
        x = feo.Tensor.uniform(feo.Shape(ldims), 7, dtype=dt, first_index=rank * int(np.prod(ldims)), ctx=ctx) \
            if dt == "f32" else feo.Tensor.fill(feo.Shape(ldims), rank + 1, dtype=dt, ctx=ctx)
        res = {}
        for label, eng in engines.items():
            st = sharded.ShardedTensor(eng, x, gdims, rank, world)
            res[label] = timed(lambda: st.reduce(kind, axes))
        if kind in ("sum", "max"):
            # the same two sequences with the Python object layer stripped: preallocated output, direct C-ABI calls
            eng = engines["nccl"]
            nout = int(np.prod(sharded.reduce_dims(gdims, axes)))
            out = eng.empty([nout], x.dtype)
            tout = eng.as_torch(out)
            lib, code, kcode = ctx.lib, x.data.code, L.KIND_CODES[kind]
            shp, ax, one = L.sizes(ldims), L.sizes(axes), np.ones(1, dtype=x.dtype)
            op = dist.ReduceOp.SUM if kind == "sum" else dist.ReduceOp.MAX

            def lean_nccl():
                lib.feo_reduce_partial(ctx.handle, kcode, code, C.c_void_p(out.data.ptr), C.c_void_p(x.data.ptr), len(ldims), shp, len(axes), ax)
                dist.all_reduce(tout, op=op)

            def lean_peer():
                lib.feo_reduce_sharded(comm.handle, kcode, code, C.c_void_p(out.data.ptr), C.c_void_p(x.data.ptr), len(ldims), shp, len(axes), ax,
                                       one.ctypes.data_as(C.c_void_p))
            res["lean_nccl"] = timed(lean_nccl)
            res["lean_peer"], nv = counted(lean_peer)
            if nv:
                nv["expected_rx_MiB"] = round((world - 1) * nout * np.dtype(x.dtype).itemsize / 2 ** 20, 2)
                res_extra = {"nvlink_lean_peer": nv}
            else:
                res_extra = {}
        if rank == 0:
            print(json.dumps(dict({"case": name, "world": world, "local_dims": ldims, "speedup": round(res["nccl"] / res["peer"], 3)},
                                  **{"us_" + k: round(v, 2) for k, v in res.items()}, **(res_extra if kind in ("sum", "max") else {}))), flush=True)
        del x
    n_local = 1 << 20 if not args.only_matmul else 1 << 10
    p = feo.Tensor.uniform(feo.Shape([n_local]), 1, dtype="f32", first_index=rank * n_local, ctx=ctx)
    q = feo.Tensor.uniform(feo.Shape([n_local]), 2, dtype="f32", first_index=rank * n_local, ctx=ctx)
    res = {}
    for label, eng in engines.items():
        sp = sharded.ShardedTensor(eng, p, [n_local * world], rank, world)
        sq = sharded.ShardedTensor(eng, q, [n_local * world], rank, world)
        res[label] = timed(lambda: sp.dot(sq))
    if rank == 0:
        print(json.dumps({"case": "dot_1Mi_per_rank", "world": world, "us_nccl": round(res["nccl"], 2), "us_peer": round(res["peer"], 2),
                          "speedup": round(res["nccl"] / res["peer"], 3)}), flush=True)
    if args.big:
        # config 4 row-sharded with B row-sharded too: NCCL all-gather of B + matmul vs the gather fused into the pre-pass
        n = 8192
        rows_ = n // world
        A = feo.Tensor.uniform(feo.Shape([rows_, n]), 11, dtype="f32", first_index=rank * rows_ * n, ctx=ctx)
        buf = sharded.PeerBuffer(ctx, rows_ * n * 4)
        Bl = buf.tensor([rows_, n], "f32")
        ctx.check(ctx.lib.feo_fill_uniform(ctx.handle, Bl.data.code, C.c_void_p(Bl.data.ptr), rows_ * n, 12, rank * rows_ * n, -1.0, 1.0))
        res = {}
        for label, eng in reversed(list(engines.items())):  # peer first: the library's routes before any NCCL all-gather of B ran
            sa = sharded.ShardedTensor(eng, A, [n, n], rank, world)
            sb = sharded.ShardedTensor(eng, Bl, [n, n], rank, world)
            if label == "peer":
                sb.peer = buf
            if label == "peer":
                res[label], nv_mm = counted(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
            else:
                res[label] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
            if label == "peer":  # default = auto (copy engines here); the other routes for comparison
                ctx.tune(L.KNOB_GATHER_OVERLAP, 3)
                res["peer_copy_engines"] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
                for shape in args.pull_shapes:  # copy-engine route, streams x piece size (knob PULL_SHAPE = 100 * streams + k-blocks per piece)
                    ctx.tune(L.KNOB_PULL_SHAPE, shape)
                    res[f"peer_copy_engines_{shape // 100}streams_{shape % 100}kb"] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
                ctx.tune(L.KNOB_PULL_SHAPE, 0)
                ctx.tune(L.KNOB_GATHER_OVERLAP, 4)
                res["peer_in_kernel_pull"] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
                ctx.tune(L.KNOB_GATHER_OVERLAP, 2)
                res["peer_gather_kernel_side_stream"] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
                ctx.tune(L.KNOB_GATHER_OVERLAP, 1)
                res["peer_gather_kernel_then_matmul"] = timed(lambda: sa.matmul(sb, b_sharded=True, algo="tf32x3"))
                ctx.tune(L.KNOB_GATHER_OVERLAP, 0)
        Bfull = feo.Tensor.uniform(feo.Shape([n, n]), 12, dtype="f32", ctx=ctx)
        sa = sharded.ShardedTensor(engines["nccl"], A, [n, n], rank, world)
        res["replicated_b"] = timed(lambda: sa.matmul(Bfull, algo="tf32x3"))
        if rank == 0:
            flop = 2.0 * n * n * n
            print(json.dumps(dict({"case": "c4_matmul_8192_b_sharded", "world": world, "speedup": round(res["nccl"] / res["peer"], 3),
                                   "tflops_peer_whole_job": round(flop / res["peer"] / 1e6, 1),
                                   "tflops_nccl_whole_job": round(flop / res["nccl"] / 1e6, 1),
                                   "tflops_replicated_b_whole_job": round(flop / res["replicated_b"] / 1e6, 1),
                                   "nvlink_peer": dict(nv_mm or {}, expected_rx_MiB=round((world - 1) * rows_ * n * 4 / 2 ** 20, 1))},
                                  **{"us_" + k: round(v, 1) for k, v in res.items()})), flush=True)
        dist.barrier()
        buf.close()
    ctx.sync()
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
