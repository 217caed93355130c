#!/usr/bin/env python
"""Throughput of the hot-path kernels on shapes that are NOT the BASELINE configs (odd extents, tiny inner axes, many
alternating kept / reduced groups, integer and 8-byte types) -- to find the planner's slow corners.

    python tools/shape_sweep.py [--iters 5] [--only reduce|bcast|ewise]

One line per case: ms (CUDA events on the launching stream) and GB/s over the algorithmic bytes (every operand read
once, the result written once).  About 1 GiB of traffic per case.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--bpsm", type=int, default=0, help="FEO_KNOB_REDUCE_BLOCKS_PER_SM")
    ap.add_argument("--match", default="", help="run only the cases whose label contains this text")
    ap.add_argument("--tile", type=int, default=0, help="FEO_KNOB_BCAST_TILE")
    args = ap.parse_args()
    import numpy as np
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = feo.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    ctx.tune(L.KNOB_BCAST_TILE, args.tile)
    ctx.tune(L.KNOB_REDUCE_BLOCKS_PER_SM, args.bpsm)

    def dev(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def fill(t, dt, n, seed):
        lo, hi = (-1.0, 1.0)
        if dt in ("i32", "i64"):
            lo, hi = -1000, 1000
        if dt == "u32":
            lo, hi = 0, 2000
        ctx.check(lib.feo_fill_uniform(h, L.code_of(dt), vp(t.data_ptr()), n, seed, 0, lo, hi))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    results = []

    def report(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        results.append({"case": name, "ms": round(ms, 4), "GB/s": round(gbs, 1)})
        print(f"{name:70s} {ms:9.4f} ms {gbs:9.1f} GB/s", flush=True)

    reduce_cases = [
        ("f32", [1000, 1000, 256], [1]), ("f32", [1000, 1000, 256], [0, 2]), ("f32", [4096, 4096, 16], [1]),
        ("f32", [8191, 8191, 3], [2]), ("f32", [8191, 8191, 3], [0]), ("f32", [8191, 8191, 3], [1]), ("f32", [8191, 8191, 3], [0, 1]),
        ("f32", [65536, 4099], [1]), ("f32", [65536, 4099], [0]), ("f32", [4099, 65536], [0]), ("f32", [3, 89478485], [0]),
        ("f32", [89478485, 3], [0]), ("f32", [89478485, 3], [1]), ("f32", [16, 16, 16, 16, 16, 16, 16], [1, 3, 5]),
        ("f32", [16, 16, 16, 16, 16, 16, 16], [0, 2, 4, 6]), ("f32", [64, 64, 64, 1024], [1, 3]), ("f32", [64, 64, 64, 1024], [0, 2]),
        ("f32", [268435456], [0]), ("f32", [2, 134217728], [0]), ("f32", [134217728, 2], [1]),
        ("f64", [4096, 4096, 8], [0]), ("f64", [4096, 4096, 8], [2]), ("f64", [4096, 4096, 8], [0, 1, 2]),
        ("i32", [4096, 4096, 16], [0]), ("i32", [4096, 4096, 16], [2]), ("i64", [4096, 4096, 8], [1]),
    ]
    if args.only in ("", "reduce"):
        for dt, shape, axes in reduce_cases:
            if args.match and args.match not in f"reduce {dt} {shape} axes {axes}":
                continue
            es = L.np_dtype(dt).itemsize
            nin = int(np.prod(shape))
            nout = int(np.prod([e for d, e in enumerate(shape) if d not in axes])) if len(axes) < len(shape) else 1
            x, o = dev(nin * es), dev(max(nout, 1) * es)
            fill(x, dt, nin, 1)
            for kind in ("sum", "var", "max"):
                ms = timed(lambda: ctx.check(lib.feo_reduce(h, L.KIND_CODES[kind], L.code_of(dt), vp(o.data_ptr()), vp(x.data_ptr()),
                                                            len(shape), L.sizes(shape), len(axes), L.sizes(axes))))
                report(f"reduce {kind:4s} {dt} {shape} axes {axes}", ms, es * (nin + nout))
            del x, o
            torch.cuda.empty_cache()

    # broadcast: (dtype, out shape, a strides as 'which dims a spans', b dims)
    bcast_cases = [
        ("f32", [4096, 4096, 16], [1, 0, 1], [0, 1, 1]), ("f32", [16384, 16384], [1, 0], [0, 1]), ("f32", [16384, 16384], [1, 1], [0, 1]),
        ("f32", [16384, 16384], [1, 1], [1, 0]), ("f32", [8191, 8191, 3], [1, 1, 1], [0, 0, 1]), ("f32", [8191, 8191, 3], [1, 0, 1], [0, 1, 1]),
        ("f32", [8191, 8191, 3], [1, 1, 0], [1, 1, 1]), ("f32", [1000, 1000, 256], [1, 1, 1], [1, 1, 0]), ("f32", [1000, 1000, 256], [0, 1, 0], [1, 0, 1]),
        ("f32", [64, 64, 64, 1024], [1, 0, 1, 0], [0, 1, 0, 1]), ("f32", [65536, 4099], [1, 0], [0, 1]), ("f32", [4099, 65536], [1, 0], [0, 1]),
        ("f64", [4096, 4096, 8], [1, 0, 1], [0, 1, 1]), ("i32", [4096, 4096, 16], [1, 0, 1], [0, 1, 1]), ("i64", [8192, 8192], [1, 0], [0, 1]),
        ("f32", [16, 16, 16, 16, 16, 16, 16], [1, 0, 1, 0, 1, 0, 1], [0, 1, 0, 1, 0, 1, 0]),
        ("f32", [1048576, 256], [1, 1], [0, 1]), ("f32", [65536, 4096], [1, 1], [0, 1]), ("f32", [4096, 65536], [1, 1], [0, 1]),
        ("f32", [4096, 65536], [0, 1], [1, 1]), ("f64", [16384, 8192], [1, 1], [0, 1]),
    ]
    if args.only in ("", "bcast"):
        for dt, shape, am, bm in bcast_cases:
            if args.match and args.match not in f"bcast mul {dt} {shape} a{am} b{bm}":
                continue
            es = L.np_dtype(dt).itemsize

            def strides(mask):
                st, run = [0] * len(shape), 1
                for d in range(len(shape) - 1, -1, -1):
                    if mask[d]:
                        st[d] = run
                        run *= shape[d]
                return st, run
            sa, na = strides(am)
            sb, nb = strides(bm)
            nout = int(np.prod(shape))
            a, b, o = dev(na * es), dev(nb * es), dev(nout * es)
            fill(a, dt, na, 2)
            fill(b, dt, nb, 3)
            for op in ("mul",):
                ms = timed(lambda: ctx.check(lib.feo_ewise_broadcast(h, L.OP_CODES[op], L.code_of(dt), vp(o.data_ptr()), vp(a.data_ptr()),
                                                                     vp(b.data_ptr()), len(shape), L.sizes(shape), L.sizes(sa), L.sizes(sb))))
                report(f"bcast {op} {dt} {shape} a{am} b{bm}", ms, es * (na + nb + nout))
            del a, b, o
            torch.cuda.empty_cache()

    if args.only in ("", "ewise"):
        for dt, n in (("f32", 268435457), ("f64", 134217729), ("i32", 268435456), ("i64", 134217728), ("u32", 100000003)):
            es = L.np_dtype(dt).itemsize
            a, b, o = dev(n * es), dev(n * es), dev(n * es)
            fill(a, dt, n, 4)
            fill(b, dt, n, 5)
            for op in ("add", "div"):
                ms = timed(lambda: ctx.check(lib.feo_ewise(h, L.OP_CODES[op], L.code_of(dt), vp(o.data_ptr()), vp(a.data_ptr()), vp(b.data_ptr()), n)))
                report(f"ewise {op} {dt} n={n}", ms, 3 * es * n)
            two = np.array([2], dtype=L.np_dtype(dt))
            ms = timed(lambda: ctx.check(lib.feo_ewise_scalar(h, L.OP_CODES["mul"], L.code_of(dt), vp(o.data_ptr()), vp(a.data_ptr()),
                                                              two.ctypes.data_as(vp), n)))
            report(f"ewise_scalar mul {dt} n={n}", ms, 2 * es * n)
            del a, b, o
            torch.cuda.empty_cache()
    print(json.dumps(results))


if __name__ == "__main__":
    main()
