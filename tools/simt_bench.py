#!/usr/bin/env python
"""SIMT matmul variants (knob MATMUL_SIMT) on one GPU, round-robin, CUDA events; one JSON line per (dtype, shape, variant).

    python tools/simt_bench.py [--shape 8192x8192x8192] [--variants 0,6,7,8,9,10] [--dtypes f32,i32] [--iters 3] [--rounds 3]

Every variant must produce the same bits as variant 0 (same k order per thread): checked on a sample of rows.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="8192x8192x8192")
    ap.add_argument("--variants", default="0,6,7,8,9,10")
    ap.add_argument("--dtypes", default="f32")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    ctx = feo.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    vp = C.c_void_p
    m, k, n = (int(x) for x in args.shape.split("x"))
    variants = [int(v) for v in args.variants.split(",")]
    for dt in args.dtypes.split(","):
        A = feo.Tensor.uniform(feo.shape(m, k), 40, dtype=dt, ctx=ctx)
        B = feo.Tensor.uniform(feo.shape(k, n), 41, dtype=dt, ctx=ctx)
        Cm = feo.Tensor.zeros(feo.shape(m, n), dt, ctx=ctx)
        code = L.code_of(dt)
        fn = lambda: ctx.check(ctx.lib.feo_matmul(ctx.handle, L.MATMUL_SIMT, code, vp(Cm.data.ptr), vp(A.data.ptr), vp(B.data.ptr), m, k, n))
        times = {v: [] for v in variants}
        sample = {}
        for rnd in range(args.rounds):
            for v in variants:
                ctx.tune(L.KNOB_MATMUL_SIMT, v)
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                times[v].append(e0.elapsed_time(e1) / args.iters)
                if rnd == 0:
                    sample[v] = np.concatenate([Cm.to_numpy().reshape(m, n)[r] for r in (0, 77, m - 1)])
        for v in variants:
            ms = float(np.median(times[v]))
            print(json.dumps({"dtype": dt, "shape": args.shape, "variant": v, "ms_median": round(ms, 3), "ms_min": round(min(times[v]), 3),
                              "Tmac2/s": round(2.0 * m * k * n / ms / 1e9, 2), "same_bits_as_first": bool(np.array_equal(sample[v], sample[variants[0]]))}), flush=True)
    ctx.tune(L.KNOB_MATMUL_SIMT, 0)


if __name__ == "__main__":
    main()
