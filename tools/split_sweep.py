#!/usr/bin/env python
"""Small-input reductions are latency-bound: time one reduction for several forced split counts (FEO_KNOB_REDUCE_SPLIT),
with the fused last-arriver combine and with the two-launch combine.

    python tools/split_sweep.py --shape 64,128,256 --axes 0,2 --kind sum,var [--splits 0,1,2,4,8,16]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="64,128,256")
    ap.add_argument("--axes", default="0,2")
    ap.add_argument("--kind", default="sum,var")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--splits", default="0,1,2,4,8,16")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--graph", type=int, default=20, help="calls per captured CUDA graph (takes the host's launch cost out); 0 = eager")
    args = ap.parse_args()
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = feo.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    dt, es = L.code_of(args.dtype), L.np_dtype(args.dtype).itemsize
    shape = [int(s) for s in args.shape.split(",")]
    axes = [int(s) for s in args.axes.split(",")]
    nin = 1
    for e in shape:
        nin *= e
    nout = 1
    for d, e in enumerate(shape):
        if d not in axes:
            nout *= e
    x = torch.empty(nin * es, dtype=torch.uint8, device="cuda")
    o = torch.empty(max(nout, 1) * es, dtype=torch.uint8, device="cuda")
    ctx.check(lib.feo_fill_uniform(h, dt, vp(x.data_ptr()), nin, 1, 0, -1.0, 1.0))
    warm = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    for _ in range(200):  # ramp the clocks up before the first timed configuration
        warm.zero_()
    torch.cuda.synchronize()
    for kind in args.kind.split(","):
        def fn():
            ctx.check(lib.feo_reduce(h, L.KIND_CODES[kind], dt, vp(o.data_ptr()), vp(x.data_ptr()), len(shape), L.sizes(shape), len(axes), L.sizes(axes)))
        for fuse in (0, 1):
            for s in [int(v) for v in args.splits.split(",")]:
                ctx.tune(L.KNOB_REDUCE_SPLIT, s)
                ctx.tune(L.KNOB_REDUCE_FUSE, fuse)
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
                run, per = fn, 1
                if args.graph:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                        for _ in range(args.graph):
                            fn()
                    run, per = g.replay, args.graph
                    run()
                    torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(max(args.iters // per, 5)):
                    run()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / (max(args.iters // per, 5) * per) * 1e3
                print(f"reduce {kind} {shape} axes {axes} split={s} {'two-launch' if fuse else 'fused'}: {us:7.2f} us", flush=True)


if __name__ == "__main__":
    main()
