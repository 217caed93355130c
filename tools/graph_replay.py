#!/usr/bin/env python
"""Launch-bound chains through a CUDA graph: the library's compute calls are plain stream-ordered launches (no host
synchronisation, no allocation once the workspace is warm), so a host program can capture a chain of them on the
context's stream and replay it.  Config-1-sized chain: z = x + y; w = z * x; s = sum(w, [0, 2]); m = mean(w, [0, 2]);
v = var(w, [0, 2]); d = z / y   (6 calls, 8 MiB operands).

    python tools/graph_replay.py [--iters 200]
Prints one JSON line: us per chain eager vs replayed, and checks that the replayed results equal the eager ones bit for bit.
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
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import numpy as np
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = feo.Context(0)
    ctx.set_stream(stream.cuda_stream)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    shape, axes = [64, 128, 256], [0, 2]
    n = int(np.prod(shape))
    x = feo.Tensor.uniform(feo.Shape(shape), 1, dtype="f32", ctx=ctx)
    y = feo.Tensor.uniform(feo.Shape(shape), 2, lo=1.0, hi=2.0, dtype="f32", ctx=ctx)
    z, w, d = (feo.Tensor.zeros(feo.Shape(shape), "f32", ctx=ctx) for _ in range(3))
    s, m, v = (feo.Tensor.zeros(feo.shape(128), "f32", ctx=ctx) for _ in range(3))
    sh, ax = L.sizes(shape), L.sizes(axes)

    def chain():
        ctx.check(lib.feo_ewise(h, L.ADD, L.F32, vp(z.data.ptr), vp(x.data.ptr), vp(y.data.ptr), n))
        ctx.check(lib.feo_ewise(h, L.MUL, L.F32, vp(w.data.ptr), vp(z.data.ptr), vp(x.data.ptr), n))
        ctx.check(lib.feo_reduce(h, L.KIND_CODES["sum"], L.F32, vp(s.data.ptr), vp(w.data.ptr), 3, sh, 2, ax))
        ctx.check(lib.feo_reduce(h, L.KIND_CODES["mean"], L.F32, vp(m.data.ptr), vp(w.data.ptr), 3, sh, 2, ax))
        ctx.check(lib.feo_reduce(h, L.KIND_CODES["var"], L.F32, vp(v.data.ptr), vp(w.data.ptr), 3, sh, 2, ax))
        ctx.check(lib.feo_ewise(h, L.DIV, L.F32, vp(d.data.ptr), vp(z.data.ptr), vp(y.data.ptr), n))

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / args.iters

    chain()  # warms the reduction workspace: nothing is allocated inside the capture
    ctx.sync()
    eager = [t.to_numpy().copy() for t in (z, w, s, m, v, d)]
    us_eager = timed(chain)
    for t in (z, w, s, m, v, d):
        ctx.check(lib.feo_fill(h, L.F32, vp(t.data.ptr), np.zeros(1, np.float32).ctypes.data_as(vp), t.size()))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        chain()
    graph.replay()
    torch.cuda.synchronize()
    for t, want in zip((z, w, s, m, v, d), eager):
        np.testing.assert_array_equal(t.to_numpy(), want)
    us_graph = timed(graph.replay)
    print(json.dumps({"chain": "add, mul, sum, mean, var over [0,2], div on f32 [64,128,256]", "calls": 6, "us_eager": round(us_eager, 2),
                      "us_graph_replay": round(us_graph, 2), "bit_identical": True}))


if __name__ == "__main__":
    main()
