#!/usr/bin/env python
"""Runs one hot-path case a few times on cuda:0 (for ncu captures and quick A/B timing).

    python tools/run_case.py reduce --shape 16384,16384,8 --axes 2 --kind sum --iters 3
    python tools/run_case.py bcast --rows 16 --iters 3 [--tile 2]
    python tools/run_case.py outer --n 65536
    python tools/run_case.py dot --n 1000000000 [--batch 1000]
    python tools/run_case.py matmul --n 8192 --algo simt
    python tools/run_case.py ewise --n 268435456
Prints one line per case: ms (CUDA events on the launching stream), GB/s or TFLOP/s over the algorithmic bytes.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["reduce", "bcast", "outer", "dot", "matmul", "ewise", "ceilings"])
    ap.add_argument("--shape", default="16384,16384,8")
    ap.add_argument("--axes", default="0")
    ap.add_argument("--kind", default="sum")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--rows", type=int, default=16)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--algo", default="simt")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--bpsm", type=int, default=0)
    ap.add_argument("--lchunk", type=int, default=0)
    ap.add_argument("--kc", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=0, help="1 = disable programmatic dependent launch")
    ap.add_argument("--slabs", type=int, default=1)
    ap.add_argument("--nofuse", type=int, default=0, help="1 = two-launch combine for split reductions")
    ap.add_argument("--overlap", type=int, default=0, help="FEO_KNOB_GATHER_OVERLAP (2 = overlapped B pre-pass on one GPU)")
    ap.add_argument("--simt", type=int, default=0, help="FEO_KNOB_MATMUL_SIMT (1 = register double buffering, 2..5 = pipeline variants)")
    ap.add_argument("--l2keep", type=int, default=0, help="1 = disable the evict_last policy on broadcast operands")
    args = ap.parse_args()

    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = feo.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.tune(L.KNOB_BCAST_TILE, args.tile)
    ctx.tune(L.KNOB_REDUCE_SPLIT, args.split)
    ctx.tune(L.KNOB_REDUCE_BLOCKS_PER_SM, args.bpsm)
    ctx.tune(L.KNOB_BCAST_LCHUNK, args.lchunk)
    ctx.tune(L.KNOB_MATMUL_TF32_KC, args.kc)
    ctx.tune(L.KNOB_PDL, args.pdl)
    ctx.tune(L.KNOB_REDUCE_FUSE, args.nofuse)
    ctx.tune(L.KNOB_BCAST_L2KEEP, args.l2keep)
    ctx.tune(L.KNOB_GATHER_OVERLAP, args.overlap)
    ctx.tune(L.KNOB_MATMUL_SIMT, args.simt)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    dt = L.code_of(args.dtype)
    es = L.np_dtype(args.dtype).itemsize

    def dev(nelem):
        return torch.empty(nelem * es, dtype=torch.uint8, device="cuda")

    def fill(t, n, seed, lo=-1.0, hi=1.0):
        if args.dtype in ("i32", "i64", "u32"):
            lo, hi = (0, 2000) if args.dtype == "u32" else (-1000, 1000)
        ctx.check(lib.feo_fill_uniform(h, dt, vp(t.data_ptr()), n, seed, 0, lo, hi))

    def timed(fn, what, nbytes=None, flop=None):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        extra = f"{nbytes / ms / 1e6:9.1f} GB/s" if nbytes else f"{flop / ms / 1e9:8.2f} TFLOP/s"
        print(f"{what}: {ms:.4f} ms  {extra}", flush=True)

    if args.case == "reduce":
        shape = [int(s) for s in args.shape.split(",")]
        axes = [int(s) for s in args.axes.split(",")] if args.axes != "all" else []
        nin = 1
        for e in shape:
            nin *= e
        nout = 1
        for d, e in enumerate(shape):
            if d not in axes:
                nout *= e
        if not axes or len(axes) == len(shape):
            nout = 1
        x, o = dev(nin), dev(nout)
        fill(x, nin, 1)
        for kind in args.kind.split(","):
            timed(lambda: ctx.check(lib.feo_reduce(h, L.KIND_CODES[kind], dt, vp(o.data_ptr()), vp(x.data_ptr()), len(shape),
                                                   L.sizes(shape), len(axes), L.sizes(axes))),
                  f"reduce {kind} {shape} axes {axes} {args.dtype}", nbytes=es * (nin + nout))
    elif args.case == "bcast":
        D, R = 4096, args.rows
        S = args.slabs
        a, b, o = dev(S * R * D), dev(D * D), dev(S * R * D * D)
        fill(a, S * R * D, 2); fill(b, D * D, 3)

        def slabs():
            for s in range(S):
                ctx.check(lib.feo_ewise_broadcast(h, L.MUL, dt, vp(o.data_ptr() + es * s * R * D * D), vp(a.data_ptr() + es * s * R * D),
                                                  vp(b.data_ptr()), 3, L.sizes([R, D, D]), L.sizes([D, 0, 1]), L.sizes([0, D, 1])))
        timed(slabs, f"bcast mul {S} x [{R},1,{D}]*[1,{D},{D}] tile={args.tile} pdl_off={args.pdl} l2keep_off={args.l2keep}", nbytes=S * es * (R * D + D * D + R * D * D))
    elif args.case == "outer":
        n = args.n
        u, v, o = dev(n), dev(n), dev(n * n)
        fill(u, n, 4); fill(v, n, 5)
        timed(lambda: ctx.check(lib.feo_outer(h, dt, vp(o.data_ptr()), vp(u.data_ptr()), n, vp(v.data_ptr()), n)),
              f"outer {n}x{n}", nbytes=es * (2 * n + n * n))
    elif args.case == "dot":
        n = args.n
        p, q, r = dev(n), dev(n), dev(max(args.batch, 1))
        fill(p, n, 6); fill(q, n, 7)
        if args.batch:
            timed(lambda: ctx.check(lib.feo_dot_batched(h, dt, vp(r.data_ptr()), vp(p.data_ptr()), vp(q.data_ptr()), args.batch, n // args.batch)),
                  f"dot_batched {args.batch}x{n // args.batch}", nbytes=es * 2 * n)
        else:
            timed(lambda: ctx.check(lib.feo_dot(h, dt, vp(r.data_ptr()), vp(p.data_ptr()), vp(q.data_ptr()), n)), f"dot {n}", nbytes=es * 2 * n)
    elif args.case == "matmul":
        m = args.n
        A, B, Cm = dev(m * m), dev(m * m), dev(m * m)
        fill(A, m * m, 8); fill(B, m * m, 9)
        timed(lambda: ctx.check(lib.feo_matmul(h, L.ALGO_CODES[args.algo], dt, vp(Cm.data_ptr()), vp(A.data_ptr()), vp(B.data_ptr()), m, m, m)),
              f"matmul {m}^3 {args.dtype} {args.algo} kc={args.kc} simt={args.simt}", flop=2.0 * m ** 3)
    elif args.case == "ceilings":
        # library kernels as yardsticks (not product): pure write, pure read, copy
        n = 1 << 28
        x = torch.empty(n, dtype=torch.float32, device="cuda")
        y = torch.empty(n, dtype=torch.float32, device="cuda")
        timed(lambda: x.zero_(), "torch zero_ 1 GiB (pure write)", nbytes=4 * n)
        timed(lambda: y.copy_(x), "torch copy_ 1 GiB (read+write)", nbytes=8 * n)
        timed(lambda: x.sum(), "torch sum 1 GiB (pure read)", nbytes=4 * n)
        del x, y
        m = 8192
        A = torch.randn(m, m, device="cuda"); B = torch.randn(m, m, device="cuda")
        torch.backends.cuda.matmul.allow_tf32 = True
        timed(lambda: torch.matmul(A, B), "cuBLAS tf32 8192^3 (1 pass: the TF32 pipe yardstick)", flop=2.0 * m ** 3)
        torch.backends.cuda.matmul.allow_tf32 = False
        timed(lambda: torch.matmul(A, B), "cuBLAS fp32 8192^3", flop=2.0 * m ** 3)
        Ab, Bb = A.bfloat16(), B.bfloat16()
        timed(lambda: torch.matmul(Ab, Bb), "cuBLAS bf16 8192^3", flop=2.0 * m ** 3)
    elif args.case == "ewise":
        n = args.n
        x, y, o = dev(n), dev(n), dev(n)
        fill(x, n, 10); fill(y, n, 11)
        timed(lambda: ctx.check(lib.feo_ewise(h, L.ADD, dt, vp(o.data_ptr()), vp(x.data_ptr()), vp(y.data_ptr()), n)), f"ewise add {n}", nbytes=3 * es * n)


if __name__ == "__main__":
    main()
