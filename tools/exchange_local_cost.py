#!/usr/bin/env python
"""What a sharded reduction costs on ONE GPU before any NVLink traffic: the plain reduction, the partial written into the
exchange window, and the whole feo_reduce_sharded call with a world of one (local kernel + exchange kernel, signalling
itself).  Per-rank block of config 3 at 8 GPUs: [2048, 16384, 8] f32 (1 GiB).  C-ABI loop, CUDA events, one JSON line per case.

    python tools/exchange_local_cost.py [--iters 30] [--rows 2048]
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
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--rows", type=int, default=2048)
    ap.add_argument("--world", type=int, default=8, help="the world the divisor / counts are computed for (the exchange itself stays local)")
    args = ap.parse_args()
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L
    from feotensor_b200 import sharded

    torch.cuda.set_device(0)
    ctx = feo.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    win = vp()
    ctx.check(lib.feo_peer_alloc(h, 4 << 20, C.byref(win)))
    comm = sharded.PeerComm(ctx, rank=0, world=1, windows=[win.value])
    dims = [args.rows, 16384, 8]
    n = int(np.prod(dims))
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    ctx.check(lib.feo_fill_uniform(h, L.F32, vp(x.data_ptr()), n, 7, 0, -1.0, 1.0))
    out = torch.empty(16384 * 8 * 2, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gdims = [args.rows * args.world, 16384, 8]

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / args.iters

    shape = L.sizes(dims)
    for kind in ("sum", "var", "max"):
        for axes in ([0], [0, 2], [0, 1, 2]):
            ax = L.sizes(axes)
            code = L.KIND_CODES[kind]
            div = sharded.reduce_divisor("f32", gdims, axes)
            divp = np.array([div], np.float32)
            slot = comm.slot()[0]
            plain = timed(lambda: ctx.check(lib.feo_reduce(h, code, L.F32, vp(out.data_ptr()), vp(x.data_ptr()), 3, shape, len(axes), ax)))
            part = timed(lambda: ctx.check(lib.feo_reduce_partial(h, code, L.F32, vp(slot), vp(x.data_ptr()), 3, shape, len(axes), ax)))
            whole = timed(lambda: ctx.check(lib.feo_reduce_sharded(comm.handle, code, L.F32, vp(out.data_ptr()), vp(x.data_ptr()), 3, shape,
                                                                   len(axes), ax, divp.ctypes.data_as(vp))))
            gb = 4.0 * n / 1e3
            print(json.dumps({"case": f"{kind}_{axes}", "dims": dims, "us_plain": round(plain, 1), "us_partial_into_window": round(part, 1),
                              "us_sharded_world1": round(whole, 1), "GB/s_plain": round(gb / plain, 0), "GB/s_sharded_world1": round(gb / whole, 0)}), flush=True)
    comm.close(barrier=False)


if __name__ == "__main__":
    main()
