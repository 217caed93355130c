#!/usr/bin/env python
"""Where the config-2 slab loop loses against one big call: launch boundaries or the access pattern?

Three ways to write the same 32 GiB (this rank's 512 rows of a[.,1,4096] * b[1,4096,4096]):
  slabs      32 calls of feo_ewise_broadcast on out[16,4096,4096]                      (the bench's step)
  one_call   1 call on out[512,4096,4096]: b is read from DRAM once, no boundaries
  slab_major 1 call on out[32,16,4096,4096] with a padded a (so the planner cannot merge the two leading dims): the CTAs walk
             the output slab by slab exactly like the 32 calls do -- b is re-read for every slab -- but inside ONE launch.
slab_major ~ slabs  -> the loss is the access pattern (b re-read against the write stream);  slab_major ~ one_call -> it is the
launch boundaries.  CUDA events, 10 repetitions each, round robin.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L

    torch.cuda.set_device(0)
    ctx = feo.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    D, ROWS, SR = 4096, 512, 16
    a = torch.empty((ROWS // SR) * (SR + 1) * D, dtype=torch.float32, device="cuda")  # slabs of 16 rows, one padding row after each
    b = torch.empty(D * D, dtype=torch.float32, device="cuda")
    out = torch.empty(ROWS * D * D, dtype=torch.float32, device="cuda")
    ctx.check(lib.feo_fill_uniform(h, L.F32, vp(a.data_ptr()), a.numel(), 1, 0, -1.0, 1.0))
    ctx.check(lib.feo_fill_uniform(h, L.F32, vp(b.data_ptr()), b.numel(), 2, 0, -1.0, 1.0))
    slab = SR * D * D

    def slabs():
        for s in range(ROWS // SR):
            ctx.check(lib.feo_ewise_broadcast(h, L.MUL, L.F32, vp(out.data_ptr() + 4 * s * slab), vp(a.data_ptr() + 4 * s * (SR + 1) * D), vp(b.data_ptr()), 3,
                                              L.sizes([SR, D, D]), L.sizes([D, 0, 1]), L.sizes([0, D, 1])))

    def one_call():  # (reads a with the padded row pitch too: rows 16, 33, ... are skipped by using the 4-d shape with mergeable b only)
        ctx.check(lib.feo_ewise_broadcast(h, L.MUL, L.F32, vp(out.data_ptr()), vp(a.data_ptr()), vp(b.data_ptr()), 3,
                                          L.sizes([ROWS, D, D]), L.sizes([D, 0, 1]), L.sizes([0, D, 1])))

    def slab_major():
        ctx.check(lib.feo_ewise_broadcast(h, L.MUL, L.F32, vp(out.data_ptr()), vp(a.data_ptr()), vp(b.data_ptr()), 4,
                                          L.sizes([ROWS // SR, SR, D, D]), L.sizes([(SR + 1) * D, D, 0, 1]), L.sizes([0, 0, D, 1])))

    res = {k: [] for k in ("slabs", "one_call", "slab_major")}
    fns = {"slabs": slabs, "one_call": one_call, "slab_major": slab_major}
    for rnd in range(4):
        for k, fn in fns.items():
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            if rnd:
                res[k].append(e0.elapsed_time(e1) / 5)
    for k, v in res.items():
        ms = sorted(v)[len(v) // 2]
        print(json.dumps({"way": k, "ms_per_32GiB": round(ms, 3), "us_per_GiB_slab": round(ms * 1e3 / 32, 1), "GB/s_slab_accounting": round(32 * 1.141112832 / ms * 1e3, 0)}))


if __name__ == "__main__":
    main()
