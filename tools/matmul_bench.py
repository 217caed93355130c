#!/usr/bin/env python
"""3xTF32 matmul variants on one GPU: the CTA-pair kernel (raw-B split in the kernel / pre-pass) at several tile widths and chunk lengths, the single-CTA kernel.

    python tools/matmul_bench.py [--shapes 8192x8192x8192,1024x8192x8192] [--iters 10]

Prints one JSON line per (shape, variant): ms (CUDA events on the launching stream, pre-pass included), effective TFLOP/s
(2MNK / time) and a sampled parity check against f64.  1024 x 8192 x 8192 is one rank's share of config 4 at 8 GPUs.
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
    ap.add_argument("--shapes", default="8192x8192x8192,1024x8192x8192,2048x8192x8192,4096x8192x8192")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--variants", default="pair,pair_prepass,pair256,pair224,single")
    args = ap.parse_args()
    import torch

    import feotensor_b200 as feo
    import feotensor_b200._lib as L
    from oracle import feo_oracle as oracle  # checker only
    torch.cuda.set_device(0)
    ctx = feo.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    vp = C.c_void_p
    # name -> (MATMUL_TF32_PAIR, MATMUL_TF32_BN, MATMUL_TF32_KC, MATMUL_TF32_BRAW)
    variants = {"pair": (0, 0, 0, 0), "pair256": (0, 256, 0, 0), "pair224": (0, 224, 0, 0), "pair192": (0, 192, 0, 0), "pair160": (0, 160, 0, 0),
                "pair128": (0, 128, 0, 0), "single": (1, 0, 0, 0), "pair_kc1": (0, 0, 1, 0), "pair_kc3": (0, 0, 3, 0), "pair_kc4": (0, 0, 4, 0),
                "pair_kc8": (0, 0, 8, 0), "pair_kc256": (0, 0, 256, 0),
                # B split + transposed by a pre-pass (round-1 data path) instead of split from raw tiles inside the kernel
                "pair_prepass": (0, 0, 0, 1), "pair_prepass256": (0, 256, 0, 1), "pair_prepass_kc4": (0, 0, 4, 1)}
    for shp in args.shapes.split(","):
        m, k, n = (int(x) for x in shp.split("x"))
        A = feo.Tensor.uniform(feo.shape(m, k), 40, dtype="f32", ctx=ctx)
        B = feo.Tensor.uniform(feo.shape(k, n), 41, dtype="f32", ctx=ctx)
        Cm = feo.Tensor.zeros(feo.shape(m, n), "f32", ctx=ctx)
        rng = np.random.default_rng(3)
        rr, cc = rng.integers(0, m, 24), rng.integers(0, n, 24)
        Ar = np.stack([oracle.fill_uniform("f32", k, 40, int(r) * k) for r in rr])
        Bc = np.stack([oracle.fill_uniform("f32", k, 41, int(c), stride=n) for c in cc], axis=1)
        exact = np.einsum("ik,ki->i", Ar.astype(np.float64), Bc.astype(np.float64))
        terms = np.einsum("ik,ki->i", np.abs(Ar).astype(np.float64), np.abs(Bc).astype(np.float64))
        names = args.variants.split(",")
        times = {nm: [] for nm in names}
        errs = {}
        for rnd in range(args.rounds):  # round-robin over the variants: clocks drift with temperature and the power cap
            for name in names:
                pair, bn, kc, braw = variants[name]
                ctx.tune(L.KNOB_MATMUL_TF32_BRAW, braw)
                ctx.tune(L.KNOB_MATMUL_TF32_PAIR, pair)
                ctx.tune(L.KNOB_MATMUL_TF32_BN, bn)
                ctx.tune(L.KNOB_MATMUL_TF32_KC, kc)
                fn = lambda: ctx.check(ctx.lib.feo_matmul(ctx.handle, L.MATMUL_TF32X3, L.F32, vp(Cm.data.ptr), vp(A.data.ptr), vp(B.data.ptr), m, k, n))
                if rnd == 0:
                    ctx.check(ctx.lib.feo_fill(ctx.handle, L.F32, vp(Cm.data.ptr), np.zeros(1, np.float32).ctypes.data_as(vp), m * n))
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                times[name].append(e0.elapsed_time(e1) / args.iters)
                if rnd == 0:
                    got = np.array([Cm.data[int(r) * n + int(c)] for r, c in zip(rr, cc)], dtype=np.float64)
                    errs[name] = float(np.max(np.abs(got - exact) / terms))
        for name in names:
            ms = float(np.median(times[name]))
            print(json.dumps({"shape": shp, "variant": name, "ms_median": round(ms, 4), "ms_min": round(min(times[name]), 4),
                              "TFLOP/s": round(2.0 * m * k * n / ms / 1e9, 1), "max_err_over_terms": errs[name], "ok": bool(errs[name] <= 2.0 ** -20)}), flush=True)
        del A, B, Cm
    ctx.tune(L.KNOB_MATMUL_TF32_PAIR, 0)
    ctx.tune(L.KNOB_MATMUL_TF32_BN, 0)
    ctx.tune(L.KNOB_MATMUL_TF32_KC, 0)
    ctx.tune(L.KNOB_MATMUL_TF32_BRAW, 0)


if __name__ == "__main__":
    main()
