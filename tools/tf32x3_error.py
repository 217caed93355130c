#!/usr/bin/env python
"""Error anatomy of the 3xTF32 tcgen05 matmul vs K and vs the accumulation-chunk length kc (diagnostic).

The tensor core truncates when it adds into the TMEM accumulator, so the error of a chain of kc k-blocks (12 kc accumulates)
is biased toward zero and grows with kc; the kernel drains every chunk into f32 registers with round-to-nearest.  This prints,
per K, the error on mixed-sign and on all-positive data (where the bias is one-sided) for several kc, next to the SIMT FP32
kernel (the reference's own arithmetic, in another order)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feotensor_b200 as feo
import feotensor_b200._lib as L

ctx = feo.Context.default()
rng = np.random.default_rng(1)
for k in (256, 4096, 16384):
    m, n = 256, 512
    A = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    B = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    for label, Ax, Bx in (("mixed", A, B), ("positive", np.abs(A), np.abs(B))):
        tA, tB = feo.Matrix(feo.shape(m, k), Ax), feo.Matrix(feo.shape(k, n), Bx)
        exact = Ax.astype(np.float64) @ Bx.astype(np.float64)
        terms = np.abs(Ax.astype(np.float64)) @ np.abs(Bx.astype(np.float64))
        for algo, kc in [("simt", 0)] + [("tf32x3", c) for c in (1, 2, 3, 4, 6, 8, 16)]:
            ctx.tune(L.KNOB_MATMUL_TF32_KC, kc)
            got = tA.matmul(tB, algo=algo).to_numpy().astype(np.float64)
            err = got - exact
            print(f"K={k:6d} {label:8s} {algo:7s} kc={kc:2d} max|err|/terms={np.max(np.abs(err) / terms):.3e}  mean(err/terms)={np.mean(err / terms):+.3e}  "
                  f"rms(err/terms)={np.sqrt(np.mean((err / terms) ** 2)):.3e}  frac toward zero={np.mean(np.sign(err) == -np.sign(exact)):.3f}")
ctx.tune(L.KNOB_MATMUL_TF32_KC, 0)
