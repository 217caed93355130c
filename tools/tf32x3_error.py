#!/usr/bin/env python
"""Error anatomy of the 3xTF32 tcgen05 matmul vs K (diagnostic; prints one line per K)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feotensor_b200 as feo

rng = np.random.default_rng(1)
for k in (32, 256, 1024, 4096, 16384):
    m, n = 128, 256
    A = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    B = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    tA, tB = feo.Matrix(feo.shape(m, k), A), feo.Matrix(feo.shape(k, n), B)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    terms = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
    for algo in ("tf32x3", "simt"):
        got = tA.matmul(tB, algo=algo).to_numpy().astype(np.float64)
        err = got - exact
        toward_zero = np.mean(np.sign(err) == -np.sign(exact))
        print(f"K={k:6d} {algo:7s} max|err|/terms={np.max(np.abs(err) / terms):.3e}  rms err={np.sqrt(np.mean(err ** 2)):.3e}  "
              f"mean(err*sign(exact))={np.mean(err * np.sign(exact)):+.3e}  frac toward zero={toward_zero:.3f}  rms|exact|={np.sqrt(np.mean(exact ** 2)):.2f}")
    # all-positive data: every partial sum is positive, so truncation bias is one-sided
    Ap, Bp = np.abs(A), np.abs(B)
    got = feo.Matrix(feo.shape(m, k), Ap).matmul(feo.Matrix(feo.shape(k, n), Bp), algo="tf32x3").to_numpy().astype(np.float64)
    ex = Ap.astype(np.float64) @ Bp.astype(np.float64)
    print(f"K={k:6d} positive data: mean rel err={np.mean((got - ex) / ex):+.3e}  max rel={np.max(np.abs(got - ex) / ex):.3e}")
