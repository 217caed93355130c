#!/usr/bin/env python
"""The process model of SURVEY.md 8e -- ONE process, ONE host thread, G devices, ncclCommInitAll, grouped calls on
per-device streams -- through the C ABI only (no torch, no torch.distributed): feo_nccl_comm_create_all,
feo_nccl_group_start / feo_nccl_group_end, feo_nccl_allreduce / feo_nccl_allgather around the local kernels.

    python tests/multigpu/nccl_single_process.py [--gpus 2]

Checks sum / mean / max / var over the sharded leading axis, the sharded dot and the all-gathered matmul against the CPU
oracle (the reference's arithmetic): integers and max bit for bit, floats within tolerance, every device the same bits.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    args = ap.parse_args()
    import feotensor_b200 as feo
    import feotensor_b200._lib as L
    from oracle import feo_oracle as o  # checker only

    G = args.gpus
    ctxs = [feo.Context(d) for d in range(G)]
    lib = ctxs[0].lib
    vp = C.c_void_p
    handles = (vp * G)(*[c.handle for c in ctxs])
    comms = (vp * G)()
    ctxs[0].check(lib.feo_nccl_comm_create_all(handles, G, comms))
    comms = [vp(comms[d]) for d in range(G)]
    rng = np.random.default_rng(7)
    checks = 0

    def grouped(fn):
        assert lib.feo_nccl_group_start() == L.OK
        for d in range(G):
            fn(d)
        assert lib.feo_nccl_group_end() == L.OK

    for dt in ("f32", "f64", "i32", "i64", "u32"):
        T = o.NP_DTYPES[dt]
        code = L.code_of(dt)
        gen = (lambda shape: rng.uniform(-1, 1, shape).astype(T)) if dt in ("f32", "f64") else \
              (lambda shape: rng.integers(0 if dt == "u32" else -1000, 1000, shape).astype(T))
        x = gen((8 * G, 24, 40))
        rows = x.shape[0] // G
        shards = [feo.Tensor(feo.Shape([rows, 24, 40]), x[d * rows:(d + 1) * rows], ctx=ctxs[d]) for d in range(G)]
        ldims = [rows, 24, 40]
        for axes in ([0], [0, 2], [0, 1, 2]):
            nout = int(np.prod(o.reduce_shape(x.shape, axes)))
            for kind in ("sum", "mean", "max", "min", "var"):
                width = 2 * nout if kind == "var" else nout
                part = [feo.Tensor.zeros(feo.shape(width), dt, ctx=ctxs[d]) for d in range(G)]
                for d in range(G):  # phase 1: local one-pass partials
                    ctxs[d].check(lib.feo_reduce_partial(ctxs[d].handle, L.KIND_CODES[kind], code, vp(part[d].data.ptr), vp(shards[d].data.ptr), 3,
                                                         L.sizes(ldims), len(axes), L.sizes(axes)))
                div = np.array([o.reduce_divisor(dt, x.shape, axes)], T)
                if kind != "var":  # phase 2: one grouped ncclAllReduce; phase 3: mean's division by the global divisor
                    red = L.SUM if kind in ("sum", "mean") else L.KIND_CODES[kind]
                    grouped(lambda d: ctxs[d].check(lib.feo_nccl_allreduce(comms[d], red, code, vp(part[d].data.ptr), nout)))
                    if kind == "mean":
                        for d in range(G):
                            ctxs[d].check(lib.feo_ewise_scalar(ctxs[d].handle, L.DIV, code, vp(part[d].data.ptr), vp(part[d].data.ptr),
                                                               div.ctypes.data_as(vp), nout))
                    got = [p.to_numpy() for p in part]
                else:      # var: grouped ncclAllGather of the partials, then the rank-ordered merge
                    gath = [feo.Tensor.zeros(feo.shape(G, width), dt, ctx=ctxs[d]) for d in range(G)]
                    grouped(lambda d: ctxs[d].check(lib.feo_nccl_allgather(comms[d], vp(gath[d].data.ptr), vp(part[d].data.ptr), width * T().itemsize)))
                    outs = [feo.Tensor.zeros(feo.shape(nout), dt, ctx=ctxs[d]) for d in range(G)]
                    count = int(np.prod([ldims[a] for a in axes]))
                    for d in range(G):
                        ctxs[d].check(lib.feo_var_merge(ctxs[d].handle, code, vp(outs[d].data.ptr), vp(gath[d].data.ptr), G, nout, count,
                                                        div.ctypes.data_as(vp)))
                    got = [t.to_numpy() for t in outs]
                want = o.reduce(kind, x, axes).reshape(-1)
                msg = f"{dt} {kind} {axes}"
                for d in range(1, G):
                    np.testing.assert_array_equal(got[d], got[0], err_msg=msg + " (devices differ)")
                if dt in ("i32", "i64", "u32") or kind in ("max", "min"):
                    np.testing.assert_array_equal(got[0], want, err_msg=msg)
                else:
                    np.testing.assert_allclose(got[0], want, err_msg=msg, **(dict(rtol=3e-4, atol=2e-5) if dt == "f32" else dict(rtol=1e-11, atol=1e-12)))
                checks += 1
        # matmul, A and B row-sharded: grouped all-gather of B, then the local kernel (matrix.rs:76-90)
        A, B = gen((16 * G, 8 * G)), gen((8 * G, 32))
        Bl = [feo.Tensor(feo.shape(8, 32), B[d * 8:(d + 1) * 8], ctx=ctxs[d]) for d in range(G)]
        Bf = [feo.Tensor.zeros(feo.shape(8 * G, 32), dt, ctx=ctxs[d]) for d in range(G)]
        grouped(lambda d: ctxs[d].check(lib.feo_nccl_allgather(comms[d], vp(Bf[d].data.ptr), vp(Bl[d].data.ptr), 8 * 32 * T().itemsize)))
        want = o.matmul(A, B)
        for d in range(G):
            Ad = feo.Matrix(feo.shape(16, 8 * G), A[d * 16:(d + 1) * 16], ctx=ctxs[d])
            got = Ad.matmul(feo.Matrix.from_tensor(Bf[d]), algo="simt").to_numpy()
            if dt in ("i32", "i64", "u32"):
                np.testing.assert_array_equal(got, want[d * 16:(d + 1) * 16])
            else:
                np.testing.assert_allclose(got, want[d * 16:(d + 1) * 16], rtol=1e-4, atol=1e-5)
        checks += 1
    for d in range(G):
        ctxs[d].sync()
        lib.feo_nccl_comm_destroy(comms[d])
    print(f"nccl_single_process ok: {G} devices in one process, {checks} checks")


if __name__ == "__main__":
    main()
