"""GPU parity tests: the CUDA path, called through the C ABI (via the reference-API mirror), against
the CPU oracle on the same inputs and against the reference's golden vectors.

Bar (BASELINE.json north_star): bit-exact for integer dtypes, elementwise ops, outer product,
max/min and all shape bookkeeping; float sum/mean/var/dot/matmul within the reassociation bound
    |gpu - ref| <= 2 * gamma_n * sum|terms|,   gamma_n = n*u / (1 - n*u),  u = 2^-24 (f32) / 2^-53 (f64)
(both summation orders are within gamma_n * sum|terms| of the exact sum).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOATS = ["f32", "f64"]
INTS = ["i32", "i64", "u32"]
ALL = FLOATS + INTS
U = {"f32": 2.0 ** -24, "f64": 2.0 ** -53}


@pytest.fixture(scope="module")
def feo():
    import feotensor_b200 as f
    f.Context.default()  # raises without a GPU: these tests never fall back
    return f


def npdt(oracle, dt):
    return oracle.NP_DTYPES[dt]


def rand(oracle, rng, dt, shape, lo=-1.0, hi=1.0):
    if dt in FLOATS:
        return rng.uniform(lo, hi, shape).astype(npdt(oracle, dt))
    lo_i, hi_i = (0, 2000) if dt == "u32" else (-1000, 1000)
    return rng.integers(lo_i, hi_i, shape).astype(npdt(oracle, dt))


def gamma(n, dt):
    nu = n * U[dt]
    assert nu < 1
    return nu / (1 - nu)


def assert_sum_like(got, ref, abs_terms_sum, n, dt, what=""):
    """Reassociation bound for sums of n terms whose absolute values add up to abs_terms_sum."""
    bound = 2 * gamma(n, dt) * np.asarray(abs_terms_sum, dtype=np.float64) + 1e-300
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert np.all(err <= bound), f"{what}: max err {err.max():.3e} bound {bound.flat[np.argmax(err - bound)]:.3e}"


# ---- the reference's own known-answer tests, on the GPU ----------------------------------------------
def _mk(feo, via, shape, data, dt):
    if via == "vector":
        return feo.Vector(data, dtype=dt)
    if via == "matrix":
        return feo.Matrix(feo.Shape(shape), data, dtype=dt)
    return feo.Tensor(feo.Shape(shape), data, dtype=dt)


def _ints_ok(*seqs):
    return all(float(x).is_integer() for s in seqs for x in (s if isinstance(s, list) else [s]))


def test_reference_kats_reductions(feo, kat, oracle):
    n = 0
    for c in [c for c in kat if c["op"] == "reduce"]:
        dts = list(FLOATS)
        if _ints_ok(c["data"], c["expect"]):
            dts += ["i32", "i64"] + ([] if any(x < 0 for x in c["data"] + c["expect"]) else ["u32"])
        for dt in dts:
            x = _mk(feo, c["via"], c["shape"], c["data"], dt)
            res = getattr(x, c["kind"])() if c["via"] == "vector" else getattr(x, c["kind"])(c["axes"])
            assert res.shape().dims == c["expect_shape"], c["ref"]
            np.testing.assert_array_equal(res.to_numpy().reshape(-1), np.array(c["expect"], npdt(oracle, dt)), err_msg=c["ref"] + dt)
            n += 1
    assert n > 100


def test_reference_kats_elementwise(feo, kat, oracle):
    for c in [c for c in kat if c["op"] in ("ewise", "ewise_scalar")]:
        nums = c["a"] + c["expect"] + (c["b"] if c["op"] == "ewise" else [c["scalar"]])
        dts = list(FLOATS)
        if _ints_ok(nums):
            dts += ["i32", "i64"] + ([] if any(x < 0 for x in nums) else ["u32"])
        for dt in dts:
            a = _mk(feo, c["lhs"], c["shape"], c["a"], dt)
            b = _mk(feo, c["rhs"], c["shape"], c["b"], dt) if c["op"] == "ewise" else c["scalar"]
            res = {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b}[c["opname"]]
            assert res.shape().dims == c["shape"], c["ref"]
            np.testing.assert_array_equal(res.to_numpy().reshape(-1), np.array(c["expect"], npdt(oracle, dt)), err_msg=c["ref"] + dt)
            # result newtype follows the reference's impl table (tensor.rs:375-505, matrix.rs:113-214, vector.rs:101-202)
            want = c["lhs"] if c["lhs"] != "tensor" else (c["rhs"] if c["op"] == "ewise" else "tensor")
            assert type(res).__name__.lower() == want, c["ref"]


def test_reference_kats_products(feo, kat, oracle):
    for dt in ALL:
        T = npdt(oracle, dt)
        for c in [c for c in kat if c["op"] == "prod"]:
            a = feo.Tensor(feo.Shape(c["a_shape"]), c["a"], dtype=dt)
            b = feo.Tensor(feo.Shape(c["b_shape"]), c["b"], dtype=dt)
            r = a.prod(b)
            assert r.shape().dims == c["expect_shape"]
            np.testing.assert_array_equal(r.to_numpy().reshape(-1), np.array(c["expect"], T), err_msg=c["ref"])
        for c in [c for c in kat if c["op"] == "matmul"]:
            A = feo.Matrix(feo.Shape(c["a_shape"]), c["a"], dtype=dt)
            B = feo.Matrix(feo.Shape(c["b_shape"]), c["b"], dtype=dt)
            r = A.matmul(B)
            assert r.shape().dims == c["expect_shape"]
            np.testing.assert_array_equal(r.to_numpy().reshape(-1), np.array(c["expect"], T), err_msg=c["ref"])
        for c in [c for c in kat if c["op"] == "matvec"]:
            r = feo.Matrix(feo.Shape(c["a_shape"]), c["a"], dtype=dt).vecmul(feo.Vector(c["v"], dtype=dt))
            assert r.shape().dims == c["expect_shape"]
            np.testing.assert_array_equal(r.to_numpy(), np.array(c["expect"], T), err_msg=c["ref"])
        for c in [c for c in kat if c["op"] == "vecmat"]:
            r = feo.Vector(c["v"], dtype=dt).matmul(feo.Matrix(feo.Shape(c["a_shape"]), c["a"], dtype=dt))
            np.testing.assert_array_equal(r.to_numpy(), np.array(c["expect"], T), err_msg=c["ref"])
        for c in [c for c in kat if c["op"] == "dot"]:
            r = feo.Vector(c["a"], dtype=dt).vecmul(feo.Vector(c["b"], dtype=dt))
            assert r.shape().dims == c["expect_shape"]
            np.testing.assert_array_equal(r.to_numpy(), np.array(c["expect"], T), err_msg=c["ref"])


def test_reference_kats_constructors_and_access(feo, kat, oracle):
    for c in kat:
        op = c["op"]
        if op == "new":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            assert t.shape() == feo.Shape(c["shape"])
            np.testing.assert_array_equal(t.data.to_numpy(), np.array(c["data"]))
        elif op == "new_err":
            with pytest.raises(feo.ShapeError):
                feo.Tensor(feo.Shape(c["shape"]), c["data"])
        elif op == "fill":
            cls = {"tensor": feo.Tensor, "matrix": feo.Matrix, "vector": feo.Vector}[c.get("via", "tensor")]
            t = getattr(cls, c["ctor"])(feo.Shape(c["shape"]), c["value"], dtype=c["dtype"]) if c["ctor"] == "fill" else \
                getattr(cls, c["ctor"])(feo.Shape(c["shape"]), dtype=c["dtype"])
            assert t.shape().dims == c["shape"] and t.size() == int(np.prod(c["shape"]))
            np.testing.assert_array_equal(t.to_numpy().reshape(-1), np.full(t.size(), c["value"], npdt(oracle, c["dtype"])))
        elif op == "eye":
            np.testing.assert_array_equal(feo.Matrix.eye(feo.Shape(c["shape"])).to_numpy().reshape(-1), np.array(c["expect"]))
        elif op == "get":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            for co, want in zip(c["coords"], c["expect"]):
                assert t.get(feo.Coordinate(co)) == want
        elif op == "set":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            for co, v in zip(c["coords"], c["values"]):
                t.set(feo.Coordinate(co), v)
            for co, v in zip(c["coords"], c["values"]):
                assert t.get(feo.Coordinate(co)) == v
        elif op == "get_err":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            for co in c["coords"]:
                with pytest.raises(feo.ShapeError):
                    t.get(feo.Coordinate(co))
        elif op == "set_err":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            for co, v in zip(c["coords"], c["values"]):
                with pytest.raises(feo.ShapeError):
                    t.set(feo.Coordinate(co), v)
        elif op == "from_tensor_err":
            t = feo.Tensor(feo.Shape(c["shape"]), c["data"])
            with pytest.raises(feo.ShapeError):
                (feo.Matrix if c["kind"] == "matrix" else feo.Vector).from_tensor(t)
    # matrix.rs:324-371 / vector.rs:304-331: Index / IndexMut / set through Deref
    m = feo.Matrix(feo.shape(2, 2), [1.0, 2.0, 3.0, 4.0])
    m[feo.coord(1, 0)] = 5.0
    assert [m[feo.coord(i, j)] for i in range(2) for j in range(2)] == [1.0, 2.0, 5.0, 4.0]
    v = feo.Vector([1.0, 2.0, 3.0, 4.0])
    v[2] = 5.0
    assert v[2] == 5.0 and v.size() == 4


def test_panic_preconditions(feo):
    a = feo.Tensor(feo.shape(2, 3), np.zeros(6, np.float32))
    b = feo.Tensor(feo.shape(3, 2), np.zeros(6, np.float32))
    with pytest.raises(feo.ReferencePanic):  # tensor.rs:366
        a + b
    for bad in ([1, 0], [0, 0], [2], [0, 1, 1]):
        with pytest.raises(feo.ReferencePanic):
            a.sum(bad)
    with pytest.raises(feo.ReferencePanic):  # matrix.rs:78
        feo.Matrix.from_tensor(a).matmul(feo.Matrix.from_tensor(a))
    with pytest.raises(feo.ReferencePanic):  # vector.rs:72
        feo.Vector([1.0, 2.0]).vecmul(feo.Vector([1.0, 2.0, 3.0]))
    with pytest.raises(feo.ReferencePanic):  # matrix.rs:93
        feo.Matrix.from_tensor(a).vecmul(feo.Vector([1.0, 2.0]))


# ---- elementwise: bit-exact on random data, every dtype, ragged sizes ---------------------------------------
@pytest.mark.parametrize("dt", ALL)
def test_elementwise_bit_exact(feo, oracle, dt):
    rng = np.random.default_rng(100)
    for n in (1, 3, 255, 1024, 4097, (1 << 20) + 3):
        a = rand(oracle, rng, dt, n)
        b = rand(oracle, rng, dt, n, 1.0, 2.0) if dt in FLOATS else (np.abs(rand(oracle, rng, dt, n)) + 1).astype(a.dtype)
        ta, tb = feo.Tensor(feo.shape(n), a), feo.Tensor(feo.shape(n), b)
        for op in ("add", "sub", "mul", "div"):
            got = {"add": ta + tb, "sub": ta - tb, "mul": ta * tb, "div": ta / tb}[op].to_numpy()
            np.testing.assert_array_equal(got, oracle.ewise(op, a, b), err_msg=f"{dt} {op} n={n}")
            s = b[0]
            got = {"add": ta + s, "sub": ta - s, "mul": ta * s, "div": ta / s}[op].to_numpy()
            np.testing.assert_array_equal(got, oracle.ewise_scalar(op, a, s), err_msg=f"{dt} scalar {op} n={n}")


def test_elementwise_special_values_and_wrap(feo, oracle):
    a = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, 3.4e38, 1.17549435e-38], np.float32)
    b = np.array([0.0, 0.0, 0.0, np.inf, 1.0, 2.0, 2.0, 3.4e38, 3.0], np.float32)
    ta, tb = feo.Tensor(feo.shape(a.size), a), feo.Tensor(feo.shape(b.size), b)
    for op, r in (("add", ta + tb), ("sub", ta - tb), ("mul", ta * tb), ("div", ta / tb)):
        got, want = r.to_numpy(), oracle.ewise(op, a, b)
        np.testing.assert_array_equal(got.view(np.uint32)[~np.isnan(want)], want.view(np.uint32)[~np.isnan(want)], err_msg=op)
        assert np.array_equal(np.isnan(got), np.isnan(want))
    ia = np.array([2 ** 31 - 1, -2 ** 31, 7, -7], np.int32)
    ib = np.array([1, -1, 2, 2], np.int32)
    ti, tj = feo.Tensor(feo.shape(4), ia), feo.Tensor(feo.shape(4), ib)
    np.testing.assert_array_equal((ti + tj).to_numpy(), oracle.ewise("add", ia, ib))
    np.testing.assert_array_equal((ti * tj).to_numpy(), oracle.ewise("mul", ia, ib))
    np.testing.assert_array_equal((ti / np.int32(2)).to_numpy(), oracle.ewise_scalar("div", ia, 2))  # truncation toward zero
    # x/0 and MIN/-1: the reference panics at the operator; the kernel writes 0, raises the sticky flag, and the next
    # synchronisation point (download / sync) reports it as FEO_ERR_ARITH -> ReferencePanic, clearing the flag
    ctx = feo.Context.default()
    assert ctx.arith_flag() == 0
    q = ti / feo.Tensor(feo.shape(4), np.array([1, -1, 0, 1], np.int32))
    with pytest.raises(feo.ReferencePanic):
        q.to_numpy()
    assert ctx.arith_flag() == 0  # cleared by the report
    np.testing.assert_array_equal(q.to_numpy(), np.array([2 ** 31 - 1, 0, 0, -7], np.int32))
    q = ti / feo.Tensor(feo.shape(4), np.array([1, 1, 0, 1], np.int32))
    assert ctx.arith_flag(clear=False) == 1  # the explicit poll sees it too
    with pytest.raises(feo.ReferencePanic):
        ctx.sync()
    ctx.sync()
    # a zero integer scalar divisor is rejected on the host, nothing is launched (tensor.rs:465-475 panics at element 0)
    for dt_ in (np.int32, np.int64, np.uint32):
        with pytest.raises(feo.ReferencePanic):
            feo.Tensor(feo.shape(3), np.array([1, 2, 3], dt_)) / dt_(0)
    assert np.isinf((feo.Tensor(feo.shape(2), np.array([1.0, -1.0], np.float32)) / np.float32(0)).to_numpy()).all()  # floats: IEEE
    feo.Tensor(feo.shape(4), np.arange(4, dtype=np.int32)).mean([0]).to_numpy()  # an ordinary integer mean raises nothing
    # i64: operands inside the i32 range take a 32-bit division, everything else the 64-bit one -- same quotients either way
    la = np.array([7, -7, 2 ** 31 - 1, -2 ** 31, -2 ** 31, 2 ** 31, -2 ** 31 - 1, 2 ** 62, -2 ** 63, -2 ** 63, 10 ** 15, -10 ** 15, 5, 0],
                  np.int64)
    lb = np.array([2, 2, -1, -1, 1, -1, 3, -3, 1, 2, 7, 10 ** 6, 2 ** 40, -9], np.int64)
    tl, tm = feo.Tensor(feo.shape(la.size), la), feo.Tensor(feo.shape(lb.size), lb)
    np.testing.assert_array_equal((tl / tm).to_numpy(), oracle.ewise("div", la, lb))
    assert ctx.arith_flag() == 0
    q = tl / feo.Tensor(feo.shape(la.size), np.array([1, 0, 1, 1, 1, 1, 1, 1, -1, 1, 1, 1, 1, 1], np.int64))
    assert ctx.arith_flag() == 1 and ctx.arith_flag() == 0
    r = q.to_numpy()
    assert r[1] == 0 and r[8] == 0 and r[0] == 7 and r[7] == 2 ** 62


def test_elementwise_unaligned_pointers(feo, oracle):
    """Raw C-ABI call on pointers offset by one element: exercises the scalar path."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(7)
    n = 10001
    a, b = rand(oracle, rng, "f32", n + 1), rand(oracle, rng, "f32", n + 1)
    ta, tb, to = feo.Tensor(feo.shape(n + 1), a), feo.Tensor(feo.shape(n + 1), b), feo.Tensor.zeros(feo.shape(n + 1), "f32")
    ctx.check(ctx.lib.feo_ewise(ctx.handle, L.MUL, L.F32, C.c_void_p(to.data.ptr + 4), C.c_void_p(ta.data.ptr + 4),
                                C.c_void_p(tb.data.ptr + 4), n))
    np.testing.assert_array_equal(to.to_numpy()[1:], oracle.ewise("mul", a[1:], b[1:]))
    assert to.to_numpy()[0] == 0


BCAST_CASES = [
    ((8, 1, 64), (1, 12, 64)),          # config-2 pattern: both operands invariant somewhere (8x4 tiles)
    ((3, 1, 64), (1, 5, 64)),           # fewer than 8 rows: 4x4 tiles with ragged edges
    ((16, 1, 4096), (1, 9, 4096)),      # full-width inner rows
    ((37, 1), (1, 128)),                # outer-product pattern (a splat along the inner axis)
    ((1, 130), (33, 1)),                # mirrored outer product
    ((33, 64), (1, 64)),                # matrix + row bias: only b invariant
    ((1, 64), (33, 64)),                # only a invariant
    ((5, 7, 12), (5, 7, 12)),           # same shape -> collapses to the flat kernel
    ((5, 1, 7, 1, 8), (1, 3, 7, 2, 8)), # rank 5, several broadcast dims
    ((4, 1, 6), (1, 5, 6)),             # inner not a multiple of 4 (f32): scalar fallback
    ((7, 1, 1), (1, 1, 9)),             # extent-1 dims dropped
    ((2, 3, 1, 16), (2, 1, 5, 16)),
    ((1,), (1,)),
    ((6, 1, 32), (6, 4, 1)),            # b splat along the inner axis, a invariant along dim 1
    ((300, 64), (300, 1)),              # matrix (op) column vector: no invariant dim, 4 rows in flight per thread
    ((70, 33, 256), (70, 33, 1)),       # the same with a short inner extent (spare thread rows walk the outer dim)
    ((500, 3), (1, 3)),                 # inner extent 3: flat-index kernel, vectors straddle rows
    ((500, 3), (500, 1)), ((1, 3), (500, 3)),
    ((41, 1, 7), (1, 13, 7)),           # flat-index kernel, rank 3
    ((9, 1, 5, 3), (1, 4, 5, 1)),       # rank 4, inner splat
    ((257, 1), (1, 4099)),              # outer product with an odd row length
    ((3, 1, 5, 1, 7, 3), (1, 2, 1, 4, 7, 3)),  # rank 5 after collapsing: generic fallback
    ((8, 1, 8, 1, 8, 1, 8), (1, 8, 1, 8, 1, 8, 1)),  # alternating short axes: the vector planner defers to the flat-index kernel
    ((4, 1, 6, 1, 5, 16), (1, 3, 1, 2, 1, 16)),      # ... with a shared inner axis
]


@pytest.mark.parametrize("dt", ALL)
def test_broadcast_bit_exact(feo, oracle, dt):
    rng = np.random.default_rng(200)
    for sa, sb in BCAST_CASES:
        a = rand(oracle, rng, dt, sa)
        b = rand(oracle, rng, dt, sb, 1.0, 2.0) if dt in FLOATS else (np.abs(rand(oracle, rng, dt, sb)) + 1).astype(a.dtype)
        ta, tb = feo.Tensor(feo.Shape(sa), a), feo.Tensor(feo.Shape(sb), b)
        for op in ("mul", "add", "sub", "div"):
            got = ta.broadcast(op, tb)
            want = oracle.ewise_broadcast(op, a, b)
            assert tuple(got.shape().dims) == want.shape
            np.testing.assert_array_equal(got.to_numpy(), want, err_msg=f"{dt} {op} {sa} {sb}")
    with pytest.raises(feo.ShapeError):
        feo.Tensor(feo.shape(2, 3), np.zeros(6, np.float32)).broadcast("mul", feo.Tensor(feo.shape(4, 3), np.zeros(12, np.float32)))


def test_broadcast_tile_variants_agree(feo, oracle):
    """Every tiling of the broadcast kernel (knob 0) produces the same bits."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(201)
    a, b = rand(oracle, rng, "f32", (19, 1, 256)), rand(oracle, rng, "f32", (1, 23, 256))
    ta, tb = feo.Tensor(feo.shape(19, 1, 256), a), feo.Tensor(feo.shape(1, 23, 256), b)
    want = oracle.ewise_broadcast("mul", a, b)
    try:
        for knob in (0, 2, 3, 4, 5, 6, 7):
            for lchunk in (0, 1, 5, 64):
                ctx.tune(L.KNOB_BCAST_TILE, knob)
                ctx.tune(L.KNOB_BCAST_LCHUNK, lchunk)
                np.testing.assert_array_equal(ta.broadcast("mul", tb).to_numpy(), want, err_msg=f"knob {knob} lchunk {lchunk}")
                np.testing.assert_array_equal(ta.broadcast("sub", tb).to_numpy(), oracle.ewise_broadcast("sub", a, b))
                # matrix (op) row vector: plain 4-rows-in-flight path (auto, 5) vs hold + stream (6)
                m, r = rand(oracle, rng, "f32", (37, 192)), rand(oracle, rng, "f32", (1, 192))
                np.testing.assert_array_equal(feo.Tensor(feo.shape(37, 192), m).broadcast("div", feo.Tensor(feo.shape(1, 192), r)).to_numpy(),
                                              oracle.ewise_broadcast("div", m, r), err_msg=f"knob {knob}")
    finally:
        ctx.tune(L.KNOB_BCAST_TILE, 0)
        ctx.tune(L.KNOB_BCAST_LCHUNK, 0)


# ---- reductions ---------------------------------------------------------------------------------------------------
REDUCE_CASES = [
    # (shape, axes)                                  kernel it lands on
    ((64, 128, 256), [0, 2]),                        # BASELINE config 1: [R,K,R] rows geometry
    ((64, 128, 256), [0]),                           # cols
    ((64, 128, 256), [2]),                           # rows, warp per output
    ((64, 128, 256), [1]),                           # [K,R,K] cols with k1 > 1
    ((64, 128, 256), [0, 1]), ((64, 128, 256), [1, 2]), ((64, 128, 256), [0, 1, 2]), ((64, 128, 256), []),
    ((300, 200, 8), [0]), ((300, 200, 8), [2]), ((300, 200, 8), [0, 2]), ((300, 200, 8), [1, 2]),  # config-3 pattern (micro r2=8)
    ((1000, 16), [1]), ((1000, 4), [1]), ((1000, 2), [1]), ((77, 32), [1]), ((50, 3, 2), [0, 2]),   # micro widths
    ((1000, 3), [1]), ((1000, 5), [1]), ((129, 7, 6), [0, 2]),                                     # micro, run-time r2
    ((5, 1000), [1]), ((5, 1001), [1]), ((3, 40000), [1]), ((1, 100003), [1]),                     # rows, ragged, split columns
    ((100003,), [0]), ((1 << 21,), []),                                                            # full reductions, split
    ((4097, 33), [0]), ((20000, 4), [0]), ((7, 9), [0]), ((1, 9), [0]), ((9, 1), [0]),             # cols: ragged / tiny
    ((3, 4, 5, 6, 7), [0, 2, 4]), ((3, 4, 5, 6, 7), [1, 3]), ((2, 3, 4, 5, 6, 7), [0, 2, 4]),      # >4 groups: generic
    ((6, 5, 4, 3), [1, 3]), ((6, 5, 4, 3), [0, 2]), ((6, 5, 4, 3), [1, 2]),
    ((16, 16, 16, 16, 16), [1, 3]), ((16, 16, 16, 16, 16), [0, 2, 4]), ((9, 3, 17, 8, 5, 12), [1, 3, 5]),  # outermost-first multi-pass
    ((2, 20, 3, 9, 33, 4), [1, 3, 5]), ((12, 2, 3, 40, 7), [0, 2, 4]),                                 # ... and its suffix alternative
    ((5000, 3), [0]), ((5000, 16), [0]), ((40, 300, 16), [1]), ((40, 300, 3), [1]), ((7, 2, 12), [1]),   # cols2d: short k2 rows
    ((9, 5000, 7), [1]), ((3, 100, 500), [1]), ((100, 128 * 4), [0]),                                 # cols2d at its lx limits / cols
    ((2, 5000), [0]), ((3, 4099), [0]), ((5, 8192), [0]), ((3, 3000, 1025), [1]),                     # cols with few rows, scalar columns
    ((3, 16385), [0]), ((7, 40001), [0]), ((2, 3, 17001), [1]),                                       # ... 8 scalar columns per thread
    ((300, 4099), [1]), ((64, 1 << 15), [1]), ((17, 100001), [1]),                                    # rows, unaligned lengths
    ((200, 5, 11), [0, 2]), ((4000, 9), [1]), ((300, 31), [1]), ((300, 17, 3), [2]),                  # micro, run-time r2 incl. > 8
    ((1,), [0]), ((1, 1, 1), [1]), ((2, 1, 3), [1]),
]


def _terms(x, axes):
    ax = tuple(axes) if axes else None
    return np.atleast_1d(np.abs(x.astype(np.float64)).sum(axis=ax)), int(np.prod([x.shape[a] for a in axes]) if axes else x.size)


@pytest.mark.parametrize("dt", ALL)
def test_reductions_vs_oracle(feo, oracle, dt):
    rng = np.random.default_rng(300)
    for shape, axes in REDUCE_CASES:
        x = rand(oracle, rng, dt, shape)
        t = feo.Tensor(feo.Shape(shape), x)
        for kind in ("sum", "mean", "var", "max", "min"):
            got = getattr(t, kind)(axes)
            want = oracle.reduce(kind, x, axes)
            assert tuple(got.shape().dims) == want.shape, (shape, axes, kind)
            g = got.to_numpy()
            msg = f"{dt} {kind} {shape} {axes}"
            if dt in INTS or kind in ("max", "min"):
                np.testing.assert_array_equal(g, want, err_msg=msg)
                continue
            abs_sum, n = _terms(x, axes)
            nd = float(oracle.reduce_divisor(dt, shape, axes))
            if kind == "sum":
                assert_sum_like(g.reshape(-1), want.reshape(-1), abs_sum.reshape(-1), n, dt, msg)
            elif kind == "mean":
                assert_sum_like(g.reshape(-1), want.reshape(-1), abs_sum.reshape(-1) / nd * (1 + 4 * U[dt]), n + 1, dt, msg)
            else:
                # var: compare with the exact (f64 / extended) value; both the two-pass reference and the
                # one-pass Welford result are within a few gamma_n of sum (x-mu)^2 / n
                xx = x.astype(np.longdouble)
                ax = tuple(axes) if axes else None
                exact = np.atleast_1d(xx.var(axis=ax)).astype(np.float64).reshape(-1)
                scale = np.atleast_1d((xx * xx).sum(axis=ax)).astype(np.float64).reshape(-1) / nd
                bound = 8 * gamma(n + 2, dt) * scale + 1e-300
                assert np.all(np.abs(g.reshape(-1).astype(np.float64) - exact) <= bound), msg
                assert np.all(np.abs(want.reshape(-1).astype(np.float64) - exact) <= bound), msg + " (oracle)"


NONFINITE_CASES = [((6, 40), [1]), ((6, 40), [0]), ((6, 40), []), ((3, 5, 64), [0, 2]), ((300, 3), [1]), ((2, 5000), [1]),
                   ((64, 128, 256), [0, 2]), ((64, 128, 256), [0]), ((4, 70000), [1]), ((9, 31), [1]), ((5000, 6), [0])]


def _one_target_slice(shape, axes):
    """Index tuple selecting every element that reduces into ONE output (the first target), for `axes` removed."""
    if not axes:
        return tuple(slice(None) for _ in shape)
    return tuple(slice(None) if d in axes else min(1, shape[d] - 1) for d in range(len(shape)))


@pytest.mark.parametrize("dt", FLOATS)
def test_max_min_with_infinities_match_the_reference(feo, oracle, dt):
    """Slices that are entirely -inf / +inf, and stray infinities elsewhere: no NaN is involved, so the reference's init
    fold (tensor.rs:234-237) is well defined and the GPU must return the oracle's values bit for bit -- a -FLT_MAX / +FLT_MAX
    seed would not.  Both the scalar path (axes == []) and the axis path, small and split geometries."""
    rng = np.random.default_rng(310)
    for shape, axes in NONFINITE_CASES:
        for fill in (-np.inf, np.inf):
            x = rand(oracle, rng, dt, shape)
            x[_one_target_slice(shape, axes)] = fill
            flat = x.reshape(-1)
            flat[rng.integers(0, flat.size, 5)] = -fill     # stray opposite infinities
            t = feo.Tensor(feo.Shape(shape), x)
            for kind in ("max", "min"):
                got, want = getattr(t, kind)(axes).to_numpy(), oracle.reduce(kind, x, axes)
                np.testing.assert_array_equal(got, want, err_msg=f"{dt} {kind} {shape} {axes} fill {fill}")
            s = t.sum(axes).to_numpy()
            assert np.array_equal(np.isfinite(s), np.isfinite(oracle.reduce("sum", x, axes))), f"{dt} sum {shape} {axes}"


@pytest.mark.parametrize("dt", FLOATS)
def test_nan_inputs_neither_crash_nor_poison_other_outputs(feo, oracle, dt):
    """NaN ordering is unspecified by SURVEY 8a-quirks (the reference's result depends on WHERE in the whole tensor the NaN
    sits).  What the GPU path guarantees, and this pins: a NaN never wins max / min (the reference's strict compares skip
    it too, tensor.rs:248), an all-NaN slice yields the seed (-inf / +inf), a NaN reaches only the sum / mean / var outputs
    it belongs to, and every output whose slice holds no NaN is exactly what it is without the NaN anywhere."""
    rng = np.random.default_rng(311)
    for shape, axes in NONFINITE_CASES:
        x = rand(oracle, rng, dt, shape)
        clean = feo.Tensor(feo.Shape(shape), x)
        y = x.copy()
        y[_one_target_slice(shape, axes)] = np.nan                     # one output's whole slice
        flat = y.reshape(-1)
        flat[rng.integers(0, flat.size, 7)] = np.nan                   # and a few in the middle of others
        t = feo.Tensor(feo.Shape(shape), y)
        ax = tuple(axes) if axes else None
        has_nan = np.atleast_1d(np.isnan(y).any(axis=ax))
        all_nan = np.atleast_1d(np.isnan(y).all(axis=ax))
        for kind, npf, seed in (("max", np.fmax, -np.inf), ("min", np.fmin, np.inf)):
            got = getattr(t, kind)(axes).to_numpy().reshape(-1)
            want = np.atleast_1d(npf.reduce(y, axis=ax) if ax is not None else npf.reduce(y.reshape(-1))).reshape(-1).copy()
            want[all_nan.reshape(-1)] = seed
            np.testing.assert_array_equal(got, want, err_msg=f"{dt} {kind} {shape} {axes}")
        for kind in ("sum", "mean", "var"):
            got = getattr(t, kind)(axes).to_numpy().reshape(-1)
            ref = getattr(clean, kind)(axes).to_numpy().reshape(-1)
            assert np.array_equal(np.isnan(got), has_nan.reshape(-1)), f"{dt} {kind} {shape} {axes}: NaN reached the wrong outputs"
            keep = ~has_nan.reshape(-1)
            np.testing.assert_array_equal(got[keep], ref[keep], err_msg=f"{dt} {kind} {shape} {axes}: clean outputs changed")
        # signed zeros: compared by value (which zero wins a tie is unspecified)
        z = np.zeros(shape, npdt(oracle, dt))
        z.reshape(-1)[::2] = -0.0
        tz = feo.Tensor(feo.Shape(shape), z)
        assert np.all(tz.max(axes).to_numpy() == 0) and np.all(tz.min(axes).to_numpy() == 0) and np.all(tz.sum(axes).to_numpy() == 0)


def test_reduction_split_counts_agree(feo, oracle):
    """Forcing the split count exercises stage 2; integers must stay bit-exact for every split."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(301)
    cases = [((513, 260), [0]), ((300, 64, 8), [0, 2]), ((9, 70000), [1]), ((200000,), [0])]
    try:
        for shape, axes in cases:
            x = rand(oracle, rng, "i64", shape)
            xf = rand(oracle, rng, "f32", shape)
            t, tf = feo.Tensor(feo.Shape(shape), x), feo.Tensor(feo.Shape(shape), xf)
            for S, fuse_off in ((0, 0), (1, 0), (2, 2), (3, 1), (7, 2), (7, 1), (16, 2), (16, 1), (0, 1), (0, 2)):
                ctx.tune(L.KNOB_REDUCE_SPLIT, S)
                ctx.tune(L.KNOB_REDUCE_FUSE, fuse_off)  # fused last-arriver combine vs the two-launch combine: same bits
                for kind in ("sum", "var", "max", "min", "mean"):
                    np.testing.assert_array_equal(getattr(t, kind)(axes).to_numpy(), oracle.reduce(kind, x, axes), err_msg=f"S={S} {kind} {shape}")
                np.testing.assert_array_equal(tf.max(axes).to_numpy(), oracle.reduce("max", xf, axes))
                # float results are deterministic for a given configuration (no atomics on data): same bits every time
                np.testing.assert_array_equal(tf.sum(axes).to_numpy(), tf.sum(axes).to_numpy())
                np.testing.assert_array_equal(tf.var(axes).to_numpy(), tf.var(axes).to_numpy())
                abs_sum, n = _terms(xf, axes)
                assert_sum_like(tf.sum(axes).to_numpy().reshape(-1), oracle.reduce("sum", xf, axes).reshape(-1), abs_sum, n, "f32")
    finally:
        ctx.tune(L.KNOB_REDUCE_SPLIT, 0)
        ctx.tune(L.KNOB_REDUCE_FUSE, 0)


def test_outputs_stay_inside_their_buffers(feo, oracle):
    """Every kernel family writes exactly its output: the result sits between two canary bands in a larger device buffer
    (and the operands sit at the very end of theirs, so a read past them would at least pick up the next allocation)."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
    rng = np.random.default_rng(77)
    PAD = 64  # elements; keeps 16-byte alignment for every dtype

    def guarded(n, dt):
        buf = feo.Tensor.fill(feo.shape(n + 2 * PAD), 77, dtype=dt)
        return buf, buf.data.ptr + PAD * np.dtype(oracle.NP_DTYPES[dt]).itemsize

    def check(buf, n, want, msg):
        got = buf.to_numpy()
        np.testing.assert_array_equal(got[:PAD], 77, err_msg=msg + " (wrote before the output)")
        np.testing.assert_array_equal(got[PAD + n:], 77, err_msg=msg + " (wrote past the output)")
        return got[PAD:PAD + n]

    for dt in ("f32", "i64"):
        code = L.code_of(dt)
        for shape, axes in REDUCE_CASES:
            x = rand(oracle, rng, dt, shape)
            t = feo.Tensor(feo.Shape(shape), x)
            for kind in ("sum", "var", "max"):
                want = oracle.reduce(kind, x, axes).reshape(-1)
                buf, optr = guarded(want.size, dt)
                ctx.check(lib.feo_reduce(h, L.KIND_CODES[kind], code, vp(optr), vp(t.data.ptr), len(shape), L.sizes(shape), len(axes), L.sizes(axes)))
                got = check(buf, want.size, want, f"reduce {dt} {kind} {shape} {axes}")
                if dt == "i64" or kind == "max":
                    np.testing.assert_array_equal(got, want)
        for sa, sb in BCAST_CASES:
            a = rand(oracle, rng, dt, sa)
            b = (np.abs(rand(oracle, rng, dt, sb)) + 1).astype(a.dtype)
            want = oracle.ewise_broadcast("mul", a, b)
            ta, tb = feo.Tensor(feo.Shape(sa), a), feo.Tensor(feo.Shape(sb), b)
            oshape = list(want.shape)

            def strides(src):
                st, run = [0] * len(src), 1
                for d in range(len(src) - 1, -1, -1):
                    st[d] = run if src[d] != 1 or oshape[d] == 1 else 0
                    run *= src[d]
                return st
            buf, optr = guarded(want.size, dt)
            ctx.check(lib.feo_ewise_broadcast(h, L.MUL, code, vp(optr), vp(ta.data.ptr), vp(tb.data.ptr), len(oshape), L.sizes(oshape),
                                              L.sizes(strides(sa)), L.sizes(strides(sb))))
            np.testing.assert_array_equal(check(buf, want.size, want, f"bcast {dt} {sa} {sb}"), want.reshape(-1))


def test_integer_variance_one_pass_equals_the_two_passes(feo, oracle):
    """Integer var runs as one pass (wrapping sum and sum of squares; the ring identity sum (x - m)^2 = S2 - 2 m S1 + n m^2
    holds modulo 2^k whatever overflows).  It must give the reference's two-pass result bit for bit -- also when nearly
    every intermediate wraps -- and agree with the literal two-pass kernels (knob REDUCE_FUSE = 3)."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(303)
    shapes = [((300, 200, 8), [0]), ((300, 200, 8), [1]), ((300, 200, 8), [0, 2]), ((5000, 3), [0]), ((5000, 3), [1]), ((64, 4099), [1]),
              ((3, 4, 5, 6, 7), [0, 2, 4]), ((100003,), [0]), ((40, 300, 16), [1]), ((2, 5000), [0])]
    ranges = {"i32": (-2 ** 31, 2 ** 31 - 1), "u32": (0, 2 ** 32 - 1), "i64": (-2 ** 63, 2 ** 63 - 1)}
    try:
        for dt, (lo, hi) in ranges.items():
            for shape, axes in shapes:
                for wide in (False, True):
                    x = rng.integers(lo, hi, shape, dtype=oracle.NP_DTYPES[dt], endpoint=True) if wide else rand(oracle, rng, dt, shape)
                    t = feo.Tensor(feo.Shape(shape), x)
                    want = oracle.reduce("var", x, axes)
                    ctx.tune(L.KNOB_REDUCE_FUSE, 0)
                    one = t.var(axes).to_numpy()
                    ctx.tune(L.KNOB_REDUCE_FUSE, 3)
                    two = t.var(axes).to_numpy()
                    np.testing.assert_array_equal(one, want, err_msg=f"one-pass {dt} {shape} {axes} wide={wide}")
                    np.testing.assert_array_equal(two, want, err_msg=f"two-pass {dt} {shape} {axes} wide={wide}")
    finally:
        ctx.tune(L.KNOB_REDUCE_FUSE, 0)
    ctx.arith_flag(clear=True)


def test_many_group_reductions_multipass_and_generic_agree(feo, oracle):
    """More than four alternating kept / reduced groups: the multi-pass route (sum / mean / max / min) and the
    one-kernel generic fallback (knob REDUCE_SPLIT = -1; always used for var) give the oracle's integers bit for bit."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(302)
    cases = [((3, 4, 5, 6, 7), [0, 2, 4]), ((3, 4, 5, 6, 7), [1, 3]), ((2, 3, 4, 5, 6, 7), [0, 2, 4]), ((2, 3, 4, 5, 6, 7), [1, 3, 5]),
             ((4, 3, 5, 2, 6, 3, 7), [0, 2, 4, 6]), ((4, 3, 5, 2, 6, 3, 7), [1, 3, 5]), ((8, 9, 10, 11, 12), [0, 2, 4])]
    try:
        for shape, axes in cases:
            x = rand(oracle, rng, "i32", shape)
            xf = rand(oracle, rng, "f64", shape)
            t, tf = feo.Tensor(feo.Shape(shape), x), feo.Tensor(feo.Shape(shape), xf)
            for knob in (0, -1):
                ctx.tune(L.KNOB_REDUCE_SPLIT, knob)
                for kind in ("sum", "mean", "var", "max", "min"):
                    np.testing.assert_array_equal(getattr(t, kind)(axes).to_numpy(), oracle.reduce(kind, x, axes), err_msg=f"{knob} {kind} {shape} {axes}")
                np.testing.assert_array_equal(tf.max(axes).to_numpy(), oracle.reduce("max", xf, axes))
                np.testing.assert_allclose(tf.sum(axes).to_numpy(), oracle.reduce("sum", xf, axes), rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(tf.mean(axes).to_numpy(), oracle.reduce("mean", xf, axes), rtol=1e-12, atol=1e-13)
                # float var: canonical Welford pass + pair merge (knob 0) or the one-kernel generic fallback (knob -1)
                np.testing.assert_allclose(tf.var(axes).to_numpy(), oracle.reduce("var", xf, axes), rtol=1e-10, atol=1e-13)
    finally:
        ctx.tune(L.KNOB_REDUCE_SPLIT, 0)


def test_mean_divisor_quirk_on_device(feo, oracle):
    """Vector::mean() takes the axes=[] branch (vector.rs:51): n counts the whole size in T and saturates at
    2^24 for f32 (tensor.rs:125-129).  All-ones data makes the sum exact, so the quirk is visible bit for bit."""
    n = (1 << 24) + 64
    v = feo.Tensor.fill(feo.shape(n), 0.5, "f32")
    # every partial sum is a multiple of 0.5 below 2^24, hence exact in any order: sum = n/2; divisor = 2^24, not n
    assert v.sum([]).to_numpy()[0] == np.float32(n // 2)
    assert v.mean([]).to_numpy()[0] == np.float32(n // 2) / np.float32(1 << 24)
    assert v.mean([0]).to_numpy()[0] == np.float32(n // 2) / np.float32(1 << 24)
    assert v.mean([]).to_numpy()[0] != np.float32(0.5)
    assert oracle.reduce_divisor("f32", (n,), []) == np.float32(1 << 24)


def test_partial_reductions_and_var_merge(feo, oracle):
    """The building blocks of the sharded path: per-shard partials combined by feo_var_merge equal the
    single-device result within the float bound (bit-exact for the integer two-pass scheme)."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    lib = ctx.lib
    rng = np.random.default_rng(302)
    G, rows, K = 4, 64, 96
    x = rand(oracle, rng, "f32", (G * rows, K))
    shards = [feo.Tensor(feo.shape(rows, K), x[g * rows:(g + 1) * rows]) for g in range(G)]
    partials = feo.Tensor.zeros(feo.shape(G, 2, K), "f32")
    for g, s in enumerate(shards):
        ctx.check(lib.feo_reduce_partial(ctx.handle, L.VAR, L.F32, C.c_void_p(partials.data.ptr + g * 2 * K * 4),
                                         C.c_void_p(s.data.ptr), 2, L.sizes([rows, K]), 1, L.sizes([0])))
    out = feo.Tensor.zeros(feo.shape(K), "f32")
    div = np.array([oracle.reduce_divisor("f32", (G * rows, K), [0])], np.float32)
    ctx.check(lib.feo_var_merge(ctx.handle, L.F32, C.c_void_p(out.data.ptr), C.c_void_p(partials.data.ptr), G, K, rows,
                                div.ctypes.data_as(C.c_void_p)))
    exact = x.astype(np.float64).var(axis=0)
    np.testing.assert_allclose(out.to_numpy(), exact, rtol=2e-5)
    np.testing.assert_allclose(out.to_numpy(), feo.Tensor(feo.shape(G * rows, K), x).var([0]).to_numpy(), rtol=2e-5)
    # integer two-pass across shards: sum partials -> truncated mean -> squared deviations -> divide
    xi = rand(oracle, rng, "i32", (G * rows, K))
    ishards = [feo.Tensor(feo.shape(rows, K), xi[g * rows:(g + 1) * rows]) for g in range(G)]
    total = None
    for s in ishards:
        p = feo.Tensor.zeros(feo.shape(K), "i32")
        ctx.check(lib.feo_reduce_partial(ctx.handle, L.SUM, L.I32, C.c_void_p(p.data.ptr), C.c_void_p(s.data.ptr), 2,
                                         L.sizes([rows, K]), 1, L.sizes([0])))
        total = p if total is None else total + p
    n = np.int32(oracle.reduce_divisor("i32", (G * rows, K), [0]))
    mean = total / n
    sq = None
    for s in ishards:
        p = feo.Tensor.zeros(feo.shape(K), "i32")
        ctx.check(lib.feo_reduce_sqdev(ctx.handle, L.I32, C.c_void_p(p.data.ptr), C.c_void_p(s.data.ptr), C.c_void_p(mean.data.ptr),
                                       2, L.sizes([rows, K]), 1, L.sizes([0])))
        sq = p if sq is None else sq + p
    np.testing.assert_array_equal((sq / n).to_numpy(), oracle.reduce("var", xi, [0]))
    # ... and in one pass: VAR partials of integers are wrapping (S1 | S2) pairs, merged exactly by feo_var_merge
    for dt, code in (("i32", L.I32), ("i64", L.I64), ("u32", L.U32)):
        info = np.iinfo(oracle.NP_DTYPES[dt])
        xw = rng.integers(info.min, info.max, (G * rows, K), dtype=oracle.NP_DTYPES[dt], endpoint=True)  # nearly everything wraps
        es = xw.itemsize
        parts = feo.Tensor.zeros(feo.shape(G, 2, K), dt)
        for g in range(G):
            sh = feo.Tensor(feo.shape(rows, K), xw[g * rows:(g + 1) * rows])
            ctx.check(lib.feo_reduce_partial(ctx.handle, L.VAR, code, C.c_void_p(parts.data.ptr + g * 2 * K * es), C.c_void_p(sh.data.ptr),
                                             2, L.sizes([rows, K]), 1, L.sizes([0])))
        outi = feo.Tensor.zeros(feo.shape(K), dt)
        divi = np.array([oracle.reduce_divisor(dt, (G * rows, K), [0])], oracle.NP_DTYPES[dt])
        ctx.check(lib.feo_var_merge(ctx.handle, code, C.c_void_p(outi.data.ptr), C.c_void_p(parts.data.ptr), G, K, rows,
                                    divi.ctypes.data_as(C.c_void_p)))
        np.testing.assert_array_equal(outi.to_numpy(), oracle.reduce("var", xw, [0]), err_msg=dt)
    ctx.arith_flag(clear=True)


# ---- products ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ALL)
def test_outer_product_bit_exact(feo, oracle, dt):
    rng = np.random.default_rng(400)
    for na, nb in ((1, 1), (3, 5), (64, 64), (129, 1000), (1000, 129), (33, 4096)):
        a, b = rand(oracle, rng, dt, na), rand(oracle, rng, dt, nb)
        r = feo.Vector(a).prod(feo.Vector(b))
        assert r.shape().dims == [na, nb]
        np.testing.assert_array_equal(r.to_numpy(), oracle.prod(a, b), err_msg=f"{dt} {na}x{nb}")
    a, b = rand(oracle, rng, dt, (3, 4)), rand(oracle, rng, dt, (2, 8))
    r = feo.Tensor(feo.shape(3, 4), a).prod(feo.Tensor(feo.shape(2, 8), b))
    assert r.shape().dims == [3, 4, 2, 8]
    np.testing.assert_array_equal(r.to_numpy(), oracle.prod(a, b))


@pytest.mark.parametrize("dt", ALL)
def test_dot_matvec_vecmat(feo, oracle, dt):
    rng = np.random.default_rng(500)
    for n in (1, 2, 31, 1000, 4099, 1 << 18):
        a, b = rand(oracle, rng, dt, n), rand(oracle, rng, dt, n)
        got = feo.Vector(a).vecmul(feo.Vector(b))
        want = oracle.dot(a, b)
        assert got.shape().dims == [1]
        if dt in INTS:
            np.testing.assert_array_equal(got.to_numpy(), want)
        else:
            assert_sum_like(got.to_numpy(), want, np.abs(a.astype(np.float64) * b).sum(), n + 1, dt, f"dot n={n}")
    for m, k in ((1, 1), (7, 5), (33, 1000), (500, 36), (64, 4096), (3, 5000), (2, 4099), (5000, 3)):
        A, v, w = rand(oracle, rng, dt, (m, k)), rand(oracle, rng, dt, k), rand(oracle, rng, dt, m)
        tA = feo.Matrix(feo.shape(m, k), A)
        g1, w1 = tA.vecmul(feo.Vector(v)).to_numpy(), oracle.matvec(A, v)
        g2, w2 = feo.Vector(w).matmul(tA).to_numpy(), oracle.vecmat(w, A)
        B2 = rand(oracle, rng, dt, (m, k))
        g3, w3 = feo.dot_batched(tA, feo.Matrix(feo.shape(m, k), B2)).to_numpy(), oracle.dot_batched(A, B2)
        if dt in INTS:
            np.testing.assert_array_equal(g1, w1); np.testing.assert_array_equal(g2, w2); np.testing.assert_array_equal(g3, w3)
        else:
            assert_sum_like(g1, w1, (np.abs(A.astype(np.float64)) * np.abs(v)).sum(axis=1), k + 1, dt, f"matvec {m}x{k}")
            assert_sum_like(g2, w2, (np.abs(A.astype(np.float64)) * np.abs(w)[:, None]).sum(axis=0), m + 1, dt, f"vecmat {m}x{k}")
            assert_sum_like(g3, w3, (np.abs(A.astype(np.float64)) * np.abs(B2)).sum(axis=1), k + 1, dt, f"bdot {m}x{k}")


@pytest.mark.parametrize("dt", ALL)
def test_matmul_simt_vs_oracle(feo, oracle, dt):
    rng = np.random.default_rng(600)
    for m, k, n in ((1, 1, 1), (2, 3, 4), (17, 33, 9), (128, 128, 128), (256, 64, 384), (130, 70, 257), (64, 1000, 48)):
        A, B = rand(oracle, rng, dt, (m, k)), rand(oracle, rng, dt, (k, n))
        got = feo.Matrix(feo.shape(m, k), A).matmul(feo.Matrix(feo.shape(k, n), B), algo="simt")
        want = oracle.matmul(A, B)
        assert got.shape().dims == [m, n]
        if dt in INTS:
            np.testing.assert_array_equal(got.to_numpy(), want, err_msg=f"{dt} {m}x{k}x{n}")
        else:
            terms = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
            assert_sum_like(got.to_numpy(), want, terms, k + 1, dt, f"matmul {m}x{k}x{n}")


def test_matmul_simt_pipeline_variants_agree(feo, oracle):
    """Aligned SIMT matmul: the cp.async pipeline (auto and every stage / k-vector variant) accumulates in the same
    per-thread k order as the register-double-buffered kernel, so all of them agree bit for bit -- and with the oracle for
    integers.  Sizes cover fewer k-panels than pipeline stages, a single panel, and many tiles."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(601)
    try:
        for dt in ("f32", "i32", "u32", "f64", "i64"):
            for m, k, n in ((128, 16, 128), (128, 32, 256), (256, 48, 128), (384, 1024, 256), (1024, 512, 640), (128, 8, 128), (256, 24, 384)):
                if k % 16 and dt in ("f32", "i32", "u32"):
                    continue  # 4-byte panels are 16 deep: these shapes take the guarded kernel anyway
                A, B = rand(oracle, rng, dt, (m, k)), rand(oracle, rng, dt, (k, n))
                ta, tb = feo.Matrix(feo.shape(m, k), A), feo.Matrix(feo.shape(k, n), B)
                ctx.tune(L.KNOB_MATMUL_SIMT, 1)
                base = ta.matmul(tb, algo="simt").to_numpy()
                if dt in INTS:
                    np.testing.assert_array_equal(base, oracle.matmul(A, B), err_msg=f"{dt} {m}x{k}x{n}")
                for v in (0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13):
                    ctx.tune(L.KNOB_MATMUL_SIMT, v)
                    np.testing.assert_array_equal(ta.matmul(tb, algo="simt").to_numpy(), base, err_msg=f"{dt} {m}x{k}x{n} variant {v}")
    finally:
        ctx.tune(L.KNOB_MATMUL_SIMT, 0)


def test_device_generator_matches_host_twin(feo, oracle):
    for dt, lo, hi in (("f32", -1.0, 1.0), ("f32", 1.0, 2.0), ("f64", -1.0, 1.0), ("i32", -1000, 1000), ("i64", -1000, 1000), ("u32", 0, 2000)):
        t = feo.Tensor.uniform(feo.shape(1000, 37), seed=42, lo=lo, hi=hi, dtype=dt, first_index=12345)
        want = oracle.fill_uniform(dt, 37000, 42, 12345, lo, hi)
        np.testing.assert_array_equal(t.to_numpy().reshape(-1), want)
        assert want.min() >= lo and want.max() < hi
    # strided regeneration of single columns
    np.testing.assert_array_equal(oracle.fill_uniform("f32", 10, 42, 5, -1, 1, stride=37),
                                  oracle.fill_uniform("f32", 37 * 10, 42, 0, -1, 1)[5::37])


TF32X3_CASES = [(128, 256, 256), (256, 96, 512), (128, 4096, 256), (384, 1024, 768), (1024, 1024, 1024),
                (130, 70, 257), (100, 1000, 300), (1, 3000, 1000), (257, 33, 513), (511, 129, 255), (64, 4099, 64)]  # ragged edges


def test_matmul_tf32x3_vs_oracle(feo, oracle):
    """tcgen05 split-TF32 matmul: |gpu - exact| <= max(2*|ref - exact|, 2^-20 * sum|a||b|), where ref is
    the oracle's sequential f32 result and exact the f64 product; also inside the generic reassociation bound."""
    rng = np.random.default_rng(700)
    for m, k, n in TF32X3_CASES:
        A, B = rand(oracle, rng, "f32", (m, k)), rand(oracle, rng, "f32", (k, n))
        tA, tB = feo.Matrix(feo.shape(m, k), A), feo.Matrix(feo.shape(k, n), B)
        got = tA.matmul(tB, algo="tf32x3").to_numpy()
        exact = A.astype(np.float64) @ B.astype(np.float64)
        terms = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
        err = np.abs(got - exact)
        if m * k * n <= 384 * 1024 * 768:
            ref = oracle.matmul(A, B)
            bound = np.maximum(2 * np.abs(ref - exact), 2.0 ** -20 * terms)
            assert np.all(err <= bound), f"{m}x{k}x{n}: max err {err.max():.3e}, worst ratio {(err / bound).max():.2f}"
            assert_sum_like(got, ref, terms, k + 1, "f32", f"tf32x3 {m}x{k}x{n}")
        assert np.all(err <= 2.0 ** -20 * terms), f"{m}x{k}x{n}: max rel err {(err / terms).max():.3e}"
        # AUTO picks the tensor-core path for supported f32 shapes and agrees with it bit for bit
        np.testing.assert_array_equal(tA.matmul(tB).to_numpy(), got)
    # exact small-integer data: every product and partial sum is exact in TF32 x f32
    A = rng.integers(-8, 9, (128, 64)).astype(np.float32)
    B = rng.integers(-8, 9, (64, 256)).astype(np.float32)
    got = feo.Matrix(feo.shape(128, 64), A).matmul(feo.Matrix(feo.shape(64, 256), B), algo="tf32x3").to_numpy()
    np.testing.assert_array_equal(got, A @ B)
    # long K on positive data: the chunked accumulation keeps the tensor core's truncation bias from growing with K
    A, B = np.abs(rand(oracle, rng, "f32", (128, 16384))), np.abs(rand(oracle, rng, "f32", (16384, 256)))
    got = feo.Matrix(feo.shape(128, 16384), A).matmul(feo.Matrix(feo.shape(16384, 256), B), algo="tf32x3").to_numpy()
    exact = A.astype(np.float64) @ B.astype(np.float64)
    assert np.max(np.abs(got - exact) / exact) < 3e-6
    # tiny products are refused for the explicit algorithm (AUTO runs them on the SIMT kernel)
    with pytest.raises(feo.BackendError):
        feo.Matrix(feo.shape(4, 4), np.zeros(16, np.float32)).matmul(feo.Matrix(feo.shape(4, 4), np.zeros(16, np.float32)), algo="tf32x3")


def test_matmul_tf32x3_kernel_variants(feo, oracle):
    """The CTA-pair kernel (tcgen05 cta_group::2, persistent, run-time tile width BN) at every BN -- with B read as raw f32
    tiles and split inside the kernel (MN-major operand; the default when N % 4 == 0) and with the split + transpose pre-pass
    (knob MATMUL_TF32_BRAW = 1) -- and the single-CTA kernel it replaced (knob MATMUL_TF32_PAIR = 1): each inside the same
    f32 bound and all bit-identical, on shapes that give a pair several tiles (persistence: barrier phases carry over),
    ragged edges in M, N and K (N = 4 x odd: the last 32-column box of raw B overhangs; K % 32 != 0: TMA zero-fills the
    missing k-rows), an odd number of 128-row blocks (the odd CTA's rows fall outside M) and a single k-block."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(701)
    shapes = [(1024, 512, 8192 + 224), (2304, 160, 4000), (128, 64, 700), (385, 32, 257), (3000, 96, 3001), (640, 8192, 512), (515, 70, 1012)]
    try:
        for m, k, n in shapes:
            A, B = rand(oracle, rng, "f32", (m, k)), rand(oracle, rng, "f32", (k, n))
            tA, tB = feo.Matrix(feo.shape(m, k), A), feo.Matrix(feo.shape(k, n), B)
            exact = A.astype(np.float64) @ B.astype(np.float64)
            terms = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
            results = {}
            for label, pair, bn, braw in (("pair auto", 0, 0, 0), ("pair 256", 0, 256, 0), ("pair 224", 0, 224, 0), ("pair 192", 0, 192, 0),
                                          ("pair 160", 0, 160, 0), ("pair 128", 0, 128, 0), ("pair 96", 0, 96, 0), ("pair 32", 0, 32, 0),
                                          ("single", 1, 0, 0), ("pair prepass", 0, 0, 1), ("pair 224 prepass", 0, 224, 1), ("pair 96 prepass", 0, 96, 1)):
                ctx.tune(L.KNOB_MATMUL_TF32_PAIR, pair)
                ctx.tune(L.KNOB_MATMUL_TF32_BN, bn)
                ctx.tune(L.KNOB_MATMUL_TF32_BRAW, braw)
                got = tA.matmul(tB, algo="tf32x3").to_numpy()
                err = np.abs(got - exact)
                assert np.all(err <= 2.0 ** -20 * terms), f"{label} {m}x{k}x{n}: max rel err {(err / terms).max():.3e} at {np.unravel_index(np.argmax(err / terms), err.shape)}"
                results[label] = got
            # the k order and the chunking are the same whatever the tile width: every variant produces the same bits
            for label, got in results.items():
                np.testing.assert_array_equal(got, results["pair auto"], err_msg=f"{label} {m}x{k}x{n}")
    finally:
        ctx.tune(L.KNOB_MATMUL_TF32_PAIR, 0)
        ctx.tune(L.KNOB_MATMUL_TF32_BN, 0)
        ctx.tune(L.KNOB_MATMUL_TF32_BRAW, 0)


def test_broadcast_cache_policy_knobs_keep_the_bits(feo, oracle):
    """The L2 policies of the tiled broadcast kernel are performance knobs only: evict_last hints (default), none, a
    persisting access-policy window on the re-read operand, and the L2 prefetch of its rows for later CTAs must all give the
    reference's bits (one IEEE multiply per element), on a shape where both operands are invariant somewhere (config 2's)
    and with the prefetch distance larger and smaller than the number of q-blocks."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    rng = np.random.default_rng(808)
    a = rand(oracle, rng, "f32", (24, 1, 512))
    b = rand(oracle, rng, "f32", (1, 700, 512))
    ta, tb = feo.Tensor(feo.shape(24, 1, 512), a), feo.Tensor(feo.shape(1, 700, 512), b)
    want = oracle.ewise_broadcast("mul", a, b)
    try:
        for l2keep, prefetch in ((0, 0), (1, 0), (2, 0), (16, 0), (0, 7), (0, 128), (0, 4096), (2, 64)):
            ctx.tune(L.KNOB_BCAST_L2KEEP, l2keep)
            ctx.tune(L.KNOB_BCAST_PREFETCH, prefetch)
            np.testing.assert_array_equal(ta.broadcast("mul", tb).to_numpy(), want, err_msg=f"l2keep={l2keep} prefetch={prefetch}")
    finally:
        ctx.tune(L.KNOB_BCAST_L2KEEP, 0)
        ctx.tune(L.KNOB_BCAST_PREFETCH, 0)


def test_pow_and_display(feo, kat, oracle):
    """The section-8(f) rows: pow (tensor.rs:325-333; exact on the reference's own vectors, a few ulp elsewhere) and
    display (tensor.rs:507-551; host-side formatting)."""
    for c in [c for c in kat if c["op"] == "pow"]:
        for dt in FLOATS:
            x = _mk(feo, c.get("via", "tensor"), c["shape"], c["a"], dt)
            r = x.pow(c["power"])
            assert type(r) is type(x) and r.shape().dims == c["shape"]
            np.testing.assert_array_equal(r.to_numpy().reshape(-1), np.array(c["expect"], npdt(oracle, dt)), err_msg=c["ref"])
    rng = np.random.default_rng(800)
    for dt, rtol in (("f32", 1e-6), ("f64", 1e-14)):
        a = rng.uniform(0.1, 4.0, 10007).astype(npdt(oracle, dt))
        t = feo.Tensor(feo.shape(a.size), a)
        for p in (2.0, 0.5, -1.0, 1.0, 0.0):   # correctly rounded multiply / sqrt / divide: within 1 ulp of glibc's (<1 ulp) powf
            np.testing.assert_array_max_ulp(t.pow(p).to_numpy(), oracle.pow(a, p), maxulp=1)
        for p in (3.0, 1.7, -2.5, 0.3333):
            np.testing.assert_allclose(t.pow(p).to_numpy(), oracle.pow(a, p), rtol=rtol, err_msg=f"{dt} pow {p}")
    with pytest.raises(TypeError):
        feo.Tensor(feo.shape(2), np.array([1, 2], np.int32)).pow(2)
    for c in [c for c in kat if c["op"] == "display"]:
        assert feo.Tensor(feo.Shape(c["shape"]), c["data"]).display() == c["text"], c["ref"]
    assert feo.Tensor(feo.shape(3), np.array([1.5, -0.25, 1e21], np.float64)).display() == "[1.5, -0.25, 1000000000000000000000]"
    assert feo.Tensor(feo.shape(2), np.array([7, -3], np.int32)).display() == "[7, -3]"


def test_overlapped_launches_respect_hazards(feo, oracle):
    """Independent back-to-back broadcast launches overlap (programmatic dependent launch); launches that read or
    overwrite what an in-flight launch produces must stay ordered.  Integer data: any mis-ordering changes the bits."""
    import feotensor_b200._lib as L
    ctx = feo.Context.default()
    lib = ctx.lib
    rng = np.random.default_rng(900)
    R, Q, D = 16, 96, 256
    a = rand(oracle, rng, "i32", (R, 1, D))
    b = rand(oracle, rng, "i32", (1, Q, D))
    ta, tb = feo.Tensor(feo.shape(R, 1, D), a), feo.Tensor(feo.shape(1, Q, D), b)
    ab = oracle.ewise_broadcast("add", a, b)
    for knob in (0, 1):  # PDL on / off give the same bits
        ctx.tune(L.KNOB_PDL, knob)
        # (1) RAW chain: x_{k+1} = x_k[R,Q,D] + b (b broadcast along dim 0); every launch reads the previous output
        x = ta.broadcast("add", tb)
        want = ab.copy()
        for _ in range(40):
            x = x.broadcast("add", tb)
            want = oracle.ewise_broadcast("add", want, b)
        np.testing.assert_array_equal(x.to_numpy(), want)
        # (2) 64 independent slab launches into one buffer, then overwrite every slab (WAW) and read it back
        out = feo.Tensor.zeros(feo.shape(64 * R, Q, D), "i32")
        slab = R * Q * D
        shape, sa, sb = L.sizes([R, Q, D]), L.sizes([D, 0, 1]), L.sizes([0, D, 1])
        for op in (L.ADD, L.MUL, L.SUB):
            for s in range(64):
                ctx.check(lib.feo_ewise_broadcast(ctx.handle, op, L.I32, C.c_void_p(out.data.ptr + 4 * s * slab), C.c_void_p(ta.data.ptr),
                                                  C.c_void_p(tb.data.ptr), 3, shape, sa, sb))
        got = out.to_numpy().reshape(64, R, Q, D)
        want_sub = oracle.ewise_broadcast("sub", a, b)
        for s in (0, 1, 31, 63):
            np.testing.assert_array_equal(got[s], want_sub)
        # (3) WAR: a launch overwrites the operand an earlier, possibly still running launch reads
        src = feo.Tensor(feo.shape(R, 1, D), a)
        o1 = src.broadcast("add", tb)                       # reads src
        ctx.check(lib.feo_ewise_broadcast(ctx.handle, L.MUL, L.I32, C.c_void_p(src.data.ptr), C.c_void_p(ta.data.ptr),
                                          C.c_void_p(ta.data.ptr), 3, L.sizes([R, 1, D]), L.sizes([D, 0, 1]), L.sizes([D, 0, 1])))  # overwrites src
        np.testing.assert_array_equal(o1.to_numpy(), ab)
        np.testing.assert_array_equal(src.to_numpy(), oracle.ewise("mul", a, a))
    ctx.tune(L.KNOB_PDL, 0)
