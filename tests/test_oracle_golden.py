"""Pins the CPU oracle against every known-answer test the reference holds for the hot path.

The cases in tests/golden/reference_kat.json are transcriptions of the reference's inline #[test]
functions (tests/golden/make_reference_kat.py, each with its file:line).  The reference asserts
exact equality, so do we -- in f64 (the literal type the reference tests use), in f32, and in the
integer dtypes wherever every number of the case is an integer.
"""
import numpy as np
import pytest

FLOATS = ["f64", "f32"]
INTS = ["i32", "i64"]


def _np(oracle, name):
    return oracle.NP_DTYPES[name]


def _all_int(*seqs):
    return all(float(x).is_integer() for s in seqs for x in (s if isinstance(s, list) else [s]))


def _dtypes_for(case_numbers_int, signed_needed):
    dts = list(FLOATS)
    if case_numbers_int:
        dts += INTS
        if not signed_needed:
            dts.append("u32")
    return dts


def _cases(kat, op):
    return [c for c in kat if c["op"] == op]


def test_kat_file_is_complete(kat):
    ops = {c["op"] for c in kat}
    for needed in ("reduce", "ewise", "ewise_scalar", "prod", "matmul", "matvec", "vecmat", "dot", "fill"):
        assert needed in ops
    assert len(kat) >= 110


def test_reduce_kats(oracle, kat):
    n = 0
    for c in _cases(kat, "reduce"):
        ints = _all_int(c["data"], c["expect"])
        neg = any(x < 0 for x in c["data"] + c["expect"])
        for dt in _dtypes_for(ints, neg):
            x = np.array(c["data"], dtype=_np(oracle, dt)).reshape(c["shape"])
            for faithful in (False, True):
                got = oracle.reduce(c["kind"], x, c["axes"], faithful=faithful, recompute_mean_per_target=faithful)
                assert list(got.shape) == c["expect_shape"], c["ref"]
                assert got.dtype == x.dtype
                np.testing.assert_array_equal(got.reshape(-1), np.array(c["expect"], dtype=x.dtype), err_msg=c["ref"])
                n += 1
    assert n > 100


def test_ewise_kats(oracle, kat):
    for c in _cases(kat, "ewise"):
        ints = _all_int(c["a"], c["b"], c["expect"])
        neg = any(x < 0 for x in c["a"] + c["b"] + c["expect"])
        for dt in _dtypes_for(ints, neg):
            a = np.array(c["a"], dtype=_np(oracle, dt)).reshape(c["shape"])
            b = np.array(c["b"], dtype=_np(oracle, dt)).reshape(c["shape"])
            for faithful in (False, True):
                got = oracle.ewise(c["opname"], a, b, faithful=faithful)
                np.testing.assert_array_equal(got.reshape(-1), np.array(c["expect"], dtype=a.dtype), err_msg=c["ref"])
                assert got.shape == a.shape


def test_ewise_scalar_kats(oracle, kat):
    for c in _cases(kat, "ewise_scalar"):
        ints = _all_int(c["a"], c["scalar"], c["expect"])
        neg = any(x < 0 for x in c["a"] + c["expect"])
        for dt in _dtypes_for(ints, neg):
            a = np.array(c["a"], dtype=_np(oracle, dt)).reshape(c["shape"])
            got = oracle.ewise_scalar(c["opname"], a, c["scalar"])
            np.testing.assert_array_equal(got.reshape(-1), np.array(c["expect"], dtype=a.dtype), err_msg=c["ref"])


def test_prod_kats(oracle, kat):
    for c in _cases(kat, "prod"):
        for dt in FLOATS + INTS + ["u32"]:
            a = np.array(c["a"], dtype=_np(oracle, dt)).reshape(c["a_shape"])
            b = np.array(c["b"], dtype=_np(oracle, dt)).reshape(c["b_shape"])
            got = oracle.prod(a, b)
            assert list(got.shape) == c["expect_shape"], c["ref"]
            np.testing.assert_array_equal(got.reshape(-1), np.array(c["expect"], dtype=a.dtype), err_msg=c["ref"])
            # the reference's own stated equivalence (tensor.rs:310)
            np.testing.assert_array_equal(got, np.tensordot(a, b, axes=0))


def test_matmul_family_kats(oracle, kat):
    for dt in FLOATS + INTS + ["u32"]:
        T = _np(oracle, dt)
        for c in _cases(kat, "matmul"):
            A = np.array(c["a"], dtype=T).reshape(c["a_shape"])
            B = np.array(c["b"], dtype=T).reshape(c["b_shape"])
            for faithful in (False, True):
                got = oracle.matmul(A, B, faithful=faithful)
                assert list(got.shape) == c["expect_shape"]
                np.testing.assert_array_equal(got.reshape(-1), np.array(c["expect"], dtype=T), err_msg=c["ref"])
        for c in _cases(kat, "matvec"):
            A = np.array(c["a"], dtype=T).reshape(c["a_shape"])
            got = oracle.matvec(A, np.array(c["v"], dtype=T))
            np.testing.assert_array_equal(got, np.array(c["expect"], dtype=T), err_msg=c["ref"])
        for c in _cases(kat, "vecmat"):
            A = np.array(c["a"], dtype=T).reshape(c["a_shape"])
            got = oracle.vecmat(np.array(c["v"], dtype=T), A)
            np.testing.assert_array_equal(got, np.array(c["expect"], dtype=T), err_msg=c["ref"])
        for c in _cases(kat, "dot"):
            got = oracle.dot(np.array(c["a"], dtype=T), np.array(c["b"], dtype=T))
            assert list(got.shape) == c["expect_shape"]
            np.testing.assert_array_equal(got, np.array(c["expect"], dtype=T), err_msg=c["ref"])


def test_bookkeeping_kats(oracle, kat):
    for c in kat:
        op = c["op"]
        if op == "shape_new":
            assert oracle.shape_check(c["dims"])
        elif op == "shape_new_err":
            assert not oracle.shape_check(c["dims"])
        elif op in ("shape_size",):
            assert oracle.shape_size(c["dims"]) == c["size"]
        elif op == "size":
            assert oracle.shape_size(c["shape"]) == c["size"]
        elif op == "index_iterator":
            assert [list(t) for t in oracle.index_iterator(c["dims"])] == c["expect"], c["ref"]
        elif op == "get":
            flat = c["data"]
            for coord, want in zip(c["coords"], c["expect"]):
                rc, idx = oracle.flatten(coord, c["shape"])
                assert rc == 0 and flat[idx] == want
        elif op in ("get_err", "set_err"):
            for coord in c["coords"]:
                rc, _ = oracle.flatten(coord, c["shape"])
                assert rc >= 2, c["ref"]
        elif op == "new_err":
            assert len(c["data"]) != oracle.shape_size(c["shape"])
        elif op == "from_tensor_err":
            assert len(c["shape"]) != (2 if c["kind"] == "matrix" else 1)


def test_flatten_error_codes(oracle):
    # storage.rs:19-22 order mismatch, :24-30 out of bounds names the first offending dimension
    assert oracle.flatten([0, 0, 0], [2, 2])[0] == 1
    assert oracle.flatten([0, 2], [2, 2])[0] == 3
    assert oracle.flatten([2, 2], [2, 2])[0] == 2
    assert oracle.flatten([1, 2, 3], [2, 3, 4]) == (0, 1 * 12 + 2 * 4 + 3)


@pytest.mark.parametrize("dt", FLOATS)
def test_oracle_matches_numpy_on_exact_data(oracle, dt):
    """Independent cross-check: on small-integer data every float op is exact, so the
    oracle, whatever its order, must equal numpy bit for bit."""
    rng = np.random.default_rng(7)
    T = _np(oracle, dt)
    x = rng.integers(-8, 9, size=(3, 4, 5)).astype(T)
    for axes in ([0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2], []):
        npaxes = tuple(axes) if axes else None
        np.testing.assert_array_equal(oracle.reduce("sum", x, axes).reshape(-1), np.atleast_1d(x.sum(axis=npaxes)).reshape(-1))
        np.testing.assert_array_equal(oracle.reduce("max", x, axes).reshape(-1), np.atleast_1d(x.max(axis=npaxes)).reshape(-1))
        np.testing.assert_array_equal(oracle.reduce("min", x, axes).reshape(-1), np.atleast_1d(x.min(axis=npaxes)).reshape(-1))
    y = rng.integers(1, 5, size=(3, 4, 5)).astype(T)
    np.testing.assert_array_equal(oracle.ewise("add", x, y), x + y)
    np.testing.assert_array_equal(oracle.ewise("sub", x, y), x - y)
    np.testing.assert_array_equal(oracle.ewise("mul", x, y), x * y)
    np.testing.assert_array_equal(oracle.ewise("div", x, y), x / y)
    A = rng.integers(-4, 5, size=(5, 7)).astype(T)
    B = rng.integers(-4, 5, size=(7, 3)).astype(T)
    np.testing.assert_array_equal(oracle.matmul(A, B), A @ B)
    np.testing.assert_array_equal(oracle.matvec(A, B[:, 0].copy()), A @ B[:, 0])
    np.testing.assert_array_equal(oracle.vecmat(A[:, 0].copy(), A), A[:, 0] @ A)


def test_pow_kats(oracle, kat):
    """tensor.rs:1335-1365, matrix.rs:645, vector.rs:596 -- exact on these inputs with a correctly rounded libm."""
    n = 0
    for c in _cases(kat, "pow"):
        for dt in FLOATS:
            a = np.array(c["a"], dtype=_np(oracle, dt))
            np.testing.assert_array_equal(oracle.pow(a, c["power"]), np.array(c["expect"], dtype=a.dtype), err_msg=c["ref"])
            n += 1
    assert n == 10
