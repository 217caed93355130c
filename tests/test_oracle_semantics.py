"""Oracle semantics beyond the reference's own tests: the behaviour ledger of SURVEY.md 8(a)-quirks.

These are the cases the reference leaves unpinned (integer dtypes, wrap, counting divisor, sequential
order); the restated oracle is their definition of truth, so each is checked here against an
independent, deliberately naive Python/numpy re-derivation of the cited reference lines.
"""
import numpy as np
import pytest


def _seq_sum_f32(vals):
    acc = np.float32(0.0)
    for v in vals:
        acc = np.float32(acc + np.float32(v))
    return acc


def test_sum_is_sequential_in_storage_order(oracle):
    # tensor.rs:85 -- fold(zero, acc + x): not numpy's pairwise order.
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 4097).astype(np.float32)
    assert oracle.reduce("sum", x, [])[0] == _seq_sum_f32(x)
    assert oracle.reduce("sum", x, [0])[0] == _seq_sum_f32(x)


def test_axis_sum_order_follows_removed_dims_row_major(oracle):
    # tensor.rs:94-105 -- for each target, removed coordinates in row-major order of remove_shape.
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (5, 3, 7)).astype(np.float32)
    got = oracle.reduce("sum", x, [0, 2])
    for k in range(3):
        want = _seq_sum_f32(x[i, k, j] for i in range(5) for j in range(7))
        assert got[k] == want


def test_mean_divisor_counts_in_type(oracle):
    # tensor.rs:112-130 -- n is built by adding T::one(); f32 stops at 2^24.
    assert oracle.count_in_type("f32", 1000) == np.float32(1000)
    assert oracle.count_in_type("f32", 2 ** 24) == np.float32(2 ** 24)
    assert oracle.count_in_type("f32", 2 ** 24 + 5) == np.float32(2 ** 24)
    assert oracle.count_in_type("f32", 2 ** 31) == np.float32(2 ** 24)
    assert oracle.count_in_type("f64", 2 ** 31) == np.float64(2 ** 31)
    assert oracle.count_in_type("i32", 12345) == 12345
    # product of per-axis counts, taken in T (tensor.rs:123)
    assert oracle.reduce_divisor("f32", (16384, 16384, 8), [0, 1, 2]) == np.float32(2 ** 31)
    # axes == [] counts the whole size (tensor.rs:125-129): saturates for f32
    assert oracle.reduce_divisor("f32", (2 ** 25,), []) == np.float32(2 ** 24)


def test_var_is_population_two_pass(oracle):
    # tensor.rs:171-177 / 196-202
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (6, 5)).astype(np.float64)
    got = oracle.reduce("var", x, [0])
    for j in range(5):
        n = 0.0
        for _ in range(6):
            n += 1.0
        s = 0.0
        for i in range(6):
            s = s + x[i, j]
        mean = s / n
        acc = 0.0
        for i in range(6):
            c = x[i, j] - mean
            acc = acc + c * c
        assert got[j] == acc / n
    np.testing.assert_allclose(got, x.var(axis=0), rtol=1e-13)


def test_integer_var_and_mean_truncate(oracle):
    x = np.array([[1, 2, 4], [2, 3, 9]], dtype=np.int32)
    np.testing.assert_array_equal(oracle.reduce("mean", x, [0]), np.array([1, 2, 6], np.int32))  # 3/2, 5/2, 13/2
    # var: mean truncated first, then sum of squares / n, all in integers
    np.testing.assert_array_equal(oracle.reduce("var", x, [0]), np.array([(0 + 1) // 2, (0 + 1) // 2, (4 + 9) // 2], np.int32))
    neg = np.array([-7, 2], dtype=np.int64)
    assert oracle.reduce("mean", neg, [])[0] == -2  # -5/2 truncates toward zero like Rust


def test_integers_wrap_like_release_rust(oracle):
    a = np.array([2 ** 31 - 1, -2 ** 31], dtype=np.int32)
    b = np.array([1, -1], dtype=np.int32)
    np.testing.assert_array_equal(oracle.ewise("add", a, b), np.array([-2 ** 31, 2 ** 31 - 1], np.int32))
    np.testing.assert_array_equal(oracle.ewise("mul", a, np.array([2, 2], np.int32)), np.array([-2, 0], np.int32))
    u = np.array([0, 5], dtype=np.uint32)
    np.testing.assert_array_equal(oracle.ewise_scalar("sub", u, 1), np.array([2 ** 32 - 1, 4], np.uint32))
    big = np.full(5, 2 ** 62, dtype=np.int64)
    assert oracle.reduce("sum", big, [])[0] == np.int64((5 * 2 ** 62) % 2 ** 64)  # = 2**62


def test_integer_division_panics_are_errors(oracle):
    a = np.array([1, 2], dtype=np.int32)
    with pytest.raises(oracle.OracleError) as e:
        oracle.ewise("div", a, np.array([1, 0], np.int32))
    assert e.value.code == oracle.ERR_ARITH
    with pytest.raises(oracle.OracleError):
        oracle.ewise("div", np.array([-2 ** 31], np.int32), np.array([-1], np.int32))
    # floats follow IEEE instead
    f = oracle.ewise("div", np.array([1.0, -1.0, 0.0], np.float32), np.zeros(3, np.float32))
    assert np.isposinf(f[0]) and np.isneginf(f[1]) and np.isnan(f[2])


def test_max_min_init_quirk_gives_true_extrema_for_finite_data(oracle):
    # tensor.rs:234-238 initialises with min(0, all x); the result is still the true maximum.
    x = np.array([[-5.0, -2.0], [-3.0, -9.0]], dtype=np.float32)
    np.testing.assert_array_equal(oracle.reduce("max", x, [0]), np.array([-3.0, -2.0], np.float32))
    np.testing.assert_array_equal(oracle.reduce("min", -x, [1]), np.array([2.0, 3.0], np.float32))
    xi = np.array([[5, 2], [3, 9]], dtype=np.uint32)
    np.testing.assert_array_equal(oracle.reduce("min", xi, [0]), np.array([3, 2], np.uint32))


def test_axes_must_be_strictly_ascending_in_range(oracle):
    x = np.zeros((2, 3, 4), np.float32)
    for bad in ([2, 0], [1, 0], [0, 0], [3], [0, 5]):
        with pytest.raises(oracle.OracleError) as e:
            oracle.reduce("sum", x, bad)
        assert e.value.code == oracle.ERR_SHAPE


def test_full_reduction_has_shape_one(oracle):
    x = np.ones((2, 3), np.float32)
    assert oracle.reduce("sum", x, [0, 1]).shape == (1,)
    assert oracle.reduce("sum", x, []).shape == (1,)
    assert oracle.reduce_shape((4, 5, 6), [1]) == (4, 6)


@pytest.mark.parametrize("dt", ["f32", "f64", "i32", "i64", "u32"])
def test_faithful_variant_is_bit_identical(oracle, dt):
    rng = np.random.default_rng(3)
    T = oracle.NP_DTYPES[dt]
    if dt.startswith("f"):
        x = rng.uniform(-1, 1, (4, 3, 5, 2)).astype(T)
    else:
        x = rng.integers(0 if dt == "u32" else -1000, 1000, (4, 3, 5, 2)).astype(T)
    for axes in ([0], [3], [1, 2], [0, 3], [0, 1, 2, 3], [0, 2, 3], []):
        for kind in ("sum", "mean", "var", "max", "min"):
            a = oracle.reduce(kind, x, axes)
            b = oracle.reduce(kind, x, axes, faithful=True, recompute_mean_per_target=True)
            np.testing.assert_array_equal(a, b)
    A = x.reshape(12, 10)
    B = x.reshape(10, 12)
    np.testing.assert_array_equal(oracle.matmul(A, B), oracle.matmul(A, B, faithful=True))
    rows = np.array([0, 5, 11]); cols = np.array([3, 0, 11])
    np.testing.assert_array_equal(oracle.matmul_entries(A, B, rows, cols), oracle.matmul(A, B)[rows, cols])


def test_broadcast_is_materialise_then_operator(oracle):
    rng = np.random.default_rng(4)
    a = rng.uniform(-1, 1, (4, 1, 6)).astype(np.float32)
    b = rng.uniform(1, 2, (1, 5, 6)).astype(np.float32)
    for op in ("add", "sub", "mul", "div"):
        got = oracle.ewise_broadcast(op, a, b)
        am, bm = oracle.materialize(a, (4, 5, 6)), oracle.materialize(b, (4, 5, 6))
        np.testing.assert_array_equal(am, np.broadcast_to(a, (4, 5, 6)))
        np.testing.assert_array_equal(got, oracle.ewise(op, am, bm))
    with pytest.raises(oracle.OracleError):
        oracle.ewise_broadcast("mul", np.zeros((2, 3), np.float32), np.zeros((4, 3), np.float32))


def test_dot_matvec_are_sequential(oracle):
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, 1000).astype(np.float32)
    b = rng.uniform(-1, 1, 1000).astype(np.float32)
    acc = np.float32(0)
    for i in range(1000):
        acc = np.float32(acc + np.float32(a[i] * b[i]))
    assert oracle.dot(a, b)[0] == acc
    A = rng.uniform(-1, 1, (3, 1000)).astype(np.float32)
    got = oracle.matvec(A, b)
    for r in range(3):
        assert got[r] == oracle.dot(A[r].copy(), b)[0]
    np.testing.assert_array_equal(oracle.dot_batched(A, np.stack([b] * 3)), got)
