"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU) and exports every symbol
include/feotensor_b200.h declares; host-only bookkeeping entry points agree with the oracle and the
reference's golden vectors; the product fails loudly without a device and never touches the oracle.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "feotensor_b200.h")).read()
    return sorted(set(re.findall(r"FEO_API[^;(]*?\b(feo_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    import feotensor_b200
    assert os.path.exists(feotensor_b200.LIB_PATH), "run __graft_entry__.build()"
    assert feotensor_b200.LIB_PATH.startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    import feotensor_b200._lib as L
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(L.SIGNATURES) == declared
    assert lib.feo_abi_version() == 1
    assert [lib.feo_dtype_size(i) for i in range(5)] == [4, 8, 4, 8, 4]


def test_rust_ffi_declares_exactly_the_header():
    """rust/ cannot be compiled in this image (no toolchain), so its `extern "C"` block is GENERATED from the header
    (tools/gen_rust_ffi.py) and this test fails the moment the two drift apart."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    assert sorted(re.findall(r"pub fn (feo_[a-z0-9_]+)\(", ffi)) == _declared_symbols()
    # every ffi:: function the crate calls exists, and the status codes agree with the header
    used = set()
    for fn in os.listdir(os.path.join(ROOT, "rust", "src")):
        used |= set(re.findall(r"ffi::(feo_[a-z0-9_]+)\(", open(os.path.join(ROOT, "rust", "src", fn)).read()))
    assert used and used <= set(_declared_symbols()), used - set(_declared_symbols())
    header = open(os.path.join(ROOT, "include", "feotensor_b200.h")).read()
    for name, value in re.findall(r"pub const (FEO_[A-Z0-9_]+): c_int = (\d+);", ffi):
        assert re.search(rf"\b{name} = {value}\b", header), f"{name} = {value} is not what the header says"


def test_library_is_sm100a_and_uses_no_cpu_fallback():
    import feotensor_b200
    out = subprocess.run(["cuobjdump", "--list-elf", feotensor_b200.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
        assert not re.search(r"sm_(?!100a)\d+", out.stdout)
    # the product never links or loads the oracle
    ldd = subprocess.run(["ldd", feotensor_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    for fn in os.listdir(os.path.join(ROOT, "feotensor_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "feotensor_b200", fn)).read().replace("no oracle", "").replace(
                "imports the oracle", "").replace("import the oracle", ""), fn


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_no_device_means_loud_failure():
    import feotensor_b200 as f
    with pytest.raises(f.BackendError):
        f.Context(0)
    with pytest.raises(f.BackendError):
        f.Tensor(f.shape(2), [1.0, 2.0])


def test_host_bookkeeping_matches_reference_kats(kat, oracle):
    import feotensor_b200 as f
    for c in kat:
        op = c["op"]
        if op == "shape_new":
            assert f.Shape(c["dims"]).dims == c["dims"]
        elif op == "shape_new_err":
            with pytest.raises(f.ShapeError):
                f.Shape(c["dims"])
        elif op == "shape_size":
            assert f.Shape(c["dims"]).size() == c["size"]
        elif op == "shape_display":
            assert str(f.Shape(c["dims"])) == c["text"]
        elif op == "shape_stack":
            assert f.Shape(c["a"]).stack(f.Shape(c["b"])).dims == c["expect"]
        elif op == "coord_order":
            assert f.Coordinate(c["indices"]).order() == c["order"]
        elif op == "coord_insert":
            assert f.Coordinate(c["indices"]).insert(c["index"], c["value"]) == f.Coordinate(c["expect"])
        elif op == "coord_display":
            assert str(f.Coordinate(c["indices"])) == c["text"]
        elif op == "index_iterator":
            assert [list(x) for x in f.IndexIterator(f.Shape(c["dims"]))] == c["expect"]
    # coordinate.rs:12-15, coord! repeat form
    with pytest.raises(f.ShapeError):
        f.Coordinate([])
    assert f.coord(1, count=3) == f.Coordinate([1, 1, 1])
    assert str(f.ShapeError("x")) == "ShapeError: x"


def test_flatten_reduce_shape_divisor_match_oracle(oracle):
    import feotensor_b200._lib as L
    lib = L.load()
    rng = np.random.default_rng(11)
    for _ in range(200):
        rank = int(rng.integers(1, 6))
        shape = [int(x) for x in rng.integers(1, 6, rank)]
        coord = [int(rng.integers(0, s + 1)) for s in shape]  # sometimes out of bounds
        idx, detail = C.c_size_t(0), C.c_int(0)
        rc = lib.feo_flatten(rank, L.sizes(coord), rank, L.sizes(shape), C.byref(idx), C.byref(detail))
        orc, oidx = oracle.flatten(coord, shape)
        assert (rc == 0) == (orc == 0)
        if rc == 0:
            assert idx.value == oidx
        else:
            assert detail.value == orc - 2
        naxes = int(rng.integers(0, rank + 1))
        axes = sorted(rng.choice(rank, naxes, replace=False).tolist())
        out_rank, out_shape = C.c_int(0), (C.c_size_t * 8)()
        assert lib.feo_reduce_shape(rank, L.sizes(shape), naxes, L.sizes(axes), C.byref(out_rank), out_shape) == 0
        assert tuple(out_shape[i] for i in range(out_rank.value)) == oracle.reduce_shape(shape, axes)
    # order mismatch
    idx, detail = C.c_size_t(0), C.c_int(0)
    assert lib.feo_flatten(3, L.sizes([0, 0, 0]), 2, L.sizes([2, 2]), C.byref(idx), C.byref(detail)) == L.ERR_SHAPE
    assert detail.value == -1
    # bad axes are refused (the reference panics)
    out_rank, out_shape = C.c_int(0), (C.c_size_t * 8)()
    for bad in ([2, 0], [1, 1], [3]):
        assert lib.feo_reduce_shape(3, L.sizes([2, 3, 4]), len(bad), L.sizes(bad), C.byref(out_rank), out_shape) == L.ERR_SHAPE
    # the counting divisor in closed form == the oracle's literal counting loop
    cases = [("f32", (16384, 16384, 8), [0, 1, 2]), ("f32", (2 ** 25,), []), ("f32", (2 ** 24 + 7, 3), [0]),
             ("f64", (2 ** 20, 5), [0, 1]), ("i32", (70000, 70000), [0, 1]), ("u32", (2 ** 20, 2 ** 13), [0, 1]),
             ("i64", (2 ** 20, 2 ** 13), []), ("f32", (64, 128, 256), [0, 2])]
    for dt, shape, axes in cases:
        out = np.zeros(1, oracle.NP_DTYPES[dt])
        assert lib.feo_reduce_divisor(L.code_of(dt), len(shape), L.sizes(shape), len(axes), L.sizes(axes),
                                      out.ctypes.data_as(C.c_void_p)) == 0
        assert out[0] == oracle.reduce_divisor(dt, shape, axes), (dt, shape, axes)


def test_flat_index_magic_division_is_exact_below_2_31():
    """The flat-index broadcast kernel (csrc/ewise_impl.cuh: FastDiv) divides 32-bit indices by the collapsed extents with
    q = (umulhi(n, mul) + n) >> shift.  The host-side constants, restated here, must give floor(n / d) for every n < 2^31
    (the launcher takes the path only below that), for small, large, power-of-two and awkward divisors."""
    rng = np.random.default_rng(0)

    def magic(d):
        sh = 0
        while sh < 32 and (1 << sh) < d:
            sh += 1
        return ((((1 << 32) * ((1 << sh) - d)) // d) + 1) & 0xFFFFFFFF, sh

    divisors = list(range(1, 300)) + [int(x) for x in rng.integers(1, 2 ** 31 - 1, 500)] + [2 ** 31 - 1, 2 ** 30, 2 ** 30 + 1, 4099, 89478485]
    for d in divisors:
        mul, sh = magic(d)
        n = np.concatenate([rng.integers(0, 2 ** 31 - 8, 500, dtype=np.uint64),
                            np.array([0, 1, d - 1, d, min(d + 1, 2 ** 31 - 8), 2 ** 31 - 9, 2 ** 31 - 8], dtype=np.uint64)])
        t = (n * np.uint64(mul)) >> np.uint64(32)
        q = ((t + n) & np.uint64(0xFFFFFFFF)) >> np.uint64(sh)
        np.testing.assert_array_equal(q, n // np.uint64(d), err_msg=f"d={d}")


def test_tuning_knobs_agree_between_the_library_and_the_python_mirror():
    """The knob numbers are part of the (experimental) ABI: csrc/feo_common.cuh's enum and _lib.py must list the same ones."""
    import feotensor_b200._lib as L
    text = open(os.path.join(ROOT, "feotensor_b200", "csrc", "feo_common.cuh")).read()
    body = text[text.index("enum feo_knob {"):]
    body = body[:body.index("};")]
    enum = {name: int(val) for name, val in re.findall(r"FEO_(KNOB_[A-Z0-9_]+)\s*=\s*(\d+)", body)}
    count = enum.pop("KNOB_COUNT")
    assert sorted(enum.values()) == list(range(count)), "knob numbers are dense"
    mirror = {k: v for k, v in vars(L).items() if k.startswith("KNOB_")}
    assert mirror == enum
