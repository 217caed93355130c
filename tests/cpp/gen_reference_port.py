#!/usr/bin/env python
"""Generates a C++ port of the reference's known-answer tests (tests/golden/reference_kat.json) against the
C++ host API (include/feotensor.hpp).  The generated program runs every case for double and float (and int32
where all literals are integers), prints a summary and exits non-zero on any failure.

    python tests/cpp/gen_reference_port.py build/cpp/test_reference_port.cpp
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def lit(v):
    return repr(float(v)) if not float(v).is_integer() else str(int(v))


def vec(vals):
    return "std::vector<T>{" + ", ".join(f"T({lit(v)})" for v in vals) + "}"


def dims(d):
    return "feo::Shape(std::vector<size_t>{" + ", ".join(str(int(x)) for x in d) + "})"


def axes(a):
    return "feo::Axes{" + ", ".join(str(int(x)) for x in a) + "}"


def make(kind, name, shape, data):
    if kind == "vector":
        return f"feo::Vector<T> {name}({vec(data)});"
    cls = "Matrix" if kind == "matrix" else "Tensor"
    return f"feo::{cls}<T> {name}({dims(shape)}, {vec(data)});"


def ints_ok(*seqs):
    return all(float(x).is_integer() for s in seqs for x in (s if isinstance(s, list) else [s]))


def main(out_path):
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))["cases"]
    body_all, body_int = [], []
    ops = {"add": "+", "sub": "-", "mul": "*", "div": "/"}
    for c in cases:
        ref, op, blk, int_ok = c["ref"], c["op"], [], False
        if op == "reduce":
            blk.append(make(c["via"], "x", c["shape"], c["data"]))
            call = f"x.{c['kind']}()" if c["via"] == "vector" else f"x.{c['kind']}({axes(c['axes'])})"
            blk.append(f"auto r = {call};")
            blk.append(f"check_shape(r.shape(), {dims(c['expect_shape'])}, \"{ref}\", fails);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = ints_ok(c["data"], c["expect"])
        elif op in ("ewise", "ewise_scalar"):
            blk.append(make(c["lhs"], "a", c["shape"], c["a"]))
            if op == "ewise":
                blk.append(make(c["rhs"], "b", c["shape"], c["b"]))
                blk.append(f"auto r = std::move(a) {ops[c['opname']]} std::move(b);")
                int_ok = ints_ok(c["a"], c["b"], c["expect"])
            else:
                blk.append(f"auto r = std::move(a) {ops[c['opname']]} T({lit(c['scalar'])});")
                int_ok = ints_ok(c["a"], c["scalar"], c["expect"])
            blk.append(f"check_shape(r.shape(), {dims(c['shape'])}, \"{ref}\", fails);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
        elif op == "prod":
            blk.append(make("tensor", "a", c["a_shape"], c["a"]))
            blk.append(make("tensor", "b", c["b_shape"], c["b"]))
            blk.append("auto r = a.prod(b);")
            blk.append(f"check_shape(r.shape(), {dims(c['expect_shape'])}, \"{ref}\", fails);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = True
        elif op == "matmul":
            blk.append(make("matrix", "a", c["a_shape"], c["a"]))
            blk.append(make("matrix", "b", c["b_shape"], c["b"]))
            blk.append("auto r = a.matmul(b);")
            blk.append(f"check_shape(r.shape(), {dims(c['expect_shape'])}, \"{ref}\", fails);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = True
        elif op == "matvec":
            blk.append(make("matrix", "a", c["a_shape"], c["a"]))
            blk.append(make("vector", "v", None, c["v"]))
            blk.append("auto r = a.vecmul(v);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = True
        elif op == "vecmat":
            blk.append(make("matrix", "a", c["a_shape"], c["a"]))
            blk.append(make("vector", "v", None, c["v"]))
            blk.append("auto r = v.matmul(a);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = True
        elif op == "dot":
            blk.append(make("vector", "a", None, c["a"]))
            blk.append(make("vector", "b", None, c["b"]))
            blk.append("auto r = a.vecmul(b);")
            blk.append(f"check_shape(r.shape(), {dims(c['expect_shape'])}, \"{ref}\", fails);")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
            int_ok = True
        elif op == "pow":
            blk.append(make(c.get("via", "tensor"), "x", c["shape"], c["a"]))
            blk.append(f"auto r = x.pow(T({lit(c['power'])}));")
            blk.append(f"check_vec(r.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
        elif op == "new_err":
            blk.append(f"expect_shape_error([&] {{ feo::Tensor<T> t({dims(c['shape'])}, {vec(c['data'])}); }}, \"{ref}\", fails);")
        elif op == "fill":
            cls = {"tensor": "Tensor", "matrix": "Matrix", "vector": "Vector"}[c.get("via", "tensor")]
            blk.append(f"auto t = feo::{cls}<T>::fill({dims(c['shape'])}, T({lit(c['value'])}));")
            n = 1
            for d in c["shape"]:
                n *= d
            blk.append(f"check_vec(t.to_vec(), std::vector<T>({n}, T({lit(c['value'])})), \"{ref}\", fails);")
        elif op == "eye":
            blk.append(f"auto m = feo::Matrix<T>::eye({dims(c['shape'])});")
            blk.append(f"check_vec(m.to_vec(), {vec(c['expect'])}, \"{ref}\", fails);")
        elif op == "get":
            blk.append(make("tensor", "t", c["shape"], c["data"]))
            for co, w in zip(c["coords"], c["expect"]):
                blk.append(f"check(t.get(feo::Coordinate(std::vector<size_t>{{{', '.join(map(str, co))}}})) == T({lit(w)}), \"{ref}\", fails);")
        elif op in ("get_err", "set_err"):
            blk.append(make("tensor", "t", c["shape"], c["data"]))
            for co in c["coords"]:
                acc = "t.get(c)" if op == "get_err" else "t.set(c, T(1))"
                blk.append(f"expect_shape_error([&] {{ feo::Coordinate c(std::vector<size_t>{{{', '.join(map(str, co))}}}); {acc}; }}, \"{ref}\", fails);")
        elif op == "set":
            blk.append(make("tensor", "t", c["shape"], c["data"]))
            for co, v in zip(c["coords"], c["values"]):
                blk.append(f"t.set(feo::Coordinate(std::vector<size_t>{{{', '.join(map(str, co))}}}), T({lit(v)}));")
            for co, v in zip(c["coords"], c["values"]):
                blk.append(f"check(t.get(feo::Coordinate(std::vector<size_t>{{{', '.join(map(str, co))}}})) == T({lit(v)}), \"{ref}\", fails);")
        elif op == "from_tensor_err":
            cls = "Matrix" if c["kind"] == "matrix" else "Vector"
            blk.append(f"expect_shape_error([&] {{ feo::{cls}<T>::from_tensor(feo::Tensor<T>({dims(c['shape'])}, {vec(c['data'])})); }}, \"{ref}\", fails);")
        else:
            continue
        code = "    {  // " + ref + "\n        " + "\n        ".join(blk) + "\n        ++ran;\n    }\n"
        body_all.append(code)
        if int_ok and not any(float(x) < 0 for k in ("data", "a", "b", "expect", "v") for x in (c.get(k) or [])):
            body_int.append(code)
    src = r'''// GENERATED by tests/cpp/gen_reference_port.py from tests/golden/reference_kat.json -- do not edit.
#include <cmath>
#include <cstdio>
#include <functional>
#include "feotensor.hpp"

static void check(bool ok, const char *what, int &fails) { if (!ok) { std::printf("FAIL %s\n", what); ++fails; } }
static void check_shape(const feo::Shape &got, const feo::Shape &want, const char *what, int &fails) {
    if (got != want) { std::printf("FAIL %s: shape %s != %s\n", what, got.to_string().c_str(), want.to_string().c_str()); ++fails; }
}
template <class T> static void check_vec(const std::vector<T> &got, const std::vector<T> &want, const char *what, int &fails) {
    bool ok = got.size() == want.size();
    for (size_t i = 0; ok && i < got.size(); ++i) ok = (got[i] == want[i]);  // the reference asserts exact equality
    if (!ok) { std::printf("FAIL %s: values differ\n", what); ++fails; }
}
static void expect_shape_error(const std::function<void()> &f, const char *what, int &fails) {
    try { f(); } catch (const feo::ShapeError &) { return; } catch (...) {}
    std::printf("FAIL %s: expected ShapeError\n", what); ++fails;
}
static void expect_panic(const std::function<void()> &f, const char *what, int &fails) {
    try { f(); } catch (const feo::Panic &) { return; } catch (...) {}
    std::printf("FAIL %s: expected Panic\n", what); ++fails;
}

template <class T> static int run_float_cases(int &fails) {
    int ran = 0;
''' + "".join(body_all) + r'''    return ran;
}
template <class T> static int run_int_cases(int &fails) {
    int ran = 0;
''' + "".join(body_int) + r'''    return ran;
}

int main() {
    int fails = 0, ran = 0;
    ran += run_float_cases<double>(fails);
    ran += run_float_cases<float>(fails);
    ran += run_int_cases<int32_t>(fails);
    ran += run_int_cases<int64_t>(fails);
    ran += run_int_cases<uint32_t>(fails);
    // bookkeeping (shape.rs, coordinate.rs, iter.rs tests) and panic preconditions
    check(feo::Shape({2, 3, 4}).size() == 24, "shape.rs:87", fails);
    check(feo::Shape({2, 3, 4}).to_string() == "(2, 3, 4)", "shape.rs:93", fails);
    check(feo::Shape({2, 3}).stack(feo::Shape({4, 5})) == feo::Shape({2, 3, 4, 5}), "shape.rs:108", fails);
    expect_shape_error([] { feo::Shape s({2, 0, 4}); }, "shape.rs:80", fails);
    check(feo::Coordinate({1, 2, 3}).insert(1, 4) == feo::Coordinate({1, 4, 2, 3}), "coordinate.rs:96", fails);
    check(feo::Coordinate({1, 2, 3}).to_string() == "(1, 2, 3)", "coordinate.rs:118", fails);
    check(feo::Coordinate::repeat(1, 3) == feo::Coordinate({1, 1, 1}), "coordinate.rs:124", fails);
    expect_shape_error([] { feo::Coordinate c(std::vector<size_t>{}); }, "coordinate.rs:12", fails);
    {
        feo::IndexIterator it(feo::Shape({2, 3}));
        feo::Coordinate c({0});
        std::vector<feo::Coordinate> seen;
        while (it.next(c)) seen.push_back(c);
        check(seen.size() == 6 && seen[0] == feo::Coordinate({0, 0}) && seen[1] == feo::Coordinate({0, 1}) && seen[3] == feo::Coordinate({1, 0}) &&
                  seen[5] == feo::Coordinate({1, 2}), "iter.rs:56", fails);
        feo::IndexIterator empty(feo::Shape(std::vector<size_t>{}));
        check(!empty.next(c), "iter.rs:82", fails);
    }
    expect_panic([] { auto r = feo::Tensor<float>(feo::Shape({2, 3}), std::vector<float>(6, 1.f)) + feo::Tensor<float>(feo::Shape({3, 2}), std::vector<float>(6, 1.f)); },
                 "tensor.rs:366 shape assert", fails);
    expect_panic([] { feo::Tensor<float>(feo::Shape({2, 3}), std::vector<float>(6, 1.f)).sum({1, 0}); }, "unsorted axes", fails);
    expect_panic([] { feo::Matrix<float> a(feo::Shape({2, 3}), std::vector<float>(6, 1.f)); a.matmul(a); }, "matrix.rs:78", fails);
    expect_panic([] { feo::Vector<float>(std::vector<float>{1, 2}).vecmul(feo::Vector<float>(std::vector<float>{1, 2, 3})); }, "vector.rs:72", fails);
    {   // broadcasting (new surface): a[2,1,4] * b[1,3,4]
        feo::Tensor<float> a(feo::Shape({2, 1, 4}), std::vector<float>{1, 2, 3, 4, 5, 6, 7, 8});
        feo::Tensor<float> b(feo::Shape({1, 3, 4}), std::vector<float>{1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3});
        auto r = a.broadcast(FEO_MUL, b);
        check_shape(r.shape(), feo::Shape({2, 3, 4}), "broadcast shape", fails);
        check_vec(r.to_vec(), std::vector<float>{1, 2, 3, 4, 2, 4, 6, 8, 3, 6, 9, 12, 5, 6, 7, 8, 10, 12, 14, 16, 15, 18, 21, 24}, "broadcast values", fails);
    }
    std::printf("reference port: %d cases run, %d failures\n", ran, fails);
    return fails ? 1 : 0;
}
'''
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    open(out_path, "w").write(src)
    print(f"wrote {out_path}: {len(body_all)} float cases x2, {len(body_int)} integer cases x3")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "build", "cpp", "test_reference_port.cpp"))
