#!/usr/bin/env python
"""Writes tests/golden/reference_kat.json: the reference's own known-answer tests for the hot path.

The reference (Rust crate feotensor) cannot be compiled in this image (no cargo/rustc), so its
golden vectors are TRANSCRIBED from its inline #[test] functions; every case carries the
`file:line test_name` it was taken from.  When /root/reference is present (it is in the build
container, not on the GPU box) `--verify` re-reads each cited test body and checks that it starts at
the cited line and that every number of the transcribed case occurs in it.

    python tests/golden/make_reference_kat.py            # rewrite reference_kat.json
    python tests/golden/make_reference_kat.py --verify   # also cross-check against /root/reference/src

All reference literals are untyped floats, i.e. f64; they are small integers / dyadic fractions and
therefore exactly representable in f32 and (except the fractional ones) in the integer dtypes too.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"

T12 = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0, 12.0]
ALT12 = [1.0, -2.0, 3.0, -4.0, 5.0, -6.0, 7.0, -8.0, 9.0, -10.0, 11.0, -12.0]

CASES = []


def red(ref, kind, shape, data, axes, eshape, expect, via="tensor"):
    CASES.append(dict(ref=ref, op="reduce", kind=kind, shape=shape, data=data, axes=axes,
                      expect_shape=eshape, expect=expect, via=via))


def ew(ref, opname, shape, a, b, expect, lhs="tensor", rhs="tensor"):
    CASES.append(dict(ref=ref, op="ewise", opname=opname, shape=shape, a=a, b=b, expect=expect, lhs=lhs, rhs=rhs))


def ews(ref, opname, shape, a, s, expect, lhs="tensor"):
    CASES.append(dict(ref=ref, op="ewise_scalar", opname=opname, shape=shape, a=a, scalar=s, expect=expect, lhs=lhs))


# ---- tensor.rs -----------------------------------------------------------------------------------
CASES.append(dict(ref="tensor.rs:559 test_new_tensor", op="new", shape=[2, 2], data=[1.0, 2.0, 3.0, 4.0]))
CASES.append(dict(ref="tensor.rs:570 test_new_tensor_shape_data_mismatch", op="new_err", shape=[2, 2],
                  data=[1.0, 2.0, 3.0]))
CASES.append(dict(ref="tensor.rs:580 test_zeros_tensor", op="fill", shape=[2, 3], value=0.0, dtype="f32", ctor="zeros"))
CASES.append(dict(ref="tensor.rs:589 test_ones_tensor", op="fill", shape=[2, 3], value=1.0, dtype="f32", ctor="ones"))
CASES.append(dict(ref="tensor.rs:598 test_fill_tensor", op="fill", shape=[2, 3], value=7.0, dtype="f32", ctor="fill"))
CASES.append(dict(ref="tensor.rs:615 test_tensor_size", op="size", shape=[2, 3], size=6))
CASES.append(dict(ref="tensor.rs:623 test_tensor_get", op="get", shape=[2, 2], data=[1.0, 2.0, 3.0, 4.0],
                  coords=[[0, 0], [0, 1], [1, 0], [1, 1]], expect=[1.0, 2.0, 3.0, 4.0]))
CASES.append(dict(ref="tensor.rs:635 test_tensor_set", op="set", shape=[2, 2], data=[1.0, 2.0, 3.0, 4.0],
                  coords=[[0, 0], [0, 1], [1, 0], [1, 1]], values=[5.0, 6.0, 7.0, 8.0]))
CASES.append(dict(ref="tensor.rs:652 test_tensor_get_out_of_bounds", op="get_err", shape=[2, 2],
                  data=[1.0, 2.0, 3.0, 4.0], coords=[[2, 0], [0, 2], [2, 2]]))
CASES.append(dict(ref="tensor.rs:663 test_tensor_set_out_of_bounds", op="set_err", shape=[2, 2],
                  data=[1.0, 2.0, 3.0, 4.0], coords=[[2, 0], [0, 2], [2, 2]], values=[5.0, 6.0, 7.0]))

red("tensor.rs:674 test_tensor_sum_no_axis_1d", "sum", [5], [1.0, 2.0, 3.0, 4.0, 5.0], [], [1], [15.0])
red("tensor.rs:686 test_tensor_sum_no_axis_2d", "sum", [2, 3], T12[:6], [], [1], [21.0])
red("tensor.rs:698 test_tensor_sum_no_axis_3d", "sum", [2, 2, 3], T12, [], [1], [78.0])
red("tensor.rs:712 test_tensor_sum_one_axis_1d", "sum", [5], [1.0, 2.0, 3.0, 4.0, 5.0], [0], [1], [15.0])
red("tensor.rs:724 test_tensor_sum_one_axis_2d", "sum", [2, 3], T12[:6], [0], [3], [5.0, 7.0, 9.0])
red("tensor.rs:736 test_tensor_sum_one_axis_3d", "sum", [2, 2, 3], T12, [0], [2, 3],
    [8.0, 10.0, 12.0, 14.0, 16.0, 18.0])
red("tensor.rs:753 test_tensor_sum_multiple_axes_2d", "sum", [2, 3], T12[:6], [0, 1], [1], [21.0])
red("tensor.rs:765 test_tensor_sum_multiple_axes_3d", "sum", [2, 2, 3], T12, [0, 1], [3], [22.0, 26.0, 30.0])

red("tensor.rs:779 test_tensor_mean_no_axis_2d", "mean", [2, 3], T12[:6], [], [1], [3.5])
red("tensor.rs:791 test_tensor_mean_one_axis_2d", "mean", [2, 3], T12[:6], [0], [3], [2.5, 3.5, 4.5])
red("tensor.rs:803 test_tensor_mean_one_axis_3d", "mean", [2, 2, 3], T12, [0], [2, 3],
    [4.0, 5.0, 6.0, 7.0, 8.0, 9.0])
red("tensor.rs:820 test_tensor_mean_multiple_axes_2d", "mean", [2, 3], T12[:6], [0, 1], [1], [3.5])
red("tensor.rs:832 test_tensor_mean_multiple_axes_3d", "mean", [2, 2, 3], T12, [0, 1], [3], [5.5, 6.5, 7.5])

red("tensor.rs:846 test_tensor_var_no_axis_2d", "var", [2, 3], [1.0, 1.0, 1.0, 7.0, 7.0, 7.0], [], [1], [9.0])
red("tensor.rs:858 test_tensor_var_one_axis_2d", "var", [2, 3], T12[:6], [0], [3], [2.25, 2.25, 2.25])
red("tensor.rs:870 test_tensor_var_one_axis_3d", "var", [2, 2, 3], T12, [0], [2, 3], [9.0] * 6)
red("tensor.rs:887 test_tensor_var_multiple_axes_2d", "var", [2, 3], [1.0, 1.0, 1.0, 7.0, 7.0, 7.0], [0, 1], [1], [9.0])
red("tensor.rs:899 test_tensor_var_multiple_axes_3d", "var", [2, 2, 3],
    [1.0, 3.0, 5.0, 7.0, 9.0, 11.0, 13.0, 15.0, 17.0, 19.0, 21.0, 23.0], [0, 1], [3], [45.0, 45.0, 45.0])

red("tensor.rs:913 test_tensor_max_no_axis_1d", "max", [5], [1.0, -2.0, 3.0, -4.0, 5.0], [], [1], [5.0])
red("tensor.rs:925 test_tensor_max_one_axis_2d", "max", [2, 3], ALT12[:6], [0], [3], [1.0, 5.0, 3.0])
red("tensor.rs:937 test_tensor_max_multiple_axes_3d", "max", [2, 2, 3], ALT12, [0, 1], [3], [7.0, 11.0, 9.0])
red("tensor.rs:951 test_tensor_min_no_axis_1d", "min", [5], [1.0, -2.0, 3.0, -4.0, 5.0], [], [1], [-4.0])
red("tensor.rs:963 test_tensor_min_one_axis_2d", "min", [2, 3], ALT12[:6], [0], [3], [-4.0, -2.0, -6.0])
red("tensor.rs:975 test_tensor_min_multiple_axes_3d", "min", [2, 2, 3], ALT12, [0, 1], [3], [-10.0, -8.0, -12.0])

CASES.append(dict(ref="tensor.rs:989 test_tensor_prod_1d_1d", op="prod", a_shape=[3], a=[1.0, 2.0, 3.0],
                  b_shape=[2], b=[4.0, 5.0], expect_shape=[3, 2], expect=[4.0, 5.0, 8.0, 10.0, 12.0, 15.0]))
CASES.append(dict(ref="tensor.rs:1008 test_tensor_prod_2d_1d", op="prod", a_shape=[2, 2], a=[1.0, 2.0, 3.0, 4.0],
                  b_shape=[2], b=[5.0, 6.0], expect_shape=[2, 2, 2],
                  expect=[5.0, 6.0, 10.0, 12.0, 15.0, 18.0, 20.0, 24.0]))
CASES.append(dict(ref="tensor.rs:1027 test_tensor_prod_2d_2d", op="prod", a_shape=[2, 2], a=[1.0, 2.0, 3.0, 4.0],
                  b_shape=[2, 2], b=[5.0, 6.0, 7.0, 8.0], expect_shape=[2, 2, 2, 2],
                  expect=[5.0, 6.0, 7.0, 8.0, 10.0, 12.0, 14.0, 16.0, 15.0, 18.0, 21.0, 24.0, 20.0, 24.0, 28.0, 32.0]))

ews("tensor.rs:1049 test_add_tensor", "add", [4], [1.0, 2.0, 3.0, 4.0], 3.0, [4.0, 5.0, 6.0, 7.0])
ew("tensor.rs:1061 test_add_tensors", "add", [2, 2], [1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [6.0, 8.0, 10.0, 12.0])
ews("tensor.rs:1075 test_sub_tensor", "sub", [4], [5.0, 6.0, 7.0, 8.0], 3.0, [2.0, 3.0, 4.0, 5.0])
ew("tensor.rs:1088 test_sub_tensors", "sub", [2, 2], [5.0, 6.0, 7.0, 8.0], [1.0, 2.0, 3.0, 4.0], [4.0, 4.0, 4.0, 4.0])
ews("tensor.rs:1102 test_mul_tensor", "mul", [4], [1.0, 2.0, 3.0, 4.0], 2.0, [2.0, 4.0, 6.0, 8.0])
ews("tensor.rs:1115 test_div_tensor", "div", [4], [4.0, 6.0, 8.0, 10.0], 2.0, [2.0, 3.0, 4.0, 5.0])
ew("tensor.rs:1128 test_vec_vec_mul_single", "mul", [1], [2.0], [5.0], [10.0])
ew("tensor.rs:1143 test_vec_vec_mul", "mul", [4], [1.0, 2.0, 3.0, 4.0], [2.0, 3.0, 4.0, 5.0], [2.0, 6.0, 12.0, 20.0])
ew("tensor.rs:1158 test_matrix_matrix_mul", "mul", [2, 3], T12[:6], [7.0, 8.0, 9.0, 10.0, 11.0, 12.0],
   [7.0, 16.0, 27.0, 40.0, 55.0, 72.0])
ew("tensor.rs:1177 test_add_tensor_vector", "add", [4], [1.0, 2.0, 3.0, 4.0], [2.0, 3.0, 4.0, 5.0],
   [3.0, 5.0, 7.0, 9.0], rhs="vector")
ew("tensor.rs:1192 test_sub_tensor_vector", "sub", [4], [2.0, 3.0, 4.0, 5.0], [1.0, 2.0, 3.0, 4.0],
   [1.0, 1.0, 1.0, 1.0], rhs="vector")
ew("tensor.rs:1207 test_mul_tensor_vector", "mul", [4], [2.0, 3.0, 4.0, 5.0], [1.0, 2.0, 3.0, 4.0],
   [2.0, 6.0, 12.0, 20.0], rhs="vector")
ew("tensor.rs:1222 test_div_tensor_vector", "div", [4], [2.0, 4.0, 6.0, 8.0], [1.0, 2.0, 3.0, 4.0],
   [2.0, 2.0, 2.0, 2.0], rhs="vector")
ew("tensor.rs:1237 test_add_tensor_matrix", "add", [2, 2], [1.0, 2.0, 3.0, 4.0], [2.0, 3.0, 4.0, 5.0],
   [3.0, 5.0, 7.0, 9.0], rhs="matrix")
ew("tensor.rs:1252 test_sub_tensor_matrix", "sub", [2, 2], [2.0, 3.0, 4.0, 5.0], [1.0, 2.0, 3.0, 4.0],
   [1.0, 1.0, 1.0, 1.0], rhs="matrix")
ew("tensor.rs:1267 test_mul_tensor_matrix", "mul", [2, 2], [2.0, 3.0, 4.0, 5.0], [1.0, 2.0, 3.0, 4.0],
   [2.0, 6.0, 12.0, 20.0], rhs="matrix")
ew("tensor.rs:1282 test_div_tensor_matrix", "div", [2, 2], [2.0, 4.0, 6.0, 8.0], [1.0, 2.0, 3.0, 4.0],
   [2.0, 2.0, 2.0, 2.0], rhs="matrix")
CASES.append(dict(ref="tensor.rs:1297 test_display_1d_tensor", op="display", shape=[3], data=[1.0, 2.0, 3.0], text="[1, 2, 3]"))
CASES.append(dict(ref="tensor.rs:1306 test_display_2d_tensor", op="display", shape=[2, 2], data=[1.0, 2.0, 3.0, 4.0],
                  text="[[1, 2],\n [3, 4]]"))
CASES.append(dict(ref="tensor.rs:1315 test_display_3d_tensor", op="display", shape=[2, 2, 2],
                  data=[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0], text="[[[1, 2],\n  [3, 4]],\n\n [[5, 6],\n  [7, 8]]]"))
CASES.append(dict(ref="tensor.rs:1324 test_display_4d_tensor", op="display", shape=[2, 2, 2, 2],
                  data=[float(i) for i in range(1, 17)],
                  text="[[[[1, 2],\n   [3, 4]],\n\n  [[5, 6],\n   [7, 8]]],\n\n\n [[[9, 10],\n   [11, 12]],\n\n  [[13, 14],\n   [15, 16]]]]"))
CASES.append(dict(ref="tensor.rs:1335 test_pow_tensor_square", op="pow", shape=[2, 2], a=[1.0, 2.0, 3.0, 4.0],
                  power=2.0, expect=[1.0, 4.0, 9.0, 16.0]))
CASES.append(dict(ref="tensor.rs:1345 test_pow_tensor_sqrt", op="pow", shape=[2, 2], a=[1.0, 4.0, 9.0, 16.0],
                  power=0.5, expect=[1.0, 2.0, 3.0, 4.0]))
CASES.append(dict(ref="tensor.rs:1355 test_pow_tensor_negative_exponent", op="pow", shape=[2, 2],
                  a=[1.0, 2.0, 4.0, 8.0], power=-1.0, expect=[1.0, 0.5, 0.25, 0.125]))

# ---- matrix.rs -----------------------------------------------------------------------------------
M1234 = [1.0, 2.0, 3.0, 4.0]
M2345 = [2.0, 3.0, 4.0, 5.0]
CASES.append(dict(ref="matrix.rs:273 test_from_tensor_fail", op="from_tensor_err", kind="matrix", shape=[2, 2, 2],
                  data=[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]))
CASES.append(dict(ref="matrix.rs:282 test_fill", op="fill", shape=[2, 2], value=3.0, dtype="f64", ctor="fill", via="matrix"))
CASES.append(dict(ref="matrix.rs:293 test_eye", op="eye", shape=[3, 3],
                  expect=[1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]))
red("matrix.rs:373 test_sum", "sum", [2, 2], M1234, [0, 1], [1], [10.0], via="matrix")
red("matrix.rs:383 test_mean", "mean", [2, 2], M1234, [0, 1], [1], [2.5], via="matrix")
red("matrix.rs:393 test_var", "var", [2, 2], M1234, [0, 1], [1], [1.25], via="matrix")
red("matrix.rs:403 test_min", "min", [2, 2], M1234, [0, 1], [1], [1.0], via="matrix")
red("matrix.rs:413 test_max", "max", [2, 2], [-1.0, -2.0, -3.0, -4.0], [0, 1], [1], [-1.0], via="matrix")
CASES.append(dict(ref="matrix.rs:423 test_matmul", op="matmul", a_shape=[2, 2], a=M1234, b_shape=[2, 2], b=M2345,
                  expect_shape=[2, 2], expect=[10.0, 13.0, 22.0, 29.0]))
CASES.append(dict(ref="matrix.rs:438 test_vecmul", op="matvec", a_shape=[2, 2], a=M1234, v=[1.0, 2.0],
                  expect_shape=[2], expect=[5.0, 11.0]))
ews("matrix.rs:451 test_add_scalar", "add", [2, 2], M1234, 2.0, [3.0, 4.0, 5.0, 6.0], lhs="matrix")
ew("matrix.rs:464 test_add_matrix", "add", [2, 2], M1234, M2345, [3.0, 5.0, 7.0, 9.0], lhs="matrix", rhs="matrix")
ew("matrix.rs:479 test_add_matrix_tensor", "add", [2, 2], M1234, M2345, [3.0, 5.0, 7.0, 9.0], lhs="matrix")
ews("matrix.rs:494 test_sub_scalar", "sub", [2, 2], M1234, 2.0, [-1.0, 0.0, 1.0, 2.0], lhs="matrix")
ew("matrix.rs:507 test_sub_matrix", "sub", [2, 2], M1234, M2345, [-1.0] * 4, lhs="matrix", rhs="matrix")
ew("matrix.rs:522 test_sub_matrix_tensor", "sub", [2, 2], M1234, M2345, [-1.0] * 4, lhs="matrix")
ews("matrix.rs:537 test_mul_scalar", "mul", [2, 2], M1234, 2.0, [2.0, 4.0, 6.0, 8.0], lhs="matrix")
ew("matrix.rs:550 test_mul_matrix", "mul", [2, 2], M1234, M2345, [2.0, 6.0, 12.0, 20.0], lhs="matrix", rhs="matrix")
ew("matrix.rs:565 test_mul_matrix_tensor", "mul", [2, 2], M1234, M2345, [2.0, 6.0, 12.0, 20.0], lhs="matrix")
ews("matrix.rs:580 test_div_scalar", "div", [2, 2], [4.0, 6.0, 8.0, 10.0], 2.0, [2.0, 3.0, 4.0, 5.0], lhs="matrix")
ew("matrix.rs:593 test_div_matrix", "div", [2, 2], [4.0, 6.0, 8.0, 10.0], M2345, [2.0] * 4, lhs="matrix", rhs="matrix")
ew("matrix.rs:608 test_div_matrix_tensor", "div", [2, 2], [4.0, 6.0, 8.0, 10.0], M2345, [2.0] * 4, lhs="matrix")
CASES.append(dict(ref="matrix.rs:645 test_pow_matrix", op="pow", shape=[2, 2], a=M2345, power=2.0,
                  expect=[4.0, 9.0, 16.0, 25.0], via="matrix"))

# ---- vector.rs -----------------------------------------------------------------------------------
CASES.append(dict(ref="vector.rs:262 test_from_tensor_fail", op="from_tensor_err", kind="vector", shape=[2, 2],
                  data=M1234))
CASES.append(dict(ref="vector.rs:271 test_fill", op="fill", shape=[4], value=3.0, dtype="f64", ctor="fill", via="vector"))
red("vector.rs:334 test_sum", "sum", [4], M1234, [], [1], [10.0], via="vector")
red("vector.rs:343 test_mean", "mean", [4], M1234, [], [1], [2.5], via="vector")
red("vector.rs:352 test_var", "var", [4], M1234, [], [1], [1.25], via="vector")
red("vector.rs:361 test_min", "min", [4], M1234, [], [1], [1.0], via="vector")
red("vector.rs:370 test_max", "max", [4], [-1.0, -2.0, -3.0, -4.0], [], [1], [-1.0], via="vector")
CASES.append(dict(ref="vector.rs:379 test_vecmul", op="dot", a=M1234, b=M2345, expect_shape=[1], expect=[40.0]))
CASES.append(dict(ref="vector.rs:390 test_matmul", op="vecmat", v=[1.0, 2.0], a_shape=[2, 2], a=M1234,
                  expect_shape=[2], expect=[7.0, 10.0]))
CASES.append(dict(ref="vector.rs:402 test_prod", op="prod", a_shape=[4], a=M1234, b_shape=[4], b=M2345,
                  expect_shape=[4, 4],
                  expect=[2.0, 3.0, 4.0, 5.0, 4.0, 6.0, 8.0, 10.0, 6.0, 9.0, 12.0, 15.0, 8.0, 12.0, 16.0, 20.0]))
ews("vector.rs:426 test_add_scalar", "add", [4], M1234, 2.0, [3.0, 4.0, 5.0, 6.0], lhs="vector")
ew("vector.rs:437 test_add_vector", "add", [4], M1234, M2345, [3.0, 5.0, 7.0, 9.0], lhs="vector", rhs="vector")
ew("vector.rs:452 test_add_vector_tensor", "add", [4], M1234, M2345, [3.0, 5.0, 7.0, 9.0], lhs="vector")
ews("vector.rs:467 test_sub_scalar", "sub", [4], M1234, 2.0, [-1.0, 0.0, 1.0, 2.0], lhs="vector")
ew("vector.rs:480 test_sub_vector", "sub", [4], M1234, M2345, [-1.0] * 4, lhs="vector", rhs="vector")
ew("vector.rs:495 test_sub_vector_tensor", "sub", [4], M1234, M2345, [-1.0] * 4, lhs="vector")
ews("vector.rs:510 test_mul_scalar", "mul", [4], M1234, 2.0, [2.0, 4.0, 6.0, 8.0], lhs="vector")
ew("vector.rs:523 test_mul_vector", "mul", [4], M1234, M2345, [2.0, 6.0, 12.0, 20.0], lhs="vector", rhs="vector")
ew("vector.rs:538 test_mul_vector_tensor", "mul", [4], M1234, M2345, [2.0, 6.0, 12.0, 20.0], lhs="vector")
ews("vector.rs:553 test_div_scalar", "div", [4], [4.0, 6.0, 8.0, 10.0], 2.0, [2.0, 3.0, 4.0, 5.0], lhs="vector")
ew("vector.rs:566 test_div_vector", "div", [4], [4.0, 6.0, 8.0, 10.0], M2345, [2.0] * 4, lhs="vector", rhs="vector")
ew("vector.rs:581 test_div_vector_tensor", "div", [4], [4.0, 6.0, 8.0, 10.0], M2345, [2.0] * 4, lhs="vector")
CASES.append(dict(ref="vector.rs:596 test_pow_vector", op="pow", shape=[4], a=M2345, power=2.0,
                  expect=[4.0, 9.0, 16.0, 25.0], via="vector"))

# ---- bookkeeping: shape.rs, coordinate.rs, iter.rs, storage.rs -----------------------------------
CASES.append(dict(ref="shape.rs:73 test_shape_new", op="shape_new", dims=[2, 3, 4]))
CASES.append(dict(ref="shape.rs:80 test_shape_new_with_zero_dimension", op="shape_new_err", dims=[2, 0, 4]))
CASES.append(dict(ref="shape.rs:87 test_shape_size", op="shape_size", dims=[2, 3, 4], size=24))
CASES.append(dict(ref="shape.rs:93 test_shape_display", op="shape_display", dims=[2, 3, 4], text="(2, 3, 4)"))
CASES.append(dict(ref="shape.rs:108 test_shape_extend", op="shape_stack", a=[2, 3], b=[4, 5], expect=[2, 3, 4, 5]))
CASES.append(dict(ref="coordinate.rs:80 test_order", op="coord_order", indices=[1, 2, 3], order=3))
CASES.append(dict(ref="coordinate.rs:96 test_insert", op="coord_insert", indices=[1, 2, 3], index=1, value=4,
                  expect=[1, 4, 2, 3]))
CASES.append(dict(ref="coordinate.rs:118 test_display", op="coord_display", indices=[1, 2, 3], text="(1, 2, 3)"))
CASES.append(dict(ref="iter.rs:56 test_index_iterator", op="index_iterator", dims=[2, 3],
                  expect=[[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2]]))
CASES.append(dict(ref="iter.rs:70 test_index_iterator_single_dimension", op="index_iterator", dims=[4],
                  expect=[[0], [1], [2], [3]]))
CASES.append(dict(ref="iter.rs:82 test_index_iterator_empty_tensor", op="index_iterator", dims=[], expect=[]))


def _numbers(text):
    return [float(x) for x in re.findall(r"(?<![\w.])-?\d+(?:\.\d+)?", text)]


def _case_numbers(case):
    out = []
    for k, v in case.items():
        if k in ("ref", "op", "kind", "opname", "lhs", "rhs", "via", "dtype", "ctor", "text"):
            continue
        stack = [v]
        while stack:
            x = stack.pop()
            if isinstance(x, (list, tuple)):
                stack.extend(x)
            elif isinstance(x, (int, float)) and not isinstance(x, bool):
                out.append(float(x))
    return out


def verify():
    bad = 0
    for case in CASES:
        loc, name = case["ref"].split(" ")
        fname, line = loc.split(":")
        lines = open(os.path.join(REF, fname)).read().split("\n")
        line = int(line)
        window = "\n".join(lines[line - 2: line + 1])
        if f"fn {name}(" not in window:
            print(f"MISCITED {case['ref']}: fn not at that line")
            bad += 1
            continue
        # body: until the next #[test] attribute or end of file
        start = line - 1
        end = start + 1
        while end < len(lines) and "#[test]" not in lines[end]:
            end += 1
        body_nums = set(_numbers("\n".join(lines[start:end])))
        missing = [x for x in _case_numbers(case) if x not in body_nums]
        if case.get("via") == "vector" and "shape" in case:  # Vector::new(&data): the length is implied
            missing = [x for x in missing if x != float(case["shape"][0])]
        # sizes such as 6 (= 2*3) or implied zeros/ones of eye/zeros tests are derived, not literal
        if case["op"] in ("fill", "eye", "size", "shape_size", "coord_order"):
            missing = []
        if missing:
            print(f"MISMATCH {case['ref']}: numbers {sorted(set(missing))} not found in the cited test body")
            bad += 1
    print(f"verified {len(CASES) - bad}/{len(CASES)} cases against {REF}")
    return bad


def main():
    if "--verify" in sys.argv:
        if not os.path.isdir(REF):
            print("reference sources not present; nothing to verify against")
        elif verify():
            sys.exit(1)
    path = os.path.join(HERE, "reference_kat.json")
    with open(path, "w") as f:
        json.dump({"source": "Rust-Scientific-Computing/feotensor inline #[test] functions (transcribed)",
                   "dtype_of_literals": "f64", "cases": CASES}, f, indent=1)
    print(f"wrote {len(CASES)} cases to {path}")


if __name__ == "__main__":
    main()
