"""Dtype table of the nx-cuda host mirror.

Tags are Dtype.Packed.tag (reference: packages/nx/lib/core/dtype.mli:76-101,
backend_c/nx_c.h:130-168). `np` is the numpy type used for HOST storage of a
dtype: f16/bf16/fp8 travel as raw bit patterns (uint16/uint8), bool as uint8
0/1, exactly like the reference's extended bigarray kinds
(buffer/nx_buffer_stubs.h:18-25).
"""
from __future__ import annotations

import numpy as np


class Dtype:
    __slots__ = ("name", "tag", "np", "itemsize", "cls")

    def __init__(self, name, tag, np_type, itemsize, cls):
        self.name, self.tag, self.np, self.itemsize, self.cls = name, tag, np.dtype(np_type), itemsize, cls

    def __repr__(self):
        return f"Dtype.{self.name}"


float16 = Dtype("f16", 0, np.uint16, 2, "float")
float32 = Dtype("f32", 1, np.float32, 4, "float")
float64 = Dtype("f64", 2, np.float64, 8, "float")
bfloat16 = Dtype("bf16", 3, np.uint16, 2, "float")
float8_e4m3 = Dtype("f8e4m3", 4, np.uint8, 1, "float")
float8_e5m2 = Dtype("f8e5m2", 5, np.uint8, 1, "float")
int4 = Dtype("i4", 6, np.uint8, 0, "packed")
uint4 = Dtype("u4", 7, np.uint8, 0, "packed")
int8 = Dtype("i8", 8, np.int8, 1, "sint")
uint8 = Dtype("u8", 9, np.uint8, 1, "uint")
int16 = Dtype("i16", 10, np.int16, 2, "sint")
uint16 = Dtype("u16", 11, np.uint16, 2, "uint")
int32 = Dtype("i32", 12, np.int32, 4, "sint")
uint32 = Dtype("u32", 13, np.uint32, 4, "uint")
int64 = Dtype("i64", 14, np.int64, 8, "sint")
uint64 = Dtype("u64", 15, np.uint64, 8, "uint")
complex64 = Dtype("c32", 16, np.complex64, 8, "complex")
complex128 = Dtype("c64", 17, np.complex128, 16, "complex")
bool_ = Dtype("bool", 18, np.uint8, 1, "bool")

ALL = [float16, float32, float64, bfloat16, float8_e4m3, float8_e5m2, int4, uint4, int8, uint8,
       int16, uint16, int32, uint32, int64, uint64, complex64, complex128, bool_]
BY_NAME = {d.name: d for d in ALL}
BY_TAG = {d.tag: d for d in ALL}


def of(x) -> Dtype:
    if isinstance(x, Dtype):
        return x
    return BY_NAME[x]
