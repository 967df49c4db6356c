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


def _fp8_encode(x: float, mb: int, bias: int, maxbits: int, ovf: int, infb: int) -> int:
    """float -> fp8 bits, RNE with subnormals (reference: buffer/nx_buffer_stubs.h:189-278)."""
    f = np.float32(x)
    b = int(np.array([f], dtype=np.float32).view(np.uint32)[0])
    if (b & 0x7FFFFFFF) > 0x7F800000:
        return 0x7F
    sign = (b >> 31) << 7
    if (b & 0x7FFFFFFF) == 0x7F800000:
        return sign | infb
    ex = ((b >> 23) & 0xFF) - 127
    emin = 1 - bias
    if ex >= emin:
        sh = 23 - mb
        sig = b & 0x7FFFFF
        q, rem, half = sig >> sh, sig & ((1 << sh) - 1), 1 << (sh - 1)
        if rem > half or (rem == half and (q & 1)):
            q += 1
        bits = ((ex + bias) << mb) + q
        return sign | (ovf if bits >= maxbits else bits)
    shift = (23 - mb) + (emin - ex)
    if shift > 24:
        return sign
    sig = (b & 0x7FFFFF) | 0x800000
    q, rem, half = sig >> shift, sig & ((1 << shift) - 1), 1 << (shift - 1)
    if rem > half or (rem == half and (q & 1)):
        q += 1
    return sign | q


def encode_scalar(dt: Dtype, value) -> np.ndarray:
    """One element of `dt` in its storage representation, from a Python number (the
    `'a` of `full : context -> ('a,'b) Dtype.t -> int array -> 'a -> t`). Integers
    already in storage form (numpy integer scalars for f16/bf16/fp8 bit patterns)
    pass through unchanged."""
    out = np.zeros(1, dtype=dt.np)
    raw_bits = isinstance(value, (np.integer,)) and dt.name in ("f16", "bf16", "f8e4m3", "f8e5m2")
    if dt.name == "f16" and not raw_bits:
        out[0] = np.array([value], dtype=np.float16).view(np.uint16)[0]
    elif dt.name == "bf16" and not raw_bits:
        b = int(np.array([value], dtype=np.float32).view(np.uint32)[0])
        out[0] = ((b >> 16) | 0x40) if (b & 0x7FFFFFFF) > 0x7F800000 else ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) & 0xFFFF
    elif dt.name == "f8e4m3" and not raw_bits:
        out[0] = _fp8_encode(value, 3, 7, 0x7F, 0x7F, 0x7F)
    elif dt.name == "f8e5m2" and not raw_bits:
        out[0] = _fp8_encode(value, 2, 15, 0x7C, 0x7C, 0x7C)
    elif dt.name == "bool":
        out[0] = 1 if value else 0
    elif dt.cls in ("sint", "uint"):
        out[0] = np.array([int(value) & ((1 << (8 * dt.itemsize)) - 1)], dtype=np.uint64).astype(dt.np)[0] \
            if not isinstance(value, np.generic) else value
    else:
        out[0] = value
    return out
