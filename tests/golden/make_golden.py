"""Generates tests/golden/nx_reference_golden.npz from the reference's own C backend
(oracle/_ref/libnxref.so, compiled unmodified from /root/reference by oracle/Makefile).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py
The fixture pins the C restatement (oracle/nxo.c) and the CUDA path on machines where
/root/reference does not exist. Inputs are the contract suite's pools / layout matrix
(packages/nx/test/backend_contract.ml:125-289, 423-468) plus its cast edge cases; outputs
are stored as raw storage bytes, errors as (class, message).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from tests import harness as H  # noqa: E402
from tests.test_gpu_map import _cast_inputs  # noqa: E402

ALL = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX) + ["bool"]


def cases():
    """Yields (key, thunk(module) -> HostView). Shared with the checker test."""
    for dt in ALL:
        lay = dict(H.layouts(dt))
        lb = dict(H.layouts(dt, rot=5))
        for op in ref.UNARY:
            for name in ("contig", "transpose", "flip"):
                yield f"un|{op}|{dt}|{name}", (lambda m, op=op, hv=lay[name]: m.unary(op, hv))
        for op in ref.BINARY:
            for na, nb in (("contig", "transpose"), ("broadcast", "contig"), ("flip", "slice")):
                yield f"bin|{op}|{dt}|{na},{nb}", (lambda m, op=op, a=lay[na], b=lb[nb]: m.binary(op, a, b))
        for op in ref.CMP:
            yield f"cmp|{op}|{dt}", (lambda m, op=op, a=lay["contig"], b=lb["transpose"]: m.compare(op, a, b))
        for op in ("sum", "prod", "max", "min"):
            for name in ("contig", "transpose", "permute3"):
                nd = len(lay[name].shape)
                for axes in ([0], [nd - 1], list(range(nd))):
                    yield (f"red|{op}|{dt}|{name}|{axes}",
                           (lambda m, op=op, hv=lay[name], axes=axes: m.reduce(op, hv, axes)))
        for op in ("argmax", "argmin"):
            for name in ("contig", "transpose", "flip"):
                for axis in (0, 1):
                    yield (f"arg|{op}|{dt}|{name}|{axis}",
                           (lambda m, op=op, hv=lay[name], axis=axis: m.argreduce(op, hv, axis, False)))
        data = _cast_inputs(dt)
        hv = H.HostView(data.copy(), dt, [data.size])
        for dst in ALL:
            yield f"cast|{dt}|{dst}", (lambda m, hv=hv, dst=dst: m.cast(hv, dst))


def run(mod, thunk):
    try:
        return thunk(mod).numpy()
    except Exception as e:  # RefError from either oracle module
        return ("err", e.kind, e.msg)


if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    n_err = 0
    for key, thunk in cases():
        r = run(ref, thunk)
        if isinstance(r, tuple):
            out[key] = np.frombuffer(f"{r[1]}|{r[2]}".encode(), dtype=np.uint8)
            n_err += 1
        else:
            out[key] = H.raw(r).copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_reference_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} cases ({n_err} expected errors) to {path}: {os.path.getsize(path)} bytes")
