"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and
exports every symbol include/nxcuda.h declares; the product path fails loudly
(never falls back) when no CUDA device is present."""
import os
import re

import pytest

from raven_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nxcuda.h")).read()
    return sorted(set(re.findall(r"NXC_API[^;]*?\b(nxc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libnxcuda.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "raven_b200/_lib.py binds a different set than nxcuda.h declares"


def test_status_classifier_matches_reference_funnel():
    lib = _lib.load()
    inv = [b"reduction over empty axis has no identity", b"reduce axes must be strictly increasing and in range",
           b"axis out of range", b"output rank inconsistent with the operation",
           b"output has a broadcast (zero) stride", b"shape mismatch"]
    fail = [b"dtype not supported for this operation", b"packed dtype not supported for this operation",
            b"unsupported bigarray kind", b"out of memory", b"argreduce axis length exceeds INT32_MAX",
            b"matmul operands must share one dtype"]
    assert all(lib.nxc_status_is_invalid_argument(s) == 1 for s in inv)
    assert all(lib.nxc_status_is_invalid_argument(s) == 0 for s in fail)


def test_elem_sizes():
    lib = _lib.load()
    from raven_b200 import dtype as D
    for d in D.ALL:
        assert lib.nxc_elem_size(d.tag) == d.itemsize


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import raven_b200.backend as B
    with pytest.raises(_lib.Failure, match="no CUDA device"):
        B.create_context()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "raven_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "libnxo" not in text and "libnxref" not in text, f"{f} references an oracle library"
