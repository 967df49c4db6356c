"""The linalg tier-1 restatement (oracle/nxo.py: cholesky, triangular_solve, qr after the
reference's unblocked kernels) against golden vectors produced by the reference's own
nx_c_tri.c / nx_c_qr.c (tests/golden/nx_reference_linalg.npz), and against the reference
binary where present. Tolerances are relative to the largest output magnitude: a few ulp of the
compute type times the problem size (the reference's blocked paths reassociate), and the storage
type's rounding for the 16-bit floats."""
import os

import numpy as np
import pytest

from oracle import nxo, ref
from oracle.hostview import HostView
from tests.golden.make_golden_linalg import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nx_reference_linalg.npz")
TOL = {"f32": 4e-4, "f64": 1e-11, "c32": 4e-4, "c64": 1e-11, "bf16": 6e-2, "f16": 8e-3}


def _check(got, want, key):
    assert got.shape == want.shape, key
    if want.size:
        err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
        assert err <= TOL[key.split("|")[1]], f"{key}: {err:.3e}"


def test_linalg_restatement_matches_reference_golden_vectors():
    gold = np.load(GOLD)
    seen = set()
    for key, thunk in cases():
        _check(thunk(nxo), gold[key], key)
        seen.add(key)
    assert len(seen) == len(gold.files) and len(seen) >= 400


@pytest.mark.skipif(not ref.available(), reason="reference binary not built")
def test_linalg_error_classes_match_reference():
    spd = HostView.from_array(np.eye(3), "f64")
    for mod in (ref, nxo):
        with pytest.raises(mod.RefError) as e:
            mod.cholesky(HostView.from_array(-np.eye(3), "f64"))
        assert e.value.kind == "Failure" and e.value.msg.endswith("matrix is not positive definite")
        with pytest.raises(mod.RefError) as e:
            mod.cholesky(HostView.from_array(np.ones((2, 3)), "f64"))
        assert e.value.kind == "Invalid_argument" and e.value.msg.endswith("matrix must be square")
        with pytest.raises(mod.RefError) as e:
            mod.cholesky(HostView.from_array(np.ones((2, 2), dtype=np.int32), "i32"))
        assert e.value.kind == "Invalid_argument" and e.value.msg.endswith("linalg requires a float or complex dtype")
        with pytest.raises(mod.RefError) as e:
            mod.triangular_solve(HostView.from_array(np.zeros((3, 3)), "f64"), spd)
        assert e.value.kind == "Failure" and e.value.msg.endswith("triangular matrix is singular")
