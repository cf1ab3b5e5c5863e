"""Pins the oracle's colour maths against the reference's own known-answer vectors
(test/tst_ColourDifference.h:26-108, extracted by tests/golden/make_colour_vectors.py) and, when
oracle/_ref/libref_core.so is present, against the reference's ColourDifference.cpp compiled unmodified."""
import ctypes
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "colour_vectors.json")))
REF_SO = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libref_core.so")


@pytest.mark.parametrize("name", ["rgb_euclidean", "cie76", "ciede2000"])
def test_known_answers(oracle, name):
    fn = {"rgb_euclidean": oracle.rgb_euclidean, "cie76": oracle.cie76, "ciede2000": oracle.ciede2000}[name]
    s = GOLD["sets"][name]
    assert len(s["vectors"]) == {"rgb_euclidean": 8, "cie76": 6, "ciede2000": 34}[name]
    for v in s["vectors"]:
        assert abs(fn(v["first"], v["second"]) - v["difference"]) <= s["tolerance"], v


def _ref():
    if not os.path.exists(REF_SO):
        pytest.skip("libref_core.so not built (reference sources absent)")
    L = ctypes.CDLL(REF_SO)
    L.ref_diff_f32.restype = ctypes.c_double
    L.ref_diff_f32.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    return L


@pytest.mark.parametrize("diff_type", [0, 1, 2])
def test_oracle_equals_compiled_reference(oracle, diff_type):
    """Seeded version of the reference's *_CPUvsCUDA random-pixel tests (tst_ColourDifference.h:315-387):
    L in [0,100], a,b in [-128,127], RGB in [0,255]; here restatement vs the reference's own object code."""
    L = _ref()
    rng = np.random.default_rng(100 + diff_type)
    n = 1 << 14
    if diff_type == 0:
        a = rng.uniform(0, 255, (n, 3)).astype(np.float32)
        b = rng.uniform(0, 255, (n, 3)).astype(np.float32)
    else:
        lo, hi = np.array([0, -128, -128]), np.array([100, 127, 127])
        a = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        b = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    # degenerate pairs: identical, achromatic, opposite hue
    a[:8] = b[:8]
    a[8:16, 1:] = 0
    b[16:24, 1:] = -a[16:24, 1:]
    mine = oracle.diff_batch(diff_type, a, b)
    fp = ctypes.POINTER(ctypes.c_float)
    ref = np.array([L.ref_diff_f32(diff_type, a[i].ctypes.data_as(fp), b[i].ctypes.data_as(fp)) for i in range(n)])
    # same formula, same f64 arithmetic: allow only last-bit differences from expression scheduling
    np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-12)
