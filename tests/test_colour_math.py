"""FP32 colour formulas of the CUDA kernels (csrc/colour_math.cuh) checked on the CPU against the
f64 oracle: golden vectors of the reference (test/tst_ColourDifference.h) and seeded random pixels
(the reference's *_CPUvsCUDA tests allow 1e-4 absolute per pixel, tst_ColourDifference.h:315-387)."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "colour_vectors.json")))


@pytest.fixture(scope="module")
def cm():
    out = os.path.join(HERE, "helpers", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcolour_math_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "helpers", "colour_math_check.cpp"), "-lm"])
    L = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    for f in (L.cm_euclid_batch, L.cm_ciede2000_batch, L.cm_ciede2000_batch_x2):
        f.argtypes = [fp, fp, ctypes.c_long, fp]

    def run(fn, a, b):
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
        o = np.empty(a.shape[0], np.float32)
        fn(a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.shape[0], o.ctypes.data_as(fp))
        return o.astype(np.float64)
    return {"euclid": lambda a, b: run(L.cm_euclid_batch, a, b), "ciede2000": lambda a, b: run(L.cm_ciede2000_batch, a, b),
            "ciede2000_x2": lambda a, b: run(L.cm_ciede2000_batch_x2, a, b)}


@pytest.mark.parametrize("name", ["rgb_euclidean", "cie76", "ciede2000"])
def test_known_answers_fp32(cm, oracle, name):
    s = GOLD["sets"][name]
    a = np.array([v["first"] for v in s["vectors"]], np.float32)
    b = np.array([v["second"] for v in s["vectors"]], np.float32)
    want = np.array([v["difference"] for v in s["vectors"]])
    got = cm["ciede2000" if name == "ciede2000" else "euclid"](a, b)
    err = np.abs(got - want)
    tol = 1e-4 + 1e-6 * np.abs(want)
    if name == "ciede2000":
        # Sharma's pairs 9-16 sit ON the mean-hue discontinuity (hues 180deg apart to within 0.03deg): which
        # side a pair falls on changes with the f32 rounding of the inputs -- the f64 oracle itself returns
        # 7.2195 instead of 7.1792 for pair 10 once its inputs are f32 (as they are in the generator).
        # For those pairs either side of the discontinuity is accepted, at 1e-3.
        knife = np.zeros(len(want), bool)
        knife[8:16] = True
        alt = oracle.diff_batch(2, a, b)
        others = np.concatenate([want[8:16], alt[8:16]])
        kerr = np.min(np.abs(got[knife, None] - others[None, :]), axis=1)
        assert (kerr < 1e-3).all(), kerr
        err = err[~knife]
        tol = tol[~knife]
    # reference tolerance for its own CUDA kernels on these vectors: 1e-4 (tst_ColourDifference.h:233-309)
    assert (err <= tol).all(), err.max()


def _lab(rng, n):
    lo, hi = np.array([0, -128, -128]), np.array([100, 127, 127])
    return rng.uniform(lo, hi, (n, 3)).astype(np.float32)


def test_ciede2000_random_vs_oracle(cm, oracle):
    rng = np.random.default_rng(7)
    n = 1 << 18
    a, b = _lab(rng, n), _lab(rng, n)
    # near-neutral, identical, one-achromatic and close-hue pairs are where the algebra is delicate
    a[:4096, 1:] *= 0.01
    b[4096:8192] = a[4096:8192]
    a[8192:12288, 1:] = 0
    b[12288:16384] = a[12288:16384] + rng.normal(0, 0.5, (4096, 3)).astype(np.float32)
    # nearly (not exactly) opposite hues, 0.05..5 degrees away from the 180-degree mean-hue discontinuity
    # (ON the discontinuity the reference's own result flips with the last bit of atan2)
    ang = np.radians(rng.uniform(0.05, 5.0, 4096) * rng.choice([-1, 1], 4096))
    va = a[16384:20480, 1:].astype(np.float64) * rng.uniform(0.5, 1.5, (4096, 1))
    b[16384:20480, 1] = -(va[:, 0] * np.cos(ang) - va[:, 1] * np.sin(ang))
    b[16384:20480, 2] = -(va[:, 0] * np.sin(ang) + va[:, 1] * np.cos(ang))
    got = cm["ciede2000"](a, b)
    want = oracle.diff_batch(2, a, b)
    err = np.abs(got - want)
    assert np.isfinite(got).all()
    # per-pixel: the reference accepts 1e-4 absolute between its own CPU and CUDA paths; here 1e-4 absolute
    # plus 2e-5 relative (FP32 arithmetic with approximate MUFU ops on values up to ~180)
    assert (err <= 1e-4 + 2e-5 * want).all(), (err - 2e-5 * want).max()
    big = want > 1.0
    assert np.quantile(err[big] / want[big], 0.999) < 1e-5
    # the sums the generator compares: relative error of a 16384-pixel sum
    s_got = got.reshape(-1, 16384).sum(1)
    s_want = want.reshape(-1, 16384).sum(1)
    assert np.max(np.abs(s_got - s_want) / s_want) < 1e-5


def test_euclid_random_vs_oracle(cm, oracle):
    rng = np.random.default_rng(8)
    n = 1 << 16
    a = rng.integers(0, 256, (n, 3)).astype(np.float32)
    b = rng.integers(0, 256, (n, 3)).astype(np.float32)
    b[:100] = a[:100]
    got = cm["euclid"](a, b)
    want = oracle.diff_batch(0, a, b)
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-6)


def test_packed_form_equals_scalar_form(cm):
    """mm_ciede2000_stored_v<mm_f2> (the FFMA2 kernel path: one cell pixel against two library pixels) is the same source as
    the scalar form; on the CPU both must agree bit for bit."""
    rng = np.random.default_rng(9)
    n = 1 << 16
    a, b = _lab(rng, n), _lab(rng, n)
    a[1::2] = a[0::2]  # the packed form shares the first colour between its two lanes
    a[:2000, 1:] = 0
    b[2000:4000, 1:] = 0
    assert np.array_equal(cm["ciede2000"](a, b), cm["ciede2000_x2"](a, b))
