"""Self-consistency of the oracle's generator restatement (CPUPhotomosaicGenerator.cpp:33-225) on small inputs:
the reference's early exit must not change the result, selection from the D matrix must reproduce the raster loop,
and the repeat rule must match the reference's kernel-test restatement (test/tst_CUDAKernel.h:64-87)."""
import numpy as np
import pytest


def _case(oracle, diff, detail, rr, ra, seed=5):
    from mosaicmagnifique_b200 import synthetic
    main = synthetic.make_main_image(150, 210, seed, block=32)
    lib = synthetic.make_library(40, 32, seed + 1)
    g = oracle.CellGroup.make(oracle.CellShape.square(32), detail, 0)
    st = oracle.grid_state(g, main)
    return main, lib, g, st


@pytest.mark.parametrize("diff", [0, 2])
def test_early_exit_is_equivalent(oracle, diff):
    pytest.importorskip("cv2")
    main, lib, g, st = _case(oracle, diff, 50, 2, 400)
    a = oracle.generate(main, lib, g, st, diff, 0, 2, 400, want_D=True, early_exit=True)[0]
    b = oracle.generate(main, lib, g, st, diff, 0, 2, 400, want_D=True, early_exit=False)[0]
    assert np.array_equal(a.grid, b.grid)
    assert a.visited <= a.nominal and a.nominal == b.nominal
    sel = oracle.select_from_D(a.D, st[0], 2, 400)
    assert np.array_equal(sel, a.grid)
    assert (a.grid[st[0] < 0] == -1).all() and (a.grid[st[0] >= 0] >= 0).all()


def test_repeat_rule_matches_reference_kernel_test(oracle):
    """Restates the host loop of CUDAKernel.CalculateRepeats (tst_CUDAKernel.h:64-87) for every cell of a 5x5 grid."""
    import ctypes
    rng = np.random.default_rng(1)
    size, n_lib, r, add = 5, 10, 2, 500
    best = rng.integers(0, n_lib, (size, size)).astype(np.int64)
    ids = np.zeros(64, np.int64)
    pen = np.zeros(64, np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    for cy in range(size):
        for cx in range(size):
            n = oracle.lib().mo_calculate_repeats(best.ctypes.data_as(i64p), size, size, cx, cy, r, add,
                                                  ids.ctypes.data_as(i64p), pen.ctypes.data_as(i64p))
            mine = np.zeros(n_lib, np.int64)
            mine[ids[:n]] = pen[:n]
            ref = np.zeros(n_lib, np.int64)
            pos = cy * size + cx
            for y in range(max(0, cy - r), min(size - 1, cy + r) + 1):
                for x in range(max(0, cx - r), min(size - 1, cx + r) + 1):
                    if y * size + x < pos:
                        ref[best[y, x]] += add
            assert np.array_equal(mine, ref), (cx, cy)


def test_grid_state_splits_by_entropy(oracle):
    pytest.importorskip("cv2")
    from mosaicmagnifique_b200 import synthetic
    main = synthetic.make_main_image(256, 384, 51, block=32)
    g = oracle.CellGroup.make(oracle.CellShape.square(64), 100, 2)
    st = oracle.grid_state(g, main)
    assert len(st) == 3
    assert [s.shape for s in st] == [(6, 8), (10, 14), (18, 26)]
    n = [(s >= 0).sum() for s in st]
    assert n[0] > 0 and n[1] > 0 and n[2] > 0


# ---------------------------------------------------------------- committed generator-level golden vectors

def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_golden.npz"))


def golden_case(oracle, G, name):
    """Rebuilds the oracle-side inputs of one fixture of tests/golden/generator_golden.npz (made by make_generator_golden.py)."""
    shape = oracle.CellShape.from_mask(G[name + "/mask"])
    fields = ("row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset", "alt_col_offset",
              "alt_col_flip_h", "alt_col_flip_v", "alt_row_flip_h", "alt_row_flip_v")
    for f, v in zip(fields, G[name + "/shape"]):
        setattr(shape, f, bool(v) if "flip" in f else int(v))
    diff, detail, steps, rr, ra, scheme = (int(v) for v in G[name + "/params"])
    return shape, diff, detail, steps, rr, ra, scheme


def test_oracle_reproduces_generator_golden(oracle):
    """The oracle must reproduce its own committed outputs bit for bit (grid states and grids) and to 1e-12 (f64 sums):
    a different cv2 / numpy / compiler on the machine that runs the tests cannot silently move the parity target."""
    pytest.importorskip("cv2")
    G = _golden()
    for name in G["names"]:
        name = str(name)
        shape, diff, detail, steps, rr, ra, scheme = golden_case(oracle, G, name)
        group = oracle.CellGroup.make(shape, detail, steps)
        states = oracle.grid_state(group, G[name + "/main"])
        res = oracle.generate(G[name + "/main"], G[name + "/lib"], group, states, diff, scheme, rr, ra, want_D=True)
        assert len(states) == steps + 1
        for s, (st, r) in enumerate(zip(states, res)):
            assert np.array_equal(st, G["%s/state%d" % (name, s)])
            assert np.array_equal(r.grid, G["%s/grid%d" % (name, s)])
            np.testing.assert_allclose(r.D, G["%s/D%d" % (name, s)], rtol=1e-12, atol=0)
