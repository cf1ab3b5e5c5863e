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
