"""Parity at BASELINE.json's full config-4 size (7680x4320 main, 10,000-image library of 128 px tiles, CIEDE2000,
repeat 8/500), where the CPU oracle cannot produce the whole answer (3.3e11 CIEDE2000 evaluations ~ 11 CPU-hours):
  * difference sums of sampled (cell, library image) pairs against the f64 oracle, 1e-4 relative;
  * the wavefront selection over all 2,040 cells against the oracle's selection rule applied to the engine's own
    difference matrix (exact, it is integer/compare work);
  * invariants: every valid cell filled, border cells use their clipped bounds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config4_full_size(oracle):
    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic
    H, W, N, S, rr, ra = 4320, 7680, 10000, 128, 8, 500
    main = synthetic.make_main_image(H, W, 2004)
    lib = synthetic.make_library(N, S, 1004)
    gen = PhotomosaicGenerator(0)
    gen.setMainImage(main)
    gen.setLibrary(lib)
    gen.setColourDifference(2)
    cg = CellGroup()
    cg.setCellShape(CellShape(S))
    gen.setCellGroup(cg)
    state = gen.computeGridState()[0]
    assert state.shape == (36, 62) and int((state >= 0).sum()) == 2040
    gen.setRepeat(rr, ra)
    gen.setKeepDifferences(True)
    assert gen.generateBestFits()
    grid = gen.getBestFits()[0]
    D = gen.getDifferences(0)
    tm = gen.getTimings()
    gen.close()
    assert D.shape == (2040, N) and np.isfinite(D).all() and (D > 0).all()
    assert ((grid >= 0) == (state >= 0)).all()

    # (1) sampled difference sums vs the f64 oracle: 3 cells (interior, right border, clipped bottom row) x 400 images
    og = oracle.CellGroup.make(oracle.CellShape.square(S), 100, 0)
    mains = [oracle.to_working_space(main, oracle.CIEDE2000)]
    rng = np.random.default_rng(5)
    sample_lib = np.sort(rng.choice(N, 400, replace=False))
    lib_f = oracle.preprocess_library(lib[sample_lib], og, oracle.CIEDE2000)
    masks4 = og.detail_cells[0].masks4()
    cols = 60
    for (cy, cx) in [(7, 13), (20, 59), (33, 31)]:
        sub = np.full_like(state, -1)
        sub[cy + 2, cx + 2] = 0
        cells, bounds, flips, _ = oracle.extract_cells(mains, og, 0, sub)
        r = oracle.generate_step(oracle.CIEDE2000, cells, bounds, flips, lib_f, masks4, sub, 0, 0, want_D=True, early_exit=False)
        got = D[cy * cols + cx, sample_lib].astype(np.float64)
        err = np.abs(got - r.D[0]) / r.D[0]
        assert err.max() < 1e-4, (cy, cx, err.max())
        if cy == 33:  # bottom row: 4320 = 33.75 x 128 -> only 96 of 128 rows are inside the image
            assert tuple(bounds[0]) == (0, 0, 128, 96)

    # (2) selection over the whole grid: oracle rule on the engine's D must reproduce the engine's grid exactly
    want = oracle.select_from_D(D.astype(np.float64), state, rr, ra)
    assert np.array_equal(want, grid), int((want != grid).sum())

    # (3) size of the tie band at full scale: cells whose best two penalised candidates are within 1e-4 relative
    assert tm["pixel_diffs"] == 60 * 33 * 16384 * N + 60 * 96 * 128 * N
