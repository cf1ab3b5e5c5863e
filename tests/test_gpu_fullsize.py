"""Parity at BASELINE.json's full config-4 size (7680x4320 main, 10,000-image library of 128 px tiles, CIEDE2000,
repeat 8/500), where the CPU oracle cannot produce the whole answer (3.3e11 CIEDE2000 evaluations ~ 11 CPU-hours):
  * difference sums of sampled (cell, library image) pairs against the f64 oracle, 1e-4 relative;
  * the wavefront selection over all 2,040 cells against the oracle's selection rule applied to the engine's own
    difference matrix (exact, it is integer/compare work);
  * invariants: every valid cell filled, border cells use their clipped bounds."""
import numpy as np
import pytest

import functools

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=1)
def _cfg4_inputs():
    from mosaicmagnifique_b200 import synthetic
    return synthetic.make_main_image(4320, 7680, 2004), synthetic.make_library(10000, 128, 1004)


def test_config4_full_size(oracle):
    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic
    H, W, N, S, rr, ra = 4320, 7680, 10000, 128, 8, 500
    main, lib = _cfg4_inputs()
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


def test_config4_full_library_cells_against_reference_and_f64_oracle(oracle):
    """Full-size parity with teeth: BASELINE config 4 inputs (8K x 10,000 images, CIEDE2000, repeat 8/500), the engine's grid
    compared with
      (a) the reference's OWN object code on the first cells of the grid in raster order x the FULL library (its choice must be the
          engine's choice: same cells, same order, same repeat penalties -- TST_Generator::CompareBestFits, test/tst_Generator.h:25-63);
      (b) complete f64 difference rows (oracle port, no early exit) of 16 cells spread over the grid: all 10,000 sums within the
          spec's 1e-4 relative, their median below 2e-7 and 99 % of them within 1e-5 (the few larger ones sit on the reference
          formula's own discontinuity -- nearly opposite hues, ColourDifference.cpp:109-122 -- where f32 and f64 take different
          mean-hue branches for single pixels of the uniform-noise blocks; measured max 4.7e-5) and the teacher-forced choice -- f64 row + the repeat penalties of the engine's own grid -- equal
          to the engine's, or inside the FP32-explainable tie band (TOL, tests/helpers/parity.py).
    CPU cost ~1-2 minutes (threads; the C calls release the GIL)."""
    import os
    import threading

    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic
    from tests.helpers.parity import TIE_TOL, window_counts
    H, W, N, S, rr, ra = 4320, 7680, 10000, 128, 8, 500
    main, lib = _cfg4_inputs()
    gen = PhotomosaicGenerator(0)
    gen.setMainImage(main)
    gen.setLibrary(lib)
    gen.setColourDifference(2)
    cg = CellGroup()
    cg.setCellShape(CellShape(S))
    gen.setCellGroup(cg)
    state = gen.computeGridState()[0]
    gen.setRepeat(rr, ra)
    gen.setKeepDifferences(True)
    gen.setReportMargins(True)
    assert gen.generateBestFits()
    grid = gen.getBestFits()[0]
    D = gen.getDifferences(0)
    best, second = gen.getMargins(0)
    gen.close()

    og = oracle.CellGroup.make(oracle.CellShape.square(S), 100, 0)
    ys, xs = np.nonzero(state >= 0)
    valid = list(zip(ys.tolist(), xs.tolist()))
    n_threads = max(2, min(16, (os.cpu_count() or 2) - 1))
    sample = [valid[i] for i in np.linspace(0, len(valid) - 1, 16).astype(int)]  # includes the clipped bottom row
    mains = [oracle.to_working_space(main, oracle.CIEDE2000)]
    lib_f = oracle.preprocess_library(lib, og, oracle.CIEDE2000)
    masks4 = og.detail_cells[0].masks4()
    rows = {}

    def f64_row(cell):
        sub = np.full_like(state, -1)
        sub[cell] = 0
        cells, bounds, flips, _ = oracle.extract_cells(mains, og, 0, sub)
        r = oracle.generate_step(oracle.CIEDE2000, cells, bounds, flips, lib_f, masks4, sub, 0, 0, want_D=True, early_exit=False)
        rows[cell] = r.D[0]

    ref_out = {}

    def reference_first_cells(k):
        if not oracle.reference_generator_available():
            return
        sub = np.full_like(state, -1)
        for c in valid[:k]:
            sub[c] = 0
        g, _ = oracle.reference_generate(main, lib, og, [sub], oracle.CIEDE2000, 0, rr, ra)
        ref_out["grid"] = g[0]

    k_ref = 3
    threads = [threading.Thread(target=reference_first_cells, args=(k_ref,))]
    pending = list(sample)
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                if not pending:
                    return
                c = pending.pop()
            f64_row(c)

    threads += [threading.Thread(target=worker) for _ in range(n_threads - 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()

    # (a) the reference's own choice for the first cells of the raster order (they see the same, growing, repeat window)
    if "grid" in ref_out:
        for c in valid[:k_ref]:
            assert ref_out["grid"][c] == grid[c], (c, int(ref_out["grid"][c]), int(grid[c]))
    # (b) complete f64 rows
    index_of = {c: i for i, c in enumerate(valid)}
    worst, in_band = 0.0, 0
    for c in sample:
        want = rows[c]
        got = D[index_of[c]].astype(np.float64)
        err = np.abs(got - want) / want
        worst = max(worst, float(err.max()))
        assert err.max() < 1e-4, (c, float(err.max()))
        assert np.median(err) < 2e-7 and np.quantile(err, 0.99) < 1e-5, (c, float(np.median(err)), float(np.quantile(err, 0.99)))
        v = want + ra * window_counts(grid, c[1], c[0], rr, N)
        b = int(np.argmin(v))
        if int(grid[c]) != b:
            gap = (v[int(grid[c])] - v[b]) / v[b]
            assert gap <= TIE_TOL, (c, int(grid[c]), b, float(gap))
            in_band += 1
    rel = (second.astype(np.float64) - best) / np.maximum(best.astype(np.float64), 1e-30)
    print("config 4 full size: %d f64 rows of 10,000 sums, max relative error %.2e; %d sampled cells chose inside the tie band; cells of "
          "the whole grid with best-two margin <= 1e-4 / 1e-5 / 1e-6: %d / %d / %d of %d; reference object code agreed on the first %d cells"
          % (len(sample), worst, in_band, int((rel <= 1e-4).sum()), int((rel <= 1e-5).sum()), int((rel <= 1e-6).sum()), rel.size,
             k_ref if "grid" in ref_out else 0))
