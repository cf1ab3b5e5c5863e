"""Pins the oracle's generator restatement on the reference ITSELF: the reference's own CPUPhotomosaicGenerator.cpp
(generateBestFits / findCellBestFit / calculateRepeats, CPUPhotomosaicGenerator.cpp:33-225), ColourDifference.cpp and
GridUtility.cpp are compiled unmodified into oracle/_ref/libref_core.so (oracle/Makefile, stand-in headers oracle/shim,
harness oracle/ref_generator_harness.cpp) and run on the same preprocessed inputs as oracle.generate. The grids must be
IDENTICAL (both sides accumulate in f64 in the same order) and the emitted progress values must follow the reference's
4^(steps-1-step) weights. The option matrix follows the reference's generator tests (test/tst_Generator.h:145-439)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.reference_generator_available():
        pytest.skip("oracle/_ref/libref_core.so with the reference's CPU generator is not built (reference sources absent)")
    return oracle


def _inputs(seed, h, w, n_lib, cell):
    from mosaicmagnifique_b200 import synthetic
    return synthetic.make_main_image(h, w, seed, block=32), synthetic.make_library(n_lib, cell, seed + 1)


def _compare(o, main, lib, shape, diff, detail, steps, rr, ra, scheme=0):
    group = o.CellGroup.make(shape, detail, steps)
    states = o.grid_state(group, main)
    want, progress = o.reference_generate(main, lib, group, states, diff, scheme, rr, ra)
    got = o.generate(main, lib, group, states, diff, scheme, rr, ra, want_D=True)
    assert len(got) == len(want) == steps + 1
    n_valid = 0
    for s, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g.grid, w), "step %d: oracle and reference object code disagree" % s
        n_valid += int((w >= 0).sum())
        # the D matrix the GPU parity tests compare against must reproduce the reference's choice too
        assert np.array_equal(o.select_from_D(g.D, states[s], rr, ra), w)
    # progress(int): one emission per grid position of the padded grid minus the padding loop bounds
    # (CPUPhotomosaicGenerator.cpp:64-91), cumulative, weight 4^(S-1-step) (:55)
    expect, total = [], 0
    for s, st in enumerate(states):
        wgt = 4 ** (len(states) - 1 - s)
        for _ in range(st.shape[0] * st.shape[1]):
            total += wgt
            expect.append(total)
    assert progress == expect
    return n_valid


@pytest.mark.parametrize("diff", [0, 1, 2])
@pytest.mark.parametrize("detail", [100, 50])
def test_square_cells_with_repeats(ref, diff, detail):
    main, lib = _inputs(11 + diff, 200, 300, 60, 32)
    assert _compare(ref, main, lib, ref.CellShape.square(32), diff, detail, 0, 3, 10000) == 70


def test_no_repeats(ref):
    main, lib = _inputs(21, 160, 160, 45, 32)
    _compare(ref, main, lib, ref.CellShape.square(32), 2, 100, 0, 0, 0)


def test_flips_offsets_and_edge_cells(ref):
    from mosaicmagnifique_b200 import synthetic
    sh = ref.CellShape.from_mask(synthetic.triangle_mask(64))
    sh.row_spacing = sh.alt_row_spacing = 64
    sh.col_spacing = sh.alt_col_spacing = 32
    sh.alt_col_flip_v = True
    sh.alt_row_flip_h = True
    main, lib = _inputs(31, 230, 310, 50, 32)
    _compare(ref, main, lib, sh.resized(32), 2, 50, 0, 2, 300)
    hx = ref.CellShape.from_mask(synthetic.hexagon_mask(128))
    hx.row_spacing = hx.alt_row_spacing = 96
    hx.col_spacing = hx.alt_col_spacing = 110
    hx.alt_row_offset = 55
    main, lib = _inputs(41, 250, 330, 40, 32)
    _compare(ref, main, lib, hx.resized(32), 1, 100, 0, 2, 100)


def test_size_steps(ref):
    """Best-fit sub-cell split: three size levels, the library halved between steps (CPUPhotomosaicGenerator.cpp:95-99)."""
    main, lib = _inputs(51, 256, 384, 48, 64)
    assert _compare(ref, main, lib, ref.CellShape.square(64), 1, 100, 2, 2, 200) > 30


@pytest.mark.parametrize("scheme,detail", [(1, 50), (2, 50), (4, 100)])
def test_colour_scheme_variants_tie_order(ref, scheme, detail):
    """V > 1: library-major, variant-minor loop order with strict < (CPUPhotomosaicGenerator.cpp:137-169), including the
    aliased-buffer case at detail 100 %."""
    main, lib = _inputs(61 + scheme, 130, 170, 24, 32)
    _compare(ref, main, lib, ref.CellShape.square(32), 0, detail, 0, 1, 50, scheme=scheme)


def test_exact_ties_keep_the_lowest_index(ref):
    main, lib = _inputs(84, 96, 128, 6, 32)
    lib = np.concatenate([lib, lib, lib])
    _compare(ref, main, lib, ref.CellShape.square(32), 2, 100, 0, 0, 0)
    _compare(ref, main, lib, ref.CellShape.square(32), 0, 100, 0, 2, 0)


def test_degenerate_inputs(ref):
    main, lib = _inputs(81, 20, 27, 10, 32)  # image smaller than a cell
    _compare(ref, main, lib, ref.CellShape.square(32), 2, 100, 0, 1, 10)
    main, lib = _inputs(82, 100, 140, 1, 32)  # one library image, repeat range larger than the grid
    _compare(ref, main, lib, ref.CellShape.square(32), 1, 100, 0, 50, 100000)
    main, lib = _inputs(83, 100, 140, 20, 32)  # one detail pixel per cell
    _compare(ref, main, lib, ref.CellShape.square(32), 0, 3, 0, 2, 10)


# ---------------------------------------------------------------- GridGenerator::getGridState from the reference's object code

def _grid_cases(o):
    from mosaicmagnifique_b200 import synthetic
    tri = o.CellShape.from_mask(synthetic.triangle_mask(64))
    tri.row_spacing = tri.alt_row_spacing = 64
    tri.col_spacing = tri.alt_col_spacing = 32
    tri.alt_col_flip_v = True
    tri.alt_row_flip_h = True
    hx = o.CellShape.from_mask(synthetic.hexagon_mask(128))
    hx.row_spacing = hx.alt_row_spacing = 96
    hx.col_spacing = hx.alt_col_spacing = 110
    hx.alt_row_offset = 55
    #       shape                       detail steps  h    w   seed
    return [(o.CellShape.square(64),    100,   2,    256, 384, 51),
            (o.CellShape.square(64),    50,    2,    300, 410, 52),
            (o.CellShape.square(32),    100,   0,    130, 170, 53),
            (tri,                       50,    1,    230, 310, 54),
            (hx.resized(64),            100,   2,    250, 330, 55),
            (o.CellShape.square(40),    33,    1,    200, 260, 56),
            (o.CellShape.square(64),    100,   1,    20,  27,  57)]


def test_grid_state_matches_reference_object_code(ref, oracle):
    """The entropy-split grid state: oracle.grid_state == the reference's GridGenerator.cpp compiled unmodified,
    for square / flipped / offset shapes, 0-2 size steps, integer and fractional detail, an image smaller than a cell."""
    from mosaicmagnifique_b200 import synthetic
    for shape, detail, steps, h, w, seed in _grid_cases(oracle):
        main = synthetic.make_main_image(h, w, seed, block=32)
        group = oracle.CellGroup.make(shape, detail, steps)
        want = oracle.reference_grid_state(group, main)
        got = oracle.grid_state(group, main)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        # without a main image nothing splits (GridGenerator.cpp:158: the entropy rule needs the image)
        want0 = oracle.reference_grid_state(group, None, h, w)
        got0 = oracle.grid_state(group, None, h, w)
        assert len(got0) == len(want0) == 1 and np.array_equal(got0[0], want0[0])
    assert any(len(oracle.grid_state(oracle.CellGroup.make(s, d, st), synthetic.make_main_image(h, w, sd, block=32))) > 1
               for s, d, st, h, w, sd in _grid_cases(oracle) if st > 0), "no case exercised a split"
