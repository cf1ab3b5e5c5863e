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


# ---------------------------------------------------------------- getCellAt / buildPhotomosaic / getMaxProgress

@pytest.mark.parametrize("diff,detail,scheme", [(2, 100, 0), (0, 50, 0), (1, 30, 0), (0, 100, 2), (2, 50, 1)])
def test_get_cell_at_matches_reference_object_code(ref, diff, detail, scheme):
    """getCellAt (PhotomosaicGeneratorBase.cpp:293-329) from the reference's object code on the reference's own
    preprocessMainImage(): cells (bit for bit, f32) and detail-space bounds equal oracle.extract_cells, for interior,
    clipped and fully padded cells, fractional detail, and the aliased variants of quirk Q1 (scheme TRIADIC at detail 100)."""
    from mosaicmagnifique_b200 import synthetic
    o = ref
    main = synthetic.make_main_image(150, 210, 91 + diff, block=32)
    lib = synthetic.make_library(4, 32, 92)
    sh = o.CellShape.from_mask(synthetic.triangle_mask(64))
    sh.row_spacing = sh.alt_row_spacing = 64
    sh.col_spacing = sh.alt_col_spacing = 32
    sh.alt_col_flip_v = True
    shape = sh.resized(32)
    group = o.CellGroup.make(shape, detail, 0)
    state = o.grid_state(group, main)[0]
    mains = [o.to_working_space(v, diff) for v in o.colour_scheme_variants(main, scheme)]
    cells, bounds, _flips, coords = o.extract_cells(mains, group, 0, state)
    g = o.ReferenceGenerator(main, lib, group, diff, scheme, 0, 0)
    try:
        for i in range(len(cells)):
            c, b = g.get_cell_at(0, int(coords[i][0]), int(coords[i][1]), len(mains))
            assert b == tuple(int(v) for v in bounds[i])
            assert np.array_equal(c, cells[i]), "cell %s differs" % (coords[i],)
    finally:
        g.close()
    assert len(cells) > 20


def test_build_photomosaic_matches_reference_object_code(ref):
    """buildPhotomosaic (PhotomosaicGeneratorBase.cpp:110-207) from the reference's object code on its own best fits: the
    BGRA mosaic equals oracle.build_photomosaic byte for byte (two size levels, non-square cells, a background colour)."""
    from mosaicmagnifique_b200 import synthetic
    o = ref
    main = synthetic.make_main_image(200, 260, 95, block=32)
    lib = synthetic.make_library(30, 64, 96)
    hx = o.CellShape.from_mask(synthetic.hexagon_mask(128))
    hx.row_spacing = hx.alt_row_spacing = 96
    hx.col_spacing = hx.alt_col_spacing = 110
    hx.alt_row_offset = 55
    for shape, steps in ((o.CellShape.square(64), 1), (hx.resized(64), 1), (o.CellShape.square(64), 0)):
        group = o.CellGroup.make(shape, 100, steps)
        states = o.grid_state(group, main)
        g = o.ReferenceGenerator(main, lib, group, 1, 0, 2, 100)
        try:
            grids, _, max_progress = g.generate(states)
            mosaic = g.build_photomosaic((10, 20, 30, 0))
        finally:
            g.close()
        want = o.build_photomosaic(main.shape, lib, group, grids, background=(10, 20, 30, 0))
        assert np.array_equal(mosaic, want)
        # getMaxProgress (PhotomosaicGeneratorBase.cpp:210-214): 4^(S-1) * S * cols * rows of step 0
        S = len(states)
        assert max_progress == 4 ** (S - 1) * S * states[0].shape[0] * states[0].shape[1]


def test_committed_golden_is_reproduced_by_the_reference(ref):
    """tests/golden/generator_golden.npz (the fixtures the CUDA path is checked against on the GPU box) are oracle outputs;
    the reference's own object code must reproduce their grids from the recorded inputs."""
    from tests.test_oracle_pipeline import _golden, golden_case
    G = _golden()
    for name in G["names"]:
        name = str(name)
        shape, diff, detail, steps, rr, ra, scheme = golden_case(ref, G, name)
        group = ref.CellGroup.make(shape, detail, steps)
        states = [G["%s/state%d" % (name, s)] for s in range(steps + 1)]
        got_states = ref.reference_grid_state(group, G[name + "/main"])
        assert len(got_states) == len(states) and all(np.array_equal(a, b) for a, b in zip(got_states, states))
        grids, _ = ref.reference_generate(G[name + "/main"], G[name + "/lib"], group, states, diff, scheme, rr, ra)
        for s in range(steps + 1):
            assert np.array_equal(grids[s], G["%s/grid%d" % (name, s)])


# ---------------------------------------------------------------- randomised configurations

def _random_config(o, rng, it):
    from mosaicmagnifique_b200 import synthetic
    cell = int(rng.choice([16, 32, 64]))
    kind = int(rng.integers(0, 3))
    sh = (o.CellShape.square(cell) if kind == 0 else
          o.CellShape.from_mask(synthetic.triangle_mask(cell) if kind == 1 else synthetic.hexagon_mask(cell)))
    sh.row_spacing = int(rng.integers(cell // 2, cell + 1))
    sh.col_spacing = int(rng.integers(cell // 2, cell + 1))
    sh.alt_row_spacing = sh.row_spacing if rng.random() < 0.6 else int(rng.integers(cell // 2, cell + 1))
    sh.alt_col_spacing = sh.col_spacing if rng.random() < 0.6 else int(rng.integers(cell // 2, cell + 1))
    sh.alt_row_offset = int(rng.integers(0, cell // 2)) if rng.random() < 0.5 else 0
    sh.alt_col_offset = int(rng.integers(0, cell // 2)) if rng.random() < 0.3 else 0
    sh.alt_col_flip_h, sh.alt_col_flip_v, sh.alt_row_flip_h, sh.alt_row_flip_v = (bool(rng.random() < 0.3) for _ in range(4))
    steps = int(rng.integers(0, 3)) if cell >= 32 else int(rng.integers(0, 2))
    detail = int(rng.choice([100, 50, 25])) if steps else int(rng.choice([100, 75, 50, 33, 25, 10]))
    if steps and (cell >> steps) * detail // 100 < 1:
        detail = 100
    cfg = dict(shape=sh, steps=steps, detail=detail, diff=int(rng.integers(0, 3)), scheme=int(rng.choice([0, 0, 0, 1, 2, 4])),
               rr=int(rng.integers(0, 4)), ra=int(rng.choice([0, 10, 500, 100000])))
    h, w = int(rng.integers(40, 260)), int(rng.integers(40, 300))
    cfg["main"] = synthetic.make_main_image(h, w, 1000 + it, block=int(rng.choice([16, 32])))
    cfg["lib"] = synthetic.make_library(int(rng.integers(1, 30)), cell, 2000 + it)
    return cfg


def test_randomised_configurations_against_reference_object_code(ref):
    """60 seeded random configurations -- cell shape, independent alternate spacings / offsets / flips, 0-2 size steps,
    integer and fractional detail, all colour differences, colour schemes, repeat settings, image and library sizes: grid
    state (oracle AND the product's host model) and best-fit grid must equal the reference's object code every time.
    (540 further configurations were run once while writing this test; none differed.)"""
    import ctypes
    from mosaicmagnifique_b200 import capi
    from mosaicmagnifique_b200._capi import CellShapeC
    L = capi()
    rng = np.random.default_rng(20261017)
    n_generated = 0
    for it in range(60):
        c = _random_config(ref, rng, it)
        sh, main = c["shape"], c["main"]
        group = ref.CellGroup.make(sh, c["detail"], c["steps"])
        want_states = ref.reference_grid_state(group, main)
        got_states = ref.grid_state(group, main)
        assert len(want_states) == len(got_states) and all(np.array_equal(a, b) for a, b in zip(want_states, got_states)), it
        # the product's host model (libmosaic_b200.so, no device needed)
        cs = CellShapeC(*sh.params())
        n_steps = ctypes.c_int()
        rows, cols = (ctypes.c_int * 8)(), (ctypes.c_int * 8)()
        out = np.empty(1 << 18, np.int64)
        rc = L.mosaic_host_grid_state(ctypes.byref(cs), sh.mask.ctypes.data, 0, c["detail"], c["steps"], main.ctypes.data, main.shape[0],
                                      main.shape[1], main.strides[0], 8, ctypes.byref(n_steps), rows, cols, out.ctypes.data, out.size)
        assert rc == 0 and n_steps.value == len(want_states), it
        off = 0
        for s, wst in enumerate(want_states):
            assert np.array_equal(out[off:off + wst.size].reshape(wst.shape), wst), (it, s)
            off += wst.size
        # the halved library must meet the detail mask size at every generated step, else the reference reads out of range
        lib_ds, q4 = group.detail_cells[0].size, True
        for s in range(1, len(want_states)):
            lib_ds = int(round(0.5 * lib_ds))
            q4 = q4 and lib_ds == group.detail_cells[s].size
        if not q4:
            continue
        want, _ = ref.reference_generate(main, c["lib"], group, want_states, c["diff"], c["scheme"], c["rr"], c["ra"])
        got = ref.generate(main, c["lib"], group, want_states, c["diff"], c["scheme"], c["rr"], c["ra"], want_D=False)
        assert all(np.array_equal(g.grid, w_) for g, w_ in zip(got, want)), it
        n_generated += 1
    assert n_generated >= 40
