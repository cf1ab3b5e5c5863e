"""Drop-in on the GPU: integration/B200PhotomosaicGenerator (the binding INTEGRATION.md asks a maintainer to add) behind the
reference's REAL generator API -- PhotomosaicGeneratorBase.h / .cpp compiled unmodified -- next to the reference's own CPU
back-end, on the same inputs, in one process. This is the comparison the reference makes between its CPU and CUDA
generators (test/tst_CUDAGenerator.h:197): the best-fit grids must be equal (outside the tie band) and, the grids being
equal, buildPhotomosaic -- the reference's own code on either back-end's fits -- must give the same image."""
import numpy as np
import pytest

from tests.helpers import dropin
from tests.helpers.parity import check_grid

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["square_ciede2000", "triangle_euclid_detail50", "size_steps_cie76"])
def test_b200_backend_equals_reference_cpu_backend(oracle, case):
    if not (oracle.reference_generator_available() and dropin.available()):
        pytest.skip("oracle/_ref/libdropin_b200.so not present / not loadable")
    from mosaicmagnifique_b200 import synthetic
    tri = oracle.CellShape.from_mask(synthetic.triangle_mask(64))
    tri.row_spacing = tri.alt_row_spacing = 64
    tri.col_spacing = tri.alt_col_spacing = 32
    tri.alt_col_flip_v = True
    tri.alt_row_flip_h = True
    #          seed  h    w   lib cell shape                         diff detail steps rr ra
    seed, h, w, n_lib, cell, shape, diff, detail, steps, rr, ra = {
        "square_ciede2000": (601, 200, 300, 60, 32, oracle.CellShape.square(32), 2, 100, 0, 3, 10000),
        "triangle_euclid_detail50": (602, 230, 310, 50, 32, tri.resized(32), 0, 50, 0, 2, 300),
        "size_steps_cie76": (603, 256, 384, 48, 64, oracle.CellShape.square(64), 1, 100, 1, 2, 200)}[case]
    main = synthetic.make_main_image(h, w, seed, block=32)
    lib = synthetic.make_library(n_lib, cell, seed + 1)
    group = oracle.CellGroup.make(shape, detail, steps)
    states = oracle.reference_grid_state(group, main)
    bg = (9, 8, 7, 0)
    rc0, cpu_grids, cpu_mosaic = dropin.run(oracle, 0, main, lib, group, states, diff, 0, rr, ra, background=bg)
    rc1, b200_grids, b200_mosaic = dropin.run(oracle, 1, main, lib, group, states, diff, 0, rr, ra, background=bg)
    assert rc0 == 0 and rc1 == 0
    same = all(np.array_equal(a, b) for a, b in zip(cpu_grids, b200_grids))
    if same:
        assert np.array_equal(cpu_mosaic, b200_mosaic)
    else:
        want = oracle.generate(main, lib, group, states, diff, 0, rr, ra, want_D=True)
        for s in range(len(states)):
            _, ties, bad = check_grid(want[s].D, states[s], b200_grids[s], rr, ra)
            assert not bad, bad[:3]
    print("%s: grids %s" % (case, "identical, mosaics identical" if same else "equal outside the tie band"))


def test_cancel_slot_is_honoured_by_the_b200_backend(oracle):
    """the binding ignored cancel(). The user cancels from the progress dialog, i.e. the cancel() slot runs inside
    a progress(int) emission: both back-ends must then return false from generateBestFits() -- the reference's CPU back-end at its
    next per-cell check, the B200 back-end by draining its kernel."""
    if not (oracle.reference_generator_available() and dropin.available()):
        pytest.skip("oracle/_ref/libdropin_b200.so not present / not loadable")
    from mosaicmagnifique_b200 import synthetic
    main = synthetic.make_main_image(2160, 3840, 931)
    lib = synthetic.make_library(2500, 64, 932)
    group = oracle.CellGroup.make(oracle.CellShape.square(64), 100, 0)
    states = oracle.reference_grid_state(group, main)
    try:
        for backend in (0, 1):
            dropin.progress_clear(oracle)
            dropin.cancel_after(oracle, 2)
            rc, _, _ = dropin.run(oracle, backend, main, lib, group, states, 2, 0, 2, 300, want_mosaic=False)
            seen = dropin.progress_values(oracle)
            assert rc == 1, "backend %d ignored cancel()" % backend
            assert 2 <= len(seen) < states[0].size and seen[-1] < states[0].size, (backend, len(seen))
    finally:
        dropin.cancel_after(oracle, 0)
    # and without a cancel the B200 back-end still completes and reports the reference's final progress value
    dropin.progress_clear(oracle)
    rc, grids, _ = dropin.run(oracle, 1, main, lib, group, states, 2, 0, 2, 300, want_mosaic=False)
    seen = dropin.progress_values(oracle)
    assert rc == 0 and seen[-1] == states[0].size and (grids[0][states[0] >= 0] >= 0).all()
