"""The product's C++ host model of CellShape / CellGroup (csrc/host_model.cpp, reached without a device through
mosaic_host_cell_group_cell) against the reference's OWN CellShape.cpp / CellGroup.cpp object code, starting from shape FILES the way
the application does: loadFromFile -> resized(cell size) -> CellGroup(detail, size steps) -> getCell(step, detail).

Includes the case ADVICE r1 raised: loadFromFile does not threshold the decoded mask (CellShape.cpp:405-410), so a hand-made .mcs with
grey values is "active where non-zero" at its stored size, while every RESIZED mask is binarised at 127 (CellShape::resized ->
setCellMask). mosaic_set_cell_group_ex(..., mask_as_stored = 1) / CellShape(mask, as_stored=True) reproduce that."""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CELLS = os.path.join(HERE, "golden", "cells")


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.reference_generator_available():
        pytest.skip("oracle/_ref/libref_core.so (reference object code) is not built")
    if not hasattr(oracle._ref(), "ref_mcs_group_cell"):
        pytest.skip("prebuilt libref_core.so predates ref_mcs_group_cell")
    return oracle


def product_cell(path, cell_size, detail, steps, step, want_detail, as_stored=True):
    from mosaicmagnifique_b200 import capi
    from mosaicmagnifique_b200._capi import CellShapeC
    L = capi()
    c = CellShapeC()
    assert L.mosaic_mcs_load(path.encode(), ctypes.byref(c), None, 0, None, 0) == 0
    mask = np.empty((c.size, c.size), np.uint8)
    assert L.mosaic_mcs_load(path.encode(), ctypes.byref(c), mask.ctypes.data, mask.size, None, 0) == 0
    out = CellShapeC()
    rc = L.mosaic_host_cell_group_cell(ctypes.byref(c), mask.ctypes.data, int(as_stored), cell_size, detail, steps, step, int(want_detail),
                                       ctypes.byref(out), None, 0)
    assert rc == 0, rc
    m = np.empty((out.size, out.size), np.uint8)
    assert L.mosaic_host_cell_group_cell(ctypes.byref(c), mask.ctypes.data, int(as_stored), cell_size, detail, steps, step, int(want_detail),
                                         ctypes.byref(out), m.ctypes.data, m.size) == 0
    return [getattr(out, n) for n, _ in CellShapeC._fields_], m


@pytest.mark.parametrize("name", ["Hexagon", "Puzzle", "IsocelesTriangle-45deg", "YinAndYang"])
@pytest.mark.parametrize("cell_size,detail,steps", [(0, 100, 0), (128, 50, 0), (128, 100, 2), (100, 33, 1), (64, 50, 2)])
def test_shipped_shapes_group_cells_equal_reference(ref, name, cell_size, detail, steps):
    path = os.path.join(CELLS, name + ".mcs")
    for step in range(steps + 1):
        for want_detail in (False, True):
            want, want_m4 = ref.reference_mcs_group_cell(path, cell_size, detail, steps, step, want_detail)
            params, mask = product_cell(path, cell_size, detail, steps, step, want_detail)
            assert params == want.params(), (step, want_detail)
            assert np.array_equal(mask, want_m4[0]), (step, want_detail, int((mask != want_m4[0]).sum()))


def _grey_mcs(tmp_path):
    """a hand-made shape file whose mask holds grey values: a radial ramp 0..255 with a hole"""
    from mosaicmagnifique_b200 import capi
    from mosaicmagnifique_b200._capi import CellShapeC
    S = 96
    y, x = np.mgrid[0:S, 0:S].astype(np.float64)
    r = np.hypot(x - 47.5, y - 47.5)
    mask = np.clip(255 - r * 5.2, 0, 255).astype(np.uint8)
    mask[40:56, 40:56] = 0
    mask[10:14, 10:80] = 1      # barely non-zero: active in the reference at the stored size, gone after any resize
    c = CellShapeC(S, 90, 84, 90, 84, 42, 0, 0, 1, 0, 0)
    path = str(tmp_path / "grey.mcs")
    assert capi().mosaic_mcs_save(path.encode(), ctypes.byref(c), mask.ctypes.data, b"grey") == 0
    return path, mask


@pytest.mark.parametrize("cell_size,detail,steps", [(0, 100, 0), (0, 50, 1), (48, 100, 1), (64, 75, 0)])
def test_grey_mask_file_follows_load_semantics(ref, tmp_path, cell_size, detail, steps):
    path, stored = _grey_mcs(tmp_path)
    loaded, m4 = ref.reference_load_mcs(path)
    assert np.array_equal(m4[0], stored) and len(np.unique(stored)) > 50     # the reference really keeps the grey values
    for step in range(steps + 1):
        for want_detail in (False, True):
            want, want_m4 = ref.reference_mcs_group_cell(path, cell_size, detail, steps, step, want_detail)
            params, mask = product_cell(path, cell_size, detail, steps, step, want_detail)
            assert params == want.params()
            assert np.array_equal(mask, want_m4[0]), (step, want_detail, int((mask != want_m4[0]).sum()))
    if cell_size == 0 and detail == 100:
        # the un-resized cell is the stored mask itself: 'active' = non-zero, which the thresholding entry point would not give
        _, kept = product_cell(path, 0, 100, 0, 0, False, as_stored=True)
        _, binarised = product_cell(path, 0, 100, 0, 0, False, as_stored=False)
        assert np.array_equal(kept, stored)
        assert int((kept != 0).sum()) > int((binarised != 0).sum())


def test_python_mirror_follows_load_semantics(ref, tmp_path):
    from mosaicmagnifique_b200 import CellGroup, load_mcs
    path, stored = _grey_mcs(tmp_path)
    shape = load_mcs(path)
    assert np.array_equal(shape.getCellMask(), stored)
    for cell_size, detail, steps in ((0, 100, 1), (48, 50, 1)):
        cg = CellGroup()
        cg.setCellShape(shape.resized(cell_size) if cell_size else shape)
        cg.setDetail(detail)
        cg.setSizeSteps(steps)
        for step in range(steps + 1):
            for want_detail in (False, True):
                want, want_m4 = ref.reference_mcs_group_cell(path, cell_size, detail, steps, step, want_detail)
                got = cg.getCell(step, want_detail)
                assert np.array_equal(got.getCellMask(), want_m4[0]), (cell_size, detail, step, want_detail)
                assert [got.rowSpacing, got.colSpacing, got.alternateRowOffset] == [want.row_spacing, want.col_spacing, want.alt_row_offset]
