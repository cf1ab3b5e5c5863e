"""File formats either side of the path (.mil library, .mcs cell shape): round trips like the reference's
ImageLibrary.SaveAndLoad / CellShape.SaveAndLoad tests (test/tst_ImageLibrary.h, test/tst_CellShape.h), and the
reference's own Cells/*.mcs fixtures when the checkout is present."""
import glob
import os

import numpy as np
import pytest

pytest.importorskip("cv2")


def test_mil_round_trip(tmp_path):
    from mosaicmagnifique_b200 import formats, synthetic
    lib = synthetic.make_library(9, 24, 3)
    names = ["img %d é" % i for i in range(9)]
    p = str(tmp_path / "lib.mil")
    formats.save_mil(p, lib, names)
    got, got_names, size = formats.load_mil(p)
    assert size == 24 and got_names == names
    assert np.array_equal(got, lib)
    raw = open(p, "rb").read()
    assert raw[:4] == bytes.fromhex("ADBE2480") and raw[4:8] == (6).to_bytes(4, "big")
    with pytest.raises(ValueError):
        open(p, "wb").write(b"\x00" * 16)
        formats.load_mil(p)


def test_mcs_round_trip(tmp_path):
    from mosaicmagnifique_b200 import formats, load_mcs, synthetic
    f = {"name": "Hex", "mask": synthetic.hexagon_mask(64), "row_spacing": 48, "col_spacing": 55, "alt_row_spacing": 48,
         "alt_col_spacing": 55, "alt_row_offset": 27, "alt_col_offset": 0, "alt_col_flip_h": False, "alt_col_flip_v": True,
         "alt_row_flip_h": True, "alt_row_flip_v": False}
    p = str(tmp_path / "hex.mcs")
    formats.save_mcs(p, f)
    g = formats.load_mcs(p)
    for k, v in f.items():
        assert np.array_equal(g[k], v), k
    s = load_mcs(p)
    assert (s.getSize(), s.rowSpacing, s.colSpacing, s.alternateRowOffset) == (64, 48, 55, 27)
    assert s.alternateColFlipVertical and s.alternateRowFlipHorizontal and not s.alternateColFlipHorizontal


def test_reference_cell_fixtures(oracle):
    files = sorted(glob.glob("/root/reference/Cells/*.mcs"))
    if not files:
        pytest.skip("reference checkout not present")
    from mosaicmagnifique_b200 import formats
    assert len(files) == 10
    for p in files:
        f = formats.load_mcs(p)
        o = oracle.load_mcs(p)
        assert f["version"] == 8 and f["mask"].shape == (512, 512)
        assert np.array_equal(np.where(f["mask"] > 127, 255, 0), o.mask)
        assert [f[k] for k in ("row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset", "alt_col_offset")] == o.params()[1:7]
    hexa = formats.load_mcs("/root/reference/Cells/Hexagon.mcs")
    assert (hexa["row_spacing"], hexa["col_spacing"], hexa["alt_row_offset"]) == (385, 440, 220)
