"""CPU-side checks of the C ABI: the library loads, exports exactly what include/mosaic_b200.h declares, refuses to
run without a GPU (no CPU fallback), and its host model (geometry, mask resize) agrees with the oracle / OpenCV."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from mosaicmagnifique_b200 import capi
    return capi()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mosaic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mosaic_[a-z0-9_]+)\s*\(", text)) - {"mosaic_progress_fn"})


def test_library_exports_every_declared_symbol(L):
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "declared in mosaic_b200.h but not exported: " + s
    assert sorted(L._signatures) == syms, "python binding and header disagree"
    assert b"sm_100a" in L.mosaic_version()


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert L.mosaic_create(0, ctypes.byref(h)) == -2  # MOSAIC_ERR_CUDA
    assert not h.value
    out = np.zeros(8)
    assert L.mosaic_kernel_microbench(0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 8) == -2


def _shapes(oracle):
    sq = oracle.CellShape.square(128)
    hx = oracle.CellShape.square(128)
    hx.row_spacing, hx.alt_row_spacing, hx.col_spacing, hx.alt_col_spacing, hx.alt_row_offset = 96, 96, 110, 110, 55
    odd = oracle.CellShape.square(50)
    odd.row_spacing, odd.alt_row_spacing, odd.col_spacing, odd.alt_col_spacing = 40, 25, 33, 47
    odd.alt_row_offset, odd.alt_col_offset = 13, 7
    odd.alt_col_flip_h, odd.alt_row_flip_v, odd.alt_row_flip_h = True, True, True
    return [sq, hx, odd]


def test_geometry_matches_oracle(L, oracle):
    """GridUtility::calculateGridSize / getRectAt / getFlipStateAt incl. negative coordinates (C++ truncation)."""
    from mosaicmagnifique_b200._capi import CellShapeC
    for sh in _shapes(oracle):
        c = CellShapeC(*sh.params())
        for (w, h) in ((1920, 1080), (333, 517), (50, 50)):
            gx, gy = ctypes.c_int(), ctypes.c_int()
            L.mosaic_grid_size(ctypes.byref(c), w, h, 2, ctypes.byref(gx), ctypes.byref(gy))
            assert (gx.value, gy.value) == oracle.grid_size(sh, w, h)
        for y in range(-3, 9):
            for x in range(-3, 9):
                r = (ctypes.c_int * 4)()
                L.mosaic_rect_at(ctypes.byref(c), x, y, r)
                assert tuple(r) == oracle.rect_at(sh, x, y)
                assert L.mosaic_flip_at(ctypes.byref(c), x, y) == oracle.flip_at(sh, x, y)


def test_oracle_geometry_matches_compiled_reference(oracle):
    so = os.path.join(ROOT, "oracle", "_ref", "libref_core.so")
    if not os.path.exists(so):
        pytest.skip("libref_core.so not built")
    R = ctypes.CDLL(so)
    for sh in _shapes(oracle):
        p = (ctypes.c_int * 11)(*sh.params())
        for y in range(-3, 9):
            for x in range(-3, 9):
                r = (ctypes.c_int * 4)()
                R.ref_rect_at(p, x, y, r)
                assert tuple(r) == oracle.rect_at(sh, x, y)
                assert R.ref_flip_at(p, x, y) == oracle.flip_at(sh, x, y)
        gx, gy = ctypes.c_int(), ctypes.c_int()
        R.ref_grid_size(p, 1000, 777, 2, ctypes.byref(gx), ctypes.byref(gy))
        assert (gx.value, gy.value) == oracle.grid_size(sh, 1000, 777)


def _bound_sets():
    """Bound lists of the kind GridGenerator::getGridState feeds to mergeBounds (cell rects of split cells: touching,
    overlapping, nested, on shifted rows) plus random small rects."""
    rng = np.random.default_rng(77)
    sets = [[], [[0, 0, 10, 10]], [[0, 0, 10, 10], [10, 0, 10, 10]], [[0, 0, 10, 10], [0, 10, 10, 10], [0, 21, 10, 10]],
            [[0, 0, 20, 20], [5, 5, 3, 3]], [[5, 5, 3, 3], [0, 0, 20, 20], [20, 0, 20, 20]]]
    for _ in range(150):
        n = int(rng.integers(2, 14))
        cell = int(rng.integers(2, 9))
        rects = []
        for _ in range(n):
            if rng.random() < 0.7:  # grid-aligned cells, sometimes with the half-cell offset of alternate rows
                x = int(rng.integers(0, 6)) * cell + (cell // 2 if rng.random() < 0.3 else 0)
                y = int(rng.integers(0, 6)) * cell
                rects.append([x, y, cell, cell])
            else:
                rects.append([int(v) for v in (rng.integers(-5, 30), rng.integers(-5, 30), rng.integers(1, 15), rng.integers(1, 15))])
        sets.append(rects)
    return sets


def test_merge_bounds_matches_compiled_reference(L, oracle):
    """GridBounds::mergeBounds: the product's host model and the oracle against the reference's own GridBounds.cpp compiled
    unmodified (oracle/_ref/libref_core.so): same bounds in the same order."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_core.so")
    R = ctypes.CDLL(so) if os.path.exists(so) else None
    for rects in _bound_sets():
        flat = np.asarray(rects, np.int32).reshape(-1)
        out = np.zeros(4 * max(len(rects), 1), np.int32)
        m = L.mosaic_host_merge_bounds(flat.ctypes.data, len(rects), out.ctypes.data, len(rects))
        mine = out[:4 * m].reshape(-1, 4).tolist()
        assert mine == [list(r) for r in oracle._merge_bounds([list(r) for r in rects])]
        if R is not None:
            ref = np.zeros_like(out)
            k = R.ref_merge_bounds(flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(rects),
                                   ref.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(rects))
            assert k == m and ref[:4 * k].reshape(-1, 4).tolist() == mine
    if R is None:
        pytest.skip("libref_core.so not built: checked against the oracle only")


@pytest.mark.parametrize("src,dst,cn", [(512, 128, 1), (512, 100, 1), (128, 64, 1), (128, 25, 3), (64, 48, 3), (96, 32, 3),
                                        (100, 37, 1), (64, 64, 1)])
def test_host_resize_area_matches_opencv(L, src, dst, cn):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(src + dst)
    a = rng.integers(0, 256, (src, src, cn), dtype=np.uint8)
    if cn == 1:
        a = (a > 100).astype(np.uint8) * 255
    out = np.empty((dst, dst, cn), np.uint8)
    assert L.mosaic_host_resize_area_u8(a.ctypes.data, src, src, cn, out.ctypes.data, dst, dst) == 0
    ref = cv2.resize(a, (dst, dst), interpolation=cv2.INTER_AREA).reshape(dst, dst, cn)
    assert np.array_equal(out, ref)


def test_host_resize_area_non_square(L):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (41, 128, 3), dtype=np.uint8)
    out = np.empty((20, 64, 3), np.uint8)
    assert L.mosaic_host_resize_area_u8(a.ctypes.data, 41, 128, 3, out.ctypes.data, 20, 64) == 0
    assert np.array_equal(out, cv2.resize(a, (64, 20), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("sh,sw,dh,dw,cn", [(50, 50, 128, 128, 3), (100, 100, 128, 128, 3), (127, 127, 128, 128, 1), (64, 64, 128, 128, 3),
                                            (37, 37, 64, 64, 1), (3, 3, 10, 10, 3), (1, 1, 5, 5, 1), (20, 31, 45, 77, 3),
                                            (255, 255, 301, 301, 3), (9, 9, 27, 27, 1)])
def test_host_resize_cubic_matches_opencv(L, oracle, sh, sw, dh, dw, cn):
    """INTER_CUBIC (ImageUtility::resizeImage when growing, ImageUtility.cpp:50-51) bit-exact against OpenCV's own code
    (IPP off: oracle.resize_cubic_opencv says why), including the float-SIMD / fixed-point split of the vertical pass."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(sh * 1000 + dw)
    a = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
    out = np.empty((dh, dw, cn), np.uint8)
    assert L.mosaic_host_resize_cubic_u8(a.ctypes.data, sh, sw, cn, out.ctypes.data, dh, dw) == 0
    ref = oracle.resize_cubic_opencv(a, dh, dw).reshape(dh, dw, cn)
    assert np.array_equal(out, ref)


def test_cell_shape_growth_matches_oracle(oracle):
    """CellShape::resized to a LARGER size (CellShape.cpp:281-312 -> INTER_CUBIC -> threshold) on the library's host model."""
    pytest.importorskip("cv2")
    from mosaicmagnifique_b200 import CellShape
    yy, xx = np.mgrid[0:48, 0:48]
    mask = (((yy - 23.5) ** 2 + (xx - 20.0) ** 2) < 19.0 ** 2).astype(np.uint8) * 255
    for new in (64, 100, 129):
        ours = CellShape(mask)
        ours.rowSpacing, ours.alternateRowOffset = 40, 13
        r = ours.resized(new)
        o = oracle.CellShape.from_mask(mask)
        o.row_spacing, o.alt_row_offset = 40, 13
        ro = o.resized(new)
        assert np.array_equal(r.getCellMask(), ro.mask)
        assert (r.rowSpacing, r.colSpacing, r.alternateRowOffset) == (ro.row_spacing, ro.col_spacing, ro.alt_row_offset)


def test_image_library_container_round_trip(tmp_path):
    """ImageLibrary save / load / operator== / removeAtIndex / clear (ImageLibrary.cpp:15-39, 100-236): host bookkeeping,
    no device needed (addImage's crop + resize is GPU work, tests/test_gpu_kernels.py)."""
    from mosaicmagnifique_b200 import ImageLibrary
    from mosaicmagnifique_b200.formats import save_mil
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (5, 16, 16, 3), dtype=np.uint8)
    path = str(tmp_path / "a.mil")
    save_mil(path, imgs, ["n%d" % i for i in range(5)])
    lib = ImageLibrary(99)
    lib.loadFromFile(path)
    assert lib.getImageSize() == 16 and lib.getNames() == ["n%d" % i for i in range(5)]
    assert np.array_equal(lib.asArray(), imgs)
    path2 = str(tmp_path / "b.mil")
    lib.saveToFile(path2)
    lib2 = ImageLibrary(1)
    lib2.loadFromFile(path2)
    assert lib == lib2
    lib2.removeAtIndex(1)
    assert lib != lib2 and lib2.getNames() == ["n0", "n2", "n3", "n4"]
    lib2.clear()
    assert lib2.asArray().shape == (0, 16, 16, 3)
    with pytest.raises(ValueError):
        lib.saveToFile("")


def test_cpp_mirror_compiles_and_links(tmp_path):
    """include/mosaic_b200.hpp (the C++ host mirror of the reference classes) builds against the C ABI; the program
    itself needs a GPU (exit 3 = clean 'no device' error, 0 = ran)."""
    import subprocess
    exe = str(tmp_path / "hpp_check")
    libdir = os.path.join(ROOT, "mosaicmagnifique_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "helpers", "hpp_compile_check.cpp"),
                           "-L" + libdir, "-lmosaic_b200", "-Wl,-rpath," + libdir])
    rc = subprocess.call([exe])
    import torch
    assert rc == (0 if torch.cuda.is_available() else 3)


def test_cpp_mirror_container_methods(tmp_path):
    """CellShape::loadFromFile / saveToFile of include/mosaic_b200.hpp (host only): write a .mcs with the Python writer, load and
    re-save it through the C++ mirror, read the result back with the Python reader."""
    import subprocess
    from mosaicmagnifique_b200 import formats, synthetic
    src, dst = str(tmp_path / "in.mcs"), str(tmp_path / "out.mcs")
    fields = {"name": "hex \u00e4", "mask": synthetic.hexagon_mask(96), "row_spacing": 72, "col_spacing": 82, "alt_row_spacing": 72,
              "alt_col_spacing": 82, "alt_row_offset": 41, "alt_col_offset": 0, "alt_col_flip_h": False, "alt_col_flip_v": True,
              "alt_row_flip_h": False, "alt_row_flip_v": False}
    formats.save_mcs(src, fields)
    exe = str(tmp_path / "hpp_containers_check")
    libdir = os.path.join(ROOT, "mosaicmagnifique_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "helpers", "hpp_containers_check.cpp"),
                           "-L" + libdir, "-lmosaic_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, src, dst, str(tmp_path / "missing.mcs")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[-3:] == ["96", "72", "82"]
    back = formats.load_mcs(dst)
    assert back["name"] == fields["name"] and np.array_equal(back["mask"], fields["mask"])
    assert all(back[k] == fields[k] for k in fields if k not in ("name", "mask"))


@pytest.mark.parametrize("cell,detail,steps,shape_kind", [(64, 100, 2, "square"), (64, 50, 2, "square"), (48, 75, 1, "square"),
                                                           (64, 50, 1, "hex")])
def test_host_grid_state_matches_oracle(L, oracle, cell, detail, steps, shape_kind):
    """GridGenerator::getGridState incl. the entropy split and mergeBounds: the library's host model vs the cv2-based oracle and,
    when it is built, vs the reference's own object code."""
    pytest.importorskip("cv2")
    from mosaicmagnifique_b200 import synthetic
    from mosaicmagnifique_b200._capi import CellShapeC
    main = synthetic.make_main_image(300, 420, 17, block=32)
    if shape_kind == "hex":
        sh = oracle.CellShape.from_mask(synthetic.hexagon_mask(cell))
        sh.row_spacing = sh.alt_row_spacing = cell * 3 // 4
        sh.col_spacing = sh.alt_col_spacing = cell * 55 // 64
        sh.alt_row_offset = cell * 55 // 128
        sh.alt_row_flip_h = True
    else:
        sh = oracle.CellShape.square(cell)
    want = oracle.grid_state(oracle.CellGroup.make(sh, detail, steps), main)
    if oracle.reference_generator_available():
        # ... and the reference's own GridGenerator.cpp / CellGroup.cpp object code (oracle/_ref/libref_core.so) says the same
        ref_states = oracle.reference_grid_state(oracle.CellGroup.make(sh, detail, steps), main)
        assert len(ref_states) == len(want) and all(np.array_equal(a, b) for a, b in zip(ref_states, want))
    c = CellShapeC(*sh.params())
    n_steps = ctypes.c_int()
    rows, cols = (ctypes.c_int * 8)(), (ctypes.c_int * 8)()
    out = np.empty(1 << 16, np.int64)
    rc = L.mosaic_host_grid_state(ctypes.byref(c), sh.mask.ctypes.data, 0, detail, steps, main.ctypes.data, main.shape[0], main.shape[1],
                                  main.strides[0], 8, ctypes.byref(n_steps), rows, cols, out.ctypes.data, out.size)
    assert rc == 0
    assert n_steps.value == len(want)
    off = 0
    for s, w in enumerate(want):
        assert (rows[s], cols[s]) == w.shape
        got = out[off:off + w.size].reshape(w.shape)
        assert np.array_equal(got, w), "step %d differs in %d cells" % (s, int((got != w).sum()))
        off += w.size
    assert sum(int((w >= 0).sum()) for w in want) > 10
