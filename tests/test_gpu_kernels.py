"""Kernel-level parity on the GPU, through the C ABI's kernel entry points. Mirrors the reference's own kernel
tests (test/tst_ColourDifference.h:233-543, test/tst_CUDAKernel.h:16-272) with fixed seeds, the oracle as checker."""
import ctypes
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "colour_vectors.json")))


@pytest.fixture(scope="module")
def L():
    from mosaicmagnifique_b200 import capi
    return capi()


def _coldiff(L, t, a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    out = np.empty(a.shape[0], np.float32)
    rc = L.mosaic_kernel_colour_difference(0, t, a.ctypes.data, b.ctypes.data, a.shape[0], out.ctypes.data)
    assert rc == 0
    return out.astype(np.float64)


@pytest.mark.parametrize("name,t", [("rgb_euclidean", 0), ("cie76", 1), ("ciede2000", 2)])
def test_known_answers(L, oracle, name, t):
    """ColourDifference.*_CUDA (tst_ColourDifference.h:233-309): golden vectors through the kernel, size = 1, 1e-4."""
    s = GOLD["sets"][name]
    a = np.array([v["first"] for v in s["vectors"]], np.float32)
    b = np.array([v["second"] for v in s["vectors"]], np.float32)
    want = np.array([v["difference"] for v in s["vectors"]])
    got = _coldiff(L, t, a, b)
    err = np.abs(got - want)
    tol = 1e-4 + 2e-6 * np.abs(want)
    if name == "ciede2000":
        knife = np.zeros(len(want), bool)
        knife[8:16] = True  # pairs ON the mean-hue discontinuity: either side accepted (see tests/test_colour_math.py)
        others = np.concatenate([want[8:16], oracle.diff_batch(2, a, b)[8:16]])
        assert (np.min(np.abs(got[knife, None] - others[None, :]), axis=1) < 1e-3).all()
        err, tol = err[~knife], tol[~knife]
    assert (err <= tol).all(), err.max()


@pytest.mark.parametrize("t", [0, 1, 2])
def test_random_pixels_vs_oracle(L, oracle, t):
    """ColourDifference.*_CPUvsCUDA (tst_ColourDifference.h:315-387), 2^14 seeded pixels."""
    rng = np.random.default_rng(40 + t)
    n = 1 << 14
    if t == 0:
        a, b = rng.uniform(0, 255, (n, 3)), rng.uniform(0, 255, (n, 3))
    else:
        lo, hi = np.array([0, -128, -128]), np.array([100, 127, 127])
        a, b = rng.uniform(lo, hi, (n, 3)), rng.uniform(lo, hi, (n, 3))
    a, b = a.astype(np.float32), b.astype(np.float32)
    got = _coldiff(L, t, a, b)
    want = oracle.diff_batch(t, a, b)
    assert np.isfinite(got).all()
    assert (np.abs(got - want) <= 1e-4 + 2e-5 * want).all(), np.abs(got - want).max()


@pytest.mark.parametrize("t", [0, 2])
@pytest.mark.parametrize("edge", [False, True])
def test_image_difference_sum(L, oracle, t, edge):
    """*_CPUvsBatchCUDA / *_CUDAEdgeCase (tst_ColourDifference.h:389-543): one cell against a batch of library images,
    random mask; edge case = last quarter of the rows (and some columns) outside the target area."""
    rng = np.random.default_rng(50 + t)
    size, n_lib = 48, 37
    lo, hi = (np.array([0, 0, 0]), np.array([255, 255, 255])) if t == 0 else (np.array([0, -128, -128]), np.array([100, 127, 127]))
    cell = rng.uniform(lo, hi, (size, size, 3)).astype(np.float32)
    lib = rng.uniform(lo, hi, (n_lib, size, size, 3)).astype(np.float32)
    mask = (rng.random((size, size)) < 0.7).astype(np.uint8) * 255
    ta = np.array([0, size * 3 // 4, 5, size - 3], np.int32) if edge else None
    out = np.empty(n_lib, np.float32)
    rc = L.mosaic_kernel_image_difference_sum(0, t, cell.ctypes.data, lib.ctypes.data, n_lib, mask.ctypes.data, size,
                                              None if ta is None else ta.ctypes.data, out.ctypes.data)
    assert rc == 0
    m = mask != 0
    if edge:
        box = np.zeros_like(m)
        box[ta[0]:ta[1], ta[2]:ta[3]] = True
        m &= box
    want = np.array([oracle.diff_batch(t, cell[m], lib[i][m]).sum() for i in range(n_lib)])
    np.testing.assert_allclose(out, want, rtol=1e-5)


def _select(L, scores, grid, r, a):
    g = np.ascontiguousarray(grid, np.int64).copy()
    s = np.ascontiguousarray(scores, np.float32)
    rc = L.mosaic_kernel_select(0, s.ctypes.data, s.shape[1], g.ctypes.data, g.shape[0], g.shape[1], r, a)
    assert rc == 0
    return g


def test_select_reference_shape(L, oracle):
    """CUDAKernel.CalculateRepeats / FindLowest shape (tst_CUDAKernel.h:16-165): 5x5 grid, 10 images, range 2, +500."""
    rng = np.random.default_rng(3)
    scores = rng.uniform(0, 100, (25, 10)).astype(np.float32)
    grid = np.zeros((5, 5), np.int64)
    got = _select(L, scores, grid, 2, 500)
    want = oracle.select_from_D(scores.astype(np.float64), grid, 2, 500)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("rows,cols,n_lib,r,a,invalid", [(12, 17, 40, 3, 50, 0.0), (30, 41, 300, 8, 500, 0.2),
                                                          (9, 9, 5, 2, 1000, 0.5), (20, 25, 64, 100000, 7, 0.1),
                                                          (16, 16, 33, 0, 500, 0.1), (16, 16, 33, 4, 0, 0.0)])
def test_select_random_grids(L, oracle, rows, cols, n_lib, r, a, invalid):
    rng = np.random.default_rng(rows * 1000 + cols)
    grid = np.where(rng.random((rows, cols)) < invalid, -1, 0).astype(np.int64)
    n_valid = int((grid >= 0).sum())
    # quantised scores force exact ties: lowest index must win, as on the CPU
    scores = np.round(rng.uniform(0, 30, (n_valid, n_lib))).astype(np.float32)
    got = _select(L, scores, grid, r, a)
    want = oracle.select_from_D(scores.astype(np.float64), grid, r, a)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n_rows,n_lib,k", [(7, 1000, 145), (3, 64, 64), (5, 333, 1), (4, 5000, 13)])
def test_topk(L, n_rows, n_lib, k):
    rng = np.random.default_rng(n_lib + k)
    s = np.round(rng.uniform(0, 50, (n_rows, n_lib)), 1).astype(np.float32)  # many ties
    os_, oi = np.empty((n_rows, k), np.float32), np.empty((n_rows, k), np.int32)
    assert L.mosaic_kernel_topk(0, s.ctypes.data, n_rows, n_lib, k, os_.ctypes.data, oi.ctypes.data) == 0
    for r in range(n_rows):
        order = np.lexsort((np.arange(n_lib), s[r]))[:k]  # by (score, index)
        assert sorted(oi[r].tolist()) == sorted(order.tolist())
        assert np.array_equal(s[r][oi[r]], os_[r])


def test_bgr_to_lab_matches_opencv(L):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    px = rng.integers(0, 256, (1, 1 << 18, 3), dtype=np.uint8)
    px[0, :256] = np.arange(256, dtype=np.uint8)[:, None]  # the grey axis incl. 0 and 255
    out = np.empty((px.shape[1], 3), np.float32)
    assert L.mosaic_kernel_bgr_to_lab(0, px.ctypes.data, px.shape[1], out.ctypes.data) == 0
    ref = cv2.cvtColor(px.astype(np.float32) * np.float32(1 / 255.0), cv2.COLOR_BGR2Lab).reshape(-1, 3)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("k", [2, 4, 8])
def test_resize_area_matches_opencv(L, k):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6 + k)
    n, size = 5, 64
    src = rng.integers(0, 256, (n, size, size, 3), dtype=np.uint8)
    dst = np.empty((n, size // k, size // k, 3), np.uint8)
    assert L.mosaic_kernel_resize_area_u8(0, src.ctypes.data, n, size, k, dst.ctypes.data) == 0
    for i in range(n):
        assert np.array_equal(dst[i], cv2.resize(src[i], (size // k, size // k), interpolation=cv2.INTER_AREA))
    srcf = (rng.random((n, size, size, 3), dtype=np.float32) * 200 - 100).astype(np.float32)
    dstf = np.empty((n, size // k, size // k, 3), np.float32)
    assert L.mosaic_kernel_resize_area_f32(0, srcf.ctypes.data, n, size, k, dstf.ctypes.data) == 0
    for i in range(n):
        assert np.array_equal(dstf[i], cv2.resize(srcf[i], (size // k, size // k), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("sh,sw,dh,dw,cn", [(50, 50, 128, 128, 3), (100, 100, 128, 128, 3), (127, 127, 128, 128, 1), (37, 37, 64, 64, 1),
                                            (3, 3, 10, 10, 3), (20, 31, 45, 77, 3), (255, 255, 301, 301, 3)])
def test_resize_cubic_matches_opencv(L, oracle, sh, sw, dh, dw, cn):
    """cubic_u8_kernel against OpenCV's own INTER_CUBIC (oracle.resize_cubic_opencv), bit exact."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(sh * 1000 + dw)
    a = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
    out = np.empty((dh, dw, cn), np.uint8)
    assert L.mosaic_kernel_resize_cubic_u8(0, a.ctypes.data, sh, sw, cn, out.ctypes.data, dh, dw) == 0
    assert np.array_equal(out, oracle.resize_cubic_opencv(a, dh, dw).reshape(dh, dw, cn))


def _reference_ingest(oracle, im, size):
    """ImageLibrary::addImage's image arithmetic (ImageLibrary.cpp:62-83) with the reference's own OpenCV calls."""
    r, c = im.shape[:2]
    if c < r:
        d = (r - c) // 2
        im = im[d:c + d, :c]
    elif c > r:
        d = (c - r) // 2
        im = im[:r, d:r + d]
    return oracle.resize_image_exact(np.ascontiguousarray(im), size, size)


@pytest.mark.parametrize("rows,cols,size", [(300, 200, 64), (256, 512, 128), (40, 60, 64), (64, 64, 64), (97, 97, 128), (513, 400, 100),
                                            (31, 90, 32)])
def test_library_ingest_matches_reference_procedure(L, oracle, rows, cols, size):
    pytest.importorskip("cv2")
    rng = np.random.default_rng(rows * 7 + cols)
    im = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    out = np.empty((size, size, 3), np.uint8)
    assert L.mosaic_library_ingest(0, im.ctypes.data, rows, cols, im.strides[0], size, out.ctypes.data) == 0
    assert np.array_equal(out, _reference_ingest(oracle, im, size))
    # a strided view (row_stride > cols * 3) must give the same result
    wide = np.zeros((rows, cols + 5, 3), np.uint8)
    wide[:, :cols] = im
    out2 = np.empty_like(out)
    assert L.mosaic_library_ingest(0, wide.ctypes.data, rows, cols, wide.strides[0], size, out2.ctypes.data) == 0
    assert np.array_equal(out, out2)
    assert L.mosaic_library_ingest(0, im.ctypes.data, 0, cols, im.strides[0], size, out.ctypes.data) == -1  # empty image


def test_image_library_mirror(oracle):
    """ImageLibrary.addImage / setImageSize (ImageLibrary.cpp:42-86) through the Python mirror."""
    pytest.importorskip("cv2")
    from mosaicmagnifique_b200 import ImageLibrary
    rng = np.random.default_rng(12)
    lib = ImageLibrary(48, seed=1)
    srcs = {}
    for i, (r, c) in enumerate([(100, 80), (48, 48), (30, 45), (200, 200)]):
        im = rng.integers(0, 256, (r, c, 3), dtype=np.uint8)
        srcs["im%d" % i] = im
        idx = lib.addImage(im, "im%d" % i)
        assert lib.getNames()[idx] == "im%d" % i
    assert lib.asArray().shape == (4, 48, 48, 3)
    for name, img in zip(lib.getNames(), lib.getImages()):
        assert np.array_equal(img, _reference_ingest(oracle, srcs[name], 48))
    before = {n: im.copy() for n, im in zip(lib.getNames(), lib.getImages())}
    lib.setImageSize(32)  # resizes the stored (already 48 px) images again, like batchResizeMat on m_originalImages
    for name, img in zip(lib.getNames(), lib.getImages()):
        assert np.array_equal(img, oracle.resize_image_exact(before[name], 32, 32))
    with pytest.raises(ValueError):
        lib.addImage(np.zeros((0, 5, 3), np.uint8))


def test_microbench_runs(L):
    out = np.zeros(8, np.float64)
    assert L.mosaic_kernel_microbench(0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 8) == 0
    assert out[0] > 1e12 and out[2] > 1e11 and out[4] >= 100


@pytest.mark.parametrize("rot", [180.0, 120.0, 240.0, 150.0, 210.0, 90.0, 270.0, 30.0, 60.0])
def test_hue_rotation_matches_opencv(L, rot):
    """ColourScheme variants (ColourScheme.cpp:36-177) against the OpenCV calls the reference makes, bit exact."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(int(rot))
    img = rng.integers(0, 256, (257, 263, 3), dtype=np.uint8)
    img[0, :256] = np.arange(256, dtype=np.uint8)[:, None]  # greys: S = 0
    hsv = cv2.cvtColor(img.astype(np.float32), cv2.COLOR_BGR2HSV_FULL)
    hsv[..., 0] = np.fmod(hsv[..., 0] + np.float32(rot), np.float32(360.0))
    want = np.clip(np.rint(cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR_FULL)), 0, 255).astype(np.uint8)
    out = np.empty_like(img)
    assert L.mosaic_kernel_hue_rotate(0, img.ctypes.data, img.shape[0], img.shape[1], rot, out.ctypes.data) == 0
    assert np.array_equal(out, want), int((out != want).sum())


def test_image_library_against_reference_object_code(oracle):
    """The GPU ingest (ImageLibrary mirror -> mosaic_library_ingest) against the reference's OWN ImageLibrary.cpp object code
    (oracle/_ref/libref_core.so, prebuilt): same images for tall / wide / square / smaller / equal-size inputs, and after
    setImageSize."""
    if not oracle.reference_generator_available():
        pytest.skip("oracle/_ref/libref_core.so (reference object code) not present / not loadable")
    from mosaicmagnifique_b200 import ImageLibrary
    rng = np.random.default_rng(31)
    ours, theirs = ImageLibrary(64, seed=2), oracle.ReferenceImageLibrary(64)
    for i, (r, c) in enumerate([(300, 200), (256, 512), (40, 60), (64, 64), (97, 97), (513, 400), (31, 90)]):
        im = rng.integers(0, 256, (r, c, 3), dtype=np.uint8)
        ours.addImage(im, "im%d" % i)
        theirs.add_image(im, "im%d" % i)
    want = dict(theirs.items())
    assert sorted(ours.getNames()) == sorted(want)
    for name, img in zip(ours.getNames(), ours.getImages()):
        assert np.array_equal(img, want[name]), name
    ours.setImageSize(40)
    theirs.set_image_size(40)
    want = dict(theirs.items())
    for name, img in zip(ours.getNames(), ours.getImages()):
        assert np.array_equal(img, want[name]), name
    theirs.close()
