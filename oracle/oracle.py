"""CPU ORACLE driver (test infrastructure, NOT product code).

Restates the reference's host-side photomosaic path in Python on top of
  * cv2 (OpenCV, the reference's own third-party dependency; reference pins 4.5.2 in
    Shared.props:5, this image has 4.13) for every OpenCV call the reference makes, and
  * oracle/_ref/libmosaic_oracle.so (mosaic_oracle.c) for the f64 per-pixel loops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (mosaicmagnifique_b200) never does.

Reference citations (all under /root/reference/src):
  CellShape.resized            CellShape/CellShape.cpp:281-312, setCellMask :116-135
  CellGroup detail/size steps  CellShape/CellGroup.cpp:65-128
  resizeImage / batchResize    Other/ImageUtility.cpp:34-101
  preprocessMainImage          Photomosaic/PhotomosaicGeneratorBase.cpp:223-252
  preprocessLibraryImages      Photomosaic/PhotomosaicGeneratorBase.cpp:255-290
  getCellAt                    Photomosaic/PhotomosaicGeneratorBase.cpp:293-329
  generateBestFits             Photomosaic/CPUPhotomosaicGenerator.cpp:33-112
  ColourScheme variants        Photomosaic/ColourScheme.cpp:36-177
  getGridState / findCellState Grid/GridGenerator.cpp:29-193
  mergeBounds                  Grid/GridBounds.cpp:39-104
  .mcs container               CellShape/CellShape.cpp:363-434, Other/CustomQDataStream.h:56-87

Parity pinning: the colour maths is pinned by the reference's 48 known-answer vectors and by
the reference's own ColourDifference.cpp / GridUtility.cpp / GridBounds.cpp compiled unmodified
(libref_core.so); the generator loop (generate / generate_step / select_from_D) is pinned on the
reference's own CPUPhotomosaicGenerator.cpp, compiled unmodified into the same library and run by
reference_generate() (tests/test_oracle_ref_generator.py).
The OpenCV numerics (Lab LUT, INTER_AREA, HSV) have no golden vectors in the reference
("parity unpinned" for those); cv2 itself is used here, so the oracle is OpenCV by construction.
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
from dataclasses import dataclass, field, replace

import numpy as np

try:  # cv2 is only needed for preprocessing; the C core works without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

HERE = os.path.dirname(os.path.abspath(__file__))
PAD_GRID = 2  # Grid/GridUtility.h:33
RGB_EUCLIDEAN, CIE76, CIEDE2000 = 0, 1, 2  # Photomosaic/ColourDifference.h:13-19
# Photomosaic/ColourScheme.h:10-19
SCHEME_NONE, SCHEME_COMPLEMENTARY, SCHEME_TRIADIC, SCHEME_COMPOUND, SCHEME_TETRADIC, SCHEME_ANALAGOUS = range(6)
SCHEME_ROTATIONS = {0: (), 1: (180.0,), 2: (120.0, 240.0), 3: (150.0, 210.0), 4: (90.0, 180.0, 270.0),
                    5: (30.0, 60.0, 90.0)}
MAX_ENTROPY = 8.0  # Other/ImageUtility.h (log2(256))


def build(force: bool = False) -> None:
    """Compiles oracle/_ref/*.so with the committed Makefile (reference sources only if present)."""
    out = os.path.join(HERE, "_ref", "libmosaic_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(HERE, "mosaic_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])


_lib = None


class _Shape(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "size", "row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset",
        "alt_col_offset", "alt_col_flip_h", "alt_col_flip_v", "alt_row_flip_h", "alt_row_flip_v")]


class _Stats(ctypes.Structure):
    _fields_ = [("visited", ctypes.c_int64), ("nominal", ctypes.c_int64)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(os.path.join(HERE, "_ref", "libmosaic_oracle.so"))
        dp, ip, fp, u8p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_double, ctypes.c_int, ctypes.c_float,
                                                              ctypes.c_uint8, ctypes.c_int64))
        L.mo_rgb_euclidean.restype = ctypes.c_double
        L.mo_rgb_euclidean.argtypes = [dp, dp]
        L.mo_ciede2000.restype = ctypes.c_double
        L.mo_ciede2000.argtypes = [dp, dp]
        L.mo_diff_batch.argtypes = [ctypes.c_int, fp, fp, ctypes.c_int64, dp]
        L.mo_grid_size.argtypes = [ctypes.POINTER(_Shape), ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip]
        L.mo_rect_at.argtypes = [ctypes.POINTER(_Shape), ctypes.c_int, ctypes.c_int, ip]
        L.mo_flip_at.restype = ctypes.c_int
        L.mo_flip_at.argtypes = [ctypes.POINTER(_Shape), ctypes.c_int, ctypes.c_int]
        L.mo_cell_bounds.argtypes = [ctypes.POINTER(_Shape), ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ip, ip, ip]
        L.mo_calculate_repeats.restype = ctypes.c_int
        L.mo_calculate_repeats.argtypes = [i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, i64p, i64p]
        L.mo_generate_step.restype = ctypes.c_int
        L.mo_generate_step.argtypes = [ctypes.c_int, fp, ip, ip, ctypes.c_int, fp, ctypes.c_int64, ctypes.c_int, u8p,
                                       i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, dp, dp, ctypes.POINTER(_Stats)]
        L.mo_select_from_D.argtypes = [dp, ctypes.c_int64, i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.mo_entropy.restype = ctypes.c_double
        L.mo_entropy.argtypes = [u8p, u8p, ctypes.c_int64]
        _lib = L
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ----------------------------------------------------------------------------- colour maths

def rgb_euclidean(a, b) -> float:
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return lib().mo_rgb_euclidean(_ptr(a, ctypes.c_double), _ptr(b, ctypes.c_double))


cie76 = rgb_euclidean  # ColourDifference.h:41


def ciede2000(a, b) -> float:
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return lib().mo_ciede2000(_ptr(a, ctypes.c_double), _ptr(b, ctypes.c_double))


def diff_batch(diff_type: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    out = np.empty(a.shape[0], np.float64)
    lib().mo_diff_batch(diff_type, _ptr(a, ctypes.c_float), _ptr(b, ctypes.c_float), a.shape[0], _ptr(out, ctypes.c_double))
    return out


# ----------------------------------------------------------------------------- cell shapes

def resize_image_exact(img: np.ndarray, th: int, tw: int) -> np.ndarray:
    """ImageUtility::resizeImage with ResizeType::EXACT (ImageUtility.cpp:34-62)."""
    factor = th / img.shape[0]
    if factor == 1.0:
        factor = tw / img.shape[1]
    if factor == 1.0:
        return img
    if factor < 1:
        return cv2.resize(img, (tw, th), interpolation=cv2.INTER_AREA)
    return resize_cubic_opencv(img, th, tw)


def resize_cubic_opencv(img: np.ndarray, th: int, tw: int) -> np.ndarray:
    """cv::resize(INTER_CUBIC) by OpenCV's OWN code (imgproc/resize.cpp). IPP-enabled builds (this cv2 is one) route 8U
    cubic through a closed IPP kernel whose results differ by +-1 LSB in ~5 % of the samples from OpenCV's code, so IPP
    is switched off around the call: the parity target is the open, reproducible implementation."""
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        return cv2.resize(img, (tw, th), interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.ipp.setUseIPP(was)


@dataclass
class CellShape:
    """CellShape (CellShape/CellShape.h:98-109): binary mask + tiling parameters."""
    mask: np.ndarray  # size x size u8, 0/255
    row_spacing: int = 0
    col_spacing: int = 0
    alt_row_spacing: int = 0
    alt_col_spacing: int = 0
    alt_row_offset: int = 0
    alt_col_offset: int = 0
    alt_col_flip_h: bool = False
    alt_col_flip_v: bool = False
    alt_row_flip_h: bool = False
    alt_row_flip_v: bool = False
    name: str = ""

    @staticmethod
    def from_mask(mask: np.ndarray) -> "CellShape":
        """CellShape(const cv::Mat&) (CellShape.cpp:37-45) + setCellMask (:116-135)."""
        m = np.where(np.asarray(mask, np.uint8) > 127, 255, 0).astype(np.uint8)  # THRESH_BINARY 127
        assert m.shape[0] == m.shape[1], "non-square masks (imageToSquare PAD) are outside the oracle"
        s = m.shape[0]
        return CellShape(m, s, s, s, s)

    @staticmethod
    def square(size: int) -> "CellShape":
        """CellShape(size_t) default cell (CellShape.cpp:66-69)."""
        return CellShape.from_mask(np.full((size, size), 255, np.uint8))

    @property
    def size(self) -> int:
        return self.mask.shape[0]

    def flipped(self, h: bool, v: bool) -> np.ndarray:
        m = self.mask
        if h:
            m = m[:, ::-1]
        if v:
            m = m[::-1, :]
        return np.ascontiguousarray(m)

    def masks4(self) -> np.ndarray:
        """index = flip_h + 2*flip_v (CellGroup.cpp:146 order)."""
        return np.stack([self.flipped(False, False), self.flipped(True, False), self.flipped(False, True),
                         self.flipped(True, True)])

    def resized(self, size: int) -> "CellShape":
        """CellShape::resized (CellShape.cpp:281-312)."""
        if self.mask.size == 0 or size == self.size:
            return replace(self)
        rm = resize_image_exact(self.mask, size, size)
        out = CellShape.from_mask(rm)
        ratio = rm.shape[0] / self.mask.shape[0]
        fl = lambda v: int(np.floor(v * ratio))
        out.row_spacing = max(fl(self.row_spacing), 1)
        out.col_spacing = max(fl(self.col_spacing), 1)
        out.alt_row_spacing = max(fl(self.alt_row_spacing), 1)
        out.alt_col_spacing = max(fl(self.alt_col_spacing), 1)
        out.alt_row_offset = fl(self.alt_row_offset)
        out.alt_col_offset = fl(self.alt_col_offset)
        out.alt_col_flip_h, out.alt_col_flip_v = self.alt_col_flip_h, self.alt_col_flip_v
        out.alt_row_flip_h, out.alt_row_flip_v = self.alt_row_flip_h, self.alt_row_flip_v
        out.name = self.name
        return out

    def c_shape(self) -> _Shape:
        return _Shape(self.size, self.row_spacing, self.col_spacing, self.alt_row_spacing, self.alt_col_spacing,
                      self.alt_row_offset, self.alt_col_offset, int(self.alt_col_flip_h), int(self.alt_col_flip_v),
                      int(self.alt_row_flip_h), int(self.alt_row_flip_v))

    def params(self) -> list:
        c = self.c_shape()
        return [getattr(c, n) for n, _ in _Shape._fields_]


def load_mcs(path: str) -> CellShape:
    """.mcs v8 reader (CellShape.cpp:363-434; QDataStream big-endian, CustomQDataStream.h:56-87)."""
    d = open(path, "rb").read()
    o = 0

    def u32():
        nonlocal o
        v = struct.unpack_from(">I", d, o)[0]; o += 4
        return v

    def i32():
        nonlocal o
        v = struct.unpack_from(">i", d, o)[0]; o += 4
        return v

    magic, version = u32(), u32()
    if magic != 0x87AECFB1:
        raise ValueError("not a .mcs file")
    n = u32()
    name = "" if n == 0xFFFFFFFF else d[o:o + n].decode("utf-16-be"); o += 0 if n == 0xFFFFFFFF else n
    n = u32()
    png = d[o:o + n]; o += n
    mask = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
    if mask.ndim == 3:
        mask = mask[..., 0]
    rs, cs, ars, acs, aro, aco = (i32() for _ in range(6))
    fl = struct.unpack_from(">4?", d, o)
    shape = CellShape.from_mask(mask)
    shape.row_spacing, shape.col_spacing, shape.alt_row_spacing, shape.alt_col_spacing = rs, cs, ars, acs
    shape.alt_row_offset, shape.alt_col_offset = aro, aco
    shape.alt_col_flip_h, shape.alt_col_flip_v, shape.alt_row_flip_h, shape.alt_row_flip_v = fl
    shape.name = name
    shape._version = version
    return shape


@dataclass
class CellGroup:
    """CellGroup (CellShape/CellGroup.cpp:29-128): per size step a normal and a detail cell."""
    cells: list = field(default_factory=list)
    detail_cells: list = field(default_factory=list)
    detail: float = 1.0
    size_steps: int = 0

    @staticmethod
    def make(shape: CellShape, detail_percent: int = 100, size_steps: int = 0) -> "CellGroup":
        g = CellGroup()
        g.detail = detail_percent / 100.0
        g.size_steps = size_steps
        g.cells = [shape]
        g.detail_cells = [shape.resized(max(int(shape.size * g.detail), 1))]  # CellGroup.cpp:79-80
        size = shape.size
        for _ in range(size_steps):
            size //= 2  # CellGroup.cpp:105
            g.cells.append(g.cells[-1].resized(size))
            g.detail_cells.append(g.cells[-1].resized(max(int(size * g.detail), 1)))  # :114
        return g


# ----------------------------------------------------------------------------- grid geometry

def grid_size(shape: CellShape, w: int, h: int, pad: int = PAD_GRID):
    gx, gy = ctypes.c_int(), ctypes.c_int()
    s = shape.c_shape()
    lib().mo_grid_size(ctypes.byref(s), w, h, pad, ctypes.byref(gx), ctypes.byref(gy))
    return gx.value, gy.value


def rect_at(shape: CellShape, x: int, y: int):
    r = (ctypes.c_int * 4)()
    s = shape.c_shape()
    lib().mo_rect_at(ctypes.byref(s), x, y, r)
    return tuple(r)


def flip_at(shape: CellShape, x: int, y: int) -> int:
    s = shape.c_shape()
    return lib().mo_flip_at(ctypes.byref(s), x, y)


def cell_bounds(normal: CellShape, detail_size: int, detail: float, x: int, y: int, w: int, h: int):
    g, l, d = (ctypes.c_int * 4)(), (ctypes.c_int * 4)(), (ctypes.c_int * 4)()
    s = normal.c_shape()
    lib().mo_cell_bounds(ctypes.byref(s), detail_size, detail, x, y, w, h, g, l, d)
    return tuple(g), tuple(l), tuple(d)


# ----------------------------------------------------------------------------- grid state (GridGenerator)

def _merge_bounds(bounds: list) -> list:
    """GridBounds::mergeBounds (GridBounds.cpp:39-104); rects are [x, y, w, h]."""
    def union(a, b):
        x0, y0 = min(a[0], b[0]), min(a[1], b[1])
        x1, y1 = max(a[0] + a[2], b[0] + b[2]), max(a[1] + a[3], b[1] + b[3])
        return [x0, y0, x1 - x0, y1 - y0]

    bounds = [list(b) for b in bounds]
    merged_any = True
    while merged_any:
        merged_any = False
        i = 0
        while i + 1 < len(bounds):
            j = i + 1
            while j < len(bounds) and i + 1 < len(bounds):
                a, b = bounds[i], bounds[j]
                merge = False
                if a[0] == b[0] and a[2] == b[2]:
                    yd = b[1] - a[1]
                    merge = yd == 0 or (0 < yd <= a[3]) or (yd < 0 and -yd <= b[3])
                elif a[1] == b[1] and a[3] == b[3]:
                    xd = b[0] - a[0]
                    merge = xd == 0 or (0 < xd <= a[2]) or (xd < 0 and -xd <= b[2])
                else:
                    u = union(a, b)
                    merge = u == a or u == b
                if merge:
                    bounds[i] = union(a, b)
                    del bounds[j]
                    merged_any = True
                else:
                    j += 1
            if i + 1 < len(bounds):
                i += 1
            else:
                break
    return bounds


def entropy(gray: np.ndarray, mask: np.ndarray | None) -> float:
    gray = np.ascontiguousarray(gray, np.uint8)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    return lib().mo_entropy(_ptr(gray, ctypes.c_uint8), None if m is None else _ptr(m, ctypes.c_uint8), gray.size)


def grid_state(group: CellGroup, main_bgr: np.ndarray | None, height: int = 0, width: int = 0) -> list:
    """GridGenerator::getGridState (GridGenerator.cpp:29-110). Returns one int64 array per generated step,
    -1 = nullopt, 0 = valid."""
    gh = height if main_bgr is None else main_bgr.shape[0]
    gw = width if main_bgr is None else main_bgr.shape[1]
    active = [[0, 0, gw, gh]]
    out = []
    clamp = lambda v, lo, hi: max(lo, min(v, hi))
    for step in range(group.size_steps + 1):
        if not active:
            break
        shape = group.cells[step]
        dshape = group.detail_cells[step]
        dmasks = dshape.masks4()
        gx, gy = grid_size(shape, gw, gh)
        g = np.full((gy, gx), -1, np.int64)
        nxt = []
        for y in range(-PAD_GRID, gy - PAD_GRID):
            for x in range(-PAD_GRID, gx - PAD_GRID):
                r = rect_at(shape, x, y)
                # findCellState (GridGenerator.cpp:113-193)
                inb = False
                for b in active:
                    y0, y1 = clamp(r[1], b[1], b[1] + b[3]), clamp(r[1] + r[3], b[1], b[1] + b[3])
                    x0, x1 = clamp(r[0], b[0], b[0] + b[2]), clamp(r[0] + r[2], b[0], b[0] + b[2])
                    if y0 != y1 and x0 != x1:
                        inb = True
                        break
                if not inb:
                    continue
                split = False
                if main_bgr is not None and step < group.size_steps:
                    _, local, db = cell_bounds(shape, dshape.size, group.detail, x, y, gw, gh)
                    y0, y1 = clamp(r[1], 0, gh), clamp(r[1] + r[3], 0, gh)
                    x0, x1 = clamp(r[0], 0, gw), clamp(r[0] + r[2], 0, gw)
                    cell = main_bgr[y0:y1, x0:x1]
                    mask = dmasks[flip_at(shape, x, y)][db[1]:db[1] + db[3], db[0]:db[0] + db[2]]
                    if cell.size:
                        cell = resize_image_exact(np.ascontiguousarray(cell), mask.shape[0], mask.shape[1])
                        gray = cv2.cvtColor(cell, cv2.COLOR_BGR2GRAY)
                        split = entropy(gray, np.ascontiguousarray(mask)) >= MAX_ENTROPY * 0.7
                    # an empty cell has entropy 0 (ImageUtility.cpp:191-192)
                if split:
                    y0, y1 = clamp(r[1], 0, gh), clamp(r[1] + r[3], 0, gh)
                    x0, x1 = clamp(r[0], 0, gw), clamp(r[0] + r[2], 0, gw)
                    if y0 != y1 and x0 != x1:
                        nxt.append([x0, y0, x1 - x0, y1 - y0])
                else:
                    g[y + PAD_GRID, x + PAD_GRID] = 0
        out.append(g)
        active = _merge_bounds(nxt) if nxt else []
    return out


# ----------------------------------------------------------------------------- preprocessing

def colour_scheme_variants(img_bgr8: np.ndarray, scheme: int) -> list:
    """ColourScheme::getColourScheme* (ColourScheme.cpp:36-177): original + hue-rotated 8U copies."""
    out = [img_bgr8]
    rots = SCHEME_ROTATIONS[scheme]
    if not rots:
        return out
    hsv = cv2.cvtColor(img_bgr8.astype(np.float32), cv2.COLOR_BGR2HSV_FULL)
    for r in rots:
        h = hsv.copy()
        h[..., 0] = np.fmod(h[..., 0] + np.float32(r), np.float32(360.0))
        bgr = cv2.cvtColor(h, cv2.COLOR_HSV2BGR_FULL)
        out.append(np.clip(np.rint(bgr), 0, 255).astype(np.uint8))  # convertTo(8U) = saturate_cast(cvRound)
    return out


def to_working_space(img_bgr8: np.ndarray, diff_type: int) -> np.ndarray:
    """PhotomosaicGeneratorBase.cpp:238-247 / :274-286."""
    if diff_type in (CIE76, CIEDE2000):
        f = img_bgr8.astype(np.float32) * np.float32(1 / 255.0)  # convertTo(CV_32F, 1/255.0): f32 multiply
        return cv2.cvtColor(f, cv2.COLOR_BGR2Lab)
    return img_bgr8.astype(np.float32)


def preprocess_library(lib_bgr8: np.ndarray, group: CellGroup, diff_type: int) -> np.ndarray:
    """preprocessLibraryImages (PhotomosaicGeneratorBase.cpp:255-290). lib: N x S x S x 3 u8."""
    size0 = int(round(group.detail_cells[0].size))
    out = []
    for im in lib_bgr8:
        if group.detail != 1:
            flags = cv2.INTER_AREA if group.detail < 1 else cv2.INTER_CUBIC
            im = cv2.resize(im, (size0, size0), interpolation=flags)
        out.append(to_working_space(np.ascontiguousarray(im), diff_type))
    return np.stack(out)


def halve_library(lib_f32: np.ndarray) -> np.ndarray:
    """ImageUtility::batchResizeMat(lib, 0.5) (ImageUtility.cpp:91-101) applied per step
    (CPUPhotomosaicGenerator.cpp:95-99)."""
    s = int(round(0.5 * lib_f32.shape[1]))
    return np.stack([resize_image_exact(im, s, s) for im in lib_f32])


def extract_cells(mains_f32: list, group: CellGroup, step: int, grid: np.ndarray, shared_buffer_quirk: bool = True):
    """getCellAt (PhotomosaicGeneratorBase.cpp:293-329) for every valid cell of one step, raster order.
    Returns cells [n, V, ds, ds, 3] f32, bounds [n,4] (detail space x,y,w,h), flips [n], coords [n,2] (x,y unpadded).

    shared_buffer_quirk reproduces Q1 of SURVEY.md: the V cell Mats share ONE buffer
    (std::vector<cv::Mat>(n, cv::Mat(...)), :310) and resizeImage returns its input when the factor
    is 1 (ImageUtility.cpp:47-48), so at detail 100 % every variant equals the LAST one."""
    shape, dshape = group.cells[step], group.detail_cells[step]
    H, W = mains_f32[0].shape[:2]
    V = len(mains_f32)
    cells, bounds, flips, coords = [], [], [], []
    rows, cols = grid.shape
    for gy in range(rows):
        for gx in range(cols):
            if grid[gy, gx] < 0:
                continue
            x, y = gx - PAD_GRID, gy - PAD_GRID
            gcl, local, db = cell_bounds(shape, dshape.size, group.detail, x, y, W, H)
            vs = []
            shared = np.zeros((shape.size, shape.size, 3), np.float32)
            for v in range(V):
                buf = shared if shared_buffer_quirk else np.zeros((shape.size, shape.size, 3), np.float32)
                if local[2] > 0 and local[3] > 0:
                    buf[local[1]:local[1] + local[3], local[0]:local[0] + local[2]] = \
                        mains_f32[v][gcl[1]:gcl[3], gcl[0]:gcl[2]]
                vs.append(buf)
            if dshape.size != shape.size:
                # the shared buffer holds the last variant by the time the resizes run (:311-316)
                vs = [resize_image_exact(b, dshape.size, dshape.size) for b in vs]
            cells.append(np.stack(vs))
            bounds.append(db)
            flips.append(flip_at(shape, x, y))
            coords.append((x, y))
    n = len(cells)
    ds = dshape.size
    return (np.stack(cells) if n else np.zeros((0, V, ds, ds, 3), np.float32),
            np.asarray(bounds, np.int32).reshape(n, 4), np.asarray(flips, np.int32), np.asarray(coords, np.int32).reshape(n, 2))


# ----------------------------------------------------------------------------- generate

@dataclass
class StepResult:
    grid: np.ndarray            # rows x cols int64, -1 = nullopt
    D: np.ndarray | None        # n_valid x N f64: min over variants of the masked sums (no penalty)
    margins: np.ndarray | None  # n_valid x 2: best and second-best penalised scores
    n_valid: int
    visited: int
    nominal: int
    coords: np.ndarray          # n_valid x 2 unpadded (x, y)


def generate_step(diff_type, cells, bounds, flips, lib_f32, masks4, grid, repeat_range, repeat_addition,
                  want_D=True, early_exit=True, y_begin=0, y_end=None) -> StepResult:
    """One size step of CPUPhotomosaicGenerator::generateBestFits on preprocessed inputs."""
    L = lib()
    grid = np.ascontiguousarray(grid, np.int64).copy()
    rows, cols = grid.shape
    n_valid = int((grid >= 0).sum())
    cells = np.ascontiguousarray(cells, np.float32)
    V = cells.shape[1] if n_valid else 1
    ds = masks4.shape[1]
    lib_f32 = np.ascontiguousarray(lib_f32, np.float32)
    N = lib_f32.shape[0]
    bounds = np.ascontiguousarray(bounds, np.int32)
    flips = np.ascontiguousarray(flips, np.int32)
    masks4 = np.ascontiguousarray(masks4, np.uint8)
    D = np.full((n_valid, N), np.nan, np.float64) if want_D else None
    margins = np.full((n_valid, 2), np.nan, np.float64) if want_D else None
    st = _Stats(0, 0)
    rc = L.mo_generate_step(diff_type, _ptr(cells, ctypes.c_float), _ptr(bounds, ctypes.c_int), _ptr(flips, ctypes.c_int),
                            V, _ptr(lib_f32, ctypes.c_float), N, ds, _ptr(masks4, ctypes.c_uint8),
                            _ptr(grid, ctypes.c_int64), rows, cols, repeat_range, repeat_addition, int(early_exit),
                            y_begin, rows if y_end is None else y_end,
                            None if D is None else _ptr(D, ctypes.c_double),
                            None if margins is None else _ptr(margins, ctypes.c_double), ctypes.byref(st))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return StepResult(grid, D, margins, n_valid, st.visited, st.nominal, np.zeros((0, 2), np.int32))


def select_from_D(D: np.ndarray, grid: np.ndarray, repeat_range: int, repeat_addition: int) -> np.ndarray:
    D = np.ascontiguousarray(D, np.float64)
    grid = np.ascontiguousarray(grid, np.int64).copy()
    lib().mo_select_from_D(_ptr(D, ctypes.c_double), D.shape[1], _ptr(grid, ctypes.c_int64), grid.shape[0], grid.shape[1],
                           repeat_range, repeat_addition)
    return grid


def generate(main_bgr8: np.ndarray, lib_bgr8: np.ndarray, group: CellGroup, grid_states: list, diff_type: int = RGB_EUCLIDEAN,
             scheme: int = SCHEME_NONE, repeat_range: int = 0, repeat_addition: int = 0, want_D: bool = True,
             early_exit: bool = True, shared_buffer_quirk: bool = True) -> list:
    """CPUPhotomosaicGenerator::generateBestFits (CPUPhotomosaicGenerator.cpp:33-112): list of StepResult."""
    mains = [to_working_space(v, diff_type) for v in colour_scheme_variants(main_bgr8, scheme)]
    lib_f32 = preprocess_library(lib_bgr8, group, diff_type)
    results = []
    for step, g in enumerate(grid_states):
        cells, bounds, flips, coords = extract_cells(mains, group, step, g, shared_buffer_quirk)
        masks4 = group.detail_cells[step].masks4()
        assert lib_f32.shape[1] == masks4.shape[1], "library / detail-mask size mismatch (SURVEY Q4)"
        r = generate_step(diff_type, cells, bounds, flips, lib_f32, masks4, g, repeat_range, repeat_addition,
                          want_D=want_D, early_exit=early_exit)
        r.coords = coords
        results.append(r)
        if step + 1 < len(grid_states):
            lib_f32 = halve_library(lib_f32)
    return results


# ----------------------------------------------------------------------------- the reference's own generator object code
#
# oracle/_ref/libref_core.so holds the reference's PhotomosaicGeneratorBase.cpp, CPUPhotomosaicGenerator.cpp, GridGenerator.cpp,
# ColourDifference.cpp, ColourScheme.cpp, GridUtility.cpp, GridBounds.cpp, CellShape.cpp, CellGroup.cpp and ImageLibrary.cpp
# compiled UNMODIFIED (oracle/Makefile, stand-in headers oracle/shim,
# harness oracle/ref_generator_harness.cpp). OpenCV arithmetic inside them (cvtColor, resize, the PNG codec) is answered by
# the callbacks below with the real OpenCV (cv2). Everything else -- setters, preprocessing flow, getCellAt,
# the best-fit loops, repeats, argmin, buildPhotomosaic, getGridState -- runs from the reference's object code.

_CV_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.c_long, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long)
_CODEC_CB = ctypes.CFUNCTYPE(ctypes.c_long, ctypes.c_int, ctypes.POINTER(ctypes.c_uint8), ctypes.c_long, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, ctypes.c_long, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int))
_ref_lib = None
_ref_callbacks = None  # keeps the ctypes thunks alive
_codec_pending = {}


def _np_view(ptr, rows, cols, cv_type, step, writable=False):
    """numpy view of a shim cv::Mat (type code: depth in the low 3 bits -- 0 = 8U, 5 = 32F -- channels - 1 above)."""
    depth, cn = cv_type & 7, ((cv_type >> 3) & 511) + 1
    dt = np.uint8 if depth == 0 else np.float32
    isz = np.dtype(dt).itemsize
    nbytes = (rows - 1) * step + cols * cn * isz
    raw = np.ctypeslib.as_array(ptr, shape=(nbytes,))
    v = np.lib.stride_tricks.as_strided(raw.view(np.uint8), (rows, cols * cn * isz), (step, 1))
    return v, dt, cn


def _cv_callback(op, code, src_p, rows, cols, cv_type, step, dst_p, drows, dcols, dtype, dstep):
    try:
        raw, dt, cn = _np_view(src_p, rows, cols, cv_type, step)
        src = np.ascontiguousarray(raw).view(dt).reshape(rows, cols, cn)
        if cn == 1:
            src = src[..., 0]
        if op == 0:
            out = cv2.cvtColor(src, code)
        elif code == cv2.INTER_CUBIC:
            out = resize_cubic_opencv(src, drows, dcols)
        else:
            out = cv2.resize(src, (dcols, drows), interpolation=code)
        draw, ddt, dcn = _np_view(dst_p, drows, dcols, dtype, dstep)
        out = np.ascontiguousarray(out, ddt).reshape(drows, dcols * dcn)
        draw[:] = out.view(np.uint8).reshape(drows, -1)
        return 0
    except Exception:  # noqa: BLE001 -- must not propagate through the C frames
        import traceback
        traceback.print_exc()
        return 1


def _codec_callback(op, src_p, n, rows, cols, cv_type, step, dst_p, info_p):
    """cv::imencode(".png") / cv::imdecode(IMREAD_UNCHANGED) for the reference's CustomQDataStream (two-phase: size, then bytes)."""
    try:
        if op == 0:
            raw, dt, cn = _np_view(src_p, rows, cols, cv_type, step)
            img = np.ascontiguousarray(raw).view(dt).reshape(rows, cols, cn)
            ok, buf = cv2.imencode(".png", img[..., 0] if cn == 1 else img)
            if not ok:
                return -1
            _codec_pending["bytes"] = np.ascontiguousarray(buf).reshape(-1)
            return int(_codec_pending["bytes"].size)
        if op == 1:
            np.ctypeslib.as_array(dst_p, shape=(n,))[:] = _codec_pending.pop("bytes")
            return n
        if op == 2:
            img = cv2.imdecode(np.ctypeslib.as_array(src_p, shape=(n,)).copy(), cv2.IMREAD_UNCHANGED)
            if img is None or img.dtype != np.uint8:
                return -1
            img = img.reshape(img.shape[0], img.shape[1], -1)
            _codec_pending["pixels"] = np.ascontiguousarray(img)
            info_p[0], info_p[1], info_p[2] = img.shape[0], img.shape[1], (img.shape[2] - 1) << 3  # CV_8UC(n)
            return 0
        img = _codec_pending.pop("pixels")
        np.ctypeslib.as_array(dst_p, shape=(img.size,))[:] = img.reshape(-1)
        return 0
    except Exception:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        return -1


def reference_generator_available() -> bool:
    so = os.path.join(HERE, "_ref", "libref_core.so")
    if not os.path.exists(so):
        return False
    try:
        return hasattr(ctypes.CDLL(so), "ref_session_create")
    except OSError:
        return False


def _ref():
    global _ref_lib, _ref_callbacks
    if _ref_lib is None:
        R = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_core.so"))
        vp, i, dbl, lng = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_long
        R.ref_set_callbacks.argtypes = [_CV_CB, _CODEC_CB]
        R.ref_colour_scheme_variants.argtypes = [i, vp, i, i, lng, vp, i]
        R.ref_session_create.restype = vp
        R.ref_session_create.argtypes = [vp, i, i, lng, vp, i, i, vp, vp, i, i, i, i, i, i]
        R.ref_session_destroy.argtypes = [vp]
        R.ref_session_generate.argtypes = [vp, i, vp, vp, vp, vp, i, vp]
        R.ref_session_build.argtypes = [vp, vp, vp]
        R.ref_session_get_cell_at.argtypes = [vp, i, i, i, vp, vp]
        R.ref_grid_state.argtypes = [vp, vp, i, i, vp, i, i, lng, i, i, vp, ctypes.c_longlong, vp, vp]
        R.ref_cell_group_cell.argtypes = [vp, vp, i, i, i, i, vp, vp, lng]
        R.ref_cell_shape_resized.argtypes = [vp, vp, i, vp, vp, lng]
        R.ref_mcs_load.argtypes = [ctypes.c_char_p, vp, vp, lng, ctypes.c_char_p, i]
        if hasattr(R, "ref_mcs_group_cell"):  # (absent from a library prebuilt before round 2)
            R.ref_mcs_group_cell.argtypes = [ctypes.c_char_p, i, i, i, i, i, vp, vp, lng]
        if hasattr(R, "ref_cancel_after"):
            R.ref_cancel_after.argtypes = [i]
            R.ref_progress_get.argtypes = [vp, i]
        R.ref_mcs_save.argtypes = [ctypes.c_char_p, vp, vp, ctypes.c_char_p]
        R.ref_library_create.restype = vp
        R.ref_library_create.argtypes = [i]
        R.ref_library_destroy.argtypes = [vp]
        R.ref_library_add.restype = lng
        R.ref_library_add.argtypes = [vp, vp, i, i, lng, ctypes.c_char_p]
        R.ref_library_set_image_size.argtypes = [vp, i]
        R.ref_library_image_size.argtypes = [vp]
        R.ref_library_count.restype = lng
        R.ref_library_count.argtypes = [vp]
        R.ref_library_get.argtypes = [vp, lng, vp, ctypes.c_char_p, i]
        R.ref_library_remove.argtypes = [vp, lng]
        R.ref_library_save.argtypes = [vp, ctypes.c_char_p]
        R.ref_library_load.argtypes = [vp, ctypes.c_char_p]
        _ref_callbacks = (_CV_CB(_cv_callback), _CODEC_CB(_codec_callback))
        R.ref_set_callbacks(*_ref_callbacks)
        _ref_lib = R
    return _ref_lib


def _group_args(group: CellGroup):
    """What the application hands to CellGroup (setCellShape, setDetail, setSizeSteps): the TOP-LEVEL shape, its mask, the
    detail level in percent and the number of size steps. The per-step cells are derived by the reference's CellGroup.cpp."""
    top = group.cells[0]
    keep = {"shape": np.ascontiguousarray(top.params(), np.int32), "mask": np.ascontiguousarray(top.mask, np.uint8)}
    return keep, keep["shape"].ctypes.data, keep["mask"].ctypes.data, int(round(group.detail * 100)), int(group.size_steps)


class ReferenceGenerator:
    """One CPUPhotomosaicGenerator object of the reference, configured through its own setters."""

    def __init__(self, main_bgr8, lib_bgr8, group: CellGroup, diff_type=RGB_EUCLIDEAN, scheme=SCHEME_NONE, repeat_range=0,
                 repeat_addition=0):
        R = _ref()
        self._main = np.ascontiguousarray(main_bgr8, np.uint8)
        self._lib = np.ascontiguousarray(lib_bgr8, np.uint8)
        self.group = group
        self._keep, shape, mask, pct, steps = _group_args(group)
        self._h = R.ref_session_create(self._main.ctypes.data, self._main.shape[0], self._main.shape[1], self._main.strides[0],
                                       self._lib.ctypes.data, self._lib.shape[0], self._lib.shape[1] if len(self._lib) else 0,
                                       shape, mask, pct, steps, int(diff_type), int(scheme), int(repeat_range), int(repeat_addition))
        if not self._h:
            raise RuntimeError("reference generator: configuration rejected")

    def close(self):
        if self._h:
            _ref().ref_session_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def generate(self, grid_states: list, timing: dict | None = None):
        """setGridState + generateBestFits + getBestFits: (grids, progress values, getMaxProgress())."""
        import time
        n = len(grid_states)
        grids = [np.ascontiguousarray(g, np.int64).copy() for g in grid_states]
        rows = (ctypes.c_int * n)(*[g.shape[0] for g in grids])
        cols = (ctypes.c_int * n)(*[g.shape[1] for g in grids])
        gp = (ctypes.c_void_p * n)(*[g.ctypes.data for g in grids])
        cap = int(sum(g.size for g in grids)) + 8
        prog = (ctypes.c_int * cap)()
        maxp = ctypes.c_int()
        t0 = time.perf_counter()
        k = _ref().ref_session_generate(self._h, n, rows, cols, gp, prog, cap, ctypes.byref(maxp))
        if timing is not None:
            timing["seconds"] = time.perf_counter() - t0
        if k < 0:
            raise RuntimeError("reference generateBestFits failed (%d)" % k)
        assert _ref().ref_message_boxes() == 0
        return grids, list(prog[:min(k, cap)]), maxp.value

    def build_photomosaic(self, background=(0, 0, 0, 0)) -> np.ndarray:
        out = np.empty(self._main.shape[:2] + (4,), np.uint8)
        bg = (ctypes.c_double * 4)(*[float(v) for v in background])
        if _ref().ref_session_build(self._h, bg, out.ctypes.data) != 0:
            raise RuntimeError("reference buildPhotomosaic failed")
        return out

    def get_cell_at(self, step: int, x: int, y: int, n_variants: int):
        ds = self.group.detail_cells[step].size
        cells = np.empty((n_variants, ds, ds, 3), np.float32)
        bounds = (ctypes.c_int * 4)()
        v = _ref().ref_session_get_cell_at(self._h, step, x, y, cells.ctypes.data, bounds)
        if v != n_variants:
            raise RuntimeError("reference getCellAt: %d variants" % v)
        return cells, tuple(bounds)


def reference_generate(main_bgr8: np.ndarray, lib_bgr8: np.ndarray, group: CellGroup, grid_states: list,
                       diff_type: int = RGB_EUCLIDEAN, scheme: int = SCHEME_NONE, repeat_range: int = 0, repeat_addition: int = 0,
                       timing: dict | None = None):
    """The reference's own generator end to end (setters, preprocessing, getCellAt, best-fit loops): (grids per step,
    emitted progress values). timing["seconds"] receives the wall time of setGridState + generateBestFits + getBestFits."""
    g = ReferenceGenerator(main_bgr8, lib_bgr8, group, diff_type, scheme, repeat_range, repeat_addition)
    try:
        grids, progress, _ = g.generate(grid_states, timing)
        return grids, progress
    finally:
        g.close()


def reference_grid_state(group: CellGroup, main_bgr: np.ndarray | None, height: int = 0, width: int = 0) -> list:
    """GridGenerator::getGridState run from the reference's OWN GridGenerator.cpp (+ GridUtility.cpp, GridBounds.cpp)."""
    R = _ref()
    keep, shape, mask, pct, steps = _group_args(group)
    n = steps + 1
    gh = height if main_bgr is None else main_bgr.shape[0]
    gw = width if main_bgr is None else main_bgr.shape[1]
    cap = sum(int(np.prod(grid_size(group.cells[s], gw, gh))) for s in range(n)) + 16
    out = np.empty(cap, np.int64)
    rows, cols = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    img = None if main_bgr is None else np.ascontiguousarray(main_bgr, np.uint8)
    k = R.ref_grid_state(shape, mask, pct, steps, None if img is None else img.ctypes.data,
                         0 if img is None else img.shape[0], 0 if img is None else img.shape[1],
                         0 if img is None else img.strides[0], int(height), int(width), out.ctypes.data, cap, rows, cols)
    if k < 0:
        raise RuntimeError("reference getGridState failed (%d)" % k)
    res, off = [], 0
    for s in range(k):
        res.append(out[off:off + rows[s] * cols[s]].reshape(rows[s], cols[s]).copy())
        off += rows[s] * cols[s]
    return res


def reference_colour_scheme_variants(img_bgr8: np.ndarray, scheme: int) -> list:
    """ColourScheme::getFunction(scheme)(image) from the reference's own ColourScheme.cpp (hue rotations in float HSV_FULL; the
    two cvtColor calls per variant are the real OpenCV)."""
    img = np.ascontiguousarray(img_bgr8, np.uint8)
    out = np.empty((4,) + img.shape, np.uint8)
    v = _ref().ref_colour_scheme_variants(int(scheme), img.ctypes.data, img.shape[0], img.shape[1], img.strides[0], out.ctypes.data, 4)
    if v < 0:
        raise RuntimeError("reference colour scheme failed")
    return [out[i].copy() for i in range(v)]


def _describe_out(size_hint: int):
    params = np.zeros(11, np.int32)
    masks = np.zeros(4 * size_hint * size_hint, np.uint8)
    return params, masks


def _shape_from(params, masks, name="") -> CellShape:
    size = int(params[0])
    m4 = masks[:4 * size * size].reshape(4, size, size)
    sh = CellShape.from_mask(m4[0].copy())
    (sh.row_spacing, sh.col_spacing, sh.alt_row_spacing, sh.alt_col_spacing, sh.alt_row_offset, sh.alt_col_offset) = \
        (int(v) for v in params[1:7])
    sh.alt_col_flip_h, sh.alt_col_flip_v, sh.alt_row_flip_h, sh.alt_row_flip_v = (bool(v) for v in params[7:11])
    sh.name = name
    return sh, m4


def reference_cell_group_cell(group: CellGroup, step: int, detail: bool):
    """Cell `step` (normal or detail) of the CellGroup the reference's own CellGroup.cpp / CellShape.cpp derive from the top-level
    shape: (CellShape, masks4 as the reference's getCellMask returns them)."""
    keep, shape, mask, pct, steps = _group_args(group)
    params, masks = _describe_out(group.cells[0].size)
    if _ref().ref_cell_group_cell(shape, mask, pct, steps, step, int(detail), params.ctypes.data, masks.ctypes.data, masks.size) < 0:
        raise RuntimeError("reference CellGroup failed")
    return _shape_from(params, masks)


def reference_cell_shape_resized(shape: CellShape, new_size: int):
    p = np.ascontiguousarray(shape.params(), np.int32)
    m = np.ascontiguousarray(shape.mask, np.uint8)
    params, masks = _describe_out(max(new_size, 1))
    if _ref().ref_cell_shape_resized(p.ctypes.data, m.ctypes.data, int(new_size), params.ctypes.data, masks.ctypes.data, masks.size) < 0:
        raise RuntimeError("reference CellShape::resized failed")
    return _shape_from(params, masks)


def reference_load_mcs(path: str):
    """CellShape::loadFromFile of the reference (CellShape.cpp:363-434, through its CustomQDataStream.h)."""
    params, masks = _describe_out(2048)
    name = ctypes.create_string_buffer(1024)
    rc = _ref().ref_mcs_load(path.encode(), params.ctypes.data, masks.ctypes.data, masks.size, name, 1024)
    if rc == -1:
        raise ValueError("the reference rejected the .mcs (std::invalid_argument)")
    if rc < 0:
        raise RuntimeError("reference loadFromFile failed")
    return _shape_from(params, masks, name.value.decode())


def reference_mcs_group_cell(path: str, cell_size: int, detail_percent: int, size_steps: int, step: int, detail: bool):
    """loadFromFile -> [resized(cell_size)] -> CellGroup -> getCell(step, detail), all by the reference's own object code."""
    params, masks = _describe_out(2048)
    rc = _ref().ref_mcs_group_cell(path.encode(), int(cell_size), int(detail_percent), int(size_steps), int(step), int(detail),
                                   params.ctypes.data, masks.ctypes.data, masks.size)
    if rc < 0:
        raise RuntimeError("reference CellGroup from .mcs failed (%d)" % rc)
    return _shape_from(params, masks)


def reference_save_mcs(path: str, shape: CellShape, name: str = ""):
    p = np.ascontiguousarray(shape.params(), np.int32)
    m = np.ascontiguousarray(shape.mask, np.uint8)
    if _ref().ref_mcs_save(path.encode(), p.ctypes.data, m.ctypes.data, name.encode()) != 0:
        raise RuntimeError("reference saveToFile failed")


class ReferenceImageLibrary:
    """The reference's own ImageLibrary object (ImageLibrary.cpp compiled unmodified)."""

    def __init__(self, image_size: int):
        self._h = _ref().ref_library_create(int(image_size))

    def close(self):
        if self._h:
            _ref().ref_library_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def add_image(self, im: np.ndarray, name: str = "") -> int:
        im = np.ascontiguousarray(im, np.uint8)
        empty = im.size == 0
        idx = _ref().ref_library_add(self._h, None if empty else im.ctypes.data, 0 if empty else im.shape[0],
                                     0 if empty else im.shape[1], 0 if empty else im.strides[0], name.encode())
        if idx == -1:
            raise ValueError("t_im was empty.")
        if idx < 0:
            raise RuntimeError("reference addImage failed")
        return int(idx)

    def set_image_size(self, size: int):
        assert _ref().ref_library_set_image_size(self._h, int(size)) == 0

    def image_size(self) -> int:
        return _ref().ref_library_image_size(self._h)

    def items(self):
        """[(name, image)] in library order."""
        size, out = self.image_size(), []
        for i in range(_ref().ref_library_count(self._h)):
            img = np.empty((size, size, 3), np.uint8)
            name = ctypes.create_string_buffer(1024)
            got = _ref().ref_library_get(self._h, i, img.ctypes.data, name, 1024)
            assert got == size, (got, size)
            out.append((name.value.decode(), img))
        return out

    def remove(self, i: int):
        assert _ref().ref_library_remove(self._h, i) == 0

    def save(self, path: str):
        if _ref().ref_library_save(self._h, path.encode()) != 0:
            raise RuntimeError("reference ImageLibrary::saveToFile failed")

    def load(self, path: str):
        rc = _ref().ref_library_load(self._h, path.encode())
        if rc == -1:
            raise ValueError("the reference rejected the .mil (std::invalid_argument)")
        if rc != 0:
            raise RuntimeError("reference ImageLibrary::loadFromFile failed")


# ----------------------------------------------------------------------------- buildPhotomosaic

def build_photomosaic(main_shape, lib_bgr8: np.ndarray, group: CellGroup, grids: list, background=(0, 0, 0, 0)) -> np.ndarray:
    """PhotomosaicGeneratorBase::buildPhotomosaic (PhotomosaicGeneratorBase.cpp:110-207): H x W x 4 uint8 BGRA."""
    H, W = main_shape[:2]
    bg = np.asarray(background, np.uint8)
    mosaic = np.empty((H, W, 4), np.uint8)
    mosaic[:] = bg
    mosaic_mask = np.zeros((H, W), np.uint8)
    lib = [cv2.cvtColor(im, cv2.COLOR_BGR2BGRA) for im in lib_bgr8]  # ImageUtility::addAlphaChannel
    for step, grid in enumerate(grids):
        if step != 0:
            s = int(round(0.5 * lib[0].shape[0]))
            lib = [resize_image_exact(im, s, s) for im in lib]  # batchResizeMat(libImg)
        shape = group.cells[step]
        masks4 = shape.masks4()
        step_img = np.empty((H, W, 4), np.uint8)
        step_img[:] = bg
        step_mask = np.zeros((H, W), np.uint8)
        rows, cols = grid.shape
        for gy in range(rows):
            for gx in range(cols):
                if grid[gy, gx] < 0:
                    continue
                r = rect_at(shape, gx - PAD_GRID, gy - PAD_GRID)
                y0, y1 = max(0, min(r[1], H)), max(0, min(r[1] + r[3], H))
                x0, x1 = max(0, min(r[0], W)), max(0, min(r[0] + r[2], W))
                if y0 == y1 or x0 == x1:
                    continue
                ly, lx = y0 - r[1], x0 - r[0]
                m = masks4[flip_at(shape, gx - PAD_GRID, gy - PAD_GRID)][ly:ly + (y1 - y0), lx:lx + (x1 - x0)] != 0
                src = lib[int(grid[gy, gx])][ly:ly + (y1 - y0), lx:lx + (x1 - x0)]
                step_img[y0:y1, x0:x1][m] = src[m]
                step_mask[y0:y1, x0:x1][m] = 255
        if step != 0:
            free = mosaic_mask == 0
            mosaic[free] = step_img[free]
            mosaic_mask[free] = step_mask[free]
        else:
            mosaic, mosaic_mask = step_img, step_mask
    return mosaic
