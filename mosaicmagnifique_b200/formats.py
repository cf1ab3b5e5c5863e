"""Readers / writers for the reference's on-disk containers (host-side file formats, not on the hot path).

  .mil image library   ImageLibrary::saveToFile / loadFromFile   src/ImageLibrary/ImageLibrary.cpp:117-236
  .mcs cell shape      CellShape::saveToFile / loadFromFile      src/CellShape/CellShape.cpp:321-434
  cv::Mat in a stream  CustomQDataStream                         src/Other/CustomQDataStream.h:22-87

Both are QDataStream (Qt_5_0) streams: big-endian integers, QString = u32 byte length + UTF-16BE (0xFFFFFFFF = null),
QByteArray = u32 length + bytes (0xFFFFFFFF = null). Images are PNG-encoded (.mil >= v6, .mcs) or raw
(type, rows, cols, bytes) in older .mil files. PNG coding uses cv2 like the reference uses cv::imencode / imdecode.
"""
from __future__ import annotations

import struct

import numpy as np

MIL_MAGIC = 0xADBE2480  # ImageLibrary.h:13
MIL_VERSION = 6
MIL_VERSION_ENCODED = 6
MCS_MAGIC = 0x87AECFB1
MCS_VERSION = 8


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def u32(self) -> int:
        v = struct.unpack_from(">I", self.d, self.o)[0]
        self.o += 4
        return v

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.d, self.o)[0]
        self.o += 4
        return v

    def bool(self) -> bool:
        v = self.d[self.o] != 0
        self.o += 1
        return v

    def bytes_(self) -> bytes:
        n = self.u32()
        if n == 0xFFFFFFFF:
            return b""
        v = self.d[self.o:self.o + n]
        if len(v) != n:
            raise ValueError("truncated stream")
        self.o += n
        return v

    def qstring(self) -> str:
        return self.bytes_().decode("utf-16-be")

    def mat(self, png: bool) -> np.ndarray:
        import cv2
        if png:
            img = cv2.imdecode(np.frombuffer(self.bytes_(), np.uint8), cv2.IMREAD_UNCHANGED)
            if img is None:
                raise ValueError("image in stream is not a valid PNG")
            return img
        mat_type, rows, cols = self.u32(), self.u32(), self.u32()
        raw = self.bytes_()
        depth, cn = mat_type & 7, (mat_type >> 3) + 1
        dtype = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}[depth]
        return np.frombuffer(raw, dtype).reshape(rows, cols, cn).squeeze().copy()


class _Writer:
    def __init__(self):
        self.parts = []

    def u32(self, v):
        self.parts.append(struct.pack(">I", v))

    def i32(self, v):
        self.parts.append(struct.pack(">i", v))

    def bool(self, v):
        self.parts.append(b"\x01" if v else b"\x00")

    def bytes_(self, b: bytes):
        self.u32(len(b))
        self.parts.append(b)

    def qstring(self, s: str):
        self.bytes_(s.encode("utf-16-be"))

    def mat_png(self, img: np.ndarray):
        import cv2
        ok, buf = cv2.imencode(".png", img)
        if not ok:
            raise ValueError("PNG encoding failed")
        self.bytes_(buf.tobytes())

    def data(self) -> bytes:
        return b"".join(self.parts)


def load_mil(path: str, magic: int | None = MIL_MAGIC):
    """Returns (images N x S x S x 3 uint8 BGR, names, image_size). Images are stored already cropped square and resized
    to image_size (ImageLibrary::addImage, ImageLibrary.cpp:62-86)."""
    r = _Reader(open(path, "rb").read())
    file_magic = r.u32()
    if magic is not None and file_magic != magic:
        raise ValueError("File is not a valid .mil")
    version = r.u32()
    if version > MIL_VERSION:
        raise ValueError(".mil uses a newer file version")
    if version < 4:
        raise ValueError(".mil uses an outdated file version")
    size, n = r.u32(), r.u32()
    images, names = [], []
    for _ in range(n):
        img = r.mat(png=version >= MIL_VERSION_ENCODED)
        names.append(r.qstring())
        if img.ndim == 2:
            img = np.repeat(img[..., None], 3, axis=2)
        images.append(np.ascontiguousarray(img[..., :3]))
    # versions < 5 are shuffled on load by the reference (random_device): order is not reproducible there either
    lib = np.stack(images) if images else np.zeros((0, size, size, 3), np.uint8)
    return lib, names, size


def save_mil(path: str, images: np.ndarray, names=None, magic: int = MIL_MAGIC):
    w = _Writer()
    w.u32(magic)
    w.u32(MIL_VERSION)
    w.u32(images.shape[1] if len(images) else 0)
    w.u32(len(images))
    for i, img in enumerate(images):
        w.mat_png(img)
        w.qstring(names[i] if names else "image%d" % i)
    open(path, "wb").write(w.data())


def load_mcs(path: str):
    """Returns the fields of a .mcs cell shape: dict(name, mask, row_spacing, col_spacing, alt_row_spacing, alt_col_spacing,
    alt_row_offset, alt_col_offset, alt_col_flip_h, alt_col_flip_v, alt_row_flip_h, alt_row_flip_v, version)."""
    r = _Reader(open(path, "rb").read())
    if r.u32() != MCS_MAGIC:
        raise ValueError("File is not a valid .mcs")
    version = r.u32()
    name = r.qstring()
    mask = r.mat(png=True)
    if mask.ndim == 3:
        mask = np.ascontiguousarray(mask[..., 0])
    keys = ["row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset", "alt_col_offset"]
    out = {"name": name, "mask": mask, "version": version}
    for k in keys:
        out[k] = r.i32()
    for k in ["alt_col_flip_h", "alt_col_flip_v", "alt_row_flip_h", "alt_row_flip_v"]:
        out[k] = r.bool()
    return out


def save_mcs(path: str, fields: dict):
    w = _Writer()
    w.u32(MCS_MAGIC)
    w.u32(MCS_VERSION)
    w.qstring(fields.get("name", ""))
    w.mat_png(fields["mask"])
    for k in ["row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset", "alt_col_offset"]:
        w.i32(int(fields[k]))
    for k in ["alt_col_flip_h", "alt_col_flip_v", "alt_row_flip_h", "alt_row_flip_v"]:
        w.bool(bool(fields[k]))
    open(path, "wb").write(w.data())
