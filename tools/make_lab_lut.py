#!/usr/bin/env python3
"""Recovers OpenCV's RGB->Lab interpolation table by probing cv2 at the 33^3 grid points k/32
(at a grid point the 4-bit trilinear weights select exactly one table entry), and writes it as
mosaicmagnifique_b200/data/lab_lut_s16.bin: int16 [33 (b)][33 (g)][33 (r)][3 (L, a, b)], little endian.

cvtColor(CV_32F, COLOR_BGR2Lab) -- the call the reference makes in
src/Photomosaic/PhotomosaicGeneratorBase.cpp:241-243, 279-281 -- interpolates this table; it is data
produced by running OpenCV (4.13 here; the reference pins 4.5.2), not OpenCV source.
The table is embedded into libmosaic_b200.so (csrc/lab_lut.S)."""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mosaicmagnifique_b200", "data",
                   "lab_lut_s16.bin")


def main():
    g = np.arange(33, dtype=np.float32) / 32
    B, G, R = np.meshgrid(g, g, g, indexing="ij")
    img = np.stack([B, G, R], -1).reshape(1, -1, 3).astype(np.float32)
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab).reshape(-1, 3).astype(np.float64)
    vals = np.stack([lab[:, 0] * 16384 / 100, (lab[:, 1] + 128) * 64, (lab[:, 2] + 128) * 64], -1)
    assert np.abs(vals - np.rint(vals)).max() == 0, "cv2 output is not on the LUT's integer lattice"
    lut = np.rint(vals).astype("<i2").reshape(33, 33, 33, 3)
    lut.tofile(OUT)
    print(OUT, lut.nbytes, "bytes, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
