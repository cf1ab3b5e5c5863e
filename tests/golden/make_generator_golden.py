#!/usr/bin/env python3
"""Writes tests/golden/generator_golden.npz: small generator-level fixtures (inputs AND the oracle's outputs) so that
  * the oracle itself is pinned against drift (a different cv2 / numpy / compiler must reproduce these numbers), and
  * the CUDA path can be checked on the GPU box against committed numbers, not only against a live oracle run.
The reference holds no golden best-fit grid or difference sum, and its generator cannot be built
here (Qt + OpenCV C++), so these vectors come from the oracle restatement (oracle/oracle.py + oracle/mosaic_oracle.c, whose
colour maths, grid geometry and selection rule ARE pinned on the reference's own vectors / object code).

    python tests/golden/make_generator_golden.py          (cv2 4.13.0, numpy 2.3 at the time of writing)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mosaicmagnifique_b200 import synthetic  # noqa: E402  (numpy-only input generator)
from oracle import oracle  # noqa: E402

SHAPE_FIELDS = ("row_spacing", "col_spacing", "alt_row_spacing", "alt_col_spacing", "alt_row_offset", "alt_col_offset",
                "alt_col_flip_h", "alt_col_flip_v", "alt_row_flip_h", "alt_row_flip_v")


def cases():
    sq = oracle.CellShape.square(32)
    tri = oracle.CellShape.from_mask(synthetic.triangle_mask(64))
    tri.row_spacing = tri.alt_row_spacing = 64
    tri.col_spacing = tri.alt_col_spacing = 32
    tri.alt_col_flip_v = True
    tri.alt_row_flip_h = True
    hexa = oracle.CellShape.from_mask(synthetic.hexagon_mask(128))
    hexa.row_spacing = hexa.alt_row_spacing = 96
    hexa.col_spacing = hexa.alt_col_spacing = 110
    hexa.alt_row_offset = 55
    #      name                 h    w   lib shape             diff detail steps rr ra    scheme
    return [("square_ciede2000", 160, 224, 40, sq,              2,   100,   0,   2, 500,   0),
            ("triangle_rgb",     150, 210, 36, tri.resized(32), 0,   50,    0,   3, 1000,  0),
            ("hexagon_cie76",    170, 230, 32, hexa.resized(32), 1,  100,   1,   2, 100,   0)]


def main():
    oracle.build()
    out = {"names": np.array([c[0] for c in cases()])}
    for i, (name, h, w, n_lib, shape, diff, detail, steps, rr, ra, scheme) in enumerate(cases()):
        main_img = synthetic.make_main_image(h, w, 700 + i, block=32)
        lib = synthetic.make_library(n_lib, 32, 800 + i)
        group = oracle.CellGroup.make(shape, detail, steps)
        states = oracle.grid_state(group, main_img)
        res = oracle.generate(main_img, lib, group, states, diff, scheme, rr, ra, want_D=True)
        out[name + "/main"] = main_img
        out[name + "/lib"] = lib
        out[name + "/mask"] = shape.mask
        out[name + "/shape"] = np.array([int(getattr(shape, f)) for f in SHAPE_FIELDS], np.int64)
        out[name + "/params"] = np.array([diff, detail, steps, rr, ra, scheme], np.int64)
        for s, (st, r) in enumerate(zip(states, res)):
            out["%s/state%d" % (name, s)] = st
            out["%s/grid%d" % (name, s)] = r.grid
            out["%s/D%d" % (name, s)] = r.D
        print(name, [int((st >= 0).sum()) for st in states], "valid cells per step")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "generator_golden.npz"), **out)


if __name__ == "__main__":
    main()
