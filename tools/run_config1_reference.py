#!/usr/bin/env python3
"""BASELINE.json configs[0] -- "SampleImages main image + Library/lib.mil, square cells, RGB Euclidean, repeats on, CPU generator
(runs without a GPU)" -- run with the reference's OWN generator object code (oracle/_ref/libref_core.so) on this machine's CPU.

lib.mil is not in the reference checkout, so the library is the substitute the survey prescribes: 214
seeded centre-cropped 128 px patches of the five SampleImages, ingested by the reference's own ImageLibrary::addImage.
Main image: SampleImages/edgar-perez-424673-unsplash.jpg scaled by 0.5 as the reference's tests do (tst_Generator.h:66-137),
cell 128, detail 100 % and 50 %, repeats (20, 10000) as in tst_Generator.h:238. Writes one JSON line per detail level.

    python tools/run_config1_reference.py > profiles/r1_config1_reference_cpu.json        (needs /root/reference)
"""
import glob
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

SAMPLES = "/root/reference/SampleImages"


def main():
    oracle.build()
    files = sorted(glob.glob(os.path.join(SAMPLES, "*.jpg")))
    images = [cv2.imread(f, cv2.IMREAD_COLOR) for f in files]
    main_img = cv2.imread(os.path.join(SAMPLES, "edgar-perez-424673-unsplash.jpg"), cv2.IMREAD_COLOR)
    main_img = cv2.resize(main_img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rng = np.random.default_rng(214)
    lib = oracle.ReferenceImageLibrary(128)
    for i in range(214):
        src = images[i % len(images)]
        side = int(rng.integers(160, 640))
        y, x = int(rng.integers(0, src.shape[0] - side)), int(rng.integers(0, src.shape[1] - side + 1))
        lib.add_image(np.ascontiguousarray(src[y:y + side, x:x + side + int(rng.integers(0, 40))]), "patch%d" % i)
    items = sorted(lib.items())  # the reference inserts at random indices: fix the order by name for reproducibility
    lib.close()
    library = np.stack([im for _, im in items])
    for detail in (100, 50):
        group = oracle.CellGroup.make(oracle.CellShape.square(128), detail, 0)
        t0 = time.perf_counter()
        states = oracle.reference_grid_state(group, main_img)
        t_state = time.perf_counter() - t0
        tm = {}
        grids, _ = oracle.reference_generate(main_img, library, group, states, oracle.RGB_EUCLIDEAN, 0, 20, 10000, timing=tm)
        n_cells = int((states[0] >= 0).sum())
        ds = group.detail_cells[0].size
        # nominal pixel-differences: in-bound pixels of every valid cell x library images (square mask: all active)
        want = oracle.generate(main_img, library, group, states, oracle.RGB_EUCLIDEAN, 0, 20, 10000, want_D=False)
        assert np.array_equal(want[0].grid, grids[0]), "oracle and reference object code disagree"
        print(json.dumps({"config": "BASELINE configs[0]: %dx%d main (SampleImages x0.5), 214-image substitute library @128, square cells, "
                                    "RGB Euclidean, repeats (20, 10000), detail %d%%" % (main_img.shape[1], main_img.shape[0], detail),
                          "impl": "reference object code (oracle/_ref/libref_core.so), 1 thread", "valid_cells": n_cells,
                          "detail_size": ds, "generate_s": tm["seconds"], "grid_state_s": t_state,
                          "nominal_pixel_diffs": want[0].nominal, "visited_pixel_diffs": want[0].visited,
                          "nominal_pixel_diffs_per_s": want[0].nominal / tm["seconds"],
                          "distinct_images_used": int(len(np.unique(grids[0][grids[0] >= 0]))),
                          "machine": "build container CPU (not the GPU box)"}))


if __name__ == "__main__":
    main()
