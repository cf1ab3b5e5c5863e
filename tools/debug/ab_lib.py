"""Debug: config-4 diff kernel time with the library given in MOSAIC_B200_LIB (A/B of two builds inside one GPU call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic, capi
import numpy as np
if os.path.exists("/tmp/ab_main.npy"):
    main, lib = np.load("/tmp/ab_main.npy"), np.load("/tmp/ab_lib.npy")
else:
    main = synthetic.make_main_image(4320, 7680, 2004)
    lib = synthetic.make_library(10000, 128, 1004)
    np.save("/tmp/ab_main.npy", main); np.save("/tmp/ab_lib.npy", lib)
gen = PhotomosaicGenerator(0)
gen.setMainImage(main); gen.setLibrary(lib); gen.setColourDifference(2)
cg = CellGroup(); cg.setCellShape(CellShape(128)); gen.setCellGroup(cg)
gen.computeGridState(); gen.setRepeat(8, 500)
ts = []
for _ in range(4):
    assert gen.generateBestFits(); ts.append(gen.getTimings()["diff_ms"])
print(os.path.basename(capi()._name), "splitk", os.environ.get("MM_SPLITK", "default"), " ".join("%.2f" % t for t in ts), int(gen.getBestFits()[0].sum()), flush=True)
