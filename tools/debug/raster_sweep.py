"""Debug / tuning: config-4 difference kernel under different rasters (MM_SPLITK, MM_SB_A, MM_SB_B are read at every launch).
    python tools/debug/raster_sweep.py [reps]      -> diff_ms per setting
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:diff_sum --csv python tools/debug/raster_sweep.py 1"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
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
SETTINGS = [(0, 0, 0), (1, 16, 16), (2, 32, 8), (3, 32, 8), (4, 32, 8), (2, 16, 16), (1, 24, 12)]  # 0 = the default plan
base = None
for sk, a, b in SETTINGS:
    os.environ["MM_SPLITK"], os.environ["MM_SB_A"], os.environ["MM_SB_B"] = str(sk), str(a), str(b)
    ts = []
    for _ in range(reps):
        assert gen.generateBestFits()
        ts.append(gen.getTimings()["diff_ms"])
    g = gen.getBestFits()[0]
    chk = int(g.sum())
    print("splitk %d  sb %2d x %2d : diff_ms %s  checksum %d" % (sk, a, b, " ".join("%.2f" % t for t in ts), chk), flush=True)
