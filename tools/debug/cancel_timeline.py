"""Debug: timeline of progress emissions and of a cancel issued from the first emission."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic

main = synthetic.make_main_image(2160, 3840, 905)
lib = synthetic.make_library(2500, 64, 906)
gen = PhotomosaicGenerator(0)
gen.setMainImage(main); gen.setLibrary(lib); gen.setColourDifference(2)
cg = CellGroup(); cg.setCellShape(CellShape(64)); gen.setCellGroup(cg)
state = gen.computeGridState(); gen.setRepeat(2, 300)
for i in range(2):
    t0 = time.perf_counter(); ok = gen.generateBestFits(); print("plain generate", ok, (time.perf_counter() - t0) * 1e3, gen.getTimings()["diff_ms"])
ev = []
t0 = [0.0]
gen.setProgressCallback(lambda v: ev.append(((time.perf_counter() - t0[0]) * 1e3, v)))
t0[0] = time.perf_counter(); ok = gen.generateBestFits(); t1 = (time.perf_counter() - t0[0]) * 1e3
print("with progress", ok, t1, len(ev), ev[:5], ev[-3:])
ev.clear()
def cb(v):
    ev.append(((time.perf_counter() - t0[0]) * 1e3, v))
    if len(ev) == 1:
        gen.cancel()
gen.setProgressCallback(cb)
t0[0] = time.perf_counter(); ok = gen.generateBestFits(); t1 = (time.perf_counter() - t0[0]) * 1e3
print("cancel at first emission", ok, "returned after", t1, "events", ev[:5])
