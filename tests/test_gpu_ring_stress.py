"""Stress / determinism test of the TMA + mbarrier rings of the two difference kernels.

racecheck cannot see the ordering that `mbarrier.arrive.expect_tx` / `cp.async.bulk ... complete_tx` / `try_wait.parity` establish
between the bulk-copy writes and the consumers' shared-memory reads, so the protocol is exercised instead: a second build of the SAME
sources (`make -C mosaicmagnifique_b200/csrc stress` -> libmosaic_b200_stress.so) with
  * 2-stage rings (MM_STAGES=2, MM_ESTAGES=2): every stage is reused after ONE other chunk,
  * 32-pixel chunks for the CIEDE2000 kernel (MM_KP=32): 2,048 chunks per 256 px cell (Euclidean kernel: 16-pixel chunks, 4,096),
  * pseudo-random sleeps in the producer lane and in every consumer warp (MM_STRESS_SKEW=1), so the warps drift apart as far as
    the full / empty barriers let them.
A hole in the protocol (a stage overwritten before all eight consumer warps released it, or read before its bytes landed) would show
as run-to-run differences or as wrong sums. Required: three runs bit-identical; sums equal to the f64 oracle within the parity
tolerance and to the production build within FP32 re-association error."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRESS_SO = os.path.join(ROOT, "mosaicmagnifique_b200", "libmosaic_b200_stress.so")

SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic, capi
assert %(tag)r in capi()._name
main = synthetic.make_main_image(768, 1024, 77, block=64)
lib = synthetic.make_library(40, 256, 78)
out = {}
for diff in (2, 0):
    gen = PhotomosaicGenerator(0)
    gen.setMainImage(main); gen.setLibrary(lib); gen.setColourDifference(diff)
    cg = CellGroup(); cg.setCellShape(CellShape(256)); gen.setCellGroup(cg)
    gen.computeGridState(); gen.setRepeat(1, 100); gen.setKeepDifferences(True)
    runs = []
    for _ in range(3):
        assert gen.generateBestFits()
        runs.append((gen.getDifferences(0).copy(), gen.getBestFits()[0].copy()))
    for D, g in runs[1:]:
        assert np.array_equal(D.view(np.uint32), runs[0][0].view(np.uint32)), "difference sums changed between runs"
        assert np.array_equal(g, runs[0][1])
    out["D%%d" %% diff] = runs[0][0]
    out["g%%d" %% diff] = runs[0][1]
    gen.close()
np.savez(%(out)r, **out)
print("ok")
"""


def _run(tmp_path, so, tag):
    out = str(tmp_path / (tag + ".npz"))
    env = dict(os.environ, MOSAIC_B200_LIB=so)
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "tag": os.path.basename(so), "out": out}], env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-3000:]
    return np.load(out)


def test_two_stage_ring_with_skewed_warps_is_deterministic_and_exact(oracle, tmp_path):
    if not os.path.exists(STRESS_SO):
        pytest.skip("libmosaic_b200_stress.so not built (make -C mosaicmagnifique_b200/csrc stress)")
    from mosaicmagnifique_b200 import library_path, synthetic
    from tests.helpers.parity import check_differences
    stress = _run(tmp_path, STRESS_SO, "stress")
    prod = _run(tmp_path, library_path(), "prod")
    main = synthetic.make_main_image(768, 1024, 77, block=64)
    lib = synthetic.make_library(40, 256, 78)
    og = oracle.CellGroup.make(oracle.CellShape.square(256), 100, 0)
    states = oracle.grid_state(og, main)
    for diff in (2, 0):
        want = oracle.generate(main, lib, og, states, diff, 0, 1, 100, want_D=True)[0]
        Ds, Dp = stress["D%d" % diff], prod["D%d" % diff]
        assert Ds.shape == want.D.shape == Dp.shape == (12, 40)
        check_differences(Ds, want.D)
        check_differences(Dp, want.D)
        # same terms, other chunking / summation order: FP32 re-association only
        assert (np.abs(Ds.astype(np.float64) - Dp) / Dp).max() < 2e-6
        assert np.array_equal(stress["g%d" % diff], want.grid) and np.array_equal(prod["g%d" % diff], want.grid)
