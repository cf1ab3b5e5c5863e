"""cancel() and progress(int) at the reference's granularity (CPUPhotomosaicGenerator.cpp:52-88).

The reference polls m_wasCanceled per step, row and cell and emits progress after every grid position with weight
4^(steps-1-step). The engine's difference kernel is one launch per step, so
  * cancel: a flag in mapped pinned memory, read by every CTA when it starts -> a cancelled launch drains in milliseconds and
    generateBestFits() returns false; the flag is sticky like m_wasCanceled;
  * progress: the host polls a device counter of finished tiles while the launch runs and emits values the reference would emit
    too (base + weight * whole grid positions), increasing, ending on the reference's own totals."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _big(diff=2):
    """~60 ms of difference kernel: 4K main, 64 px cells (2,040 valid), 2,500 images."""
    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic
    main = synthetic.make_main_image(2160, 3840, 905)
    lib = synthetic.make_library(2500, 64, 906)
    gen = PhotomosaicGenerator(0)
    gen.setMainImage(main)
    gen.setLibrary(lib)
    gen.setColourDifference(diff)
    cg = CellGroup()
    cg.setCellShape(CellShape(64))
    gen.setCellGroup(cg)
    state = gen.computeGridState()
    gen.setRepeat(2, 300)
    return gen, state


def test_progress_values_are_reference_values():
    gen, state = _big()
    rows, cols = state[0].shape
    seen = []
    gen.setProgressCallback(seen.append)
    assert gen.generateBestFits()
    assert gen.getMaxProgress() == rows * cols
    assert seen[-1] == rows * cols                      # one size step: weight 1, the step total is the last value
    assert all(b > a for a, b in zip(seen, seen[1:]))   # strictly increasing
    assert all(1 <= v <= rows * cols for v in seen)     # each one is a cumulative position count the reference emits as well
    assert len(seen) >= 4, seen                         # the kernel was observed while it ran, not only at its end
    gen.setProgressCallback(None)
    gen.close()


def test_progress_weights_over_size_steps(oracle):
    """Three size levels: the values at the end of each step are the reference's cumulative totals sum_s 4^(S-1-s) * rows_s * cols_s,
    and every value in between is base + weight * k."""
    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, synthetic
    main = synthetic.make_main_image(512, 768, 915, block=32)
    lib = synthetic.make_library(300, 64, 916)
    gen = PhotomosaicGenerator(0)
    gen.setMainImage(main)
    gen.setLibrary(lib)
    gen.setColourDifference(1)
    cg = CellGroup()
    cg.setCellShape(CellShape(64))
    cg.setSizeSteps(2)
    gen.setCellGroup(cg)
    state = gen.computeGridState()
    seen = []
    gen.setProgressCallback(seen.append)
    assert gen.generateBestFits()
    S = len(state)
    base, totals, allowed = 0, [], set()
    for s, st in enumerate(state):
        w = 4 ** (S - 1 - s)
        for k in range(1, st.size + 1):
            allowed.add(base + w * k)
        base += w * st.size
        totals.append(base)
    assert set(totals) <= set(seen) and seen[-1] == totals[-1]
    assert set(seen) <= allowed
    assert all(b > a for a, b in zip(seen, seen[1:]))
    # the reference's own sequence (object code), when it is available here: ours is a subsequence of it
    if oracle.reference_generator_available():
        og = oracle.CellGroup.make(oracle.CellShape.square(64), 100, 2)
        _, ref_progress = oracle.reference_generate(main, lib[:4], og, state, 1, 0, 0, 0)
        assert set(seen) <= set(ref_progress) and seen[-1] == ref_progress[-1]
    gen.close()


def test_cancel_from_the_progress_callback_drains_the_kernel():
    gen, state = _big()
    assert gen.generateBestFits()
    want = gen.getBestFits()[0].copy()
    full_ms = gen.getTimings()["total_ms"]
    t = {}

    def on_progress(v):
        if "cancel" not in t:
            t["cancel"] = time.perf_counter()
            t["value"] = v
            gen.cancel()

    gen.setProgressCallback(on_progress)
    ok = gen.generateBestFits()
    t_ret = time.perf_counter()
    assert ok is False                                   # MOSAIC_ERR_CANCELLED -> false, as the reference
    assert t["value"] < state[0].size                    # cancelled while the kernel was running
    drain_ms = (t_ret - t["cancel"]) * 1e3
    assert drain_ms < max(10.0, 0.35 * full_ms), (drain_ms, full_ms)
    # sticky like m_wasCanceled (never reset by the reference): the next call returns at once
    t0 = time.perf_counter()
    assert gen.generateBestFits() is False
    assert (time.perf_counter() - t0) * 1e3 < 5.0
    gen.setProgressCallback(None)
    gen.resetCancel()
    assert gen.generateBestFits()
    assert np.array_equal(gen.getBestFits()[0], want)
    print("cancel: %.2f ms from cancel() to return (uncancelled generate %.1f ms)" % (drain_ms, full_ms))
    gen.close()


def test_cancel_from_another_thread_and_before_the_call():
    gen, state = _big(diff=0)
    gen.cancel()                                         # before the call: lost by round 1's engine, honoured now
    assert gen.generateBestFits() is False
    gen.resetCancel()
    assert gen.generateBestFits()
    full_ms = gen.getTimings()["total_ms"]
    timer = threading.Timer(full_ms * 0.3e-3, gen.cancel)
    timer.start()
    ok = gen.generateBestFits()
    timer.join()
    # the timer may fire after a very fast run finished; when it fired in time the call must report the cancellation
    if not ok:
        assert gen.generateBestFits() is False
    gen.resetCancel()
    assert gen.generateBestFits()
    gen.close()
