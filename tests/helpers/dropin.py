"""ctypes driver of oracle/_ref/libdropin_b200.so (tests/helpers/dropin_harness.cpp): the reference's generator API with a
selectable back-end -- the reference's own CPUPhotomosaicGenerator (0) or integration/B200PhotomosaicGenerator (1)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "oracle", "_ref", "libdropin_b200.so")


def available() -> bool:
    if not os.path.exists(SO):
        return False
    try:
        ctypes.CDLL(SO)
        return True
    except OSError:
        return False


def run(oracle, backend, main, lib, group, grid_states, diff, scheme, rr, ra, background=(0, 0, 0, 0), device=0, want_mosaic=True):
    """Returns (status, grids, mosaic). status 0 = ok, 1 = generateBestFits() returned false."""
    oracle._ref()  # installs the OpenCV callbacks in libref_core.so (the same loaded instance libdropin_b200.so links to)
    from oracle.oracle import _group_args
    L = ctypes.CDLL(SO)
    vp, i, lng = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
    L.dropin_run.argtypes = [i, i, vp, i, i, lng, vp, i, i, vp, vp, i, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    main = np.ascontiguousarray(main, np.uint8)
    lib = np.ascontiguousarray(lib, np.uint8)
    keep, shape, mask, pct, steps = _group_args(group)
    n = len(grid_states)
    grids = [np.ascontiguousarray(g, np.int64).copy() for g in grid_states]
    rows = (ctypes.c_int * n)(*[g.shape[0] for g in grids])
    cols = (ctypes.c_int * n)(*[g.shape[1] for g in grids])
    gp = (ctypes.c_void_p * n)(*[g.ctypes.data for g in grids])
    bg = (ctypes.c_double * 4)(*[float(v) for v in background])
    mosaic = np.zeros(main.shape[:2] + (4,), np.uint8) if want_mosaic else None
    rc = L.dropin_run(backend, device, main.ctypes.data, main.shape[0], main.shape[1], main.strides[0], lib.ctypes.data, lib.shape[0],
                      lib.shape[1], shape, mask, pct, steps, int(diff), int(scheme), int(rr), int(ra), n, rows, cols, gp, bg,
                      None if mosaic is None else mosaic.ctypes.data)
    if rc < 0:
        raise RuntimeError("dropin_run failed (%d)" % rc)
    return rc, grids, mosaic


def cancel_after(oracle, n_emissions: int):
    """Plays the user pressing Cancel at the n-th progress(int) emission (0 = never): the signal body of the harness calls the
    generator's own cancel() slot, as a connected QProgressDialog does."""
    oracle._ref().ref_cancel_after(int(n_emissions))


def progress_values(oracle):
    R = oracle._ref()
    n = R.ref_progress_get(None, 0)
    buf = (ctypes.c_int * max(n, 1))()
    R.ref_progress_get(buf, n)
    return list(buf[:n])


def progress_clear(oracle):
    oracle._ref().ref_progress_clear()
