"""The drop-in harness on the CPU: the reference's generator API with the reference's own CPU back-end reproduces
oracle.reference_generate, and the B200 back-end behind the same API refuses to run without a CUDA device (no CPU fallback)."""
import numpy as np
import pytest

from tests.helpers import dropin


@pytest.fixture(scope="module")
def env(oracle):
    if not (oracle.reference_generator_available() and dropin.available()):
        pytest.skip("oracle/_ref/libdropin_b200.so not built (reference sources or product library absent)")
    return oracle


def _case(o):
    from mosaicmagnifique_b200 import synthetic
    main = synthetic.make_main_image(130, 170, 501, block=32)
    lib = synthetic.make_library(20, 32, 502)
    group = o.CellGroup.make(o.CellShape.square(32), 100, 0)
    return main, lib, group, o.grid_state(group, main)


def test_reference_cpu_backend_through_the_generator_api(env):
    main, lib, group, states = _case(env)
    rc, grids, mosaic = dropin.run(env, 0, main, lib, group, states, 2, 0, 2, 300, background=(1, 2, 3, 0))
    assert rc == 0
    want, _ = env.reference_generate(main, lib, group, states, 2, 0, 2, 300)
    assert np.array_equal(grids[0], want[0])
    assert np.array_equal(mosaic, env.build_photomosaic(main.shape, lib, group, grids, background=(1, 2, 3, 0)))


def test_b200_backend_has_no_cpu_fallback(env):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by tests/test_gpu_dropin.py")
    main, lib, group, states = _case(env)
    rc, grids, _ = dropin.run(env, 1, main, lib, group, states, 2, 0, 2, 300, want_mosaic=False)
    assert rc == 1  # generateBestFits() == false, like the reference's CUDA back-end without a device
    assert np.array_equal(grids[0], np.ascontiguousarray(states[0], np.int64))  # m_bestFits untouched
