"""Sharded (multi-GPU) path checked on the GPU(s) present.
* world = 2..3 emulated on ONE device: each "rank" is its own generator handle with its own block of grid rows; the
  candidate lists are concatenated in rank order (what the NCCL all-gather does) and every rank runs the selection.
  The result must equal the unsharded generate() bit for bit (same kernels, same sums).
* with >= 2 GPUs the real thing runs under torchrun with NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(gen, main, lib, diff, rr, ra, steps=0, detail=100, cell=32):
    from mosaicmagnifique_b200 import CellGroup, CellShape
    gen.setMainImage(main)
    gen.setLibrary(lib)
    gen.setColourDifference(diff)
    cg = CellGroup()
    cg.setCellShape(CellShape(cell))
    cg.setDetail(detail)
    cg.setSizeSteps(steps)
    gen.setCellGroup(cg)
    gen.setRepeat(rr, ra)
    return gen.computeGridState()


@pytest.mark.parametrize("world,diff,rr,ra,steps", [(2, 2, 3, 400, 0), (3, 0, 2, 10000, 0), (2, 1, 2, 100, 1), (2, 2, 0, 0, 0)])
def test_emulated_ranks_equal_unsharded(world, diff, rr, ra, steps):
    import torch
    from mosaicmagnifique_b200 import PhotomosaicGenerator, synthetic
    from mosaicmagnifique_b200.parallel import device_view
    main = synthetic.make_main_image(260, 330, 70 + world, block=32)
    lib = synthetic.make_library(70, 64 if steps else 32, 71)
    cell = 64 if steps else 32
    ref = PhotomosaicGenerator(0)
    state = _setup(ref, main, lib, diff, rr, ra, steps, 100, cell)
    assert ref.generateBestFits()
    want = ref.getBestFits()
    ref.close()

    gens = []
    for r in range(world):
        g = PhotomosaicGenerator(0)
        _setup(g, main, lib, diff, rr, ra, steps, 100, cell)
        g.setShard(r, world)
        g.generateCandidates()
        gens.append(g)
    dev = torch.device("cuda", 0)
    for step in range(len(state)):
        infos = [g.candidateInfo(step) for g in gens]
        k = infos[0]["k"]
        assert all(i["k"] == k for i in infos)
        assert infos[0]["first_cell"] == 0 and sum(i["n_cells"] for i in infos) == infos[0]["n_valid"]
        for a, b in zip(infos[:-1], infos[1:]):
            assert a["first_cell"] + a["n_cells"] == b["first_cell"]
        scores = torch.cat([device_view(i["scores_ptr"], (i["n_cells"], k), torch.float32, dev).clone() for i in infos]).contiguous()
        idx = torch.cat([device_view(i["indices_ptr"], (i["n_cells"], k), torch.int32, dev).clone() for i in infos]).contiguous()
        torch.cuda.synchronize()
        for g in gens:
            if infos[0]["n_valid"]:
                g.selectFromCandidates(step, scores.data_ptr(), idx.data_ptr(), k)
    for g in gens:
        got = g.getBestFits()
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        g.close()


def test_torchrun_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--workload", "cfg4-small", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert '"n_gpus": 2' in out.stdout
