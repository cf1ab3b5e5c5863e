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


@pytest.mark.parametrize("world,diff,rr,ra,detail,shape", [(2, 2, 3, 400, 100, None), (3, 0, 0, 0, 50, "Puzzle"), (4, 1, 2, 300, 50, None),
                                                           (8, 2, 2, 500, 50, "Hexagon")])
def test_emulated_ranks_single_collective_and_sharded_inputs(world, diff, rr, ra, detail, shape):
    """The production exchange (parallel.generate_sharded) emulated on one device: every rank uploads ONLY the main-image rows its
    cells read and 1/world of the library (reduced to the detail size on the GPU), the library slices and the candidate blocks are
    concatenated in rank order exactly as the two all-gathers do, and every rank selects from the gathered blocks. Result: bit-identical
    to the unsharded generate() on full inputs."""
    import torch
    from mosaicmagnifique_b200 import CellGroup, CellShape, PhotomosaicGenerator, load_mcs, synthetic
    from mosaicmagnifique_b200.parallel import device_view, split_rows
    cell = 32
    main = synthetic.make_main_image(300, 410, 170 + world, block=32)
    lib = synthetic.make_library(61, cell, 171)
    if shape:
        sh = load_mcs(os.path.join(ROOT, "tests", "golden", "cells", shape + ".mcs")).resized(cell)
    else:
        sh = CellShape(cell)

    def configure(g):
        g.setColourDifference(diff)
        cg = CellGroup()
        cg.setCellShape(sh)
        cg.setDetail(detail)
        g.setCellGroup(cg)
        g.setRepeat(rr, ra)

    ref = PhotomosaicGenerator(0)
    configure(ref)
    ref.setMainImage(main)
    ref.setLibrary(lib)
    state = ref.computeGridState()
    assert ref.generateBestFits()
    want = ref.getBestFits()
    ref.close()

    dev = torch.device("cuda", 0)
    n, per_lib = len(lib), -(-len(lib) // world)
    main_t = torch.from_numpy(main)
    gens, rows = [], []
    for r in range(world):
        g = PhotomosaicGenerator(0)
        configure(g)
        g.setGridState(state)
        g.setShard(r, world)
        lo, hi = g.shardRows(main.shape[0], main.shape[1])
        rows.append((lo, hi))
        g.setMainImageRowsPtr(main_t.data_ptr(), main.shape[0], main.shape[1], main_t.stride(0), lo, hi)
        a, b = min(n, r * per_lib), min(n, (r + 1) * per_lib)
        sl = np.ascontiguousarray(lib[a:b])
        g.setLibraryShardPtr(sl.ctypes.data if b > a else 0, a, b - a, n, cell, world * per_lib)
        gens.append(g)
    assert min(lo for lo, _ in rows) == 0 or world == 1
    assert any(hi - lo < main.shape[0] for lo, hi in rows)  # at least one rank uploads a proper band
    # "all-gather" of the library slices (in place in every generator's buffer)
    infos = [g.libraryDevice() for g in gens]
    stored_row = infos[0]["stored_size"] ** 2 * 3
    assert infos[0]["stored_size"] == (cell * detail) // 100
    views = [device_view(i["ptr"], (world * per_lib, stored_row), torch.uint8, dev) for i in infos]
    for r in range(world):
        for q in range(world):
            if q != r:
                views[q][r * per_lib:(r + 1) * per_lib].copy_(views[r][r * per_lib:(r + 1) * per_lib])
    torch.cuda.synchronize()
    for g in gens:
        g.generateCandidates()
    blks = [g.candidateBlock(0) for g in gens]
    per, parts = split_rows(blks[0]["n_valid"], world, 8 if diff == 2 else 64)
    assert all(b["rows_per_rank"] == per and b["k"] == blks[0]["k"] and b["bytes"] == blks[0]["bytes"] for b in blks)
    for g, (f, c) in zip(gens, parts):
        info = g.candidateInfo(0)
        assert (info["first_cell"], info["n_cells"]) == (f, c)
    gathered = torch.cat([device_view(b["ptr"], (b["bytes"] // 4,), torch.int32, dev).clone() for b in blks]).contiguous()
    torch.cuda.synchronize()
    for g in gens:
        g.selectFromGathered(0, gathered.data_ptr(), blks[0]["k"], per)
        got = g.getBestFits()
        assert np.array_equal(got[0], want[0])
        g.close()


def test_missing_rows_are_reported():
    """A sharded handle whose band does not cover its cells must fail loudly (MOSAIC_ERR_NOT_READY), never read stale rows."""
    import torch
    from mosaicmagnifique_b200 import CellGroup, CellShape, MosaicError, PhotomosaicGenerator, synthetic
    main = synthetic.make_main_image(200, 260, 5, block=32)
    lib = synthetic.make_library(20, 32, 6)
    g = PhotomosaicGenerator(0)
    g.setColourDifference(2)  # CIEDE2000 layout: cell tiles of 8, so rank 1 of 2 owns the lower half of this small grid
    cg = CellGroup()
    cg.setCellShape(CellShape(32))
    g.setCellGroup(cg)
    g.setMainImage(main)
    state = g.computeGridState()
    g.setLibrary(lib)
    g.setGridState(state)
    g.setShard(1, 2)
    lo, hi = g.shardRows(200, 260)
    assert lo > 0
    t = torch.from_numpy(main)
    g.setMainImageRowsPtr(t.data_ptr(), 200, 260, t.stride(0), lo + 8, hi)
    with pytest.raises(MosaicError) as e:
        g.generateCandidates()
    assert e.value.code == -4
    g.setMainImageRowsPtr(t.data_ptr(), 200, 260, t.stride(0), lo, hi)
    g.generateCandidates()
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
