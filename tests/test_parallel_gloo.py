"""Host-side logic of the multi-GPU path (mosaicmagnifique_b200/parallel.py) on CPU: two gloo ranks exchange their candidate
blocks with ONE all_gather_into_tensor (equal block sizes from the deterministic split) and must see the blocks of all ranks in
rank order; the split itself is checked against the rule generator.cu::make_plans implements."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_valid, tile, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mosaicmagnifique_b200.parallel import gather_blocks, split_rows
    per, parts = split_rows(n_valid, world, tile)
    # the full candidate matrix every rank would end up with, and this rank's block {scores [per][k], indices [per][k]} as int32 words
    full_s = (torch.arange(n_valid * k, dtype=torch.float32).reshape(n_valid, k) * 0.5)
    full_i = (torch.arange(n_valid * k, dtype=torch.int32).reshape(n_valid, k) * 7) % 1000
    first, count = parts[rank]
    blk_s = torch.full((per, k), -1.0)
    blk_i = torch.full((per, k), -1, dtype=torch.int32)
    blk_s[:count] = full_s[first:first + count]
    blk_i[:count] = full_i[first:first + count]
    local = torch.cat([blk_s.view(torch.int32).reshape(-1), blk_i.reshape(-1)])
    got = gather_blocks(local).reshape(world, 2, per, k)
    ok = True
    for r, (f, c) in enumerate(parts):
        ok = ok and torch.equal(got[r, 0, :c].view(torch.float32), full_s[f:f + c]) and torch.equal(got[r, 1, :c], full_i[f:f + c])
    # cell c of the raster list lives in block c // per at row c % per (what select_kernel computes)
    for c in range(n_valid):
        ok = ok and torch.equal(got[c // per, 1, c % per], full_i[c])
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([bool(ok)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_valid,tile,k", [(12, 8, 3), (7, 8, 145), (130, 64, 1), (2040, 8, 5)])
def test_gather_blocks_world2(tmp_path, n_valid, tile, k):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_valid, tile, k, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / ("ok%d.npy" % r))[0]


@pytest.mark.parametrize("n_valid,world,tile", [(2040, 8, 8), (2040, 1, 8), (11700, 8, 64), (5, 8, 8), (0, 4, 8), (64, 2, 64), (65, 2, 64)])
def test_split_rows(n_valid, world, tile):
    from mosaicmagnifique_b200.parallel import split_rows
    per, parts = split_rows(n_valid, world, tile)
    assert per % tile == 0 and per >= tile and len(parts) == world
    assert sum(c for _, c in parts) == n_valid
    pos = 0
    for r, (f, c) in enumerate(parts):
        assert f == min(n_valid, r * per) and 0 <= c <= per
        assert f == pos or c == 0
        pos += c
    # balanced to one tile: no rank owns more than ceil(tiles / world) tiles
    assert per == max(1, -(-(-(-n_valid // tile)) // world)) * tile


def test_split_rows_is_the_engines_split():
    """parallel.split_rows (what the Python side assumes about the gathered buffer) against the engine's own arithmetic
    (generator.cu::shard_split through mosaic_host_shard_split), for both kernels' cell tiles."""
    import ctypes

    from mosaicmagnifique_b200 import capi
    from mosaicmagnifique_b200.parallel import split_rows
    L = capi()
    rng = np.random.default_rng(8)
    cases = [(2040, 8), (11664, 8), (1999, 8), (5, 8), (0, 3), (64, 2), (65, 2), (13823, 4)] + [(int(rng.integers(0, 30000)), int(rng.integers(1, 17))) for _ in range(200)]
    for n_valid, world in cases:
        for diff, tile in ((2, 8), (0, 64), (1, 64)):
            per, parts = split_rows(n_valid, world, tile)
            for r in range(world):
                p, f, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
                assert L.mosaic_host_shard_split(n_valid, diff, r, world, ctypes.byref(p), ctypes.byref(f), ctypes.byref(c)) == 0
                assert (p.value, f.value, c.value) == (per, parts[r][0], parts[r][1]), (n_valid, world, diff, r)
    assert L.mosaic_host_shard_split(10, 2, 3, 3, None, None, None) == -1
