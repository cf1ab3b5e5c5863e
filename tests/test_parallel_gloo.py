"""Host-side logic of the multi-GPU path (mosaicmagnifique_b200/parallel.py) on CPU: two gloo ranks all-gather
row blocks of unequal height (the per-cell candidate lists) and must reassemble the full matrix in cell order."""
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


def _worker(rank, world, port, splits, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mosaicmagnifique_b200.parallel import gather_rows
    n_total = splits[-1]
    full_s = torch.arange(n_total * k, dtype=torch.float32).reshape(n_total, k) * 0.5
    full_i = (torch.arange(n_total * k, dtype=torch.int32).reshape(n_total, k) * 7) % 1000
    first, last = splits[rank], splits[rank + 1]
    got_s = gather_rows(full_s[first:last].clone(), first, n_total)
    got_i = gather_rows(full_i[first:last].clone(), first, n_total)
    ok = torch.equal(got_s, full_s) and torch.equal(got_i, full_i) and got_i.dtype == torch.int32
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("splits,k", [((0, 5, 12), 3), ((0, 0, 7), 145), ((0, 9, 9), 1)])
def test_gather_rows_world2(tmp_path, splits, k):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, splits, k, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / ("ok%d.npy" % r))[0]
