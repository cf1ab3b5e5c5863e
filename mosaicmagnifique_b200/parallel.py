"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the single exchange step.

SURVEY.md section 8(e): the cells x library difference tensor shards by grid rows with the library replicated; only
the per-cell candidate lists (K = min(N, 2r^2+2r+1) best (score, index) pairs, exact -- see DESIGN.md) cross the
wire, once per size step, for the order-dependent repeat pass, which every rank then runs redundantly.
The reference has no counterpart (single GPU, CUDAPhotomosaicGenerator.h:31).

The payload is a few MB (config 4: 2,040 cells x 145 x 8 B = 2.4 MB), i.e. latency-bound: one padded all_gather per
tensor. torch is plumbing here (device buffers, process group); the compute stays in libmosaic_b200.so.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _DevArray:
    """Zero-copy view of a device buffer owned by libmosaic_b200.so (__cuda_array_interface__ v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}


def device_view(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    if ptr is None or 0 in tuple(shape):
        return torch.empty(tuple(shape), dtype=dtype, device=device)
    typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def gather_rows(local: torch.Tensor, first: int, n_total: int, group=None) -> torch.Tensor:
    """All-gathers row blocks of unequal height: rank r owns rows [first_r, first_r + n_r) of an n_total-row matrix.
    Works on any backend (NCCL on GPU tensors, gloo on CPU tensors in the tests)."""
    world = dist.get_world_size(group)
    k = local.shape[1]
    meta = torch.tensor([first, local.shape[0]], dtype=torch.int64, device=local.device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu().tolist()
    max_rows = max(1, max(m[1] for m in metas))
    padded = torch.zeros((max_rows, k), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = torch.empty((n_total, k), dtype=local.dtype, device=local.device)
    covered = 0
    for (f, n), part in zip(metas, parts):
        out[f:f + n] = part[:n]
        covered += n
    if covered != n_total:
        raise RuntimeError("shards cover %d of %d cells" % (covered, n_total))
    return out


def set_library_sharded(gen, lib_host: torch.Tensor, rank: int, world: int, group=None) -> int:
    """setLibrary() for a replicated library without N full PCIe uploads: rank r copies only its 1/world slice of the
    (pinned) host library to its GPU, the slices are all-gathered over NVLink, and the generator takes the device buffer
    (mosaic_set_library accepts device pointers). Returns the bytes this rank moved host->device."""
    n, size = lib_host.shape[0], lib_host.shape[1]
    device = torch.device("cuda", gen.device)
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    row = size * size * 3
    part = torch.zeros((per, row), dtype=torch.uint8, device=device)
    part[:hi - lo].copy_(lib_host[lo:hi].reshape(hi - lo, row), non_blocking=True)
    full = torch.empty((world * per, row), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(full, part, group=group)
    torch.cuda.current_stream(device).synchronize()
    gen.setLibraryPtr(full.data_ptr(), n, size)  # copies (device to device) before `full` is released
    return (hi - lo) * row


def generate_sharded(gen, rank: int, world: int, group=None):
    """generateBestFits() across `world` GPUs. Every rank holds the same inputs in its own generator; returns the
    best-fit grids (identical on every rank). bytes_exchanged is the all-gather payload this rank received."""
    gen.setShard(rank, world)
    gen.generateCandidates()  # preprocessing + difference sums + top-K for this rank's grid rows (stream-synchronised)
    device = torch.device("cuda", gen.device)
    n_steps = len(gen.getBestFits())
    exchanged = 0
    for step in range(n_steps):
        info = gen.candidateInfo(step)
        k, n_local, n_valid = info["k"], info["n_cells"], info["n_valid"]
        scores = device_view(info["scores_ptr"], (n_local, k), torch.float32, device)
        idx = device_view(info["indices_ptr"], (n_local, k), torch.int32, device)
        all_scores = gather_rows(scores, info["first_cell"], n_valid, group).contiguous()
        all_idx = gather_rows(idx, info["first_cell"], n_valid, group).contiguous()
        torch.cuda.current_stream(device).synchronize()
        exchanged += all_scores.numel() * 4 + all_idx.numel() * 4
        if n_valid:
            gen.selectFromCandidates(step, all_scores.data_ptr(), all_idx.data_ptr(), k)
    return gen.getBestFits(), exchanged
