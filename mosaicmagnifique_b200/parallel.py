"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange steps.

SURVEY.md section 8(e): the cells x library difference tensor shards by cells (raster order, whole cell tiles) with the library
replicated; only the per-cell candidate lists (K = min(N, 2r^2+2r+1) best (score, index) pairs, exact -- see DESIGN.md) cross
the wire, once per size step, for the order-dependent repeat pass, which every rank then runs redundantly.
The reference has no counterpart (single GPU, CUDAPhotomosaicGenerator.h:31).

Collectives per generate(): ONE all_gather_into_tensor per size step. Every rank owns the same number of rows of the padded cell
list (generator.cu::make_plans computes the split on every rank identically), its candidates are one device block
{f32 scores [rows][K], i32 indices [rows][K]}, and the gathered blocks go straight into the selection kernel: no size exchange,
no device->host read, no re-packing (config 4: 2,040 cells x 145 x 8 B = 2.4 MB, latency-bound).
End-to-end inputs are sharded as well: a rank uploads only the main-image rows its cells read and 1/world of the library, which it
first reduces to the detail size on its GPU; the slices are all-gathered in place over NVLink.
torch is plumbing here (device buffers, process group); the compute stays in libmosaic_b200.so.
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
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def gather_blocks(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gathers equally sized 1-D blocks in rank order with one collective. Works on any backend (NCCL on GPU tensors, gloo
    on CPU tensors in the tests)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * local.numel(),), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.reshape(-1), group=group)
    return out


def split_rows(n_valid: int, world: int, tile: int):
    """The engine's shard split (generator.cu::make_plans): every rank owns `per` rows of the padded raster list of valid cells,
    a multiple of the cell tile. Returns (per, [(first, count) per rank])."""
    n_tiles = -(-n_valid // tile)
    per = max(1, -(-n_tiles // world)) * tile
    return per, [(min(n_valid, r * per), min(n_valid, (r + 1) * per) - min(n_valid, r * per)) for r in range(world)]


def set_library_sharded(gen, lib_host: torch.Tensor, rank: int, world: int, group=None) -> int:
    """setLibrary() for a replicated library without N full PCIe uploads: rank r copies only its 1/world slice of the (pinned)
    host library to its GPU, where it is reduced to the detail size when detail != 100 % (the generator would do that anyway),
    and the slices are all-gathered IN PLACE over NVLink inside the generator's library buffer. The cell group must be set.
    Returns the bytes this rank moved host->device."""
    n, size = lib_host.shape[0], lib_host.shape[1]
    device = torch.device("cuda", gen.device)
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    row = size * size * 3
    flat = lib_host.reshape(n, row)
    gen.setLibraryShardPtr(flat[lo:hi].data_ptr() if hi > lo else 0, lo, hi - lo, n, size, world * per)  # stream-synchronised
    info = gen.libraryDevice()
    stored_row = info["stored_size"] ** 2 * 3
    full = device_view(info["ptr"], (world * per * stored_row,), torch.uint8, device)
    dist.all_gather_into_tensor(full, full[rank * per * stored_row:(rank + 1) * per * stored_row], group=group)  # in place
    torch.cuda.current_stream(device).synchronize()
    return (hi - lo) * row


def set_main_image_sharded(gen, main_host: torch.Tensor) -> int:
    """setMainImage() for a sharded handle: uploads only the rows this rank's cells read (cell group, grid state and shard must be
    set). Returns the bytes moved host->device."""
    h, w = main_host.shape[0], main_host.shape[1]
    lo, hi = gen.shardRows(h, w)
    gen.setMainImageRowsPtr(main_host.data_ptr(), h, w, main_host.stride(0), lo, hi)
    return (hi - lo) * w * 3


def generate_sharded(gen, rank: int, world: int, group=None):
    """generateBestFits() across `world` GPUs. Every rank holds the inputs its cells need in its own generator; returns the
    best-fit grids (identical on every rank) and the bytes of the all-gather payload this rank received."""
    gen.setShard(rank, world)
    gen.generateCandidates()  # preprocessing + difference sums + top-K for this rank's cells (stream-synchronised)
    device = torch.device("cuda", gen.device)
    n_steps = len(gen.getBestFits())
    exchanged = 0
    for step in range(n_steps):
        blk = gen.candidateBlock(step)
        if blk["n_valid"] == 0:
            continue
        local = device_view(blk["ptr"], (blk["bytes"] // 4,), torch.int32, device)
        gathered = gather_blocks(local, group)
        torch.cuda.current_stream(device).synchronize()
        exchanged += gathered.numel() * 4
        gen.selectFromGathered(step, gathered.data_ptr(), blk["k"], blk["rows_per_rank"])
    return gen.getBestFits(), exchanged
