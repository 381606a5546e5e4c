"""Multi-GPU plumbing: contig sharding and the one collective of the path -- the cross-contig gather
of the per-rank call tables to rank 0.  torch.distributed only (NCCL on the GPUs,
gloo in the CPU tests); nothing here touches the data path."""
from __future__ import annotations

import numpy as np

from . import _native as N


def shard_contigs(lengths, world: int):
    """Longest-processing-time assignment of contigs to ranks (work ~ length).  -> list of index arrays."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    load = np.zeros(world, dtype=np.int64)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.asarray(sorted(p), dtype=np.int64) for p in parts]


def gather_call_tables(mine, n_rows: int, dist, rank: int, world: int):
    """mine: uint8 torch tensor holding n_rows pb200_call records (device of the process group's backend).

    all_gather of the row counts, then a gather of the rows padded to the largest count.  Returns on
    rank 0 a list of uint8 tensors (one per rank, trimmed), elsewhere None.
    """
    import torch
    dev = mine.device
    cnt = torch.tensor([n_rows], device=dev, dtype=torch.int64)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = [int(c.item()) for c in allc]
    width = max(max(counts), 1) * N.CALL.itemsize
    buf = torch.zeros(width, dtype=torch.uint8, device=dev)
    if n_rows:
        buf[:n_rows * N.CALL.itemsize] = mine[:n_rows * N.CALL.itemsize]
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return [o[:c * N.CALL.itemsize] for o, c in zip(out, counts)]


class DeviceCalls:
    """Zero-copy view of the library's device call table for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n_rows: int):
        self.__cuda_array_interface__ = {"shape": (max(n_rows, 1) * N.CALL.itemsize,), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}
