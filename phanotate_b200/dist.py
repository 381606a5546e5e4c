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
    counts = [int(v) for v in torch.stack(allc).flatten().tolist()]      # (one synchronisation, not one per rank)
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


def bind_near_gpu(index: int) -> dict:
    """Pin this process (and the threads and pinned buffers it creates afterwards) to the CPU cores of the NUMA node its
    GPU hangs off: with one process per GPU, host<->device copies then stay on the local memory controller instead of
    crossing the socket interconnect.  Best effort: returns what was done, never raises."""
    import os
    info = {"gpu": index, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read().strip())
        cpus = open(dev + "/local_cpulist").read().strip()
        info.update(numa_node=node, cpulist=cpus)
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if node >= 0 and len(ids) >= 4:
            os.sched_setaffinity(0, ids)
            info["bound"] = True
            info["cores"] = len(ids)
    except Exception as e:       # no NVML, no sysfs, a container without the topology: stay unbound
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info
