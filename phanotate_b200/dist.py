"""Multi-GPU plumbing: contig sharding and the one collective of the path -- the cross-contig gather of the per-rank call
tables to rank 0.  One process per GPU; the gather runs inside the CUDA library on raw NCCL
(csrc/comm.inc, `pb200_comm_*`): no PyTorch.  Nothing here touches the data path -- contigs are independent
(reference phanotate.py:40-56 is a loop over loci)."""
from __future__ import annotations

import ctypes
import os
import time

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


def rank_slices(counts):
    """rows per rank -> [(first row, rows)] of every rank inside the gathered table (rank order, no padding)"""
    out, at = [], 0
    for c in counts:
        out.append((at, int(c)))
        at += int(c)
    return out


def unshard_calls(gathered, counts, parts):
    """Gathered call rows of a run whose contigs were dealt to the ranks by `parts` (shard_contigs): the rows' contig
    column is the index inside the rank's shard; returns the table renumbered to the original contigs and ordered by
    (contig, position in path) -- what one process would have produced for the whole batch."""
    out = np.array(gathered, copy=True)
    for (first, n), idx in zip(rank_slices(counts), parts):
        if n:
            out["contig"][first:first + n] = np.asarray(idx, dtype=np.int64)[gathered["contig"][first:first + n]]
    order = np.argsort(out["contig"], kind="stable")
    return out[order]


class Comm:
    """The ranks' communicator: NCCL inside the library, rendezvous through a file.

        comm = Comm(engine)                      # RANK / WORLD_SIZE / MASTER_PORT from the environment (torchrun sets them)
        counts, total = comm.gather_calls()      # the engine's call table of its last run -> rank 0's HBM
        rows = comm.fetch(0, total)              # rank 0: ... -> host (numpy, pb200_call records)

    Rank 0 makes the NCCL unique id and writes it to `id_path` (default /tmp/pb200_nccl_<launcher pid>_<MASTER_PORT>.id;
    all ranks of one torchrun share the launcher as parent); the others wait for the file."""

    def __init__(self, engine, rank=None, world=None, id_path=None, timeout=300.0):
        self.e = engine
        self.lib = engine.lib
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        if id_path is None:
            id_path = os.environ.get("PB200_NCCL_ID_FILE") or "/tmp/pb200_nccl_%d_%s.id" % (
                os.getppid(), os.environ.get("MASTER_PORT", "0"))
        self.id_path = id_path
        uid = (ctypes.c_uint8 * 128)()
        if self.rank == 0:
            if self.lib.pb200_comm_unique_id(uid) != 0:
                raise RuntimeError("phanotate_b200: NCCL is not available (libnccl.so.2; set PB200_NCCL_LIB)")
            tmp = id_path + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as fh:
                fh.write(bytes(uid))
            os.replace(tmp, id_path)
        else:
            t0 = time.time()
            while True:
                try:
                    data = open(id_path, "rb").read()
                    if len(data) == 128:
                        break
                except OSError:
                    pass
                if time.time() - t0 > timeout:
                    raise RuntimeError("phanotate_b200: no NCCL id at %s after %.0f s" % (id_path, timeout))
                time.sleep(0.01)
            ctypes.memmove(uid, data, 128)
        engine._ck(self.lib.pb200_comm_init(engine.ctx, uid, self.rank, self.world))
        self.barrier()
        if self.rank == 0:
            try:
                os.unlink(id_path)
            except OSError:
                pass
        self.counts = [0] * self.world
        self.total = 0
        self._host = np.zeros(0, dtype=N.CALL)
        self._async = [np.zeros(0, dtype=N.CALL), np.zeros(0, dtype=N.CALL)]
        self._dtype = N.CALL                                       # rows of the last gather: pb200_call or pb200_call24
        self._flip, self._pending = 0, None

    def gather_calls(self, engines=None, compact=False):
        """Call tables -> rank 0's device memory.  engines: the contexts whose tables make up this rank's rows, in order
        (the lanes of a PipelinedEngine); default: the communicator's own engine.  compact: the rows travel and arrive as
        pb200_call24 records (no Decimal weight column; half the bytes).  -> (rows per rank, total rows)"""
        fn = self.lib.pb200_comm_gather_calls24 if compact else self.lib.pb200_comm_gather_calls
        self._dtype = N.CALL24 if compact else N.CALL
        counts = np.zeros(self.world, dtype=np.int64)
        total = ctypes.c_int64(0)
        if engines is None:
            rc = fn(self.e.ctx, None, None, 0, counts.ctypes.data, None, ctypes.byref(total))
        else:
            ptrs, rows = [], []
            for e in engines:
                n = e.sizes()[6]
                if n:
                    ptrs.append(int(e.lib.pb200_device_calls(e.ctx)))
                    rows.append(n)
            if not ptrs:
                ptrs, rows = [0], [0]
            p = (ctypes.c_void_p * len(ptrs))(*ptrs)
            r = np.asarray(rows, dtype=np.int64)
            rc = fn(self.e.ctx, p, r.ctypes.data, len(ptrs), counts.ctypes.data, None, ctypes.byref(total))
        self.e._ck(rc)
        self.counts, self.total = [int(c) for c in counts], int(total.value)
        return self.counts, self.total

    def fetch(self, first=0, n=None):
        """rank 0: rows [first, first+n) of the last gather as a numpy array of pb200_call records -- a view into a
        page-locked buffer the communicator keeps (valid until the next fetch; copy it to keep it)"""
        n = self.total - first if n is None else n
        if n > len(self._host) or self._host.dtype != self._dtype:
            if len(self._host):
                self.e.unpin(self._host)
            self._host = np.zeros(n + n // 4 + 1024, dtype=self._dtype)
            self.e.pin(self._host)                 # page-locked: the device->host copy is plain DMA
        out = self._host[:n]
        if n:
            self.e._ck(self.lib.pb200_comm_fetch_gathered(self.e.ctx, first, n, out.ctypes.data))
        return out

    def fetch_begin(self, first=0, n=None):
        """rank 0: start copying rows [first, first+n) of the last gather to the host on a stream of its own (it runs beside
        the next batch); fetch_wait() returns them.  Two page-locked buffers alternate, so the rows of one gather stay valid
        while the next one is on its way."""
        n = self.total - first if n is None else n
        self._flip ^= 1
        buf = self._async[self._flip]
        if n > len(buf) or buf.dtype != self._dtype:
            if len(buf):
                self.e.unpin(buf)
            buf = np.zeros(n + n // 4 + 1024, dtype=self._dtype)
            self.e.pin(buf)
            self._async[self._flip] = buf
        self._pending = buf[:n]
        self.e._ck(self.lib.pb200_comm_fetch_begin(self.e.ctx, first, n, buf.ctypes.data))

    def fetch_wait(self):
        if self._pending is None:
            return None
        self.e._ck(self.lib.pb200_comm_fetch_wait(self.e.ctx))
        out, self._pending = self._pending, None
        return out

    def allreduce(self, values, op="sum"):
        v = np.asarray(values, dtype=np.float64).copy()
        self.e._ck(self.lib.pb200_comm_allreduce(self.e.ctx, v.ctypes.data, len(v), 1 if op == "max" else 0))
        return [float(x) for x in v]

    def barrier(self):
        self.e._ck(self.lib.pb200_comm_barrier(self.e.ctx))

    def close(self):
        self.fetch_wait()
        for b in self._async:
            if len(b):
                self.e.unpin(b)
        self._async = [np.zeros(0, dtype=N.CALL), np.zeros(0, dtype=N.CALL)]
        if len(self._host):
            self.e.unpin(self._host)
            self._host = np.zeros(0, dtype=N.CALL)
        if self.e.ctx:
            self.lib.pb200_comm_destroy(self.e.ctx)


def _cpulist(text):
    ids = set()
    for part in text.split(","):
        part = part.strip()
        if "-" in part:
            a, b = part.split("-")
            ids.update(range(int(a), int(b) + 1))
        elif part:
            ids.add(int(part))
    return ids


def bind_near_gpu(index: int, world: int = 1) -> dict:
    """Pin this process (and the threads and pinned buffers it creates afterwards) to CPU cores near its GPU: with one
    process per GPU, host<->device copies then stay on the local memory controller instead of crossing the socket
    interconnect.  Sources of the GPU's cores, in order: sysfs (numa_node / local_cpulist of the PCI device), NVML's CPU
    affinity mask, `nvidia-smi topo -m`.  Without any of them (a container that hides the topology) the ranks at least get
    disjoint, equal slices of the allowed cores, so that their host threads do not migrate over each other.  Best effort:
    returns what was done, never raises; says so on stderr when nothing could be bound."""
    import sys
    info = {"gpu": index, "bound": False, "how": None}
    allowed = os.sched_getaffinity(0)
    ids = set()
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        try:
            node = int(open(dev + "/numa_node").read().strip())
            info["numa_node"] = node
            if node >= 0:
                ids = _cpulist(open(dev + "/local_cpulist").read().strip()) & allowed
                info["how"] = "sysfs"
        except OSError:
            pass
        if len(ids) < 4 or len(ids) == len(allowed):
            words = (len(os.sched_getaffinity(0)) + 63) // 64 + 4
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            got = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
            if 4 <= len(got) < len(allowed):
                ids, info["how"] = got, "nvml"
    except Exception as e:       # no NVML: try the tool
        info["error"] = "%s: %s" % (type(e).__name__, e)
    if len(ids) < 4 or len(ids) == len(allowed):
        try:
            import subprocess
            txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            for line in txt.splitlines():
                f = line.split()
                if f and f[0] == "GPU%d" % index:
                    for tok in f[1:]:
                        if tok[0].isdigit() and ("-" in tok or "," in tok):
                            got = _cpulist(tok) & allowed
                            if 4 <= len(got) < len(allowed):
                                ids, info["how"] = got, "nvidia-smi topo"
                            break
        except Exception:
            pass
    if len(ids) < 4 or len(ids) == len(allowed):
        # topology unknown: disjoint equal slices of the allowed cores per rank
        cores = sorted(allowed)
        per = len(cores) // max(world, 1)
        if world > 1 and per >= 8:               # (fewer cores than a rank has host threads: binding would only hurt)
            ids, info["how"] = set(cores[index * per:(index + 1) * per]), "equal slices (topology hidden)"
    if len(ids) >= 4 and len(ids) < len(allowed):
        try:
            os.sched_setaffinity(0, ids)
            info.update(bound=True, cores=len(ids), cpulist="%d-%d" % (min(ids), max(ids)))
        except OSError as e:
            info["error"] = str(e)
    if not info["bound"] and world > 1:
        sys.stderr.write("phanotate_b200: rank for GPU %d is NOT bound to cores near its GPU (%s)\n" % (index, info))
    return info
