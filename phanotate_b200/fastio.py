"""Host-side ingest and tabular output at batch speed: a multi-record FASTA reader that produces the
packed batch the engine takes, and a tabular writer that formats the call table of a whole batch like Locus.tabular
(reference locus.py:39-56).  Both are single passes in the C library (pb200_fasta_parse, pb200_format_tabular)."""
from __future__ import annotations

import ctypes
import gzip
import mmap
import os

import numpy as np

from . import _native as N

_lib = None


def _library(lib=None):
    global _lib
    if lib is not None:
        return lib
    if _lib is None:
        _lib = N.load()
    return _lib


def read_fasta_packed(path, lib=None, pack4=False):
    """-> (names, bases uint8[total], offsets int64[n+1]); pack4=True: bases as 4-bit letters, two per byte (pb200_pack4;
    what Engine.run_packed(..., packed4=True) takes: half the bytes over the host link).  Record name = first word of the '>' line (like the
    reader in phanotate_modules/file.py); sequence = every non-blank byte of the record's other lines, case kept
    (the library lower-cases like functions.py:144)."""
    lib = _library(lib)
    with open(path, "rb") as fh:
        head = fh.read(2)
        if str(path).endswith(".gz") or head == b"\x1f\x8b":
            fh.seek(0)
            data = np.frombuffer(gzip.decompress(fh.read()), dtype=np.uint8)
        else:
            size = os.fstat(fh.fileno()).st_size
            # the text is only read (twice, by the library's host threads): map the file instead of copying it
            data = np.frombuffer(mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ), dtype=np.uint8) if size else np.zeros(0, np.uint8)
    n = len(data)
    ptr = data.ctypes.data if n else None
    nrec = int(lib.pb200_fasta_count(ptr, n))
    bases = np.empty(max(n, 1), dtype=np.uint8)
    offsets = np.zeros(nrec + 1, dtype=np.int64)
    nb, ne = np.zeros(max(nrec, 1), dtype=np.int64), np.zeros(max(nrec, 1), dtype=np.int64)
    got = int(lib.pb200_fasta_parse(ptr, n, bases.ctypes.data, offsets.ctypes.data, nb.ctypes.data, ne.ctypes.data, nrec))
    if got != nrec:
        raise RuntimeError("FASTA parse failed")
    names = [data[a:b].tobytes().decode() for a, b in zip(nb[:nrec].tolist(), ne[:nrec].tolist())]
    total = int(offsets[-1])
    if pack4:
        packed = np.zeros((total + 1) // 2, dtype=np.uint8)
        lib.pb200_pack4(bases.ctypes.data, total, packed.ctypes.data)
        return names, packed, offsets
    return names, bases[:total], offsets


def tabular_text(res, names, lib=None) -> bytes:
    """Locus.tabular for every contig of a batch result (Result or MergedResult), as bytes."""
    lib = _library(lib)
    enc = [s.encode() for s in names]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in enc], out=off[1:])
    blob = b"".join(enc) or b"\0"
    calls = np.ascontiguousarray(res.calls)
    contigs = np.ascontiguousarray(res.contigs)
    # first call: how many bytes at most (returned negated); then one buffer that is never zero-filled or copied twice
    need = -int(lib.pb200_format_tabular(calls.ctypes.data, contigs.ctypes.data, len(enc), blob, off.ctypes.data, None, 0))
    if need <= 0:
        return b""
    out = np.empty(need, dtype=np.uint8)
    got = int(lib.pb200_format_tabular(calls.ctypes.data, contigs.ctypes.data, len(enc), blob, off.ctypes.data,
                                       out.ctypes.data, need))
    if got < 0:
        raise RuntimeError("pb200_format_tabular: buffer too small")
    return out[:got].tobytes()


def write_tabular(res, names, out, check=False, lib=None, skipped=None):
    """Writes tabular_text to a text stream.  check=True handles a contig the run could not finish when its turn comes,
    after the blocks of the contigs before it were written -- like the reference's per-locus loop: what the reference
    would have raised (KeyError / ValueError) is raised; a contig that only this implementation cannot finish
    (PhanotateError: edge weights beyond the exact range) is left out, reported in `skipped` (a list, if given) and on
    stderr, and the contigs after it are still written.  A contig without a source->target path prints its header
    and no rows, as it does with the other output formats."""
    import sys
    from .engine import PhanotateError, Result
    bad = [k for k in range(len(names)) if Result.fatal(res.contigs[k]["err"])] if check else []
    if not bad:
        out.write(tabular_text(res, names, lib).decode())
        return

    class _Part:                                   # a run of consecutive contigs
        pass
    at = 0
    for k in bad + [len(names)]:
        if k > at:
            part = _Part()
            part.contigs = res.contigs[at:k]
            part.calls = res.calls
            out.write(tabular_text(part, names[at:k], lib).decode())
        if k < len(names):
            try:
                res.check(k)
            except PhanotateError as e:
                sys.stderr.write("Warning: %s: %s; contig left out\n" % (names[k], e))
                if skipped is not None:
                    skipped.append(k)
        at = k + 1
