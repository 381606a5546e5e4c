"""Host-side ingest and tabular output at batch speed: a vectorised multi-record FASTA reader that
produces the packed batch the engine takes (no per-line Python), and a tabular writer that formats the call table
of a whole batch like Locus.tabular (reference locus.py:39-56).  Everything numeric stays in the CUDA library."""
from __future__ import annotations

import gzip

import numpy as np


def read_fasta_packed(path):
    """-> (names, bases uint8[total], offsets int64[n+1]).  Record name = first word of the '>' line (like the
    reader in phanotate_modules/file.py); sequence = every non-blank byte of the record's other lines, case kept
    (the library lower-cases like functions.py:144)."""
    with open(path, "rb") as fh:
        data = fh.read()
    if str(path).endswith(".gz") or data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    buf = np.frombuffer(data, dtype=np.uint8)
    n = len(buf)
    if n == 0:
        return [], np.zeros(0, np.uint8), np.zeros(1, np.int64)
    nl = buf == 10
    line_start = np.concatenate(([0], np.flatnonzero(nl) + 1))
    line_start = line_start[line_start < n]
    is_hdr = buf[line_start] == ord(">")
    line_id = np.cumsum(nl) - nl                         # line index of every byte (a newline belongs to its own line)
    rec_of_line = np.cumsum(is_hdr) - 1                  # record index of every line (-1 before the first header)
    keep = ~is_hdr[line_id] & (rec_of_line[line_id] >= 0) & ~nl & (buf != 13) & (buf != 32) & (buf != 9)
    bases = buf[keep]
    rec = rec_of_line[line_id[keep]]
    nrec = int(is_hdr.sum())
    offsets = np.zeros(nrec + 1, dtype=np.int64)
    np.cumsum(np.bincount(rec, minlength=nrec), out=offsets[1:])
    names = []
    hs = line_start[is_hdr]
    ends = np.flatnonzero(nl)
    he = ends[np.searchsorted(ends, hs)] if len(ends) else np.full(len(hs), n)
    for a, b in zip(hs.tolist(), he.tolist() if len(ends) else [n] * len(hs)):
        words = data[a + 1:b].split()
        names.append(words[0].decode() if words else "")
    return names, np.ascontiguousarray(bases), offsets


def write_tabular(res, names, out, check=False):
    """Locus.tabular for every contig of a batch result (Result or MergedResult): same bytes as the per-locus writer.
    check=True raises what the reference would have raised for a contig (KeyError / ValueError) when its turn comes,
    after the blocks of the contigs before it were written -- like the reference's per-locus loop."""
    calls, contigs = res.calls, res.contigs
    left = calls["left"].tolist()
    right = calls["right"].tolist()
    strand = calls["strand"].tolist()
    score = calls["score"].tolist()
    parts = []
    for k, name in enumerate(names):
        if check and int(contigs[k]["err"]):
            out.write("".join(parts))
            parts = []
            res.check(k)
        a = int(contigs[k]["call_off"])
        b = a + int(contigs[k]["n_calls"])
        parts.append("#id:\t%s\n#START\tSTOP\tFRAME\tCONTIG\tSCORE\n" % name)
        parts.extend(("%d\t%d\t+\t%s\t%E\n" % (left[i], right[i], name, score[i])) if strand[i] > 0 else
                     ("%d\t%d\t-\t%s\t%E\n" % (right[i], left[i], name, score[i])) for i in range(a, b))
    out.write("".join(parts))
