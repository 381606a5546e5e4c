"""Synthetic contig generators for the benchmark configs (BASELINE.json configs 4 and 5).

Everything here is deterministic numpy; the same generator feeds bench.py, the
parity tests and the golden-vector script (tests/golden/make_golden.py), so a
contig named ``synth4:k`` is byte-identical everywhere.
"""
from __future__ import annotations

import os
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "data")
_DONOR_FILES = ("NC_001416.1.fasta", "NC_000866.1.fasta", "phiX174.fasta")  # lambda || T4 || phiX
_COMP = bytes.maketrans(b"acgt", b"tgca")
_donor_cache = None


def read_fasta_bytes(path: str):
    """Minimal multi-record FASTA reader -> list of (name, bytes) with the case preserved."""
    out, name, parts = [], None, []
    opener = open
    if path.endswith(".gz"):
        import gzip
        opener = gzip.open
    with opener(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, b"".join(parts)))
                name, parts = line[1:].split()[0].decode() if len(line) > 1 else "", []
            elif name is not None:
                parts.append(line)
    if name is not None:
        out.append((name, b"".join(parts)))
    return out


def donor() -> bytes:
    """Lower-cased lambda || T4 || phiX (222,791 bp), the donor the synthetic windows are cut from."""
    global _donor_cache
    if _donor_cache is None:
        seqs = []
        for f in _DONOR_FILES:
            recs = read_fasta_bytes(os.path.join(_DATA, f))
            seqs.append(recs[0][1].lower())
        _donor_cache = b"".join(seqs)
    return _donor_cache


def synth4_contig(k: int, length: int = 50000) -> bytes:
    """Config-4 contig k: a circular donor window, maybe reverse-complemented, 2 % point mutations."""
    d = donor()
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([20261017, int(k)])))
    off = int(rng.integers(0, len(d)))
    dd = d + d
    while len(dd) < off + length:
        dd += d
    win = dd[off:off + length]
    if rng.random() < 0.5:
        win = win.translate(_COMP)[::-1]
    arr = np.frombuffer(win, dtype=np.uint8).copy()
    mask = rng.random(length) < 0.02
    c = rng.integers(0, 3, size=int(mask.sum()))
    alpha = np.frombuffer(b"acgt", dtype=np.uint8)
    idx = np.nonzero(mask)[0]
    cur = arr[idx]
    # the c-th of the three *other* bases in alphabetical order
    cur_i = np.searchsorted(alpha, cur)
    new_i = c + (c >= cur_i)
    arr[idx] = alpha[new_i]
    return arr.tobytes()


def synth4_batch(n: int, length: int = 50000, first: int = 0):
    """(concatenated uint8 array, int64 offsets[n+1]) for contigs first..first+n-1."""
    offs = np.arange(n + 1, dtype=np.int64) * length
    buf = np.empty(n * length, dtype=np.uint8)
    for i in range(n):
        buf[i * length:(i + 1) * length] = np.frombuffer(synth4_contig(first + i, length), dtype=np.uint8)
    return buf, offs


def synth5_contig(n_windows: int = 200, length: int = 50000) -> bytes:
    """Config 5: one long contig = concatenation of config-4 windows k = 10^6 .. 10^6+n_windows-1."""
    return b"".join(synth4_contig(1000000 + i, length) for i in range(n_windows))


def tile_batch(unique: np.ndarray, uoffs: np.ndarray, n_total: int):
    """Repeat a set of unique contigs cyclically up to n_total contigs (used to build multi-GB batches fast)."""
    nu = len(uoffs) - 1
    lens = np.diff(uoffs)
    sel = np.arange(n_total) % nu
    offs = np.zeros(n_total + 1, dtype=np.int64)
    np.cumsum(lens[sel], out=offs[1:])
    buf = np.empty(int(offs[-1]), dtype=np.uint8)
    for i in range(n_total):
        j = sel[i]
        buf[offs[i]:offs[i + 1]] = unique[uoffs[j]:uoffs[j + 1]]
    return buf, offs


def stress_contigs(n: int = 64, seed: int = 7):
    """Short contigs that exercise end effects, IUPAC codes and N-runs > 500 bp (SURVEY.md §8d stress set)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = donor()
    out = []
    iupac = np.frombuffer(b"nryswkmbvdh", dtype=np.uint8)
    for k in range(n):
        kind = k % 4
        if kind == 0:      # tiny: end effects dominate
            L = int(rng.integers(95, 400))
        elif kind == 1:
            L = int(rng.integers(400, 2500))
        else:
            L = int(rng.integers(2500, 7000))
        if k % 3 == 0:     # iid bases with random GC
            gc = rng.uniform(0.3, 0.7)
            p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
            arr = np.frombuffer(b"acgt", dtype=np.uint8)[rng.choice(4, size=L, p=p)].copy()
        else:              # donor window
            off = int(rng.integers(0, len(d) - L))
            arr = np.frombuffer(d[off:off + L], dtype=np.uint8).copy()
            if rng.random() < 0.5:
                arr = np.frombuffer(arr.tobytes().translate(_COMP)[::-1], dtype=np.uint8).copy()
        if kind >= 1 and k % 2 == 0:   # sprinkle ambiguity codes
            m = rng.random(L) < 0.004
            arr[m] = iupac[rng.integers(0, len(iupac), size=int(m.sum()))]
        if kind == 3 and L > 1500:     # an N-run longer than 500 bp
            a = int(rng.integers(200, L - 900))
            arr[a:a + int(rng.integers(520, 800))] = ord("n")
        if k % 5 == 0:                 # mixed case on input (the reference lower-cases, functions.py:144)
            up = rng.random(L) < 0.5
            arr[up] = arr[up] & 0xDF
        out.append(("stress%d" % k, arr.tobytes()))
    return out


def long_contig(nwin: int = 200) -> bytes:
    """BASELINE.json config 5: ONE contig of nwin x 50 kb (200 -> 10 Mb), the config-4 windows 10**6 .. joined."""
    return b"".join(synth4_contig(10 ** 6 + k) for k in range(nwin))
