"""The word-parallel node marks (csrc/enum_fwd.inc: st_mark_starts / st_mark_stops -- 64 positions per step through a
shifted-OR window and the carry of an addition) against the candidate-by-candidate statement of the same stage
(mark_word_typed, which PB200_SCAN_REFERENCE forces for every word) and against the oracle.

functions.py:184-251 is the reference loop both restate.  The host build runs the very same stage functions."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from phanotate_b200 import engine  # noqa: E402
from phanotate_b200 import _native as N  # noqa: E402

HOSTSIM = os.path.join(ROOT, "tests", "native", "pb200_hostsim.so")
MINLENS = [9, 30, 60, 90, 91, 92, 120, 132, 135, 300]        # 132 -> 43 codons (the window's limit), 135 and 300 -> statement


def contigs(seed, n, lo, hi):
    """random composition (stop-poor and stop-rich), some with ambiguity codes and N runs, ragged lengths"""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        L = int(rng.integers(lo, hi))
        gc = rng.uniform(0.25, 0.75)
        p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        s = np.frombuffer(b"acgt", dtype=np.uint8)[rng.choice(4, size=L, p=p)].copy()
        if k % 4 == 1:                                         # long open frames: few stops
            for a in range(0, L - 2, 3):
                if bytes(s[a:a + 3]) in (b"taa", b"tag", b"tga") and rng.random() < 0.8:
                    s[a] = ord("c")
        if k % 5 == 2:
            idx = rng.integers(0, L, size=max(1, L // 200))
            s[idx] = np.frombuffer(b"nryswkmbvdh", dtype=np.uint8)[rng.integers(0, 11, size=len(idx))]
        if k % 7 == 3 and L > 900:
            a = int(rng.integers(0, L - 700))
            s[a:a + 600] = ord("n")
        out.append(s.tobytes())
    return out


def tables(e, seqs, minlen, flags):
    r = e.run(seqs, params=engine.make_params(min_orf_len=minlen), flags=flags).fetch_all()
    return r.orfs.tobytes(), r.nodes.tobytes(), r.calls.tobytes(), r.contigs["err"].copy()


def check(e, seqs, minlens):
    for minlen in minlens:
        a = tables(e, seqs, minlen, 0)
        b = tables(e, seqs, minlen, N.SCAN_REFERENCE)
        assert (a[3] == b[3]).all() and not (a[3] & ~np.uint32(16)).any(), minlen    # 16: no path (no ORF that long)
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2], minlen


@pytest.fixture(scope="module")
def sim():
    from helpers import hostsim_path
    hostsim_path()
    e = engine.Engine(0, lib_path=HOSTSIM)
    yield e
    e.close()


def test_word_parallel_marks_equal_the_statement_on_host(sim):
    check(sim, contigs(11, 40, 1, 5000), MINLENS)


def test_word_parallel_marks_at_word_and_run_boundaries_on_host(sim):
    """contig lengths around multiples of the 64-base word and of the 8-word runs the stop-key stage hands its carries
    along (MARK_RUN), packed one behind the other so that contig ends fall on every phase of a run"""
    rng = np.random.default_rng(5)
    seqs = []
    for L in [63, 64, 65, 127, 128, 129, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 9, 10, 3, 700, 1536]:
        s = np.frombuffer(b"acgt", dtype=np.uint8)[rng.integers(0, 4, size=L)]
        seqs.append(s.tobytes())
    check(sim, seqs, [9, 30, 90])
    check(sim, seqs[::-1], [30, 90])


def test_word_parallel_marks_one_long_contig_on_host(sim):
    check(sim, contigs(12, 1, 60000, 60001) + contigs(13, 2, 20000, 30000), [30, 90, 132])


@pytest.mark.parametrize("minlen", [90, 30, 132])
def test_word_parallel_marks_against_the_oracle_on_host(sim, minlen):
    from oracle import phanotate_oracle as O
    seqs = contigs(14 + minlen, 6 if minlen == 90 else 3, 3000, 9000 if minlen == 90 else 6000)
    res = sim.run(seqs, params=engine.make_params(min_orf_len=minlen))
    for k, s in enumerate(seqs):
        want = [tuple(r[:4]) for r in O.call_contig(s.decode(), min_orf_len=minlen)[3]]
        assert [tuple(r) for r in res.call_rows(k)] == want, k


@pytest.mark.gpu
def test_word_parallel_marks_equal_the_statement_on_gpu():
    e = engine.Engine(0)
    try:
        check(e, contigs(21, 300, 1, 20000), [9, 30, 90, 92, 132, 135])
        check(e, contigs(22, 2, 400000, 500000), [90])
    finally:
        e.close()
