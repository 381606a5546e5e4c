"""The host-side ingest and tabular writer run on several host threads (pb200_fasta_parse / pb200_format_tabular cut the
text into ranges): every thread count gives the bytes the one-thread pass gives, on a file with CRLF lines, blank lines,
empty records, text before the first header and very different record sizes."""
import os

import numpy as np
import pytest

from phanotate_b200 import _native as N, fastio


@pytest.fixture(scope="module")
def lib():
    return N.load()


def odd_fasta(tmp_path):
    rng = np.random.default_rng(3)
    parts = [b"text before the first header\nacgt\n"]
    for k in range(900):
        n = int(rng.integers(0, 40000)) if k % 4 else int(rng.integers(0, 50))
        s = rng.choice(np.frombuffer(b"acgtACGTn", dtype=np.uint8), size=n).tobytes()
        w = int(rng.choice([60, 70, 80, 1000]))
        eol = b"\r\n" if k % 7 == 0 else b"\n"
        parts.append(b">rec%d description %d" % (k, k) + eol + eol.join(s[i:i + w] for i in range(0, n, w)) + eol +
                     (eol if k % 5 == 0 else b""))
    p = tmp_path / "odd.fa"
    p.write_bytes(b"".join(parts))
    return str(p)


def test_threaded_ingest_equals_single_thread_and_the_per_locus_reader(lib, tmp_path, monkeypatch):
    from phanotate_modules.file import File
    path = odd_fasta(tmp_path)
    monkeypatch.setenv("PB200_HOST_THREADS", "1")
    names, bases, offs = fastio.read_fasta_packed(path, lib)
    loci = list(File(path))
    assert [l.name() for l in loci] == names and len(names) == 900
    assert all(l.seq().encode() == bases[offs[k]:offs[k + 1]].tobytes() for k, l in enumerate(loci))
    for t in ("2", "3", "5", "16"):
        monkeypatch.setenv("PB200_HOST_THREADS", t)
        n2, b2, o2 = fastio.read_fasta_packed(path, lib)
        assert n2 == names and np.array_equal(b2, bases) and np.array_equal(o2, offs), t


def test_threaded_tabular_equals_single_thread(lib, monkeypatch):
    n = 700

    class R:
        pass
    r = R()
    rng = np.random.default_rng(9)
    per = rng.integers(0, 300, size=n)
    nc = int(per.sum())
    r.calls = np.zeros(nc, dtype=N.CALL)
    r.calls["left"] = rng.integers(1, 50000, size=nc)
    r.calls["right"] = r.calls["left"] + rng.integers(90, 3000, size=nc)
    r.calls["strand"] = 1 - 2 * rng.integers(0, 2, size=nc)
    r.calls["score"] = -np.exp(rng.uniform(-5, 80, size=nc))
    r.contigs = np.zeros(n, dtype=N.CONTIG)
    r.contigs["n_calls"] = per
    r.contigs["call_off"] = np.concatenate(([0], np.cumsum(per)[:-1]))
    names = ["contig_%d" % k for k in range(n)]
    monkeypatch.setenv("PB200_HOST_THREADS", "1")
    ref = fastio.tabular_text(r, names, lib)
    rows = ref.decode().splitlines()
    assert len(rows) == 2 * n + nc
    k3 = int(np.nonzero(per)[0][3])                      # a contig with calls: its first row, spelled out
    c = r.calls[int(r.contigs["call_off"][k3])]
    fwd = c["strand"] > 0
    want = "%d\t%d\t%s\tcontig_%d\t%E" % (c["left"] if fwd else c["right"], c["right"] if fwd else c["left"],
                                          "+" if fwd else "-", k3, float(c["score"]))
    assert want in rows
    for t in ("2", "7", "16"):
        monkeypatch.setenv("PB200_HOST_THREADS", t)
        assert fastio.tabular_text(r, names, lib) == ref, t


def test_score_text_is_printf_E_exactly(lib):
    """pb200_format_score (the tabular writer's own '%E': 128-bit integer arithmetic, half-even on the exact binary value)
    against Python's '%E' -- what the reference prints (phanotate.py:75-76) -- on random magnitudes, short decimals,
    integers, exact ties at the seventh digit, and the values that take the sprintf route (0, inf, 1e300, denormals)."""
    import ctypes
    import random
    buf = ctypes.create_string_buffer(64)

    def text(x):
        n = lib.pb200_format_score(x, buf)
        return buf.raw[:n].decode()
    rnd = random.Random(7)
    xs = [12345675.0, 1234567.5, 0.5, 1.0, 9.9999995, 99999995.0, 9999999.5, 1e-16, 1.69e38, 123456.75, 1e22, 1e23, 5e-324, 0.0,
          -0.0, float("inf"), float("-inf"), 1e300, -7.525584e174, 2.5e-5, 1.0000005, 1.0000015, -4.827981e2, -20.0]
    for _ in range(60000):
        x = -(10 ** rnd.uniform(-20, 42)) * rnd.uniform(0.1, 1)
        if rnd.random() < 0.3:
            x = -round(abs(x), rnd.randint(0, 8))
        xs.append(x)
    for k in range(1000000, 1000200):
        xs += [k + 0.5, (k + 0.5) * 8, (k + 0.5) / 1024, (k + 0.5) * 1e3]
    bad = [x for x in xs if text(x) != "%E" % x]
    assert bad == [], bad[:5]
