"""CPU-only checks of the boundary: the CUDA library exports every symbol of include/phanotate_b200.h,
struct layouts agree with the binding, the product refuses to run without a GPU, and the host-side
logic (parameter parsing, ordering views) behaves like the reference."""
import ctypes
import os
import re
import subprocess
from decimal import Decimal

import numpy as np
import pytest

import conftest
from phanotate_b200 import _native as N
from phanotate_b200 import engine, mirror
from helpers import INDEX, STRESS, golden_text, md5, seq_of

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HOSTSIM = os.path.join(ROOT, "tests", "native", "pb200_hostsim.so")


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "phanotate_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(N.EXPORTS)
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(N.LIB_PATH)                     # loading needs no GPU
    for sym in declared:
        assert hasattr(lib, sym), sym
    sz = (ctypes.c_int32 * 8)()
    lib.pb200_struct_sizes(sz)
    assert list(sz)[:7] == [N.DEC.itemsize, N.PARAMS.itemsize, N.CALL.itemsize, N.ORF.itemsize, N.NODE.itemsize,
                            N.EDGE.itemsize, N.CONTIG.itemsize]


@pytest.mark.skipif(conftest.HAVE_GPU, reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_gpu():
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def test_start_codon_weights_match_reference_normalisation():
    w = engine.parse_start_codons(engine.DEFAULT_START_CODONS)          # file_handling.py:58-62
    assert w["atg"] == Decimal(1)
    assert str(w["gtg"]) == "0.1176470588235294117647058824"
    assert str(w["ttg"]) == "0.05882352941176470588235294118"
    p = engine.make_params()
    assert int(p["n_start"][0]) == 3 and int(p["min_orf_len"][0]) == 90
    assert N.dec_to_decimal(p["start_weight"][0][1]) == w["gtg"]


@pytest.fixture(scope="module")
def sim():
    """Host build of the SAME stage functions the kernels run (tests only; see pb200.cu header)."""
    from helpers import hostsim_path
    hostsim_path()
    e = engine.Engine(0, lib_path=HOSTSIM)
    yield e
    e.close()


def _texts(res, k):
    calls = "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))
    return calls, "".join(mirror.orf_table_lines(res, k)), "".join(mirror.ContigGraph(res, k).dump_lines())


def test_stage_logic_on_host_matches_reference_goldens(sim):
    """Stage functions (host build) on the 64 stress contigs in ONE batch: calls, ORF table in
    Orfs.iter_orfs() order and the --dump edge text in Graph.iteredges() order."""
    res = sim.run([seq_of(n) for n in STRESS]).fetch_all()
    for k, name in enumerate(STRESS):
        assert int(res.contigs[k]["err"]) == 0, name
        calls, orfs, edges = _texts(res, k)
        assert calls == golden_text(name, "calls.tsv"), name
        assert orfs == golden_text(name, "orfs.csv.gz"), name
        assert edges == golden_text(name, "edges.txt.gz"), name


@pytest.mark.parametrize("name", ["phiX174", "lambda", "T4", "synth4_0"])
def test_stage_logic_on_host_fixtures(sim, name):
    res = sim.run([seq_of(name)]).fetch_all()
    calls, orfs, edges = _texts(res, 0)
    g = INDEX[name]
    assert (md5(calls), md5(orfs), md5(edges)) == (g["calls_md5"], g["orfs_md5"], g["edges_md5"])


def test_literal_bellman_ford_entry_point(sim):
    # graph with a tie: two equal-cost paths 0->1->3 and 0->2->3; strict '<' keeps the first found
    src = np.array([0, 0, 1, 2], dtype=np.int32)
    dst = np.array([1, 2, 3, 3], dtype=np.int32)
    w = np.zeros((4, 8), dtype=np.uint32)
    w[:, 0] = [5, 5, 7, 7]
    path = np.zeros(4, dtype=np.int32)
    n = ctypes.c_int32(0)
    rc = sim.lib.pb200_bellman_ford(sim.ctx, 4, 4, src.ctypes.data, dst.ctypes.data, w.ctypes.data, 0, 3,
                                    path.ctypes.data, ctypes.byref(n))
    assert rc == 0 and list(path[:n.value]) == [0, 1, 3]


def check_packed4_input(e, make_pipelined):
    """PB200_INPUT_PACKED4: the batch as 4-bit letters (pb200_pack4) gives the tables of the 1-byte letters -- mixed case,
    IUPAC codes and N runs (the stress contigs), contigs starting on odd offsets (groups of a PipelinedEngine start in the
    middle of a byte), a letter outside the alphabet flagged like the letter itself."""
    names = ["phiX174"] + STRESS[:24] + ["lambda"]
    seqs = [seq_of(n).encode() for n in names]
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    assert any(int(o) & 1 for o in offs)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    a = e.run_packed(bases, offs)
    pk = e.pack4(bases)
    assert len(pk) == (len(bases) + 1) // 2
    b = e.run_packed(pk, offs, packed4=True)
    assert np.array_equal(a.calls, b.calls) and np.array_equal(a.contigs, b.contigs)
    assert np.array_equal(e.run_packed(pk, offs, resident=True).calls, a.calls)       # the expanded letters stay resident
    p = make_pipelined()
    try:
        c = p.run_packed(pk, offs, packed4=True)
        assert np.array_equal(a.calls, c.calls) and np.array_equal(a.contigs["err"], c.contigs["err"])
        # double buffering across batches (pb200_prefetch_async): batch X's run brings batch Y's letters in; Y's run finds
        # them there; a run of something else in between must not use them
        seqs2 = seqs[5:] + seqs[:3]
        offs2 = np.zeros(len(seqs2) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs2], out=offs2[1:])
        bases2 = np.frombuffer(b"".join(seqs2), dtype=np.uint8)
        pk2 = e.pack4(bases2)
        want2 = e.run_packed(bases2, offs2).calls.copy()
        p.pin(pk)
        p.pin(pk2)
        c = p.run_packed(pk, offs, packed4=True, prefetch=(pk2, offs2))
        assert np.array_equal(a.calls, c.calls)
        assert np.array_equal(p.run_packed(pk2, offs2, packed4=True, prefetch=(pk, offs)).calls, want2)
        assert np.array_equal(p.run_packed(pk, offs, packed4=True, prefetch=(pk, offs)).calls, a.calls)
        assert np.array_equal(p.run_packed(pk2, offs2, packed4=True).calls, want2)         # (prefetched: pk, run: pk2)
        assert np.array_equal(p.run_packed(pk, offs, packed4=True).calls, a.calls)
        p.unpin(pk)
        p.unpin(pk2)
        # compact rows (pb200_call24): the same columns without the Decimal weight
        c24 = p.run_packed(pk, offs, packed4=True, compact=True)
        assert c24.calls.dtype == N.CALL24 and c24.calls.itemsize == 24 and len(c24.calls) == len(a.calls)
        for col in ("contig", "left", "right", "strand", "score"):
            assert np.array_equal(c24.calls[col], a.calls[col]), col
        assert [c24.call_rows(k) for k in range(len(seqs))] == [a.call_rows(k) for k in range(len(seqs))]
    finally:
        p.close()
    bad = seqs[0][:900] + b"x" + seqs[0][900:2000]
    r = e.run_packed(e.pack4(np.frombuffer(bad, dtype=np.uint8)), np.array([0, len(bad)], dtype=np.int64), packed4=True)
    assert int(r.contigs[0]["err"]) == N.ERR_CHAR


def test_packed4_input_on_host(sim):
    check_packed4_input(sim, lambda: engine.PipelinedEngine(0, lanes=4, lib_path=HOSTSIM))


@pytest.mark.gpu
def test_prefetch_of_a_larger_batch_on_gpu():
    """pb200_prefetch_async while a run is going, for a next batch that does not fit the packed-letter buffer yet (the
    buffer grows under the running batch), and a prefetched batch that is never run"""
    from phanotate_b200 import synth
    small_b, small_o = synth.synth4_batch(8, 20000)
    big_b, big_o = synth.synth4_batch(40, 50000, first=100)
    e = engine.Engine(0)
    p = engine.PipelinedEngine(0, lanes=4)
    try:
        want_small = e.run_packed(small_b, small_o).calls.copy()
        want_big = e.run_packed(big_b, big_o).calls.copy()
        pk_s, pk_b = p.pack4(small_b), p.pack4(big_b)
        p.pin(pk_s)
        p.pin(pk_b)
        for rep in range(3):
            a = p.run_packed(pk_s, small_o, packed4=True, prefetch=(pk_b, big_o))
            assert np.array_equal(a.calls, want_small), rep
            b = p.run_packed(pk_b, big_o, packed4=True, prefetch=(pk_s, small_o), compact=True)
            for col in ("contig", "left", "right", "strand", "score"):
                assert np.array_equal(b.calls[col], want_big[col]), (rep, col)
        a = p.run_packed(pk_b, big_o, packed4=True)                 # (the small batch was prefetched and is dropped)
        assert np.array_equal(a.calls, want_big)
        p.unpin(pk_s)
        p.unpin(pk_b)
    finally:
        p.close()
        e.close()


@pytest.mark.gpu
def test_packed4_input_on_gpu():
    e = engine.Engine(0)
    try:
        check_packed4_input(e, lambda: engine.PipelinedEngine(0, lanes=4))
    finally:
        e.close()
