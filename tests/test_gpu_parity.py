"""Parity of the CUDA path (through the C ABI) with the reference goldens and the oracle -- needs a B200."""
import os

import numpy as np
import pytest

from helpers import INDEX, STRESS, golden_text, seq_of
from phanotate_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from phanotate_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def calls_text(res, k):
    return "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))


def orf_rows(res, k):
    c = res.contigs[k]
    out = {}
    for o in res.orfs[c["orf_off"]:c["orf_off"] + c["n_orfs"]]:
        out[(int(o["start"]), int(o["stop"]), int(o["frame"]))] = (
            int(o["rbs_score"]), str(N.dec_to_decimal(o["pstop"])), str(N.dec_to_decimal(o["weight"])))
    return out


def golden_orf_rows(name):
    out = {}
    for line in golden_text(name, "orfs.csv.gz").splitlines():
        s, e, f, r, p, w = line.split(",")
        out[(int(s), int(e), int(f))] = (int(r), p, w)
    return out


@pytest.mark.parametrize("name", ["phiX174", "lambda", "T4", "synth4_0", "synth4_1"])
def test_fixture_calls_and_orf_tables(eng, name):
    res = eng.run([seq_of(name)])
    c = res.contigs[0]
    g = INDEX[name]
    assert int(c["err"]) == 0
    assert (int(c["n_orfs"]), int(c["n_nodes"]) + 2, int(c["n_calls"])) == (g["n_orfs"], g["n_nodes"], g["n_calls"])
    assert str(N.dec_to_decimal(c["pstop"])) == g["pstop"]
    assert calls_text(res, 0) == golden_text(name, "calls.tsv")          # CDS coordinates, strand, %E score
    assert orf_rows(res, 0) == golden_orf_rows(name)                      # 28-digit pstop and weight of every ORF
    assert int(c["n_ties"]) == 0


def test_stress_set_in_one_batch(eng):
    """64 contigs (tiny, IUPAC, N-runs, mixed case) in a single launch sequence: batching + edge cases."""
    res = eng.run([seq_of(n) for n in STRESS])
    for k, name in enumerate(STRESS):
        assert int(res.contigs[k]["err"]) == 0, name
        assert calls_text(res, k) == golden_text(name, "calls.tsv"), name
        assert orf_rows(res, k) == golden_orf_rows(name), name


def test_batch_equals_single(eng):
    names = ["phiX174", "stress3", "lambda", "stress17"]
    res = eng.run([seq_of(n) for n in names])
    for k, name in enumerate(names):
        assert calls_text(res, k) == golden_text(name, "calls.tsv"), name


def test_invalid_letter_sets_keyerror_bit(eng):
    res = eng.run([b"acgt" * 40 + b"x" + b"acgt" * 40, seq_of("phiX174").encode()])
    assert int(res.contigs[0]["err"]) & N.ERR_CHAR
    with pytest.raises(KeyError):
        res.check(0)
    assert calls_text(res, 1) == golden_text("phiX174", "calls.tsv")


def test_certified_integers_equal_literal_on_gpu(eng):
    """Default mode (certified integer weights, fast.cuh) vs PB200_LITERAL on 48 synthetic 50-kb contigs + fixtures:
    the integers the solver sees, the calls and the Decimal weights of the calls are identical."""
    from phanotate_b200 import synth
    seqs = [synth.synth4_contig(k) for k in range(48)] + [seq_of(n).encode() for n in ("T4", "lambda", "phiX174")]
    fast = eng.run(seqs)
    wf = fast.orf_int_weights()
    ovf = fast.overlap_int_weights()
    gsf, gdf = fast.gap_int_weights()
    lit = eng.run(seqs, literal=True)
    gsl, gdl = lit.gap_int_weights()
    assert np.array_equal(gsf, gsl) and np.array_equal(gdf, gdl)
    assert np.array_equal(ovf, lit.overlap_int_weights())
    assert fast.n_literal_overlaps < fast.n_overlaps // 50
    wl = lit.orf_int_weights()
    lit.orfs                                          # (lazy table: fetched before the context runs the next batch)
    assert wf == wl
    assert fast.n_literal_presolve < fast.n_orfs // 20 and lit.n_literal_presolve == lit.n_orfs
    for col in ("contig", "left", "right", "strand", "score"):
        assert np.array_equal(fast.calls[col], lit.calls[col]), col
    assert fast.n_literal_postsolve < 4
    assert np.array_equal(eng.run(seqs, call_weights=True).calls, lit.calls)
    fast = eng.run(seqs)
    assert int((fast.contigs["err"] != 0).sum()) == 0
    # lazy completion gives the same ORF table as the literal run
    assert np.array_equal(fast.orfs, lit.orfs)


def test_tiled_scan_equals_reference_scan(eng):
    """The block-cooperative scan kernel (scan_tile.cuh) against the per-strip statement of the stage
    (PB200_SCAN_REFERENCE): contig statistics incl. the RBS background histogram, every ORF, every call.
    The batch mixes 50-kb contigs, the fixtures and the 64 stress contigs (tiny, IUPAC, N runs), so tiles
    with one and with many contig segments are both exercised."""
    from phanotate_b200 import synth
    seqs = ([synth.synth4_contig(k) for k in range(8)] + [seq_of(n).encode() for n in STRESS] +
            [seq_of(n).encode() for n in ("T4", "phiX174", "lambda")] + [seq_of(n).encode() for n in STRESS[:20]])
    a = eng.run(seqs).fetch_all()
    b = eng.run(seqs, flags=N.SCAN_REFERENCE)
    assert np.array_equal(a.contigs, b.contigs)
    assert np.array_equal(a.calls, b.calls)
    assert np.array_equal(a.orfs, b.orfs)
    assert np.array_equal(a.nodes, b.nodes)


def test_wide_solver_on_gpu(eng):
    """256-bit distances forced for every contig give the same calls as the default 128-bit solve."""
    seqs = [seq_of(n).encode() for n in ("T4", "lambda", "phiX174")] + [seq_of(n).encode() for n in STRESS[:16]]
    a = eng.run(seqs)
    b = eng.run(seqs, flags=N.SOLVE_WIDE)
    assert np.array_equal(a.calls, b.calls)
    assert int(a.contigs["wide"].sum()) == 0 and int(b.contigs["wide"].sum()) == len(seqs)


def test_empty_and_tiny_contigs_inside_a_batch(eng):
    """Zero-length and few-base contigs between ordinary ones: they get no calls (an empty contig is flagged), their
    neighbours are unaffected -- with the tiled and with the per-strip scan."""
    seqs = [seq_of("phiX174").encode(), b"", b"acg", seq_of("stress3").encode(), b"", b"atgaaataa", seq_of("lambda").encode(), b"n"]
    for flags in (0, N.SCAN_REFERENCE):
        res = eng.run(seqs, flags=flags)
        assert calls_text(res, 0) == golden_text("phiX174", "calls.tsv")
        assert calls_text(res, 3) == golden_text("stress3", "calls.tsv")
        assert calls_text(res, 6) == golden_text("lambda", "calls.tsv")
        for k in (1, 2, 4, 5, 7):
            assert int(res.contigs[k]["n_calls"]) == 0
        assert int(res.contigs[1]["err"]) & N.ERR_RANGE and int(res.contigs[0]["err"]) == 0 and int(res.contigs[6]["err"]) == 0


def _oracle_rows(seq):
    from oracle import phanotate_oracle as O
    try:
        return [r[:4] for r in O.call_contig(seq)[3]]
    except KeyError:
        return "KeyError"


def test_random_ragged_batch_fast_paths_equal_plain_paths(eng):
    """300 random contigs of ragged length (60 .. 9000 bp, random composition, sprinkled IUPAC codes, lower/upper case):
    the default run (tiled scan, certified weights, 128-bit solve) against the plain statement of every stage at once
    (PB200_SCAN_REFERENCE | PB200_LITERAL | PB200_SOLVE_WIDE): calls, ORF table, node table."""
    rng = np.random.default_rng(20261017)
    seqs = []
    for k in range(300):
        n = int(rng.integers(60, 9000))
        gc = rng.uniform(0.25, 0.75)
        p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
        s = rng.choice(np.frombuffer(b"acgt", dtype=np.uint8), size=n, p=p)
        m = rng.random(n) < 0.002
        s[m] = rng.choice(np.frombuffer(b"nrykmswbdhv", dtype=np.uint8), size=int(m.sum()))
        if k % 3 == 0:
            s = np.frombuffer(s.tobytes().upper(), dtype=np.uint8)
        seqs.append(s.tobytes())
    a = eng.run(seqs)
    wa = a.orf_int_weights()
    a.orfs, a.nodes                                   # (lazy tables: fetched before the context runs the next batch)
    b = eng.run(seqs, literal=True, flags=N.SCAN_REFERENCE | N.SOLVE_WIDE)
    assert wa == b.orf_int_weights()
    for col in ("contig", "left", "right", "strand", "score"):
        assert np.array_equal(a.calls[col], b.calls[col]), col
    assert np.array_equal(a.orfs, b.orfs) and np.array_equal(a.nodes, b.nodes)
    assert np.array_equal(a.contigs["err"], b.contigs["err"]) and np.array_equal(a.contigs["n_calls"], b.contigs["n_calls"])
    assert a.n_calls > 1000
    # ... and against the oracle (CPU restatement of the reference, pinned to its goldens): every contig's call table
    from oracle import phanotate_oracle as O
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(min(16, os.cpu_count() or 1)) as pool:
        want = list(pool.map(_oracle_rows, [s.decode() for s in seqs], chunksize=8))
    for k in range(len(seqs)):
        if int(a.contigs[k]["err"]) & N.ERR_CHAR:
            assert want[k] == "KeyError", k
        else:
            assert a.call_rows(k) == want[k], k
    # the windowed 128-bit sweep against its plain statement: same parents, same tie counts, same calls
    p = eng.run(seqs, flags=N.SOLVE_PLAIN)
    assert np.array_equal(a.calls, p.calls) and np.array_equal(a.contigs, p.contigs)


def test_all_bench_contigs_match_the_oracle_digest(eng):
    """ALL 10,000 contigs of the bench workload (BASELINE config 4: 0.5 Gbp, ~1,000 of them with an exact tie in the
    shortest path) in one batch: the md5 of every contig's call table against the oracle's
    (tests/golden/synth4_calls_digest.json, a derived golden made by tests/golden/make_synth4_digest.py).  The default
    run over all of them; the plain 128-bit sweep and the two-contigs-per-warp variant over the first 3,000."""
    import hashlib
    import json
    import os
    from helpers import GOLDEN
    from phanotate_b200 import synth
    g = json.load(open(os.path.join(GOLDEN, "synth4_calls_digest.json")))
    for flags, half, n in ((0, False, len(g["md5_16"])), (N.SOLVE_PLAIN, False, 3000), (0, True, 3000)):
        bases, offs = synth.synth4_batch(n, 50000)
        # half: two contigs per warp, forced here through the environment
        os.environ.pop("PB200_SOLVE_HALF", None)
        if half:
            os.environ["PB200_SOLVE_HALF"] = "1"
        try:
            res = eng.run_packed(bases, offs, flags=flags)
        finally:
            os.environ.pop("PB200_SOLVE_HALF", None)
        assert int((res.contigs["err"] != 0).sum()) == 0
        assert [int(v) for v in res.contigs["n_calls"]] == g["n_calls"][:n]
        bad = []
        for k in range(n):
            text = "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))
            if hashlib.md5(text.encode()).hexdigest()[:16] != g["md5_16"][k]:
                bad.append(k)
        assert bad == [], (flags, half, bad[:10])


def test_caller_owned_device_buffers(eng):
    """PB200_INPUT_DEVICE: letters and offsets already in device memory (the scan then takes plain loads instead of bulk
    TMA copies).  A 16-byte aligned buffer and one shifted by 3 bytes give the calls of the host-buffer run."""
    import ctypes
    import torch
    from phanotate_b200.engine import Result, make_params
    names = ["phiX174", "lambda", "stress2", "stress19", "synth4_0", "stress12", "T4"]
    seqs = [seq_of(n).encode() for n in names]
    want = eng.run(seqs)
    want_rows = [want.call_rows(k) for k in range(len(seqs))]
    want_hist = want.contigs["background_rbs"].copy()
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    flat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    params = make_params()
    d_off = torch.from_numpy(offs).cuda()
    for shift in (0, 3):
        d_buf = torch.zeros(len(flat) + 64, dtype=torch.uint8, device="cuda")
        d_buf[shift:shift + len(flat)] = torch.from_numpy(flat.copy()).cuda()
        torch.cuda.synchronize()
        eng._ck(eng.lib.pb200_run(eng.ctx, ctypes.c_void_p(d_buf.data_ptr() + shift), ctypes.c_void_p(d_off.data_ptr()),
                                  len(seqs), params.ctypes.data, N.INPUT_DEVICE))
        got = Result(eng)
        assert [got.call_rows(k) for k in range(len(seqs))] == want_rows, shift
        assert np.array_equal(got.contigs["background_rbs"], want_hist), shift
        assert int((got.contigs["err"] != 0).sum()) == 0


@pytest.mark.parametrize("lanes", [1, 3, 4, 7])
def test_pipelined_engine_on_the_gpu_equals_one_context(eng, lanes):
    """The end-to-end path of bench.py: several contexts / streams / host threads on one GPU, ordered uploads from pinned
    host buffers, rows copied out lane by lane, contig numbering applied on the device.  Same tables as one batch in one
    context -- for a ragged batch (tiny contigs, tie contigs, T4), for a second, smaller batch that reuses the pinned
    output buffers, and with the batch left resident."""
    from phanotate_b200.engine import PipelinedEngine
    names = (["phiX174", "synth4_1055"] + STRESS[:20] + ["lambda", "synth4_1999"] + STRESS[20:40] +
             ["T4", "synth4_0", "synth4_2006"] + STRESS[40:])
    seqs = [seq_of(n).encode() for n in names]
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    want = eng.run_packed(bases, offs)
    pe = PipelinedEngine(0, lanes=lanes)
    try:
        pe.pin(bases)
        for rep in range(2):
            got = pe.run_packed(bases, offs)
            assert np.array_equal(got.calls, want.calls) and np.array_equal(got.contigs, want.contigs), rep
        res = pe.run_packed(bases, offs, resident=True)
        assert np.array_equal(res.calls, want.calls) and np.array_equal(res.contigs, want.contigs)
        k = len(names) // 3
        part = pe.run_packed(bases[:offs[k]], offs[:k + 1])
        assert part.n_contigs == k and np.array_equal(part.calls, want.calls[:part.n_calls])
        assert np.array_equal(part.contigs["n_calls"], want.contigs["n_calls"][:k])
        pe.unpin(bases)
    finally:
        pe.close()
