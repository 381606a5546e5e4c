"""Behaviour around contigs the run cannot finish (round-1 advisor findings): a contig without a source->target path, a
contig whose edge weights leave the exact range, stale lazy tables, and batching of very large inputs.  Host build of the
stage functions (tests only); the CLI is driven like the reference's (phanotate.py:24-77)."""
import io
import os

import numpy as np
import pytest

from helpers import hostsim_path, seq_of
from phanotate_b200 import _native as N
from phanotate_b200 import engine, fastio
from phanotate_modules import functions

FILL = "taat" * 650                       # stop codons in all six frames, no start codon: 2,600 bp without any ORF
ORF = "atg" + "gcagaaaaactggct" * 8 + "taa"


@pytest.fixture()
def sim():
    e = engine.Engine(0, lib_path=hostsim_path())
    functions.set_engine(e)
    yield e
    functions.set_engine(None)
    e.close()


def _fasta(tmp_path, recs):
    p = tmp_path / "in.fa"
    p.write_text("".join(">%s\n%s\n" % (n, s) for n, s in recs))
    return str(p)


def test_contig_without_a_path_prints_its_header_and_the_run_goes_on(sim, tmp_path, capsys):
    """one ORF more than 2 kb from both ends: no source or target edge (functions.py:444-451) -> no path; the tabular
    fast path used to abort the whole run there while the other formats printed an empty block"""
    import phanotate
    lone = FILL + ORF + FILL
    res = sim.run([lone])
    assert int(res.contigs[0]["err"]) == N.ERR_NOPATH and res.n_calls == 0
    path = _fasta(tmp_path, [("a", seq_of("phiX174")), ("lonely", lone), ("c", seq_of("phiX174"))])
    out = {}
    for fmt in ("tabular", "genbank"):
        assert phanotate.main([path, "-f", fmt]) == 0
        out[fmt] = capsys.readouterr().out
    blocks = out["tabular"].split("#id:\t")[1:]
    assert [b.split("\n")[0] for b in blocks] == ["a", "lonely", "c"]
    assert blocks[1].count("\n") == 2                       # header lines only
    assert blocks[0].replace("\ta\t", "\tc\t").split("\n")[1:] == blocks[2].split("\n")[1:] and blocks[0].count("\n") > 5
    assert out["genbank"].count("LOCUS") == 3


def _giant_orf_contig(ncodons, seed=3):
    """phiX174 flanks around one ORF of `ncodons` A/T-only codons: its weight grows like 1e58 per kb"""
    rng = np.random.default_rng(seed)
    codons = [c for c in (a + b + c for a in "at" for b in "at" for c in "at") if c not in ("taa", "tta")]
    giant = "atg" + "".join(codons[i] for i in rng.integers(0, len(codons), ncodons)) + "taa"
    return seq_of("phiX174")[:1500] + giant + seq_of("phiX174")[1500:3000]


def check_astronomic_weights(e):
    """ORF weights beyond the 256-bit integers of the ordinary solve (the reference's Decimal and the GMP-backed fastpathz
    have no limit): a 9-kb A/T-only ORF weighs 7.5e174, an 18-kb one 2.5e348 (its %E score is -INF in the reference as
    well: float() overflows).  Such contigs are solved with 2048-bit distances; calls and scores equal the oracle's."""
    from oracle import phanotate_oracle as O
    seqs = [_giant_orf_contig(3000), seq_of("phiX174"), _giant_orf_contig(6000)]
    res = e.run(seqs)
    assert [int(v) for v in res.contigs["err"]] == [0, 0, 0] and [int(v) for v in res.contigs["wide"]] == [1, 0, 1]
    for k in (0, 2):
        want = [tuple(r[:4]) for r in O.call_contig(seqs[k])[3]]
        assert res.call_rows(k) == want, k
    assert any(r[3] == "-7.525584E+174" for r in res.call_rows(0)) and any(r[3] == "-INF" for r in res.call_rows(2))
    res.check()
    assert res.n_huge_weights >= 2                         # (every start of the giant reading frame is such an ORF)
    ws = res.orf_int_weights()                             # the integers the solve used, the 2048-bit ones included
    assert min(ws) < -(1 << 1000) and sum(1 for w in ws if abs(w) >> 240) == res.n_huge_weights


def test_astronomic_orf_weights_are_solved_exactly(sim):
    check_astronomic_weights(sim)


@pytest.mark.gpu
def test_astronomic_orf_weights_are_solved_exactly_on_gpu():
    e = engine.Engine(0)
    try:
        check_astronomic_weights(e)
    finally:
        e.close()


def test_contig_beyond_the_exact_range_is_left_out_not_fatal(sim, tmp_path, capsys):
    """a 60-kb A/T-only ORF weighs ~1e1160: beyond even the 2048-bit integers.  The contig is flagged, Result.check raises
    PhanotateError for it, and the CLI reports it and still prints the other contigs."""
    import phanotate
    big = _giant_orf_contig(20000)
    res = sim.run([big])
    assert int(res.contigs[0]["err"]) & (N.ERR_OVERFLOW | N.ERR_RANGE)
    with pytest.raises(engine.PhanotateError):
        res.check(0)
    path = _fasta(tmp_path, [("a", seq_of("phiX174")), ("giant", big), ("c", seq_of("lambda"))])
    for fmt in ("tabular", "genbank"):
        assert phanotate.main([path, "-f", fmt]) == 0
        cap = capsys.readouterr()
        assert "giant" in cap.err and "left out" in cap.err
        if fmt == "tabular":
            assert [b.split("\n")[0] for b in cap.out.split("#id:\t")[1:]] == ["a", "c"]
        else:
            assert cap.out.count("LOCUS") == 2


def test_reference_exceptions_still_raise(sim, tmp_path):
    import phanotate
    path = _fasta(tmp_path, [("a", seq_of("phiX174")), ("bad", seq_of("phiX174")[:900] + "x" + seq_of("phiX174")[900:2000])])
    with pytest.raises(KeyError):                            # functions.py:20-24 (rev_comp) on a letter outside the alphabet
        phanotate.main([path])
    res = sim.run([seq_of("phiX174")])
    res.contigs["err"][0] = N.ERR_LOOKUP                     # Orfs.get_orf's ValueError (orfs.py:62-69)
    with pytest.raises(ValueError):
        res.check(0)


def test_lazy_tables_of_a_previous_batch_are_refused(sim):
    r1 = sim.run([seq_of("phiX174")])
    r2 = sim.run([seq_of("lambda")])
    with pytest.raises(engine.PhanotateError):
        r1.orfs
    with pytest.raises(engine.PhanotateError):
        r1.orf_int_weights()
    assert len(r2.orfs) == r2.n_orfs and len(r2.nodes) == r2.n_nodes
    r3 = sim.run([seq_of("phiX174")]).fetch_all()
    sim.run([seq_of("lambda")])
    assert len(r3.orfs) == r3.n_orfs and len(r3.edges) > 0  # fetched before the next run: kept


def test_cli_cuts_a_large_input_into_batches(sim, tmp_path, capsys, monkeypatch):
    import phanotate
    assert phanotate.batches([5, 5, 5, 20, 1, 1], 10) == [(0, 2), (2, 3), (3, 4), (4, 6)]
    assert phanotate.batches([], 10) == [] and phanotate.batches([3], 1) == [(0, 1)]
    recs = [("a", seq_of("phiX174")), ("b", seq_of("stress13")), ("c", seq_of("phiX174")[:3000]), ("d", seq_of("stress27"))]
    path = _fasta(tmp_path, recs)
    outs = []
    for limit in ("1000000000", "6000", "1"):
        monkeypatch.setenv("PB200_MAX_BATCH_BASES", limit)
        for fmt in ("tabular", "genbank"):
            assert phanotate.main([path, "-f", fmt]) == 0
            outs.append((fmt, capsys.readouterr().out))
    for fmt in ("tabular", "genbank"):
        texts = [t for f, t in outs if f == fmt]
        assert texts[0] == texts[1] == texts[2] and len(texts[0]) > 1000
