"""tRNA masking (reference functions.py:457-509 add_trnas + the tRNA branch of the connect loop functions.py:388-399).

Goldens: the reference's own get_graph run with stub `aragorn` / `tRNAscan-SE` executables on PATH that print canned hit
lists in the tools' formats (tests/golden/make_trna_golden.py) -- 14 cases on phiX174 and 8 stress contigs with 0, 1 and up to
8 hits on both strands, next to the contig ends, next to each other, and tRNAscan-SE hits that aragorn's hits mask.  Compared:
the --dump edge text in Graph.iteredges() order (md5) and the path (CDS and tRNA rows)."""
import hashlib
import json
import os
import stat

import numpy as np
import pytest

from helpers import GOLDEN, hostsim_path, seq_of
from phanotate_b200 import _native as N
from phanotate_b200 import engine, mirror

TRNA = json.load(open(os.path.join(GOLDEN, "trna.json")))


def _check_case(e, case, flags=0, literal=False):
    rec = TRNA[case]
    res = e.run([seq_of(rec["contig"])], trnas=[(0, a, b) for a, b in rec["trnas"]], flags=flags, literal=literal).fetch_all()
    assert int(res.contigs[0]["err"]) == 0 and res.n_trnas == len(rec["trnas"]), case
    cg = mirror.ContigGraph(res, 0)
    assert len(cg.node_names) == rec["n_nodes"] and len(cg.edge_src) == rec["n_edges"], case
    assert hashlib.md5("".join(cg.dump_lines()).encode()).hexdigest() == rec["edges_md5"], case
    rows = [[l, r, s, g, sc] for (l, r, s, sc), g in zip(res.call_rows(0), res.call_genes(0))]
    assert rows == rec["calls"], case


def _check_batch(e):
    """all 14 cases as ONE batch (tRNA contigs between plain ones): every contig's path as when it runs alone"""
    cases = sorted(TRNA)
    seqs, trnas = [], []
    for case in cases:
        seqs.append(seq_of(TRNA[case]["contig"]))
        trnas += [(len(seqs) - 1, a, b) for a, b in TRNA[case]["trnas"]]
        seqs.append(seq_of("stress3"))                      # a contig without hits in between
    res = e.run(seqs, trnas=trnas)
    assert int((res.contigs["err"] != 0).sum()) == 0
    plain = e.run([seq_of("stress3")])
    for k, case in enumerate(cases):
        rows = [[l, r, s, g, sc] for (l, r, s, sc), g in zip(res.call_rows(2 * k), res.call_genes(2 * k))]
        assert rows == TRNA[case]["calls"], case
        assert res.call_rows(2 * k + 1) == plain.call_rows(0)
    assert sum(1 for k in range(len(seqs)) for g in res.call_genes(k) if g == "tRNA") == sum(r["n_trna_calls"] for r in TRNA.values())


@pytest.fixture(scope="module")
def sim():
    e = engine.Engine(0, lib_path=hostsim_path())
    yield e
    e.close()


@pytest.mark.parametrize("case", sorted(TRNA))
def test_trna_cases_on_host(sim, case):
    _check_case(sim, case)


def test_trna_literal_and_wide_on_host(sim):
    _check_case(sim, "phiX174_many", literal=True)
    _check_case(sim, "stress6_ends", flags=N.SOLVE_WIDE)
    _check_case(sim, "stress2_rev", flags=N.SOLVE_WIDE, literal=True)


def test_trna_batch_on_host(sim):
    _check_batch(sim)


def test_hits_outside_the_contig_or_unsorted_are_refused(sim):
    res = sim.run([seq_of("stress9")], trnas=[(0, 700, 775)])          # the contig is 409 bp long
    assert int(res.contigs[0]["err"]) & N.ERR_RANGE
    sim.set_trnas([(1, 10, 80), (0, 10, 80)])                           # sorted by contig on the way in
    sim.set_trnas(None)
    with pytest.raises(engine.PhanotateError):
        sim.run([seq_of("stress9")], trnas=[(3, 10, 80)])               # no such contig in the batch


def test_cli_runs_the_trna_programs_like_the_reference(tmp_path, capsys, monkeypatch):
    """stub aragorn / tRNAscan-SE on PATH (the same canned output the goldens were made with): --dump and the genbank
    output of the CLI carry the tRNA nodes, edges and features"""
    import phanotate
    from phanotate_modules import functions
    import sys
    sys.path.insert(0, GOLDEN)
    import make_trna_golden as MG
    d = str(tmp_path / "stubs")
    os.makedirs(d)
    MG.write_stubs(d)
    monkeypatch.setenv("PB200_STUB_DIR", d)
    monkeypatch.setenv("PATH", d + os.pathsep + os.environ["PATH"])
    e = engine.Engine(0, lib_path=hostsim_path())
    functions.set_engine(e)
    try:
        for case in ("phiX174_many", "phiX174_scan_overlap", "phiX174_none"):
            contig, aragorn, scan = MG.CASES[case]
            MG.set_case(d, aragorn, scan)
            assert functions.find_trnas(seq_of(contig)) == TRNA[case]["trnas"]
            fa = tmp_path / "in.fa"
            fa.write_text(">phiX174\n%s\n" % seq_of(contig))
            phanotate.main([str(fa), "--dump"])
            cap = capsys.readouterr()
            assert hashlib.md5(cap.out.encode()).hexdigest() == TRNA[case]["edges_md5"], case
            assert "not found" not in cap.err
        contig, aragorn, scan = MG.CASES["stress6_ends"]
        MG.set_case(d, aragorn, scan)
        fa.write_text(">s6\n%s\n" % seq_of(contig))
        phanotate.main([str(fa), "-f", "genbank"])
        out = capsys.readouterr().out
        assert out.count("     tRNA            ") == TRNA["stress6_ends"]["n_trna_calls"] == 1
        phanotate.main([str(fa)])                            # tabular: CDS rows only (locus.py:41)
        rows = [l for l in capsys.readouterr().out.splitlines() if not l.startswith("#")]
        assert len(rows) == sum(1 for r in TRNA["stress6_ends"]["calls"] if r[3] == "CDS")
        # API mirror driven like the reference: Graph with the tRNA nodes and edges, other_end['t...'] entries
        from test_mirror_and_dist import drive_like_reference
        MG.set_case(d, *MG.CASES["phiX174_one"][1:])
        orf_txt, edge_txt, calls, orfs, graph = drive_like_reference("phiX174")
        assert hashlib.md5(edge_txt.encode()).hexdigest() == TRNA["phiX174_one"]["edges_md5"]
        assert orfs.other_end["t1000"] == 1070 and orfs.other_end["t1070"] == 1000
    finally:
        functions.set_engine(None)
        e.close()


# ---------------------------------------------------------------------------------------- the CUDA path
@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(TRNA))
def test_trna_cases_on_gpu(eng, case):
    _check_case(eng, case)


@pytest.mark.gpu
def test_trna_batch_literal_and_wide_on_gpu(eng):
    _check_batch(eng)
    _check_case(eng, "phiX174_many", literal=True)
    _check_case(eng, "stress6_ends", flags=N.SOLVE_WIDE)
    _check_case(eng, "phiX174_many", flags=N.SOLVE_PLAIN)
