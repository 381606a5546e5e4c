"""The CPU oracle pinned against the reference's own outputs (CPU only, no GPU)."""
import pytest

from oracle import phanotate_oracle as O
from helpers import INDEX, STRESS, golden_text, md5, seq_of

README_ROWS = [  # /root/reference/README.md:47-54, the only known-answer output in the reference tree
    (100, 627, "+", "-4.827981E+02"), (687, 1622, "+", "-4.857517E+06"),
    (1686, 3227, "+", "-3.785434E+10"), (3224, 3484, "+", "-3.779878E+02")]


def _run(name, **kw):
    orfs, nodes, edges, rows = O.call_contig(seq_of(name), **kw)
    return ("".join(O.orf_table_lines(orfs)), "".join(O.edge_dump_lines(edges)), "".join(O.calls_lines(rows)),
            orfs, nodes, edges, rows)


def test_readme_known_answer_rows():
    rows = _run("phiX174")[6]
    assert [r[:4] for r in rows[:4]] == README_ROWS


@pytest.mark.parametrize("name", ["phiX174", "lambda", "synth4_1"])
def test_full_tables_match_reference(name):
    orf_txt, edge_txt, call_txt = _run(name)[:3]
    g = INDEX[name]
    assert orf_txt == golden_text(name, "orfs.csv.gz")        # 28-digit pstop and weight per ORF
    assert md5(edge_txt) == g["edges_md5"]                    # the --dump text, order included
    assert call_txt == golden_text(name, "calls.tsv")
    assert (g["n_orfs"], g["n_edges"], g["n_calls"]) == (orf_txt.count("\n"), edge_txt.count("\n"), call_txt.count("\n"))


def test_phix_edge_dump_text():
    assert _run("phiX174")[1] == golden_text("phiX174", "edges.txt.gz")


@pytest.mark.parametrize("name", STRESS)
def test_stress_contigs(name):
    orf_txt, edge_txt, call_txt = _run(name)[:3]
    assert orf_txt == golden_text(name, "orfs.csv.gz")
    assert edge_txt == golden_text(name, "edges.txt.gz")
    assert call_txt == golden_text(name, "calls.tsv")


def test_literal_stage_e_equals_memoised():
    dna = seq_of("stress13")
    a = O.orf_table_lines(O.get_orfs(dna, literal=True))
    b = O.orf_table_lines(O.get_orfs(dna, literal=False))
    assert a == b


def test_rbs_closed_form_equals_scalar():
    for name in ("stress0", "stress9", "stress25"):
        dna = seq_of(name).lower()
        bgf, bgr = O.rbs_arrays(dna)
        for i in range(len(dna)):
            w = dna[i:i + 21]
            assert bgf[i] == O.score_rbs(w), (name, i)
            assert bgr[i] == O.score_rbs(O.rev_comp(w)), (name, i)


def test_non_iupac_letter_raises_keyerror():
    with pytest.raises(KeyError):
        O.get_orfs("acgt" * 30 + "x" + "acgt" * 30)


def test_t4_calls():
    # T4 is the dense overlapping-ORF stress case (BASELINE.json config 3): weights reach 6e28
    call_txt = _run("T4")[2]
    assert md5(call_txt) == INDEX["T4"]["calls_md5"]
    assert call_txt == golden_text("T4", "calls.tsv")
