"""Long contigs (BASELINE.json config 5: ONE contig, intra-contig solve; functions.py:360-438 + phanotate.py:56-64 on a
graph of up to 3.6e5 nodes).  The chunked solve (csrc/chunk.cuh) against

* long4  (200 kb): the reference's own get_orfs/get_graph + edge-order Bellman-Ford (tests/golden/make_long_golden.py),
* long20 / long40 / long200 (1, 2, 10 Mb = config 5 itself): the oracle with shortest_path_fast (derived goldens; the
  fast path equals the replayed Bellman-Ford on every fixture, checked below),

and against the one-warp sweep on adversarial contigs (N runs with bridges, tandem repeats, GC 20 % / 80 %), under
several chunk geometries including ones that are bound to fail their check and fall back."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, STRESS, hostsim_path, seq_of
from phanotate_b200 import _native as N
from phanotate_b200 import engine, mirror, synth

LONG = json.load(open(os.path.join(GOLDEN, "long_index.json")))


def _calls_text(res, k=0):
    return "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))


def _md5(t):
    return hashlib.md5(t.encode()).hexdigest()


def _adversarial():
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"acgt", dtype=np.uint8)

    def rnd(n, gc):
        return bytes(acgt[rng.choice(4, n, p=[(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])])
    t4 = seq_of("T4").encode()
    return {
        "T4x2": t4 * 2,
        "gc20": rnd(400000, .2),
        "gc80": rnd(300000, .8),
        "n_runs": b"".join(rnd(30000, .5) + b"n" * int(rng.integers(100, 3000)) for _ in range(8)),
        "tandem": rnd(3000, .5) * 60,
        "tandem_small": rnd(300, .45) * 700,
        "poly": b"".join(rnd(5000, .5) + b"a" * 2000 + rnd(5000, .3) + b"at" * 800 for _ in range(10)),
    }


# ---------------------------------------------------------------------------------------- CPU: oracle and stage logic
@pytest.mark.parametrize("name", ["phiX174", "lambda", "stress13", "stress27", "synth4_0"])
def test_fast_shortest_path_equals_the_replayed_bellman_ford(name):
    from oracle import phanotate_oracle as O
    seq = seq_of(name)
    orfs = O.get_orfs(seq)
    nodes, edges = O.get_graph(orfs)
    s, t = O.ONode(('source', 'source', 0, 0)), O.ONode(('target', 'target', 0, len(seq) + 1))
    a = O.shortest_path(nodes, edges, s, t)
    assert a == O.shortest_path_fast(nodes, edges, s, t)
    if len(seq) < 10000:
        assert a == O.shortest_path_literal(nodes, edges, s, t)


def test_fast_shortest_path_on_contigs_with_exact_ties():
    """the two contigs whose tie goldens come from the reference's own code (tests/golden/make_tie_golden.py)"""
    from oracle import phanotate_oracle as O
    ties = json.load(open(os.path.join(GOLDEN, "ties.json"))) if os.path.exists(os.path.join(GOLDEN, "ties.json")) else {}
    ks = [int(k.split("_")[1]) for k in ties if k.startswith("synth4_")][:4] or [26, 33]
    for k in ks:
        seq = synth.synth4_contig(k).decode()
        orfs = O.get_orfs(seq)
        nodes, edges = O.get_graph(orfs)
        s, t = O.ONode(('source', 'source', 0, 0)), O.ONode(('target', 'target', 0, len(seq) + 1))
        assert O.shortest_path(nodes, edges, s, t) == O.shortest_path_fast(nodes, edges, s, t), k


@pytest.fixture(scope="module")
def sim():
    e = engine.Engine(0, lib_path=hostsim_path())
    yield e
    e.close()


def test_chunked_stage_logic_matches_the_reference_on_200kb(sim):
    """host build of the stage functions, default geometry: call table, ORF table and edge dump of long4"""
    res = sim.run([synth.long_contig(4)]).fetch_all()
    g = LONG["long4"]
    assert res.n_chunks > 0 and res.n_chunk_fallbacks == 0 and int(res.contigs[0]["err"]) == 0
    assert _calls_text(res) == open(os.path.join(GOLDEN, "long4.calls.tsv")).read()
    assert _md5("".join(mirror.orf_table_lines(res, 0))) == g["orfs_md5"]
    assert _md5("".join(mirror.ContigGraph(res, 0).dump_lines())) == g["edges_md5"]


@pytest.mark.parametrize("geo", [(256, 768, 64, 4096), (64, 300, 16, 100), (1000, 2000, 100, 2000), (32, 16, 4, 64), (128, 0, 0, 64)])
def test_chunk_geometry_does_not_change_results(sim, geo):
    """any geometry gives the oracle's calls: geometries with a warm-up too short to forget the stand-in source fail
    their check and are solved again by one sweep"""
    sim.set_chunking(*geo)
    try:
        res = sim.run([synth.long_contig(20)])
    finally:
        sim.set_chunking()
    assert res.n_chunks > 0 and int(res.contigs[0]["err"]) == 0
    if geo[1] < 100:
        assert res.n_chunk_fallbacks == 1
    assert _md5(_calls_text(res)) == LONG["long20"]["calls_md5"]


def test_chunked_equals_one_sweep_on_adversarial_contigs(sim):
    adv = _adversarial()
    names = sorted(adv)
    one = sim.run([adv[k] for k in names], flags=N.SOLVE_NOCHUNK)
    for geo in ((256, 768, 64, 4096), (64, 128, 16, 512)):
        sim.set_chunking(*geo)
        try:
            ch = sim.run([adv[k] for k in names])
        finally:
            sim.set_chunking()
        assert ch.n_chunks > 0
        assert np.array_equal(one.calls, ch.calls), geo
        assert [int(v) for v in one.contigs["err"]] == [int(v) for v in ch.contigs["err"]]


def check_second_attempt(e):
    """sequence with few stops (GC 70-85 %) forgets the stand-in source over 30-60 kb, not 5-10: the default geometry fails
    its check there, the second attempt (four times the warm-up) passes -- no contig falls back to the one-warp sweep"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"acgt", dtype=np.uint8)
    seqs = [bytes(acgt[rng.choice(4, 400000, p=[(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])]) for gc in (.5, .75, .85)]
    seqs.append(synth.long_contig(6))
    one = e.run(seqs, flags=N.SOLVE_NOCHUNK)
    res = e.run(seqs)
    assert res.n_chunks > 0 and res.chunk_second_attempt and res.n_chunk_fallbacks == 0
    assert np.array_equal(one.calls, res.calls) and int((res.contigs["err"] != 0).sum()) == 0
    res = e.run(seqs[:1] + seqs[3:])                       # nothing fails here: one attempt
    assert res.n_chunks > 0 and not res.chunk_second_attempt and res.n_chunk_fallbacks == 0


def test_second_attempt_with_longer_warm_up(tmp_path):
    e = engine.Engine(0, lib_path=hostsim_path())          # (a fresh context: the library picks the geometry)
    try:
        check_second_attempt(e)
    finally:
        e.close()


@pytest.mark.gpu
def test_second_attempt_with_longer_warm_up_on_gpu():
    e = engine.Engine(0)
    try:
        check_second_attempt(e)
    finally:
        e.close()


def test_mixed_batch_of_short_and_long_contigs(sim):
    """long contigs in the middle of a batch of short ones: only they are chunked, every contig's calls stay what they are alone"""
    seqs = [seq_of("phiX174").encode(), synth.long_contig(3), seq_of("lambda").encode(), seq_of("T4").encode(), seq_of(STRESS[5]).encode()]
    seqs = seqs + [synth.synth4_contig(k) for k in range(1100)]      # > 1024 contigs: the long-contig threshold of big batches applies
    res = sim.run(seqs)
    assert int((res.contigs["err"] != 0).sum()) == 0
    assert res.n_chunks == sum((int(n) + 255) // 256 for n in res.contigs["n_nodes"] if n > 4096) and res.n_chunks > 0
    for k in (0, 1, 2, 3, 4, 50):
        alone = sim.run([seqs[k]], flags=N.SOLVE_NOCHUNK)
        assert res.call_rows(k) == alone.call_rows(0), k


# ---------------------------------------------------------------------------------------- GPU: the CUDA path through the C ABI
@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nwin", [4, 20, 40, 200])
def test_long_contig_calls_match_the_goldens_on_gpu(eng, nwin):
    """config 5 (nwin = 200: ONE 10-Mb contig) and three shorter ones: chunked solve and one-warp sweep, both against the golden"""
    g = LONG["long%d" % nwin]
    seq = synth.long_contig(nwin)
    for flags in (0, N.SOLVE_NOCHUNK, N.SOLVE_PLAIN):
        res = eng.run([seq], flags=flags)
        assert int(res.contigs[0]["err"]) == 0 and res.n_calls == g["n_calls"]
        if flags != N.SOLVE_NOCHUNK:
            assert res.n_chunks > 0 and res.n_chunk_fallbacks == 0
        text = _calls_text(res)
        if nwin == 4:
            assert text == open(os.path.join(GOLDEN, "long4.calls.tsv")).read()
        else:
            assert text == gzip.open(os.path.join(GOLDEN, "long%d.calls.tsv.gz" % nwin), "rb").read().decode()
        assert _md5(text) == g["calls_md5"]


@pytest.mark.gpu
def test_long4_tables_match_the_reference_on_gpu(eng):
    res = eng.run([synth.long_contig(4)]).fetch_all()
    g = LONG["long4"]
    assert _md5("".join(mirror.orf_table_lines(res, 0))) == g["orfs_md5"]
    assert _md5("".join(mirror.ContigGraph(res, 0).dump_lines())) == g["edges_md5"]


@pytest.mark.gpu
def test_chunk_geometries_and_fallback_on_gpu(eng):
    adv = _adversarial()
    names = sorted(adv)
    seqs = [adv[k] for k in names] + [synth.long_contig(10)]
    one = eng.run(seqs, flags=N.SOLVE_NOCHUNK)
    fell = 0
    for geo in ((256, 768, 64, 4096), (64, 128, 16, 512), (512, 1536, 128, 2048), (32, 16, 4, 64)):
        eng.set_chunking(*geo)
        try:
            ch = eng.run(seqs)
        finally:
            eng.set_chunking()
        assert ch.n_chunks > 0
        assert np.array_equal(one.calls, ch.calls), geo
        assert [int(v) for v in one.contigs["err"]] == [int(v) for v in ch.contigs["err"]]
        fell += ch.n_chunk_fallbacks
    assert fell > 0                                   # the fallback ran (and gave the same calls)
