"""The certified integer weights (csrc/fast.cuh) against the literal Decimal chain.

The solve only needs trunc(weight*1000) per edge (edges.py:22).  By default pb200_run decides that
integer from a closed form with a rigorous error bound and replays the reference's Decimal
arithmetic only where it is owed; PB200_LITERAL replays it for everything.  These tests check, on
the host build of the same stage functions, that both modes give identical integers, calls and
Decimal weights, and (with the oracle) that the error bound used by the filter really holds."""
import decimal
import os
import subprocess
from decimal import Decimal

import numpy as np
import pytest

from phanotate_b200 import _native as N
from phanotate_b200 import engine
from helpers import STRESS, golden_text, seq_of

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HOSTSIM = os.path.join(ROOT, "tests", "native", "pb200_hostsim.so")


@pytest.fixture(scope="module")
def sim():
    from helpers import hostsim_path
    hostsim_path()
    e = engine.Engine(0, lib_path=HOSTSIM)
    yield e
    e.close()


def _calls(res, k):
    return "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))


def _both(sim, seqs):
    fast = sim.run(seqs)
    wf = fast.orf_int_weights()                       # before .orfs triggers the lazy literal completion
    ovf = fast.overlap_int_weights()
    gsf, gdf = fast.gap_int_weights()
    fast.orfs, fast.nodes                             # (lazy tables: fetched before the context runs the next batch)
    lit = sim.run(seqs, literal=True)
    assert np.array_equal(ovf, lit.overlap_int_weights())
    gsl, gdl = lit.gap_int_weights()
    assert np.array_equal(gsf, gsl) and np.array_equal(gdf, gdl)
    assert lit.n_literal_overlaps == lit.n_overlaps and fast.n_literal_overlaps <= max(2, fast.n_overlaps // 50)
    return fast, wf, lit, lit.orf_int_weights()


@pytest.mark.parametrize("name", ["phiX174", "lambda", "T4", "synth4_0", "synth4_1"])
def test_certified_integers_equal_literal(sim, name):
    fast, wf, lit, wl = _both(sim, [seq_of(name)])
    assert wf == wl
    assert _calls(fast, 0) == _calls(lit, 0) == golden_text(name, "calls.tsv")
    assert [float(r["score"]) for r in fast.calls] == [float(r["score"]) for r in lit.calls]      # certified scores
    withw = sim.run([seq_of(name)], call_weights=True)
    assert [str(N.dec_to_decimal(r["weight"])) for r in withw.calls] == [str(N.dec_to_decimal(r["weight"])) for r in lit.calls]
    assert withw.n_literal_presolve + withw.n_literal_postsolve >= withw.n_calls
    assert lit.n_literal_presolve == lit.n_orfs
    # the filter must actually filter: almost every ORF is decided without the literal chain
    assert fast.n_literal_presolve <= max(2, fast.n_orfs // 20)
    assert fast.n_literal_postsolve <= 2                  # scores are certified too: no Decimal chain after the solve


def test_certified_stress_batch(sim):
    seqs = [seq_of(n) for n in STRESS]
    fast, wf, lit, wl = _both(sim, seqs)
    assert wf == wl
    for k, name in enumerate(STRESS):
        assert _calls(fast, k) == golden_text(name, "calls.tsv"), name


def test_integer_is_trunc_of_weight_times_1000(sim):
    res = sim.run([seq_of("lambda")])
    wf = res.orf_int_weights()
    for w, o in zip(wf, res.orfs):                    # .orfs completes the Decimal weights lazily
        d = N.dec_to_decimal(o["weight"])
        with decimal.localcontext() as ctx:
            ctx.prec = 80
            assert w == int(d * 1000)                 # edges.py:22; int() truncates toward zero


def test_closed_form_error_bound_holds_on_reference_weights():
    """|W_ref / W_closed_form - 1| <= (2.51 n + 1.51) 1e-27 (fast.cuh header), checked with the oracle's
    literal weights at 90 digits on phiX174 and lambda."""
    from oracle import phanotate_oracle as O
    worst = worst_x = Decimal(0)
    for name in ("phiX174", "lambda"):
        dna = seq_of(name).lower()
        orfs = O.get_orfs(dna)
        T, pm, pn = orfs.T, orfs.pos_max, orfs.pos_min
        sc = O.normalise_starts(O.DEFAULT_STARTS)
        for o in orfs.iter_orfs():
            fwd = o.frame > 0
            rng = range(o.start, o.stop, 3 if fwd else -3)
            cnt = {}
            for base in rng:
                a, b, c = int(T[base]), int(T[base + 1]), int(T[base + 2])
                if not fwd:
                    a, c = c, a
                k = (O.max_idx(a, b, c), O.min_idx(a, b, c))
                cnt[k] = cnt.get(k, 0) + 1
            x = 1 - o.pstop                            # the literal 28-digit value
            # second bound: x_ref against the exact rational 1 - nt na (na + 2 ng) / len^3 (fast.cuh header)
            lo, hi = (o.start - 1, o.stop + 2) if fwd else (o.stop - 1, o.start + 2)
            lo, hi = max(lo, 0), min(hi, len(dna))
            sub = dna[lo:hi] if fwd else O.rev_comp(dna[lo:hi])
            na, nt, ng, ln = sub.count("a"), sub.count("t"), sub.count("g"), len(sub)
            with decimal.localcontext() as ctx:
                ctx.prec = 90
                xt = 1 - Decimal(nt * na * (na + 2 * ng)) / Decimal(ln) ** 3
                worst_x = max(worst_x, abs(x / xt - 1) / Decimal("4.01e-27"))
            with decimal.localcontext() as ctx:
                ctx.prec = 90
                E = sum(Decimal(v) * pm[k[0]] * pn[k[1]] for k, v in cnt.items())
                S = (-(E * x.ln())).exp()
                if o.codon in sc:
                    S *= sc[o.codon]
                S *= Decimal(str(o.weight_rbs))
                rel = abs((-o.weight) / S - 1)
                bound = (Decimal("2.51") * len(rng) + Decimal("1.51")) * Decimal("1e-27")
                worst = max(worst, rel / bound)
    assert worst < Decimal("0.25")                    # observed: 0.03
    assert worst_x < Decimal("0.5")


def test_wide_and_narrow_solver_agree(sim):
    """128-bit and 256-bit distances are the same algorithm: forcing the wide solver changes nothing; and a contig
    with an astronomically heavy ORF (an 800-codon repeat without a stop, weight ~1e40) is solved wide on its own and matches the oracle."""
    import random
    from oracle import phanotate_oracle as O
    seqs = [seq_of(n) for n in ("T4", "lambda", "phiX174")]
    a = sim.run(seqs)
    b = sim.run(seqs, flags=N.SOLVE_WIDE)
    assert np.array_equal(a.calls, b.calls) and int(a.contigs["wide"].sum()) == 0 and int(b.contigs["wide"].sum()) == 3
    rng = random.Random(7)
    rnd = lambda n: "".join(rng.choice("acgt") for _ in range(n))
    heavy = rnd(400) + "atg" + "atc" * 800 + "taa" + rnd(400)
    r = sim.run([heavy, seq_of("phiX174")])
    assert int(r.contigs[0]["wide"]) == 1 and int(r.contigs[1]["wide"]) == 0 and int(r.contigs[0]["err"]) == 0
    want = [row[:4] for row in O.call_contig(heavy)[3]]
    assert r.call_rows(0) == want
    assert "".join("%d\t%d\t%s\t%s\n" % x for x in r.call_rows(1)) == golden_text("phiX174", "calls.tsv")


def test_random_ragged_batch_certified_equals_literal(sim):
    """120 random contigs of ragged length with IUPAC codes and mixed case on the host build: default == PB200_LITERAL."""
    rng = np.random.default_rng(7)
    seqs = []
    for k in range(120):
        n = int(rng.integers(60, 4000))
        gc = rng.uniform(0.25, 0.75)
        s = rng.choice(np.frombuffer(b"acgt", dtype=np.uint8), size=n, p=[(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        m = rng.random(n) < 0.002
        s[m] = rng.choice(np.frombuffer(b"nrykmswbdhv", dtype=np.uint8), size=int(m.sum()))
        seqs.append(s.tobytes().upper() if k % 3 == 0 else s.tobytes())
    fast, wf, lit, wl = _both(sim, seqs)
    assert wf == wl
    for col in ("contig", "left", "right", "strand", "score"):
        assert np.array_equal(fast.calls[col], lit.calls[col]), col
    assert np.array_equal(fast.orfs, lit.orfs)
