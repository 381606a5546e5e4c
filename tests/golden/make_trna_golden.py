#!/usr/bin/env python3
"""Goldens for tRNA masking (reference functions.py:457-509 add_trnas, and the tRNA branch of the connect loop
functions.py:388-399), from the reference's own code.

The reference shells out to `aragorn` / `tRNAscan-SE`, which are not installed here; their OUTPUT is all it uses.  This
script puts two stub executables of those names on PATH that print canned hit lists in the tools' formats (the aragorn
`-t -w` batch lines `N  tRNA-Xxx  [a,b]  ..` / `c[a,b]`, the tRNAscan-SE `--brief` tab-separated rows), then runs the
reference's unmodified get_orfs / get_graph and the exact-integer edge-order Bellman-Ford of make_golden.py.

    python tests/golden/make_trna_golden.py   ->  tests/golden/trna.json (+ trna_<case>.edges.txt.gz for the phiX174 cases)

Per case: the tRNA list as add_trnas sees it ([start, stop], start > stop on the reverse strand), md5 and size of the
--dump edge text, and the path as (left, right+2, strand, gene, %E weight) rows.
"""
import gzip
import hashlib
import json
import os
import stat
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

# case -> (contig, aragorn hits [(a, b, complement)], tRNAscan-SE hits [(begin, end)])
CASES = {
    "phiX174_none": ("phiX174", [], []),
    "phiX174_one": ("phiX174", [(1000, 1072, False)], []),
    "phiX174_rev": ("phiX174", [(2931, 3005, True)], []),
    "phiX174_many": ("phiX174", [(30, 102, False), (400, 473, True), (640, 712, False), (2000, 2076, False), (2100, 2171, True),
                                 (5290, 5362, False)], [(3600, 3672), (4480, 4405)]),
    "phiX174_scan_overlap": ("phiX174", [(1000, 1072, False)], [(1040, 1110), (3000, 3075)]),
    "stress13_two": ("stress13", [(300, 371, False), (900, 975, True)], []),
    "stress1_one": ("stress1", [(500, 572, False)], []),
    "stress2_rev": ("stress2", [(1200, 1273, True)], []),
    "stress5_two": ("stress5", [(100, 172, False), (420, 490, False)], []),
    "stress6_ends": ("stress6", [(5, 77, False)], [(2400, 2330)]),
    "stress9_one": ("stress9", [(200, 275, True)], []),
    "stress10_three": ("stress10", [(1500, 1572, False), (1600, 1672, False), (1700, 1771, True)], []),
    "stress14_one": ("stress14", [(2500, 2573, False)], []),
    "stress26_two": ("stress26", [(800, 872, True), (3000, 3072, False)], []),
    # a contig long enough for the chunked solve (5,517 nodes): with hits it takes the one-warp sweep that knows their edges
    "T4_eight": ("T4", [(70120, 70195, False), (70203, 70278, False), (70400, 70474, True), (71012, 71088, False), (71600, 71672, False),
                        (72303, 72378, False), (72411, 72484, True), (73000, 73075, False)], [(1200, 1273), (160100, 160030)]),
}

ARAGORN = """#!/bin/sh
cat "$PB200_STUB_DIR/aragorn.out"
"""
TRNASCAN = """#!/bin/sh
cat "$PB200_STUB_DIR/trnascan.out"
"""


def write_stubs(d):
    for name, body in (("aragorn", ARAGORN), ("tRNAscan-SE", TRNASCAN)):
        p = os.path.join(d, name)
        with open(p, "w") as fh:
            fh.write(body)
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)


def set_case(d, aragorn, scan):
    with open(os.path.join(d, "aragorn.out"), "w") as fh:
        fh.write(">temp\n%d genes found\n" % len(aragorn))
        for k, (a, b, comp) in enumerate(aragorn):
            fh.write("%d   tRNA-Ala   %s[%d,%d]\t34  \t(tgc)\n" % (k + 1, "c" if comp else "", a, b))
    with open(os.path.join(d, "trnascan.out"), "w") as fh:
        for k, (a, b) in enumerate(scan):
            fh.write("temp \t%d\t%d\t%d\tAla\tTGC\t0\t0\t60.1\t\n" % (k + 1, a, b))


def trna_list(aragorn, scan):
    """what add_trnas ends up with (functions.py:469-489)"""
    out, seen = [], []
    for a, b, comp in aragorn:
        out.append([b, a] if comp else [a, b])
        seen.extend(range(a, b))
    for a, b in scan:
        if a < b and not set(seen) & set(range(a, b)):
            out.append([a, b])
        elif a > b and not set(seen) & set(range(b, a)):
            out.append([a, b])
    return out


def main():
    only = sys.argv[1:]
    import make_golden as MG
    from helpers import seq_of
    d = tempfile.mkdtemp(prefix="pb200_stubs_")
    write_stubs(d)
    os.environ["PB200_STUB_DIR"] = d
    os.environ["PATH"] = d + os.pathsep + os.environ["PATH"]
    from phanotate_modules import functions
    from phanotate_modules.edges import Edge
    from phanotate_modules.nodes import Node  # noqa: F401
    assert functions.__file__.startswith("/root/reference")
    path = os.path.join(HERE, "trna.json")
    out = json.load(open(path)) if only and os.path.exists(path) else {}
    for case, (contig, aragorn, scan) in CASES.items():
        if only and case not in only:
            continue
        set_case(d, aragorn, scan)
        seq = seq_of(contig)
        orfs = functions.get_orfs(MG.LocusShim(contig, seq))
        graph = functions.get_graph(orfs)
        edge_lines = [str(e) + "\n" for e in graph.iteredges()]
        source = "Node('source','source',0,0)"
        target = "Node('target','target',0,%d)" % (len(seq) + 1)
        path, passes = MG.bellman_ford([l[:-1] for l in edge_lines], source, target)
        rows = []
        it = iter((path or [])[1:])
        for s, t in zip(it, it):
            left, right = eval(s), eval(t)
            w = graph.weight(Edge(left, right, 0))
            rows.append([left.position, right.position + 2, "+" if left.frame > 0 else "-", left.gene, "%E" % w])
        rec = {"contig": contig, "trnas": trna_list(aragorn, scan), "n_nodes": len(graph), "n_edges": len(edge_lines),
               "edges_md5": hashlib.md5("".join(edge_lines).encode()).hexdigest(), "bf_passes": passes, "calls": rows,
               "n_trna_calls": sum(1 for r in rows if r[3] == "tRNA")}
        out[case] = rec
        if contig == "phiX174":
            with gzip.GzipFile(os.path.join(HERE, "trna_%s.edges.txt.gz" % case), "wb", mtime=0) as fh:
                fh.write("".join(edge_lines).encode())
        print(case, rec["trnas"], rec["n_nodes"], rec["n_edges"], rec["edges_md5"][:8], len(rows), rec["n_trna_calls"], flush=True)
    with open(os.path.join(HERE, "trna.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
