#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own code.

Run once, in the build container (needs /root/reference, which is read-only and
does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs: the reference's unmodified ``phanotate_modules.functions.get_orfs`` and
``get_graph`` (imported from /root/reference), driven the way ``phanotate.py:40-76``
drives them.  The third-party solver ``fastpathz`` is not installable here, so
the shortest path is found by an exact-integer Bellman-Ford over the reference's
own edge strings (edges.py:17-23; weight = integer part of Decimal*1000,
phanotate.py:55) in ``graph.iteredges()`` order with strict ``<`` relaxation;
this reproduces the four known-answer rows in the reference's README.md:47-54.
tRNA tools are absent, so ``add_trnas`` returns early (functions.py:493-495).

Outputs (all small, committed):
    <name>.calls.tsv        left<TAB>right<TAB>strand<TAB>%E score  (locus.py:39-56 row order/values)
    <name>.orfs.csv.gz      start,stop,frame,rbs_score,pstop,weight  in Orfs.iter_orfs() order
    <name>.edges.txt.gz     the --dump text (phanotate.py:58), small inputs only
    index.json              sizes, md5s of the three tables per input, any exception the reference raised
"""
import gzip
import hashlib
import json
import os
import sys
import warnings
from decimal import Decimal, ROUND_DOWN
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from phanotate_b200 import synth  # noqa: E402


class LocusShim:
    """The duck type get_orfs needs (functions.py:143-150, orfs.py:8-15)."""

    def __init__(self, name, seq):
        self._name, self._seq = name, seq
        w = {"atg": Decimal("0.85"), "gtg": Decimal("0.10"), "ttg": Decimal("0.05")}
        m = max(w.values())
        self.start_codons = {k: v / m for k, v in w.items()}     # file_handling.py:58-62
        self.stop_codons = ["tag", "tga", "taa"]
        self.min_orf_len = 90

    def seq(self):
        return self._seq

    def length(self):
        return len(self._seq)

    def name(self):
        return self._name


def bellman_ford(edge_strings, source, target):
    names, idx = [], {}
    E = []
    for s in edge_strings:
        a, b, w = s.split("\t")
        for n in (a, b):
            if n not in idx:
                idx[n] = len(names)
                names.append(n)
        E.append((idx[a], idx[b], int(Decimal(w).to_integral_value(rounding=ROUND_DOWN))))
    if source not in idx or target not in idx:
        return None, 0
    dist = [None] * len(names)
    par = [-1] * len(names)
    dist[idx[source]] = 0
    passes = 0
    while True:
        passes += 1
        changed = False
        for u, v, w in E:
            du = dist[u]
            if du is None:
                continue
            nd = du + w
            if dist[v] is None or nd < dist[v]:
                dist[v] = nd
                par[v] = u
                changed = True
        if not changed:
            break
    if dist[idx[target]] is None:
        return None, passes
    path, v = [], idx[target]
    while v != -1:
        path.append(names[v])
        v = par[v]
    return path[::-1], passes


def run_reference(args):
    name, seq, keep_edges = args
    from phanotate_modules import functions
    from phanotate_modules.nodes import Node  # noqa: F401  (eval of node reprs)
    from phanotate_modules.edges import Edge
    rec = {"name": name, "L": len(seq)}
    try:
        locus = LocusShim(name, seq)
        orfs = functions.get_orfs(locus)
        graph = functions.get_graph(orfs)
    except Exception as e:  # the reference's own error behaviour is part of the contract
        rec["exception"] = type(e).__name__
        return rec, None, None, None
    orf_lines = ["%d,%d,%d,%d,%s,%s\n" % (o.start, o.stop, o.frame, o.rbs_score, o.pstop, o.weight)
                 for o in orfs.iter_orfs()]
    edge_lines = [str(e) + "\n" for e in graph.iteredges()]
    source = "Node('source','source',0,0)"
    target = "Node('target','target',0,%d)" % (len(seq) + 1)
    calls = []
    passes = 0
    if len(graph) > 2:
        path, passes = bellman_ford([l[:-1] for l in edge_lines], source, target)
        if path is None:
            rec["no_path"] = True
            path = []
        path = path[1:]
        it = iter(path)
        for s, t in zip(it, it):                      # file_handling.pairwise:24-26
            left, right = eval(s), eval(t)
            w = graph.weight(Edge(left, right, 0))
            strand = "+" if left.frame > 0 else "-"
            calls.append("%d\t%d\t%s\t%s\n" % (left.position, right.position + 2, strand, "%E" % w))
    rec.update(n_orfs=len(orf_lines), n_families=len(orfs), n_nodes=len(graph), n_edges=len(edge_lines),
               n_calls=len(calls), bf_passes=passes, pstop=str(orfs.pstop),
               orfs_md5=hashlib.md5("".join(orf_lines).encode()).hexdigest(),
               edges_md5=hashlib.md5("".join(edge_lines).encode()).hexdigest(),
               calls_md5=hashlib.md5("".join(calls).encode()).hexdigest())
    return rec, orf_lines, edge_lines if keep_edges else None, calls


def main():
    jobs = []
    for f, nm in (("phiX174.fasta", "phiX174"), ("NC_001416.1.fasta", "lambda"), ("NC_000866.1.fasta", "T4")):
        recs = synth.read_fasta_bytes(os.path.join(ROOT, "tests", "data", f))
        jobs.append((nm, recs[0][1].decode(), nm == "phiX174"))
    for k in (0, 1):
        jobs.append(("synth4_%d" % k, synth.synth4_contig(k).decode(), False))
    for nm, s in synth.stress_contigs():
        jobs.append((nm, s.decode(), True))
    with Pool(8) as pool:
        results = pool.map(run_reference, jobs, chunksize=1)
    index = {}
    for (rec, orf_lines, edge_lines, calls) in results:
        nm = rec["name"]
        index[nm] = rec
        if orf_lines is None:
            continue
        with open(os.path.join(HERE, nm + ".calls.tsv"), "w") as fh:
            fh.writelines(calls)
        with gzip.GzipFile(os.path.join(HERE, nm + ".orfs.csv.gz"), "wb", mtime=0) as fh:
            fh.write("".join(orf_lines).encode())
        if edge_lines is not None:
            with gzip.GzipFile(os.path.join(HERE, nm + ".edges.txt.gz"), "wb", mtime=0) as fh:
                fh.write("".join(edge_lines).encode())
    with open(os.path.join(HERE, "index.json"), "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)
    for nm in ("phiX174", "lambda", "T4", "synth4_0", "synth4_1"):
        print(nm, {k: index[nm].get(k) for k in ("n_orfs", "n_nodes", "n_edges", "n_calls", "bf_passes", "calls_md5")})
    exc = {k: v["exception"] for k, v in index.items() if "exception" in v}
    print("exceptions:", exc, " no_path:", [k for k, v in index.items() if v.get("no_path")])


if __name__ == "__main__":
    main()
