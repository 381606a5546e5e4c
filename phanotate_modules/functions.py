"""get_orfs / get_graph with the reference's signatures (reference functions.py:143, :307).

Both run on the GPU through phanotate_b200 (one contig = a batch of one); this module only
rebuilds the reference's Python containers from the device tables, in the reference's insertion
orders (phanotate_b200/mirror.py).  For throughput use phanotate_b200.engine.Engine on a batch of
contigs instead -- the per-object Python views below are for drop-in compatibility.
"""
import sys
from decimal import Decimal

from phanotate_b200 import _native as N
from phanotate_b200 import mirror
from phanotate_b200.engine import Engine, make_params

from .edges import Edge
from .graphs import Graph
from .nodes import Node
from .orfs import Orf, Orfs

_COMP = {'a': 't', 't': 'a', 'g': 'c', 'c': 'g', 'n': 'n', 'r': 'y', 'y': 'r', 's': 's', 'w': 'w', 'k': 'm',
         'm': 'k', 'b': 'v', 'v': 'b', 'd': 'h', 'h': 'd'}
_engine = None


def engine() -> Engine:
    """the process-wide context on cuda:0 (raises if the CUDA library or a GPU is missing)"""
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def set_engine(e):
    global _engine
    _engine = e


def rev_comp(seq):
    """reverse complement over the 15 IUPAC letters; KeyError otherwise (functions.py:19-24)"""
    return "".join(_COMP[b] for b in reversed(seq))


def get_orfs(locus):
    """functions.py:143-303: six-frame scan + scoring of one locus -> Orfs (stop -> {start -> Orf})."""
    dna = locus.seq().lower()
    params = make_params(locus.start_codons, locus.stop_codons, locus.min_orf_len)
    # literal=True: the reference's Decimal chain for every ORF inside the run, which also keeps Orf.hold (orfs.py:84)
    res = engine().run([dna.encode()], params, literal=True).fetch_all()
    res.check(0)
    holds = res.orf_holds()
    my_orfs = Orfs(locus)
    my_orfs.seq = dna
    my_orfs.contig_length = len(dna)
    c = res.contigs[0]
    my_orfs.pstop = N.dec_to_decimal(c["pstop"])
    table = res.orfs
    for i in mirror.orf_order(table):
        r = table[i]
        start, stop, frame = int(r["start"]), int(r["stop"]), int(r["frame"])
        if frame > 0:
            seq = dna[max(0, start - 1):stop + 2]
            rbs = dna[start - 21:start]
            length = stop + 2 - start + 1
        else:
            seq = rev_comp(dna[max(0, stop - 1):start + 2])
            rbs = rev_comp(dna[start + 2:start + 2 + 21])
            length = start + 2 - stop + 1
        o = Orf(start, stop, length, frame, seq, rbs, int(r["rbs_score"]), my_orfs.start_codons, my_orfs.stop_codons)
        o.pstop = N.dec_to_decimal(r["pstop"])
        o.hold = holds[i]                                          # functions.py:286-298
        o.weight = N.dec_to_decimal(r["weight"])
        o.weight_rbs = float(c["training_rbs"][o.rbs_score]) / float(c["background_rbs"][o.rbs_score])
        my_orfs._insert(o)
    my_orfs._pb200 = res
    return my_orfs


def get_graph(my_orfs):
    """functions.py:307-454: ORF, gap, overlap, bridge and terminal edges -> Graph, reference edge order."""
    res = getattr(my_orfs, "_pb200", None)
    if res is None:
        class _L:
            start_codons, stop_codons, min_orf_len = my_orfs.start_codons, my_orfs.stop_codons, my_orfs.min_orf_len

            @staticmethod
            def seq():
                return my_orfs.seq
        res = get_orfs(_L)._pb200
    cg = mirror.ContigGraph(res, 0)
    # no tRNA tools are driven from here (functions.py:457-509 needs aragorn / tRNAscan-SE); the reference
    # prints this and carries on when they are missing (functions.py:493-495)
    sys.stderr.write("Warning: tRNAscan or Aragorn were not found, proceding without tRNA masking.\n")
    G = Graph(directed=True)
    nodes = [eval(name) for name in cg.node_names]
    for n in nodes[:-2]:
        G.add_node(n)
    for a, b, w in zip(cg.edge_src, cg.edge_dst, cg.edge_w):
        G.add_edge(Edge(nodes[a], nodes[b], w))
    G.add_node(nodes[-2])
    G.add_node(nodes[-1])
    return G
