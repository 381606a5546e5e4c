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


def find_trnas(dna):
    """functions.py:457-489: the hits of aragorn (`-t -w`) and tRNAscan-SE (`-B -q --brief`) on one contig as add_trnas
    collects them -- [start, stop] pairs, start > stop on the reverse strand; a tRNAscan-SE hit is kept only where aragorn
    found nothing.  None when neither program can be started (the reference then warns and skips the masking)."""
    import tempfile
    from subprocess import PIPE, Popen
    hits, covered, ran = [], set(), 0
    with tempfile.NamedTemporaryFile(mode='wt') as f:
        f.write(">temp\n")
        f.write(dna)
        f.flush()
        try:
            text = Popen(["aragorn", "-t", "-w", f.name], stdout=PIPE, stdin=PIPE, stderr=PIPE).stdout.read().decode()
            ran += 1
            for line in text.splitlines():
                if line.startswith('>') or line.endswith('found'):
                    continue
                where = line.split()[2]                    # "[a,b]" or "c[a,b]"
                a, b = (int(x) for x in where.lstrip('c').strip('[]').split(','))
                hits.append([b, a] if where.startswith('c') else [a, b])
                covered.update(range(a, b))
        except Exception:                                  # (the reference swallows everything here: a missing tool, odd output)
            pass
        try:
            text = Popen(["tRNAscan-SE", "-B", "-q", "--brief", f.name], stdout=PIPE, stdin=PIPE, stderr=PIPE).stdout.read().decode()
            ran += 1
            for line in text.splitlines():
                a, b = (int(x) for x in line.split('\t')[2:4])
                if a != b and not covered & set(range(min(a, b), max(a, b))):
                    hits.append([a, b])
        except Exception:
            pass
    return hits if ran else None


def get_orfs(locus):
    """functions.py:143-303: six-frame scan + scoring of one locus -> Orfs (stop -> {start -> Orf}).  The device run also
    builds the graph, so the tRNA programs (functions.py:457-509) are started here and their hits go into the same run."""
    dna = locus.seq().lower()
    params = make_params(locus.start_codons, locus.stop_codons, locus.min_orf_len)
    trnas = find_trnas(dna)
    # literal=True: the reference's Decimal chain for every ORF inside the run, which also keeps Orf.hold (orfs.py:84)
    res = engine().run([dna.encode()], params, literal=True, trnas=[(0, a, b) for a, b in trnas or []]).fetch_all()
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
    my_orfs._trnas = trnas
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
    trnas = getattr(my_orfs, "_trnas", None)
    if trnas is None:
        # neither aragorn nor tRNAscan-SE could be started: the reference prints this and carries on (functions.py:493-495)
        sys.stderr.write("Warning: tRNAscan or Aragorn were not found, proceding without tRNA masking.\n")
    for start, stop in trnas or []:                        # the other_end entries add_trnas leaves behind (functions.py:496-507)
        if start < stop:
            my_orfs.other_end['t' + str(stop - 2)] = start
            my_orfs.other_end['t' + str(start)] = stop - 2
        else:
            my_orfs.other_end['t' + str(start - 2)] = stop
            my_orfs.other_end['t' + str(stop)] = start - 2
    G = Graph(directed=True)
    nodes = [eval(name) for name in cg.node_names]
    for n in nodes[:-2]:
        G.add_node(n)
    for a, b, w in zip(cg.edge_src, cg.edge_dst, cg.edge_w):
        G.add_edge(Edge(nodes[a], nodes[b], w))
    G.add_node(nodes[-2])
    G.add_node(nodes[-1])
    return G
