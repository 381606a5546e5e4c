"""Host-side views that put the device tables into the reference's Python iteration orders.

The kernels emit ORFs, nodes and edges sorted by position.  The reference's containers are dicts,
so their iteration order is insertion order, and that order is visible (``--dump``, Orfs.iter_orfs,
Graph.iteredges).  This module reorders the tables; it does no arithmetic.

  Orfs.iter_orfs()      families by first appearance, members by emission      orfs.py:17-37, functions.py:196-251
  Graph node order      first appearance over iter_orfs(), entry then exit     functions.py:311-318
  Graph.iteredges()     per node: ORF edges, bridges, connect-loop edges in
                        (outer=right node, inner=left node) order, terminals   functions.py:331-452, graphs.py:121-126
"""
from __future__ import annotations

import numpy as np

from . import _native as N

KIND_TYPE = {0: "start", 1: "stop", 2: "stop", 3: "start"}


def orf_order(orfs: np.ndarray) -> np.ndarray:
    """Indices of one contig's ORF rows in Orfs.iter_orfs() order.

    The reference emits a forward family when its stop codon is scanned (nearest start first) and a
    reverse family when the NEXT reverse stop of its frame is scanned (ascending starts); the contig
    end flushes frame 1 forward, frame 1 reverse, frame 2 ...  `trigger` encodes that scan position.
    """
    sec = np.where(orfs["frame"] > 0, -orfs["start"], orfs["start"])
    return np.lexsort((sec, orfs["trigger"]))


def node_repr(kind: int, frame: int, position: int, gene: str = "CDS") -> str:
    f = frame if kind in (0, 1) else -frame
    return "Node('%s','%s',%d,%d)" % (gene, KIND_TYPE[kind], f, position)


class ContigGraph:
    """Node and edge tables of one contig in reference order."""

    def __init__(self, res, k: int):
        c = res.contigs[k]
        self.length = int(c["length"])
        no0, nn0 = int(c["orf_off"]), int(c["node_off"])
        orfs = res.orfs[no0:no0 + int(c["n_orfs"])]
        nodes = res.nodes[nn0:nn0 + int(c["n_nodes"])]
        self.orfs = orfs
        self.order = orf_order(orfs)
        n = len(nodes)
        # node insertion index: entry then exit of every ORF in iter_orfs order, first appearance
        start_node = orfs["node"][self.order] - nn0
        stop_node = nodes["mate"][start_node] - nn0
        fwd = orfs["frame"][self.order] > 0
        entry = np.where(fwd, start_node, stop_node)
        exit_ = np.where(fwd, stop_node, start_node)
        seq = np.empty(2 * len(entry), dtype=np.int64)
        seq[0::2], seq[1::2] = entry, exit_
        _, first = np.unique(seq, return_index=True)
        ins_order = seq[np.sort(first)]                       # local node ids in insertion order
        # tRNA node pairs (add_trnas, functions.py:457-509): behind the regular nodes of the whole batch, inserted into the
        # graph after every CDS node and before source / target, entry then exit per hit
        n_reg_all = res.n_nodes
        tnodes = res.nodes[n_reg_all:] if getattr(res, "n_trnas", 0) else res.nodes[:0]
        tsel = np.nonzero(tnodes["contig"] == k)[0]            # (batch-tail indices of this contig's tRNA nodes, in hit order)
        nt2 = len(tsel)
        self.ins = np.full(n + nt2 + 2, -1, dtype=np.int64)
        self.ins[ins_order] = np.arange(len(ins_order))
        self.ins[n:n + nt2] = len(ins_order) + np.arange(nt2)
        self.ins[n + nt2], self.ins[n + nt2 + 1] = n + nt2, n + nt2 + 1      # source, target come last (functions.py:440-443)
        self.node_names = [node_repr(int(nodes["kind"][i]), int(nodes["frame"][i]), int(nodes["position"][i]))
                           for i in ins_order]
        self.node_names += [node_repr(int(tnodes["kind"][i]), 4, int(tnodes["position"][i]), "tRNA") for i in tsel]
        self.node_names += ["Node('source','source',0,0)", "Node('target','target',0,%d)" % (self.length + 1)]
        if len(ins_order) < n:                                  # (cannot happen: every node belongs to an ORF)
            raise ValueError("node without an ORF")
        self.local_nodes = nodes
        # edges of this contig
        e = res.edges
        e = e[e["contig"] == k]
        tmap = {int(n_reg_all + g): n + j for j, g in enumerate(tsel)}      # global tRNA node index -> local

        def local(col, special, special_to):
            out = np.empty(len(col), dtype=np.int64)
            for i, v in enumerate(col.tolist()):
                out[i] = special_to if v == special else (tmap[v] if v >= n_reg_all else v - nn0)
            return out
        if nt2:
            src = local(e["src"], N.NODE_SOURCE, n + nt2)
            dst = local(e["dst"], N.NODE_TARGET, n + nt2 + 1)
        else:
            src = np.where(e["src"] == N.NODE_SOURCE, n, e["src"] - nn0)
            dst = np.where(e["dst"] == N.NODE_TARGET, n + 1, e["dst"] - nn0)
        isrc, idst = self.ins[src], self.ins[dst]
        kind = e["kind"]
        orf_rank = np.empty(len(orfs), dtype=np.int64)
        orf_rank[self.order] = np.arange(len(orfs))
        # order inside a node's group: ORF edges, bridges, the tRNA edge (add_trnas runs after the bridges), connect loop, terminals
        phase = np.select([kind == 0, kind == 3, kind == 6, (kind == 1) | (kind == 2)], [0, 1, 2, 3], 4)
        k1 = np.zeros(len(e), dtype=np.int64)
        k2 = np.zeros(len(e), dtype=np.int64)
        m = kind == 0                                          # ORF edge: rank of its ORF
        if m.any():
            # the ORF of an ORF edge is the ORF of whichever end is a start node
            s_loc, d_loc = src[m], dst[m]
            s_is_start = np.isin(nodes["kind"][s_loc], (0, 3))
            o = np.where(s_is_start, nodes["orf"][s_loc], nodes["orf"][d_loc]) - no0
            k1[m] = orf_rank[o]
        m = kind == 3
        k1[m] = idst[m]
        m = kind == 1                                          # gap: outer loop = right node = dst
        k1[m], k2[m] = idst[m], isrc[m]
        m = kind == 2                                          # overlap: outer loop = right node = src
        k1[m], k2[m] = isrc[m], idst[m]
        m = kind == 4                                          # source -> entry, in node order
        k1[m] = idst[m]
        o = np.lexsort((k2, k1, phase, isrc))
        self.edge_src = isrc[o]
        self.edge_dst = idst[o]
        self.edge_w = [N.dec_to_decimal(r) for r in e["weight"][o]]
        self.edge_kind = kind[o]

    def dump_lines(self):
        """The --dump text (phanotate.py:58, edges.py:17-23)."""
        nm = self.node_names
        return ["%s\t%s\t%s\n" % (nm[a], nm[b], str(w * 1000))
                for a, b, w in zip(self.edge_src, self.edge_dst, self.edge_w)]


def orf_table_lines(res, k: int):
    """start,stop,frame,rbs_score,pstop,weight per ORF in iter_orfs order (the golden ORF table format)."""
    c = res.contigs[k]
    orfs = res.orfs[int(c["orf_off"]):int(c["orf_off"]) + int(c["n_orfs"])]
    out = []
    for i in orf_order(orfs):
        o = orfs[i]
        out.append("%d,%d,%d,%d,%s,%s\n" % (o["start"], o["stop"], o["frame"], o["rbs_score"],
                                            N.dec_to_decimal(o["pstop"]), N.dec_to_decimal(o["weight"])))
    return out
