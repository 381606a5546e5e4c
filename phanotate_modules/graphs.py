"""dict-of-dict graph container of the boundary (reference graphs.py:7-126): node -> {node -> Edge},
insertion ordered, loops and parallel edges rejected with ValueError."""
from .edges import Edge  # noqa: F401


class Graph(dict):
    def __init__(self, n=0, directed=False):
        super().__init__()
        self.n, self.directed = n, directed

    def is_directed(self):
        return self.directed

    def v(self):
        return len(self)

    def e(self):
        m = sum(len(out) for out in self.values())
        return m if self.directed else m / 2

    def add_node(self, node):
        self.setdefault(node, {})

    def has_node(self, node):
        return node in self

    def add_edge(self, edge):
        if edge.source == edge.target:
            raise ValueError("loops are forbidden:" + str(edge.source) + " " + str(edge.target))
        self.add_node(edge.source)
        self.add_node(edge.target)
        if edge.target in self[edge.source]:
            raise ValueError("parallel edges are forbidden")
        self[edge.source][edge.target] = edge
        if not self.directed:
            if edge.source in self[edge.target]:
                raise ValueError("parallel edges are forbidden")
            self[edge.target][edge.source] = ~edge

    def del_edge(self, edge):
        del self[edge.source][edge.target]
        if not self.directed:
            del self[edge.target][edge.source]

    def has_edge(self, edge):
        return edge.source in self and edge.target in self[edge.source]

    def weight(self, edge):
        """weight of the stored edge with these endpoints, else 0 (graphs.py:91-96)"""
        if self.has_edge(edge):
            return self[edge.source][edge.target].weight
        return 0

    def iternodes(self):
        return self.keys()

    def iteradjacent(self, source):
        return self[source].keys()

    def iteroutedges(self, source):
        yield from self[source].values()

    def iterinedges(self, node):
        for src in self:
            if node in self[src]:
                yield self[src][node]

    def iteredges(self):
        for src, out in self.items():
            for dst, edge in out.items():
                if self.directed or src < dst:
                    yield edge
