"""Directed weighted edge (reference edges.py:3-60).  str(edge) is the solver wire format
``repr(source) TAB repr(target) TAB str(weight*1000)`` (edges.py:17-23, phanotate.py:55-59)."""


class Edge:
    __slots__ = ("source", "target", "weight")

    def __init__(self, source, target, weight):
        self.source, self.target, self.weight = source, target, weight

    def __str__(self):
        return "%r\t%r\t%s" % (self.source, self.target, str(self.weight * 1000))

    def __repr__(self):
        return "Edge(%r, %r, %r)" % (self.source, self.target, self.weight)

    def __hash__(self):
        return hash(repr(self))

    def __invert__(self):
        return self.__class__(self.target, self.source, self.weight)

    inverted = __invert__
