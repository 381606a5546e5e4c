"""Graph node type of the boundary (reference nodes.py:2-21): identity is the eval-able repr string."""


class Node:
    __slots__ = ("gene", "type", "frame", "position")

    def __init__(self, gene, type, frame, position):
        self.gene, self.type, self.frame, self.position = gene, type, frame, position

    def __repr__(self):
        return "Node(%r,%r,%r,%r)" % (self.gene, self.type, self.frame, self.position)

    def __hash__(self):
        return hash(repr(self))

    def __eq__(self, other):
        return hash(self) == hash(other)
