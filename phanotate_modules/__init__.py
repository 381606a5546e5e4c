"""Drop-in surface of the reference's ``phanotate_modules`` package for the hot path, backed by the B200 kernels in
``phanotate_b200``.  The package exports what the reference's does (its __init__.py:2-15): four names at package level
and six submodules through ``__all__``; ``graphs``, ``locus``, ``file`` and ``feature`` are importable as well."""
from . import edges as _edges, file_handling as _fh, nodes as _nodes, orfs as _orfs

read_fasta, Edge, Node, Orf = _fh.read_fasta, _edges.Edge, _nodes.Node, _orfs.Orf
__all__ = sorted(('orfs', 'nodes', 'edges', 'functions', 'file_handling', 'gc_frame_plot'))
