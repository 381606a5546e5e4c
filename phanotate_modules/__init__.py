"""Drop-in surface of the reference's ``phanotate_modules`` package for the hot path
(reference phanotate_modules/__init__.py:2-15), backed by the B200 kernels in ``phanotate_b200``."""
from .file_handling import read_fasta
from .edges import Edge
from .nodes import Node
from .orfs import Orf

__all__ = ['file_handling', 'functions', 'edges', 'nodes', 'orfs', 'gc_frame_plot']
