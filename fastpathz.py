"""fastpathz-compatible solver module (the reference imports the third-party package of this name,
phanotate.py:9,56-64): module-global graph, edges pushed as tab-separated text, exact integers.

    fz.empty_graph(); fz.add_edge("src\\tdst\\tweight"); fz.get_path(source=..., target=...) -> [node names]

The solve runs on the GPU through pb200_bellman_ford (csrc/graph.cuh: bf_literal): literal
Bellman-Ford, edges in insertion order, strict '<', integer part of the weight string
(CHANGELOG.md:11-13,57).  For throughput the batch path (phanotate_b200.engine) never builds strings.
"""
import ctypes
from decimal import Decimal, ROUND_DOWN, localcontext

import numpy as np

_names, _index, _src, _dst, _w = [], {}, [], [], []
_engine = None


def _eng():
    global _engine
    if _engine is None:
        from phanotate_modules import functions
        _engine = functions.engine()
    return _engine


def empty_graph():
    _names.clear()
    _index.clear()
    _src.clear()
    _dst.clear()
    _w.clear()


def _node(name):
    i = _index.get(name)
    if i is None:
        i = _index[name] = len(_names)
        _names.append(name)
    return i


def add_edge(edge_string):
    a, b, w = edge_string.split("\t")
    with localcontext() as ctx:
        ctx.prec = 400
        wi = int(Decimal(w).to_integral_value(rounding=ROUND_DOWN))
    if not -(1 << 240) < wi < (1 << 240):
        raise OverflowError("edge weight does not fit 240 bits")
    _src.append(_node(a))
    _dst.append(_node(b))
    _w.append(wi)
    return None


def get_path(source=None, target=None):
    if source not in _index or target not in _index:
        return []
    n, m = len(_names), len(_w)
    limbs = np.zeros((max(m, 1), 8), dtype=np.uint32)
    for k, wi in enumerate(_w):
        v = wi & ((1 << 256) - 1)
        for j in range(8):
            limbs[k, j] = (v >> (32 * j)) & 0xFFFFFFFF
    src = np.asarray(_src or [0], dtype=np.int32)
    dst = np.asarray(_dst or [0], dtype=np.int32)
    path = np.zeros(n, dtype=np.int32)
    plen = ctypes.c_int32(0)
    e = _eng()
    e._ck(e.lib.pb200_bellman_ford(e.ctx, n, m, src.ctypes.data, dst.ctypes.data, limbs.ctypes.data,
                                   _index[source], _index[target], path.ctypes.data, ctypes.byref(plen)))
    return [_names[i] for i in path[:plen.value]]
