"""Drop-in for the reference's C extension module of this name (src/phanotate_connect.c, built by setup.py:6-21).

    import phanotate_connect as pc
    pc.add_edge(left_position, right_position)          # phanotate_connect.c:62-76
    pc.get_connected(min_distance=300) -> [(right_i, left_j, 0), ...]          # :78-121

Like the reference, the module keeps ONE global, ever-growing edge list (there is no reset call in the reference;
`clear()` here is an extra), `min_distance` is accepted and ignored (the reference compares against the literal 300), and
rows come right-entry-major in insertion order.  The join runs on the GPU through pb200_connect (csrc/connect.cuh);
there is no CPU path.
"""
import ctypes
import operator

import numpy as np

_left, _right = [], []
_engine = None


def _eng():
    global _engine
    if _engine is None:
        from phanotate_modules import functions
        _engine = functions.engine()
    return _engine


def add_edge(left_position, right_position):
    vals = []
    for v in (left_position, right_position):
        if isinstance(v, float):
            raise TypeError("integer argument expected, got float")          # PyArg_ParseTuple "ii"
        v = operator.index(v)
        if not -(1 << 31) <= v < (1 << 31):
            raise OverflowError("signed integer is greater than maximum" if v > 0 else "signed integer is less than minimum")
        vals.append(v)
    _left.append(vals[0])
    _right.append(vals[1])


def clear():
    """Not in the reference (its tables only grow): forget the edges added so far."""
    _left.clear()
    _right.clear()


def connected_arrays(left, right, engine=None):
    """rows (right_i, left_j) as an int32 array [n_rows, 2] for the edges (left[i], right[i])."""
    e = engine or _eng()
    left = np.ascontiguousarray(left, dtype=np.int32)
    right = np.ascontiguousarray(right, dtype=np.int32)
    n = len(left)
    if len(right) != n:
        raise ValueError("left and right differ in length")
    rows = ctypes.c_int64(0)
    e._ck(e.lib.pb200_connect(e.ctx, left.ctypes.data, right.ctypes.data, n, None, 0, ctypes.byref(rows)))
    out = np.zeros((rows.value, 2), dtype=np.int32)
    if rows.value:
        e._ck(e.lib.pb200_connect(e.ctx, left.ctypes.data, right.ctypes.data, n, out.ctypes.data, rows.value,
                                  ctypes.byref(rows)))
    return out


def get_connected(min_distance=300):
    if not isinstance(min_distance, int):
        raise TypeError("an integer is required (got type %s)" % type(min_distance).__name__)
    out = connected_arrays(_left, _right)
    return [(int(r), int(l), 0) for r, l in out]
