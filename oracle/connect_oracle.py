"""CPU restatement of the reference's C extension join  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Follows /root/reference/src/phanotate_connect.c: `add_edge` (:62-76) appends (key=left, value=right) to `nodes_left`
(:44-52) and (key=right, value=left) to `nodes_right` (:54-62); uthash's HASH_ADD_INT neither merges equal keys nor
reorders, so both tables iterate in insertion order.  `get_connected` (:78-121) walks all pairs, `nodes_right`
outermost, and appends (s1->key, s2->key, 0) when |s1->key - s2->key| <= 300 and s1->key != s2->value and
s1->value != s2->key (:104-113); the parsed `min_distance` is never used.

Parity status: PINNED against the compiled reference itself (oracle/_ref, built by oracle/Makefile from the sources
under /root/reference): tests/test_connect.py compares on random inputs whenever oracle/_ref is present, and against
tests/golden/connect.json (generated from oracle/_ref by tests/golden/make_connect_golden.py) everywhere.
Only tests/ may import this module.
"""
import numpy as np


def get_connected(left, right):
    """rows (right_i, left_j) as int64 array [n_rows, 2]; edges (left[i], right[i]) in add_edge order."""
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    rows = []
    for i in range(len(left)):                                   # s1 over nodes_right (:104)
        d = np.abs(right[i] - left)                              # :107-108
        ok = (d <= 300) & (right[i] != right) & (left[i] != left)    # :109  (s2->value = right_j, s2->key = left_j)
        lj = left[ok]                                            # s2 over nodes_left in insertion order (:105)
        rows.append(np.stack([np.full(len(lj), right[i]), lj], axis=1))
    return np.concatenate(rows, axis=0) if rows else np.zeros((0, 2), dtype=np.int64)
