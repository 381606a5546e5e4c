"""GC frame plot of the boundary (reference gc_frame_plot.py:7-74): `max_idx` / `min_idx` and the `GCframe` class
(`add_base`, `_close`, `get`).  On the hot path the window sums are popcounts inside the scan kernel
(csrc/scan_tile.cuh); this class is the reference's streaming interface for callers that use it directly."""
from collections import deque
import itertools


def max_idx(a, b, c):
    if a > b:
        return 1 if a > c else 3
    return 2 if b > c else 3


def min_idx(a, b, c):
    if a > b:
        return 3 if b > c else 2
    return 3 if a > c else 1


class GCframe:
    """Per-residue sliding GC count: bases go round-robin to residues 1, 2, 3; each residue keeps its last `window//3`
    bases and, per base added, the number of g/c among them (`total[residue]`).  `_close()` turns the trailing windows into
    centred ones (drops the first window//6 counts and lets the window run out at the end); `get()` closes and returns one
    triple per base position p >= 1 -- the counts of the residues of p, p+1, p+2 -- behind a dummy entry for index 0."""

    def __init__(self, window=120):
        self.window = window // 3
        self.states = itertools.cycle([1, 2, 3])
        self.bases = [None] + [deque('-' * self.window) for _ in range(3)]
        self.frequency = [None] + [dict.fromkeys('atcg-', 0) for _ in range(3)]
        self.total = [deque() for _ in range(4)]
        self.freq = []

    def _drop_oldest(self, r):
        self.frequency[r][self.bases[r].popleft()] -= 1
        self.total[r].append(self.frequency[r]['g'] + self.frequency[r]['c'])

    def add_base(self, base):
        r = next(self.states)
        self.bases[r].append(base)
        self.frequency[r][base] += 1          # KeyError for anything but a, t, c, g, '-' like the reference
        self._drop_oldest(r)

    def _close(self):
        for _ in range(self.window // 2):
            for r in (1, 2, 3):
                self.total[r].popleft()
                self._drop_oldest(r)

    def get(self):
        self._close()
        t = self.total
        self.freq.append([20, 20, 20])
        last = len(t[3]) - 1                  # residue-3 counts decide how many whole codons there are
        if last < 0:
            raise UnboundLocalError("GCframe.get() on fewer than three bases")       # (the reference fails the same way)
        seq = []                              # counts in base order: t[1][0], t[2][0], t[3][0], t[1][1], ...
        for i in range(last + 1):
            seq += [t[1][i], t[2][i], t[3][i]]
        for r in (1, 2):                      # the bases of an unfinished last codon
            if len(t[r]) > last + 1:
                seq.append(t[r][last + 1])
        self.freq.extend([seq[p], seq[p + 1], seq[p + 2]] for p in range(len(seq) - 2))
        return self.freq
