"""CPU oracle for the PHANOTATE hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A restatement (numpy + Python ``decimal``, i.e. the very libmpdec the reference
runs on) of the reference algorithm for the path BASELINE.json names: six-frame
ORF scan + RBS / start-codon / GC-frame scoring, ORF/gap/overlap graph, exact
integer shortest path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module;
the product (``phanotate_b200``) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
(a) the four known-answer rows in the reference's README.md:47-54 and
(b) golden tables generated from the reference's own ``get_orfs``/``get_graph``
    (tests/golden/make_golden.py): ORF tables (28-digit pstop and weight), the
    full ``--dump`` edge text and the call tables of phiX174, lambda, T4, two
    50-kb synthetic contigs and 64 stress contigs (IUPAC codes, N-runs, tiny).
The solver boundary (third-party ``fastpath``/``fastpathz`` >= 1.3, setup.py:53,
absent from the tree) is pinned only by the README rows; its published contract
(CHANGELOG.md:11-13,54,57: exact integers, weight x1000) is restated here.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Coordinates are 1-based as in the reference.
"""
from __future__ import annotations

import itertools
from decimal import Decimal, ROUND_DOWN

import numpy as np

# ---------------------------------------------------------------------------------------------
# alphabet                                                                   functions.py:19-24
# ---------------------------------------------------------------------------------------------
_COMP = {'a': 't', 't': 'a', 'g': 'c', 'c': 'g', 'n': 'n', 'r': 'y', 'y': 'r', 's': 's', 'w': 'w',
         'k': 'm', 'm': 'k', 'b': 'v', 'v': 'b', 'd': 'h', 'h': 'd'}
_COMP_TAB = str.maketrans("".join(_COMP.keys()), "".join(_COMP.values()))


def rev_comp(seq: str) -> str:
    """functions.py:19-24 -- KeyError on anything outside the 15 IUPAC letters."""
    for ch in seq:
        if ch not in _COMP:
            raise KeyError(ch)
    return seq.translate(_COMP_TAB)[::-1]


# ---------------------------------------------------------------------------------------------
# RBS motif scoring                                                         functions.py:48-138
# ---------------------------------------------------------------------------------------------
_SIX = ('ggacga', 'ggatga', 'ggaaga', 'ggcgga', 'ggggga', 'ggtgga')
_G5 = ('ggtgg', 'ggggg', 'ggcgg')
_T3 = ('agg', 'gag', 'gga')
_A5 = ('gaaga', 'gatga', 'gacga')
_Q4 = ('agga', 'gagg', 'ggag')
_MID, _LO, _HI, _FAR = (5, 6, 7, 8, 9, 10), (3, 4), (11, 12), (13, 14, 15)
# (score, motifs, offsets into the REVERSED window); first hit in this order wins.
RBS_RULES = [
    (27, ('ggagga',), _MID), (26, ('ggagga',), _LO), (25, ('ggagga',), _HI),
    (24, ('ggagg',), _MID), (23, ('ggagg',), _LO), (22, ('gagga',), _MID), (21, ('gagga',), _LO),
    (20, ('gagga', 'ggagg'), _HI),
    (19, _SIX, _MID), (18, _SIX, _LO), (17, _SIX, _HI),
    (16, ('ggag', 'gagg'), _MID), (15, ('agga',), _MID), (14, _G5, _MID), (13, _T3, _MID),
    (12, _Q4, _HI), (11, _Q4, _LO),
    (10, ('gagga', 'ggagg', 'ggagga'), _FAR),
    (9, _A5, _MID), (8, _G5, _LO), (7, _G5, _HI), (6, _T3, _HI), (5, _A5, _LO), (4, _A5, _HI),
    (3, _Q4, _FAR), (2, _T3 + ('ggaaga', 'ggatga', 'ggacga') + _G5, _FAR), (1, _T3, _LO),
]


def score_rbs(seq: str) -> int:
    """Scalar restatement of functions.py:48-138 (tuple membership == exact slice equality)."""
    s = seq[::-1]
    for score, motifs, offs in RBS_RULES:
        for m in motifs:
            k = len(m)
            for a in offs:
                if s[a:a + k] == m:
                    return score
    return 0


def _hits(text: str, pat: str) -> np.ndarray:
    """bool[len(text)]: text[j:j+len(pat)] == pat."""
    out = np.zeros(len(text), dtype=bool)
    j = text.find(pat)
    while j != -1:
        out[j] = True
        j = text.find(pat, j + 1)
    return out


def rbs_arrays(dna: str):
    """bgf[i] = score_rbs(dna[i:i+21]); bgr[i] = score_rbs(rev_comp(dna[i:i+21]))  (functions.py:168-169).

    Closed form: reversing the window turns "motif m at offset a of s" into "reversed(m) ends at
    text position i+W-1-a"; for the reverse strand s is the plain complement of the window, so the
    test is "complement(m) starts at i+a".  W = min(21, L-i); a slice that would run past W never
    matches.  Checked against the scalar form in tests/test_oracle.py.
    """
    L = len(dna)
    bgf = np.zeros(L, dtype=np.int8)
    bgr = np.zeros(L, dtype=np.int8)
    comp = dna.translate(_COMP_TAB)
    idx = np.arange(L)
    W = np.minimum(21, L - idx)
    for score, motifs, offs in reversed(RBS_RULES):       # ascending priority, later writes win
        hf = np.zeros(L, dtype=bool)
        hr = np.zeros(L, dtype=bool)
        for m in motifs:
            k = len(m)
            Hf = _hits(dna, m[::-1])          # reversed motif, indexed by start; ends at start+k-1
            Hr = _hits(comp, m)               # complement(text) == m  <=>  text == complement(m)
            for a in offs:
                ok = (a + k) <= W
                st = idx + W - a - k          # start of reversed motif in the text
                v = ok & (st >= 0)
                hf[v] |= Hf[st[v]]
                st2 = idx + a
                v2 = ok & (st2 < L)
                hr[v2] |= Hr[st2[v2]]
        bgf[hf] = score
        bgr[hr] = score
    return bgf, bgr


# ---------------------------------------------------------------------------------------------
# GC frame plot                                                          gc_frame_plot.py:7-74
# ---------------------------------------------------------------------------------------------
def max_idx(a, b, c):
    """gc_frame_plot.py:7-17"""
    if a > b:
        return 1 if a > c else 3
    return 2 if b > c else 3


def min_idx(a, b, c):
    """gc_frame_plot.py:18-28"""
    if a > b:
        return 3 if b > c else 2
    return 3 if a > c else 1


def gc_window_sums(sub: np.ndarray) -> np.ndarray:
    """T[q] for q in 0..L+2 (1-based q; T[0] unused): sum_{k=-19..20} gc(q+3k) clipped to [1,L].

    Closed form of GCframe.add_base/_close/get (gc_frame_plot.py:44-74): a 40-deep per-frame deque
    whose running g+c count is sampled 20 codons late.  gc_pos_freq[p] == [T[p],T[p+1],T[p+2]] for
    p in 1..L-2 and gc_pos_freq[0] == [20,20,20].
    """
    L = len(sub)
    gc = np.zeros(L + 200, dtype=np.int64)               # position q at index q+70
    gc[71:71 + L] = (sub == ord('g')) | (sub == ord('c'))
    T = np.zeros(L + 3, dtype=np.int64)
    q = np.arange(1, L + 3)
    for k in range(-19, 21):
        T[1:] += gc[q + 3 * k + 70]
    return T


# ---------------------------------------------------------------------------------------------
# containers (dict semantics are part of the behaviour: insertion order, bare-position keys)
# ---------------------------------------------------------------------------------------------
class OOrf:
    __slots__ = ("start", "stop", "frame", "length", "rbs_score", "pstop", "weight", "weight_rbs", "hold",
                 "codon", "trigger")


class OOrfs(dict):
    """stop -> {start -> OOrf}; other_end keyed by bare position (orfs.py:6-32)."""

    def __init__(self):
        super().__init__()
        self.other_end = {}
        self.pstop = None
        self.contig_length = 0
        self.seq = ""

    def add(self, o: OOrf):
        start, stop = o.start, o.stop
        if stop not in self:                              # orfs.py:19-23
            self[stop] = {start: o}
            self.other_end[stop] = start
            self.other_end[start] = stop
        elif start not in self[stop]:                     # orfs.py:24-30
            self[stop][start] = o
            self.other_end[start] = stop
            if o.frame > 0 and start < self.other_end[stop]:
                self.other_end[stop] = start
            elif o.frame < 0 and start > self.other_end[stop]:
                self.other_end[stop] = start
        else:
            raise ValueError("orf already defined")       # orfs.py:32

    def iter_orfs(self):                                  # orfs.py:34-37
        for fam in self.values():
            yield from fam.values()

    def iter_in(self):                                    # orfs.py:38-46
        for fam in self.values():
            keys = list(fam.keys())
            keys.sort(reverse=not (fam[keys[0]].frame > 0))
            yield [fam[k] for k in keys]


DEFAULT_STARTS = {"atg": Decimal("0.85"), "gtg": Decimal("0.10"), "ttg": Decimal("0.05")}
DEFAULT_STOPS = ["tag", "tga", "taa"]


def normalise_starts(w):
    """file_handling.py:58-62: weights divided by their max."""
    m = max(w.values())
    return {k: v / m for k, v in w.items()}


def substituted(dna: str) -> np.ndarray:
    """functions.py:159-163: for counting only, s,b,v -> g and every other non-acgt -> a."""
    arr = np.frombuffer(dna.encode(), dtype=np.uint8).copy()
    isg = (arr == ord('s')) | (arr == ord('b')) | (arr == ord('v'))
    acgt = (arr == ord('a')) | (arr == ord('c')) | (arr == ord('g')) | (arr == ord('t'))
    arr[~acgt] = ord('a')
    arr[isg] = ord('g')
    return arr


def p_stop_from_counts(na, nt, ng, length):
    """orfs.py:162-173 / functions.py:174-178 (left-to-right, each op rounded to 28 digits)."""
    length = Decimal(length)
    Pa, Pt, Pg = na / length, nt / length, ng / length
    return Pt * Pa * Pa + Pt * Pg * Pa + Pt * Pa * Pg


# ---------------------------------------------------------------------------------------------
# get_orfs                                                                functions.py:143-303
# ---------------------------------------------------------------------------------------------
def get_orfs(dna: str, start_codons=None, stop_codons=None, min_orf_len: int = 90, literal: bool = False):
    """Restatement of functions.get_orfs.  ``start_codons`` are already max-normalised Decimals.

    literal=True replays stage E exactly as written (two Decimal pows per codon visit,
    functions.py:286-298); literal=False memoises the six (imax,imin) factors per ORF, which is
    the same sequence of Decimal multiplications and therefore bit-identical.
    """
    if start_codons is None:
        start_codons = normalise_starts(DEFAULT_STARTS)
    if stop_codons is None:
        stop_codons = DEFAULT_STOPS
    dna = dna.lower()                                      # functions.py:144
    L = len(dna)
    for ch in set(dna):                                    # rev_comp(dna[i:i+21]) raises, functions.py:169
        if ch not in _COMP:
            raise KeyError(ch)
    orfs = OOrfs()
    orfs.seq, orfs.contig_length = dna, L

    # ---- stage A: contig statistics (functions.py:153-181)
    sub = substituted(dna)
    cnt = {b: int((sub == ord(b)).sum()) for b in "acgt"}
    fa = Decimal(cnt['a'] + cnt['t'])                      # both strands, :165-166
    fg = Decimal(cnt['g'] + cnt['c'])
    Pa = fa / (L * 2)
    Pt = fa / (L * 2)
    Pg = fg / (L * 2)
    orfs.pstop = Pt * Pa * Pa + Pt * Pg * Pa + Pt * Pa * Pg  # :178
    bgf, bgr = rbs_arrays(dna)
    background = [1.0] * 28
    hist = np.bincount(bgf, minlength=28) + np.bincount(bgr, minlength=28)
    for r in range(28):
        background[r] += float(hist[r])
    y = sum(background)
    background = [x / y for x in background]
    T = gc_window_sums(sub)

    raw = np.frombuffer(dna.encode(), dtype=np.uint8)
    cum = {b: np.concatenate(([0], np.cumsum(raw == ord(b)))) for b in "acgt"}

    def make(start, stop, length, frame, lo, hi, rbs_score, trigger):
        """lo,hi: 0-based half-open extent of orf.seq in the forward text."""
        o = OOrf()
        o.start, o.stop, o.length, o.frame, o.rbs_score, o.trigger = start, stop, length, frame, rbs_score, trigger
        na, nt, ng = (int(cum[b][hi] - cum[b][lo]) for b in "atg")
        if frame < 0:
            nc = int(cum['c'][hi] - cum['c'][lo])
            na, nt, ng = nt, na, nc                        # counts of the reverse-complemented string
        o.pstop = p_stop_from_counts(na, nt, ng, hi - lo)  # orfs.py:162-173 (length keeps non-acgt)
        if frame > 0:
            o.codon = dna[lo:lo + 3]
        else:
            o.codon = rev_comp(dna[hi - 3:hi]) if hi - 3 >= lo else rev_comp(dna[lo:hi])[:3]
        o.weight = 1
        o.weight_rbs = 1
        o.hold = 1
        orfs.add(o)
        return o

    training = [1.0] * 28
    # ---- stage B: six-frame scan (functions.py:184-251)
    stops = {1: 0, 2: 0, 3: 0, -1: 1, -2: 2, -3: 3}
    starts = {1: [], 2: [], 3: [], -1: [], -2: [], -3: []}
    for f in (1, 2, 3):
        if dna[f - 1:f + 2] not in start_codons:
            starts[f].append(f)
    rc_starts = {rev_comp(c) for c in start_codons}
    rc_stops = {rev_comp(c) for c in stop_codons}
    for i in range(1, L - 1):
        codon = dna[i - 1:i + 2]
        frame = (i - 1) % 3 + 1
        if codon in start_codons:
            starts[frame].append(i)
        elif codon in rc_starts:
            starts[-frame].append(i + 2)
        elif codon in stop_codons:
            stop = i + 2
            for start in reversed(starts[frame]):
                length = stop - start + 1
                if length >= min_orf_len:
                    rs = int(bgf[start - 21]) if start >= 21 else 0
                    make(start, stop - 2, length, frame, start - 1, stop, rs, i)
                    training[rs] += 1
            starts[frame] = []
            stops[frame] = stop
        elif codon in rc_stops:
            stop = stops[-frame]
            for start in starts[-frame]:
                length = start - stop + 1
                if length >= min_orf_len:
                    rs = int(bgr[start]) if start < L else 0
                    make(start - 2, stop, length, -frame, max(0, stop - 1), start, rs, i)
                    training[rs] += 1
            starts[-frame] = []
            stops[-frame] = i
    for frame in (1, 2, 3):                                # functions.py:229-251
        end = L - ((L - (frame - 1)) % 3)
        for start in reversed(starts[frame]):
            length = end - start + 1
            if length >= min_orf_len:
                rs = int(bgf[start - 21]) if start >= 21 else 0
                make(start, end - 2, length, frame, max(0, start - 1), end, rs, L + 2 * frame)
                training[rs] += 1
        if rev_comp(dna[end - 3:end]) not in start_codons:
            starts[-frame].append(end)
        for start in starts[-frame]:
            stop = stops[-frame]
            length = start - stop + 1
            if length >= min_orf_len:
                rs = int(bgr[start]) if start < L else 0
                make(start - 2, stop, length, -frame, max(0, stop - 1), start, rs, L + 2 * frame + 1)
                training[rs] += 1

    # ---- stage C: RBS likelihood ratio (functions.py:253-257)
    y = sum(training)
    training = [x / y for x in training]
    for o in orfs.iter_orfs():
        o.weight_rbs = training[o.rbs_score] / background[o.rbs_score]

    # ---- stage D: GC-frame training (functions.py:261-284)
    pos_max = [Decimal(1)] * 4
    pos_min = [Decimal(1)] * 4

    def cls(base, fwd):
        a, b, c = int(T[base]), int(T[base + 1]), int(T[base + 2])   # gc_pos_freq[base]
        if not fwd:
            a, c = c, a
        return max_idx(a, b, c), min_idx(a, b, c)

    for fam in orfs.iter_in():
        for o in fam:
            if o.codon == 'atg':
                if o.start < o.stop:
                    n = int((o.stop - o.start) / 8) * 3
                    rng = range(o.start + n, o.stop - 36, 3)
                    fwd = True
                elif o.stop < o.start:
                    n = int((o.start - o.stop) / 8) * 3
                    rng = range(o.start - n, o.stop + 36, -3)
                    fwd = False
                else:
                    rng, fwd = (), True
                for base in rng:
                    im, il = cls(base, fwd)
                    pos_max[im] += 1
                    pos_min[il] += 1
                break
    y = max(pos_max)
    pos_max = [x / y for x in pos_max]
    y = max(pos_min)
    pos_min = [x / y for x in pos_min]
    orfs.pos_max, orfs.pos_min = pos_max, pos_min

    # ---- stage E: per-codon product (functions.py:286-298)
    for o in orfs.iter_orfs():
        fwd = o.frame > 0
        rng = range(o.start, o.stop, 3 if fwd else -3)
        if literal:
            for base in rng:
                im, il = cls(base, fwd)
                o.hold = o.hold * (((1 - o.pstop) ** pos_max[im]) ** pos_min[il])
        else:
            memo = {}
            hold = o.hold
            for base in rng:
                key = cls(base, fwd)
                fac = memo.get(key)
                if fac is None:
                    fac = memo[key] = ((1 - o.pstop) ** pos_max[key[0]]) ** pos_min[key[1]]
                hold = hold * fac
            o.hold = hold
    # ---- Orf.score (orfs.py:122-127)
    for o in orfs.iter_orfs():
        s = 1 / o.hold
        if o.codon in start_codons:
            s = s * start_codons[o.codon]
        s = s * Decimal(str(o.weight_rbs))
        o.weight = -s
    orfs.background, orfs.training, orfs.bgf, orfs.bgr, orfs.T = background, training, bgf, bgr, T
    return orfs


# ---------------------------------------------------------------------------------------------
# edge scores                                                               functions.py:26-46,140
# ---------------------------------------------------------------------------------------------
def score_overlap(length, direction, pstop):
    o = Decimal(1 - pstop)
    score = 1 / (Decimal(o) ** Decimal(length))
    if direction == 'diff':
        score = score + (1 / Decimal('0.05'))
    return score


def score_gap(length, direction, pgap):
    g = Decimal(1 - pgap)
    if length > 300:
        return Decimal(g) ** Decimal(100) + length
    score = 1 / (Decimal(g) ** Decimal(length / 3))
    if direction == 'diff':
        score = score + (1 / Decimal('0.05'))
    return score


def ave(a):
    return Decimal(sum(a) / len(a))


# ---------------------------------------------------------------------------------------------
# get_graph                                                               functions.py:307-454
# ---------------------------------------------------------------------------------------------
class ONode(tuple):
    """(gene, type, frame, position); repr is the reference's Node repr (nodes.py:14-21)."""
    __slots__ = ()

    def __repr__(self):
        return "Node(%r,%r,%r,%r)" % tuple(self)


def edge_string(src, dst, weight) -> str:
    """edges.py:17-23"""
    return "%r\t%r\t%s" % (src, dst, str(weight * 1000))


def get_graph(orfs: OOrfs):
    """Returns (nodes in insertion order, edges as (src, dst, Decimal weight) in Graph.iteredges() order).

    The connect step is a sorted-window join instead of the reference's all-pairs loop
    (functions.py:360-438); out-edge lists are then ordered the way the double loop would have
    appended them (outer = right node, inner = left node, both in node insertion order).
    """
    L = orfs.contig_length
    pgap = orfs.pstop
    node_idx = {}
    out = []               # per node: list of (phase, k1, k2, dst, weight)

    def nid(n):
        if n not in node_idx:
            node_idx[n] = len(node_idx)
            out.append([])
        return node_idx[n]

    def add(src, dst, w, key):
        if src == dst:
            raise ValueError("loops are forbidden")        # graphs.py:67-68
        a = nid(src)
        nid(dst)
        for e in out[a]:
            if e[3] == dst:
                raise ValueError("parallel edges are forbidden")   # graphs.py:73-74
        out[a].append(key + (dst, w))

    seq = 0
    for o in orfs.iter_orfs():                             # functions.py:311-318
        if o.frame > 0:
            s, t = ONode(('CDS', 'start', o.frame, o.start)), ONode(('CDS', 'stop', o.frame, o.stop))
        else:
            s, t = ONode(('CDS', 'stop', o.frame, o.stop)), ONode(('CDS', 'start', o.frame, o.start))
        add(s, t, o.weight, (0, seq, 0))
        seq += 1
    nodes = list(node_idx.keys())
    pos = np.array([n[3] for n in nodes], dtype=np.int64)
    order = np.argsort(pos, kind="stable")
    spos = pos[order]

    def is_exit(n):
        return (n[1] == 'stop' and n[2] > 0) or (n[1] == 'start' and n[2] < 0)

    # ---- bridges over >500-bp uncovered runs (functions.py:320-354)
    covered = np.zeros(L + 1, dtype=bool)
    for fam in orfs.iter_in():
        o = fam[0]
        mi, ma = min(o.start, o.stop), max(o.start, o.stop)
        covered[mi:min(ma, L - 1)] = True
    covered[0] = False                                     # 'if(base)' is false for 0
    last = 0
    bseq = 0
    for base in np.nonzero(covered)[0]:
        base = int(base)
        if base - last > 500:
            for ri, right in enumerate(nodes):
                r = right[3]
                if not (base - 1 <= r < base + 500):
                    continue
                for li, left in enumerate(nodes):
                    l = left[3]
                    if not (last + 1 >= l > last - 500):
                        continue
                    if left[2] * right[2] > 0:
                        if (left[1] == 'stop' and right[1] == 'start' and left[2] > 0) or \
                           (left[1] == 'start' and right[1] == 'stop' and left[2] < 0):
                            add(left, right, score_gap(r - l - 3, 'same', pgap), (1, bseq, 0))
                            bseq += 1
                    else:
                        if (left[1] == 'stop' and right[1] == 'stop' and left[2] > 0) or \
                           (left[1] == 'start' and right[1] == 'start' and left[2] < 0):
                            add(left, right, score_gap(r - l - 3, 'diff', pgap), (1, bseq, 0))
                            bseq += 1
        last = base

    # ---- connect (functions.py:360-438)
    oe = orfs.other_end

    def o_of(p):                                           # functions.py:373-385
        if p in orfs and oe[p] in orfs[p]:
            return orfs[p][oe[p]].pstop
        if p in orfs:
            fam = orfs.get(oe[p])
            if fam is None or p not in fam:
                raise ValueError("orf not found")
            return fam[p].pstop
        return pgap

    gap_memo = {}

    def gap(length, direction):
        k = (length, direction)
        v = gap_memo.get(k)
        if v is None:
            v = gap_memo[k] = score_gap(length, direction, pgap)
        return v

    for ri, right in enumerate(nodes):
        r = right[3]
        lo = np.searchsorted(spos, r - 499, side="left")
        hi = np.searchsorted(spos, r, side="left")
        if hi <= lo:
            continue
        r_other = oe[r]
        for li in sorted(int(x) for x in order[lo:hi]):
            left = nodes[li]
            l = left[3]
            l_other = oe[l]
            lt, rt, lf, rf = left[1], right[1], left[2], right[2]
            if lf * rf > 0:
                if lt == 'stop' and rt == 'start':
                    if lf > 0:
                        add(left, right, gap(r - l - 3, 'same'), (2, ri, li))
                    elif lf != rf and r < l_other and r_other < l:
                        add(right, left, score_overlap(r - l + 3, 'same', ave([o_of(l), o_of(r)])), (2, ri, li))
                if lt == 'start' and rt == 'stop':
                    if lf > 0:
                        if lf != rf and r < l_other and r_other < l:
                            add(right, left, score_overlap(r - l + 3, 'same', ave([o_of(l), o_of(r)])), (2, ri, li))
                    else:
                        add(left, right, gap(r - l - 3, 'same'), (2, ri, li))
            else:
                if lt == 'stop' and rt == 'stop':
                    if rf > 0:
                        if r_other + 3 < l and r < l_other:
                            add(right, left, score_overlap(r - l + 3, 'diff', ave([o_of(l), o_of(r)])), (2, ri, li))
                    else:
                        add(left, right, gap(r - l - 3, 'diff'), (2, ri, li))
                if lt == 'start' and rt == 'start':
                    if rf > 0 and r - l > 2:
                        add(left, right, gap(r - l - 3, 'diff'), (2, ri, li))
                    elif rf < 0:
                        if r_other < l and r < l_other:
                            add(right, left, score_overlap(r - l + 3, 'diff', ave([o_of(l), o_of(r)])), (2, ri, li))

    # ---- terminals (functions.py:440-452)
    source = ONode(('source', 'source', 0, 0))
    target = ONode(('target', 'target', 0, L + 1))
    nid(source)
    nid(target)
    for ni, n in enumerate(nodes):
        p = n[3]
        if p <= 2000 and not is_exit(n):
            add(source, n, gap(p, 'same'), (3, ni, 0))
        if L - p <= 2000 and is_exit(n):
            add(n, target, gap(L - p, 'same'), (3, ni, 0))
    all_nodes = list(node_idx.keys())
    edges = []
    for a, n in enumerate(all_nodes):
        for e in sorted(out[a], key=lambda e: e[:3]):
            edges.append((n, e[3], e[4]))
    return all_nodes, edges


# ---------------------------------------------------------------------------------------------
# solve + calls                                   phanotate.py:53-76, fastpathz contract (CHANGELOG)
# ---------------------------------------------------------------------------------------------
def int_weight(w: Decimal) -> int:
    """Integer the solver sees: the integer part of Decimal*1000 (phanotate.py:55, CHANGELOG.md:13,57)."""
    return int((w * 1000).to_integral_value(rounding=ROUND_DOWN))


def shortest_path_literal(nodes, edges, source, target):
    """Exact-integer Bellman-Ford, edges in iteredges() order, strict '<', passes to fixpoint."""
    idx = {n: i for i, n in enumerate(nodes)}
    E = [(idx[a], idx[b], int_weight(w)) for a, b, w in edges]
    dist = [None] * len(nodes)
    par = [-1] * len(nodes)
    dist[idx[source]] = 0
    while True:
        changed = False
        for u, v, w in E:
            du = dist[u]
            if du is None:
                continue
            nd = du + w
            if dist[v] is None or nd < dist[v]:
                dist[v], par[v], changed = nd, u, True
        if not changed:
            break
    if dist[idx[target]] is None:
        return []
    path, v = [], idx[target]
    while v != -1:
        path.append(nodes[v])
        v = par[v]
    return path[::-1]


def shortest_path(nodes, edges, source, target):
    """The same Bellman-Ford (same relaxations in the same order, hence the same parents on exact ties), without
    the scans that cannot change anything: relaxing u->v can only succeed if dist[u] fell since the edge was last
    scanned, so a pass scans the edge groups of the nodes whose distance fell since their group was last scanned.
    iteredges() groups the edges by source node in node insertion order (graphs.py:121-126); a node improved by a
    group scanned later in the pass waits for the next pass, one improved by an earlier group is scanned in this
    pass -- exactly what the full scan does.  Linear instead of quadratic in the contig length (the number of
    passes grows with the path length), which is what makes Mb-sized goldens affordable.  Falls back to the
    literal scan if the edge list is not grouped as expected."""
    import heapq
    idx = {n: i for i, n in enumerate(nodes)}
    groups = {}
    order = []
    last = None
    for a, b, w in edges:
        u = idx[a]
        if u != last:
            if u in groups:
                return shortest_path_literal(nodes, edges, source, target)
            groups[u] = []
            order.append(u)
            last = u
        groups[u].append((idx[b], int_weight(w)))
    rank = {u: k for k, u in enumerate(order)}          # position of a node's group in the scan
    dist = [None] * len(nodes)
    par = [-1] * len(nodes)
    s = idx[source]
    dist[s] = 0
    cur, nxt = [], []
    queued = set()
    if s in rank:
        cur.append(rank[s])
        queued.add(s)
    while cur:
        while cur:
            k = heapq.heappop(cur)
            u = order[k]
            queued.discard(u)
            du = dist[u]
            for v, w in groups[u]:
                nd = du + w
                if dist[v] is None or nd < dist[v]:
                    dist[v], par[v] = nd, u
                    if v in rank and v not in queued:
                        queued.add(v)
                        if rank[v] > k:
                            heapq.heappush(cur, rank[v])
                        else:
                            nxt.append(rank[v])
        cur, nxt = nxt, []
        heapq.heapify(cur)
    if dist[idx[target]] is None:
        return []
    path, v = [], idx[target]
    while v != -1:
        path.append(nodes[v])
        v = par[v]
    return path[::-1]


def shortest_path_fast(nodes, edges, source, target):
    """What the edge-order Bellman-Ford above ends with, computed without replaying its passes (for Mb-sized contigs, where
    the replay is quadratic: every pass moves the improvements a few nodes further).  DERIVED, not literal:

    1. the distances do not depend on the scan order: label correcting in node-position order;
    2. the parent of v is the source of the FIRST scanned edge that offers v its final distance (any earlier offer is
       larger, any later one is not strictly smaller).  An edge u->v offers it the first time u's group is scanned after u
       itself became final.  If u became final while the group with scan rank k was being scanned in pass p, u's group
       (rank r) is scanned in pass p when r > k, else in pass p+1.  So T(v) = min over tight in-edges of (pass, r), computed
       in increasing T like Dijkstra (offers are strictly later than T(u)).

    tests/test_oracle.py checks it against shortest_path on every fixture and on bench contigs with exact ties."""
    import heapq
    idx = {n: i for i, n in enumerate(nodes)}
    groups, order, last = {}, [], None
    for a, b, w in edges:
        u = idx[a]
        if u != last:
            if u in groups:
                raise ValueError("edges are not grouped by source node")
            groups[u] = []
            order.append(u)
            last = u
        groups[u].append((idx[b], int_weight(w)))
    rank = {u: k for k, u in enumerate(order)}
    n = len(nodes)
    s, t = idx[source], idx[target]
    pos = [nd[3] for nd in nodes]
    dist = [None] * n
    dist[s] = 0
    heap, queued = [(pos[s], s)], {s}
    while heap:
        _, u = heapq.heappop(heap)
        queued.discard(u)
        du = dist[u]
        for v, w in groups.get(u, ()):
            nd = du + w
            if dist[v] is None or nd < dist[v]:
                dist[v] = nd
                if v not in queued:
                    queued.add(v)
                    heapq.heappush(heap, (pos[v], v))
    if dist[t] is None:
        return []
    par = [-1] * n
    T = [None] * n
    T[s] = (1, -1)
    heap = [(T[s], s)]
    done = [False] * n
    while heap:
        tu, u = heapq.heappop(heap)
        if done[u] or tu != T[u]:
            continue
        done[u] = True
        if u not in rank:
            continue
        r = rank[u]
        when = (tu[0] if r > tu[1] else tu[0] + 1, r)
        du = dist[u]
        for v, w in groups[u]:
            if du + w == dist[v] and (T[v] is None or when < T[v]):
                T[v] = when
                par[v] = u
                heapq.heappush(heap, (when, v))
    path, v = [], t
    while v != -1:
        path.append(nodes[v])
        v = par[v]
    return path[::-1]


def calls_from_path(path, edges):
    """phanotate.py:65-76 + locus.py:29-37: consecutive non-overlapping pairs after dropping the source."""
    wmap = {(a, b): w for a, b, w in edges}
    path = path[1:]
    it = iter(path)
    rows = []
    for left, right in zip(it, it):
        w = wmap.get((left, right), 0)
        rows.append((left[3], right[3] + 2, '+' if left[2] > 0 else '-', "%E" % w, w))
    return rows


def call_contig(dna: str, start_codons=None, stop_codons=None, min_orf_len: int = 90, literal=False, fast=False):
    """Whole path for one contig -> (orfs, nodes, edges, rows).  fast: shortest_path_fast (Mb-sized contigs)."""
    orfs = get_orfs(dna, start_codons, stop_codons, min_orf_len, literal)
    nodes, edges = get_graph(orfs)
    rows = []
    if len(nodes) > 2:                                     # phanotate.py:63
        source = ONode(('source', 'source', 0, 0))
        target = ONode(('target', 'target', 0, len(dna) + 1))
        rows = calls_from_path((shortest_path_fast if fast else shortest_path)(nodes, edges, source, target), edges)
    return orfs, nodes, edges, rows


def orf_table_lines(orfs):
    return ["%d,%d,%d,%d,%s,%s\n" % (o.start, o.stop, o.frame, o.rbs_score, o.pstop, o.weight)
            for o in orfs.iter_orfs()]


def edge_dump_lines(edges):
    return [edge_string(a, b, w) + "\n" for a, b, w in edges]


def calls_lines(rows):
    return ["%d\t%d\t%s\t%s\n" % r[:4] for r in rows]
