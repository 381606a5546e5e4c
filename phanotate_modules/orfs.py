"""ORF containers of the boundary (reference orfs.py:6-139).  Instances are filled from the device
ORF table by functions.get_orfs; the dict semantics (stop -> {start -> Orf}, other_end keyed by bare
position, insertion order) are the reference's."""


class Orfs(dict):
    def __init__(self, locus):
        super().__init__()
        self.pstop = 0
        self.min_orf_len = locus.min_orf_len
        self.contig_length = 0
        self.seq = ''
        self.other_end = dict()
        self.start_codons = locus.start_codons
        self.stop_codons = locus.stop_codons

    def _insert(self, o):
        """container bookkeeping of add_orf (orfs.py:17-32) for an already scored Orf"""
        start, stop = o.start, o.stop
        if stop not in self:
            self[stop] = {start: o}
            self.other_end[stop] = start
            self.other_end[start] = stop
        elif start not in self[stop]:
            self[stop][start] = o
            self.other_end[start] = stop
            if (o.frame > 0 and start < self.other_end[stop]) or (o.frame < 0 and start > self.other_end[stop]):
                self.other_end[stop] = start
        else:
            raise ValueError("orf already defined")

    def iter_orfs(self):
        for family in self.values():
            yield from family.values()

    def _sorted(self, longest_first):
        for family in self.values():
            keys = list(family.keys())
            forward = family[keys[0]].frame > 0
            keys.sort(reverse=(forward != longest_first))
            yield (family[k] for k in keys)

    def iter_in(self):
        return self._sorted(True)

    def iter_out(self):
        return self._sorted(False)

    def get_orf(self, start, stop):
        if stop not in self:
            raise ValueError(" orf with stop codon not found")
        if start not in self[stop]:
            raise ValueError("orf with start codon not found")
        return self[stop][start]


class Orf:
    def __init__(self, start, stop, length, frame, seq, rbs, rbs_score, start_codons, stop_codons):
        self.start, self.stop, self.length, self.frame = start, stop, length, frame
        self.seq, self.rbs, self.rbs_score = seq, rbs, rbs_score
        self.start_codons, self.stop_codons = start_codons, stop_codons
        self.pstop = None
        self.weight = 1
        self.weight_start = 1
        self.weight_rbs = 1
        self.hold = 1                  # product over the codons (functions.py:286-298); filled from the device (csrc/hold.cuh)
        self.gcfp_mins = 1
        self.gcfp_maxs = 1

    def start_codon(self):
        return self.seq[0:3]

    def stop_codon(self):
        return self.seq[-3:]

    def has_start(self):
        return self.start_codon() in self.start_codons

    def has_stop(self):
        return self.stop_codon() in self.stop_codons

    def score(self):
        """orfs.py:122-127: weight = -(1/hold x start-codon weight x Decimal(str(weight_rbs))).  The device already ran this
        (csrc/hold.cuh: st_orf_finish) and functions.get_orfs stored its result; calling it again recomputes the same
        Decimal from `hold`."""
        from decimal import Decimal
        s = 1 / self.hold
        if self.start_codon() in self.start_codons:
            s = s * self.start_codons[self.start_codon()]
        s = s * Decimal(str(self.weight_rbs))
        self.weight = -s

    def __repr__(self):
        return "%s(%r,%r,%r,%r,%r)" % (self.__class__.__name__, self.start, self.stop, self.frame,
                                       self.weight_rbs, self.weight)
