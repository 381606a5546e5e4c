"""Minimal CDS feature record (stands in for the third-party genbank.feature.Feature the reference
subclasses in feature.py:5-38; that package is not vendored)."""


class Feature:
    def __init__(self, key, strand, pairs, locus, tags=None):
        self.type, self.strand, self.locus = key, strand, locus
        self.pairs = tuple(tuple(p) for p in pairs)
        self.tags = dict(tags or {})
        self.weight = None

    def left(self):
        return int(self.pairs[0][0])

    def right(self):
        return int(self.pairs[-1][-1])

    def seq(self):
        dna = self.locus.seq()[self.left() - 1:self.right()].lower()
        if self.strand < 0:
            from .functions import rev_comp
            dna = rev_comp(dna)
        return dna

    def start_codon(self):
        return self.seq()[:3]

    def stop_codon(self):
        return self.seq()[-3:]

    def has_start(self):
        return self.start_codon() in self.locus.start_codons

    def has_stop(self):
        return self.stop_codon() in self.locus.stop_codons
