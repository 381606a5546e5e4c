"""Locus record + the tabular writer (reference locus.py:23-56).  FASTA-only input: the reference gets
parsing and the other writers from the third-party ``genbank`` package, which is not vendored; the
genbank / fna / faa writers below follow the examples in the reference's README.md:40-68."""
import sys
import textwrap

from .feature import Feature

_AA = 'FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG'
_CODONS = [a + b + c for a in 'tcag' for b in 'tcag' for c in 'tcag']
_TABLE = dict(zip(_CODONS, _AA))


class Locus:
    def __init__(self, name, dna):
        self._name, self._dna = name, dna
        self._features = []
        self.start_codons = self.stop_codons = self.min_orf_len = None

    def name(self):
        return self._name

    def seq(self):
        return self._dna

    def length(self):
        return len(self._dna)

    def add_feature(self, key, strand, pairs, tags=dict()):
        pairs[-1][-1] += 2                                   # node positions are codon starts (locus.py:30)
        f = Feature(key, strand, [[str(x) for x in p] for p in pairs], self, tags)
        self._features.append(f)
        return f

    def features(self, include=None):
        return [f for f in self._features if include is None or f.type in include]

    def tabular(self, outfile=sys.stdout):                   # locus.py:39-56
        outfile.write("#id:\t" + self.name() + "\n")
        outfile.write("#START\tSTOP\tFRAME\tCONTIG\tSCORE\n")
        for f in self.features(include=['CDS']):
            left, right = f.pairs[0][0], f.pairs[-1][-1]
            if f.strand < 0:
                right, left = left, right
            outfile.write("%s\t%s\t%s\t%s\t%s\n" % (left, right, chr(44 - f.strand), self.name(), f.weight))

    def _location(self, f):
        loc = "%s..%s" % (f.pairs[0][0], f.pairs[-1][-1])
        return "complement(%s)" % loc if f.strand < 0 else loc

    def genbank(self, outfile=sys.stdout):                   # README.md:40-54
        outfile.write("LOCUS       %s %20d bp \n" % (self.name(), self.length()))
        outfile.write("FEATURES             Location/Qualifiers\n")
        for f in self.features():                            # CDS calls and tRNA hits on the path (phanotate.py:71: left.gene)
            outfile.write("     %-16s%s\n" % (f.type, self._location(f)))
            for k, vals in f.tags.items():
                for v in vals:
                    outfile.write("                     /%s=%s\n" % (k, v))
        outfile.write("ORIGIN\n")
        dna = self._dna.lower()
        for i in range(0, len(dna), 60):
            outfile.write("%9d %s\n" % (i + 1, " ".join(textwrap.wrap(dna[i:i + 60], 10))))
        outfile.write("//\n")

    def _header(self, f):
        tags = " ".join("[%s=%s]" % (k, v) for k, vals in f.tags.items() for v in vals)
        return ">%s_CDS_[%s] %s\n" % (self.name(), self._location(f), tags)

    def fna(self, outfile=sys.stdout):                       # README.md:56-61
        for f in self.features(include=['CDS']):
            outfile.write(self._header(f))
            outfile.write(f.seq() + "\n")

    def faa(self, outfile=sys.stdout):                       # README.md:63-68
        for f in self.features(include=['CDS']):
            s = f.seq()
            aa = "".join(_TABLE.get(s[i:i + 3], 'X') for i in range(0, len(s) - 2, 3))
            outfile.write(self._header(f))
            outfile.write(aa + "\n")

    def gff3(self, outfile=sys.stdout):
        """`-f gff` / `-f gff3`: the reference lists these formats (README.md:40) but their writer lives in the absent
        ``genbank`` package and the reference shows no example, so this is plain GFF3 (one `CDS` row per call, the score
        column = the printed score), NOT pinned to the reference's bytes."""
        outfile.write("##gff-version 3\n")
        outfile.write("##sequence-region %s 1 %d\n" % (self.name(), self.length()))
        for k, f in enumerate(self.features(include=['CDS'])):
            outfile.write("%s\tPHANOTATE\tCDS\t%s\t%s\t%s\t%s\t0\tID=%s_CDS_%d\n" % (
                self.name(), f.pairs[0][0], f.pairs[-1][-1], f.weight, '+' if f.strand > 0 else '-', self.name(), k + 1))

    gff = gff3

    def write(self, args):
        getattr(self, args.format)(args.outfile)
