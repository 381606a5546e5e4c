"""Command line flags and small helpers of the boundary (reference file_handling.py:24-26,42-86)."""
import argparse
import os.path
import sys
from argparse import RawTextHelpFormatter
from decimal import Decimal

try:
    from importlib.metadata import version as _v
    __version__ = _v('phanotate')
except Exception:
    __version__ = 'unknown'


def pairwise(iterable):
    """non-overlapping pairs (s0,s1), (s2,s3), ...  (file_handling.py:24-26)"""
    it = iter(iterable)
    return zip(it, it)


def is_valid_file(x):
    if not os.path.exists(x):
        raise argparse.ArgumentTypeError("{0} does not exist".format(x))
    return x


def get_args(File, argv=None):
    """file_handling.py:42-68: same flags, defaults and start-codon weight normalisation."""
    usage = 'phanotate.py [-opt1, [-opt2, ...]] infile'
    p = argparse.ArgumentParser(description='PHANOTATE: A phage genome annotator', formatter_class=RawTextHelpFormatter,
                                usage=usage)
    p.add_argument('infile', type=is_valid_file, help='input file in fasta format')
    p.add_argument('-o', '--outfile', action="store", default=sys.stdout, type=argparse.FileType('w'),
                   help='where to write the output [stdout]')
    p.add_argument('-f', '--format', help='Output the features in the specified format [tabular]', type=str,
                   default='tabular', choices=File.formats[:7])
    p.add_argument('-s', '--start_codons', action="store", default="atg:0.85,gtg:0.10,ttg:0.05", dest='start_codons',
                   help='comma separated list of start codons and frequency [atg:0.85,gtg:0.10,ttg:0.05]')
    p.add_argument('-e', '--stop_codons', action="store", default="tag,tga,taa", dest='stop_codons',
                   help='comma separated list of stop codons [tag,tga,taa]')
    p.add_argument('-l', '--minlen', action="store", type=int, default=90, dest='min_orf_len', help='to store a variable')
    p.add_argument('-d', '--dump', action="store_true")
    p.add_argument('-V', '--version', action='version', version=__version__)
    args = p.parse_args(argv)
    weights = {}
    for item in args.start_codons.split(','):
        codon, w = item.split(':')
        weights[codon.lower()] = Decimal(w)
    top = max(weights.values())
    args.start_codons = {k: v / top for k, v in weights.items()}
    args.stop_codons = [c.lower() for c in args.stop_codons.split(',')]
    return args


def read_fasta(filepath):
    """{'>name': lower-cased sequence} (file_handling.py:71-86)"""
    contigs, name, parts = {}, '', []
    with open(filepath) as fh:
        for line in fh:
            if line.startswith(">"):
                contigs[name] = "".join(parts)
                name, parts = line.split()[0], []
            else:
                parts.append(line.replace("\n", "").lower())
    contigs[name] = "".join(parts)
    contigs.pop('', None)
    return contigs
