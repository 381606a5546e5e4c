#!/usr/bin/env python3
"""PHANOTATE command line (same flags and outputs as reference phanotate.py:24-77), B200 back end.

All loci of the input go to the GPU as ONE batch (contigs are independent: phanotate.py:40-56); the
per-locus loop below only formats results.  `--dump` prints the edge list of the first locus in the
reference's order and exits (phanotate.py:57-61).
"""
import sys

from phanotate_modules import file_handling
from phanotate_modules.file import File
from phanotate_b200 import mirror
from phanotate_b200.engine import make_params


def main(argv=None):
    from phanotate_modules import functions
    args = file_handling.get_args(File, argv)
    if args.format == 'fasta':
        args.format = 'fna'
    if args.format == 'tabular' and not args.dump:
        # batch fast path: vectorised FASTA ingest -> one engine run -> the tabular text of every locus
        from phanotate_b200 import fastio
        eng = functions.engine()
        names, bases, offs = fastio.read_fasta_packed(args.infile, eng.lib)
        if len(bases) == 0:
            sys.stdout.write("Error: no sequences found in infile\n")
            return 0
        res = eng.run_packed(bases, offs, make_params(args.start_codons, args.stop_codons, args.min_orf_len))
        fastio.write_tabular(res, names, args.outfile, check=True, lib=eng.lib)
        return 0
    genbank = File(args.infile)
    if not genbank.seq():
        sys.stdout.write("Error: no sequences found in infile\n")
        return 0
    loci = list(genbank)
    params = make_params(args.start_codons, args.stop_codons, args.min_orf_len)
    res = functions.engine().run([l.seq().encode() for l in loci], params)
    for k, locus in enumerate(loci):
        locus.start_codons, locus.stop_codons, locus.min_orf_len = args.start_codons, args.stop_codons, args.min_orf_len
        res.check(k)                                   # KeyError / ValueError like the reference would raise
        if args.dump:
            sys.stderr.write("Warning: tRNAscan or Aragorn were not found, proceding without tRNA masking.\n")
            sys.stdout.writelines(mirror.ContigGraph(res, k).dump_lines())
            return 0
        c = res.contigs[k]
        from phanotate_b200 import _native as N
        for r in res.calls[c["call_off"]:c["call_off"] + c["n_calls"]]:
            weight = float(r["score"])                             # '%E' % Decimal goes through float() as well
            strand = 1 if r["strand"] > 0 else -1
            pairs = [[int(r["left"]), int(r["right"]) - 2]]       # add_feature adds the 2 back (locus.py:30)
            feature = locus.add_feature('CDS', strand, pairs, {'note': ['score:%E' % weight]})
            feature.weight = '%E' % weight
        locus.write(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
