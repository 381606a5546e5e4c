#!/usr/bin/env python3
"""PHANOTATE command line (same flags and outputs as reference phanotate.py:24-77), B200 back end.

The loci of the input go to the GPU in batches of whole contigs (contigs are independent: phanotate.py:40-56), one
batch unless the input holds more than PB200_MAX_BATCH_BASES (default 2^30) bases; the per-locus loop below only formats
results.  `--dump` prints the edge list of the first locus in the reference's order and exits (phanotate.py:57-61).
"""
import os
import sys

from phanotate_modules import file_handling
from phanotate_modules.file import File
from phanotate_b200 import mirror
from phanotate_b200.engine import make_params


def batches(lengths, limit):
    """[(first, last+1)] runs of consecutive contigs whose bases stay within `limit` (a longer contig runs alone)."""
    out, a, acc = [], 0, 0
    for k, n in enumerate(lengths):
        if k > a and acc + n > limit:
            out.append((a, k))
            a, acc = k, 0
        acc += n
    if len(lengths) > a:
        out.append((a, len(lengths)))
    return out


def main(argv=None):
    from phanotate_modules import functions
    from phanotate_b200.engine import PhanotateError
    limit = int(os.environ.get("PB200_MAX_BATCH_BASES", 1 << 30))
    args = file_handling.get_args(File, argv)
    if args.format == 'fasta':
        args.format = 'fna'
    import shutil
    have_trna_tools = bool(shutil.which("aragorn") or shutil.which("tRNAscan-SE"))
    if not have_trna_tools:
        sys.stderr.write("Warning: tRNAscan or Aragorn were not found, proceding without tRNA masking.\n")
    if args.format == 'tabular' and not args.dump and not have_trna_tools:
        # batch fast path: vectorised FASTA ingest -> one engine run -> the tabular text of every locus
        from phanotate_b200 import fastio
        eng = functions.engine()
        names, bases, offs = fastio.read_fasta_packed(args.infile, eng.lib)
        if len(bases) == 0:
            sys.stdout.write("Error: no sequences found in infile\n")
            return 0
        params = make_params(args.start_codons, args.stop_codons, args.min_orf_len)
        for a, b in batches([int(offs[k + 1] - offs[k]) for k in range(len(names))], limit):
            res = eng.run_packed(bases[offs[a]:offs[b]], offs[a:b + 1] - offs[a], params)
            fastio.write_tabular(res, names[a:b], args.outfile, check=True, lib=eng.lib)
        return 0
    genbank = File(args.infile)
    if not genbank.seq():
        sys.stdout.write("Error: no sequences found in infile\n")
        return 0
    loci = list(genbank)
    params = make_params(args.start_codons, args.stop_codons, args.min_orf_len)
    for a, b in batches([len(l.seq()) for l in loci], limit):
        # tRNA masking (functions.py:457-509): the two programs run on the host, locus by locus like in the reference; their
        # hits go into the batch run
        trnas = []
        if have_trna_tools:
            for k, locus in enumerate(loci[a:b]):
                trnas += [(k, s, e) for s, e in functions.find_trnas(locus.seq().lower()) or []]
        res = functions.engine().run([l.seq().encode() for l in loci[a:b]], params, trnas=trnas)
        for k, locus in enumerate(loci[a:b]):
            locus.start_codons, locus.stop_codons, locus.min_orf_len = args.start_codons, args.stop_codons, args.min_orf_len
            try:
                res.check(k)                               # KeyError / ValueError like the reference would raise
            except PhanotateError as e:                    # a contig only this implementation cannot finish: the others still print
                sys.stderr.write("Warning: %s: %s; contig left out\n" % (locus.name(), e))
                continue
            if args.dump:
                sys.stdout.writelines(mirror.ContigGraph(res, k).dump_lines())
                return 0
            c = res.contigs[k]
            for r in res.calls[c["call_off"]:c["call_off"] + c["n_calls"]]:
                weight = float(r["score"])                         # '%E' % Decimal goes through float() as well
                strand = 1 if r["strand"] > 0 else -1
                gene = 'tRNA' if abs(int(r["strand"])) == 2 else 'CDS'     # left.gene (phanotate.py:71)
                pairs = [[int(r["left"]), int(r["right"]) - 2]]   # add_feature adds the 2 back (locus.py:30)
                feature = locus.add_feature(gene, strand, pairs, {'note': ['score:%E' % weight]})
                feature.weight = '%E' % weight
            locus.write(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
