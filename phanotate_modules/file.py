"""Multi-record FASTA reader (plain or .gz) yielding Locus objects -- stands in for the third-party
genbank.file.File the reference subclasses in file.py:4."""
import gzip

from .locus import Locus


class File:
    formats = ['tabular', 'genbank', 'fna', 'faa', 'fasta', 'gff', 'gff3']      # (get_args takes formats[:7])

    def __init__(self, path):
        self.path = path
        self._loci = []
        opener = gzip.open if str(path).endswith('.gz') else open
        name, parts = None, []
        with opener(path, 'rt') as fh:
            for line in fh:
                line = line.rstrip("\r\n")
                if line.startswith('>'):
                    if name is not None:
                        self._loci.append(Locus(name, "".join(parts)))
                    name, parts = (line[1:].split() or [''])[0], []
                elif name is not None:
                    parts.append(line.strip())
        if name is not None:
            self._loci.append(Locus(name, "".join(parts)))

    def seq(self):
        return "".join(l.seq() for l in self._loci)

    def __iter__(self):
        return iter(self._loci)

    def __len__(self):
        return len(self._loci)
