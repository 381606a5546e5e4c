#!/usr/bin/env python3
"""md5 of the ORACLE's call table (tests' checker, oracle/phanotate_oracle.py -- itself pinned to the reference goldens)
for the first N contigs of the bench workload (BASELINE.json config 4).  A DERIVED golden: it lets the GPU suite compare
thousands of contigs, ties included, without running the oracle on the GPU box.

    python tests/golden/make_synth4_digest.py [N=3000]
"""
import hashlib
import json
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def one(k):
    from phanotate_b200 import synth
    from oracle import phanotate_oracle as O
    rows = O.call_contig(synth.synth4_contig(k).decode())[3]
    return hashlib.md5("".join(O.calls_lines(rows)).encode()).hexdigest()[:16], len(rows)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    with Pool(os.cpu_count()) as pool:
        out = pool.map(one, range(n), chunksize=4)
    json.dump({"workload": "synth4 contigs 0..%d (50,000 bp each)" % (n - 1), "md5_16": [o[0] for o in out],
               "n_calls": [o[1] for o in out]}, open(os.path.join(HERE, "synth4_calls_digest.json"), "w"))
    print(n, sum(o[1] for o in out))


if __name__ == "__main__":
    main()
