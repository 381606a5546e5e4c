#!/usr/bin/env python3
"""Goldens for long contigs (BASELINE.json config 5: one contig, intra-contig solve).

    python tests/golden/make_long_golden.py            # everything (the 10-Mb row takes ~10 min)

Inputs: `long<N>` = the first N windows of synth.long_contig() (config-4 windows 10**6 .. concatenated, 50 kb each).

* long4 (200 kb): the REFERENCE's own get_orfs / get_graph (imported from /root/reference) + the exact-integer
  edge-order Bellman-Ford of make_golden.py  ->  long4.calls.tsv, md5s of the ORF table and the edge dump.
  (The reference's connect loop is O(N^2), functions.py:360-438: 200 kb is what it finishes in minutes.)
* long20 (1 Mb), long40 (2 Mb), long200 (10 Mb = config 5 itself): the oracle (oracle/phanotate_oracle.py, pinned to the
  reference goldens by tests/test_oracle.py) with shortest_path_fast -- the edge-order Bellman-Ford's result without
  replaying its quadratic passes, itself checked against the literal replay in tests/test_oracle.py.  DERIVED goldens.
  long20 is also run through the replayed Bellman-Ford (shortest_path) and must agree.
"""
import gzip
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from phanotate_b200 import synth  # noqa: E402


def oracle_rows(nwin, fast=True):
    from oracle import phanotate_oracle as O
    seq = synth.long_contig(nwin).decode()
    t = time.time()
    rows = O.call_contig(seq, fast=fast)[3]
    return "".join(O.calls_lines(rows)), len(rows), time.time() - t


def main():
    which = sys.argv[1:] or ["ref4", "20", "40", "200"]
    path = os.path.join(HERE, "long_index.json")
    index = json.load(open(path)) if os.path.exists(path) else {}
    for w in which:
        if w == "ref4":
            sys.path.insert(0, "/root/reference")
            import make_golden as MG
            t = time.time()
            rec, orf_lines, edge_lines, calls = MG.run_reference(("long4", synth.long_contig(4).decode(), False))
            rec["seconds"] = round(time.time() - t, 1)
            rec["by"] = "reference get_orfs/get_graph + edge-order Bellman-Ford (make_golden.run_reference)"
            with open(os.path.join(HERE, "long4.calls.tsv"), "w") as fh:
                fh.writelines(calls)
            index["long4"] = rec
        else:
            n = int(w)
            text, rows, dt = oracle_rows(n)
            rec = {"name": "long%d" % n, "L": 50000 * n, "n_calls": rows, "calls_md5": hashlib.md5(text.encode()).hexdigest(),
                   "seconds": round(dt, 1), "by": "oracle call_contig(fast=True)"}
            if n <= 20:
                text2, _, _ = oracle_rows(n, fast=False)
                assert text2 == text, "shortest_path_fast differs from the replayed Bellman-Ford"
                rec["by"] += ", equal to the replayed edge-order Bellman-Ford"
            with gzip.GzipFile(os.path.join(HERE, "long%d.calls.tsv.gz" % n), "wb", mtime=0) as fh:
                fh.write(text.encode())
            index["long%d" % n] = rec
        print(w, index["long4" if w == "ref4" else "long" + w], flush=True)
        with open(path, "w") as fh:
            json.dump(index, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
