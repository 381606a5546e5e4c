"""Audit (run by hand; not collected by pytest): random ragged contigs -- 60 .. 9000 bp, GC 0.25 .. 0.75, sprinkled IUPAC
codes, mixed case -- through the HOST build of the stage functions and through the oracle; reports differing call tables.

    python tests/ragged_audit.py [seed] [count]
"""
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)


def oracle_rows(seq):
    from oracle import phanotate_oracle as O
    try:
        return [tuple(r[:4]) for r in O.call_contig(seq.decode().lower())[3]]
    except Exception as e:
        return "EXC " + type(e).__name__


def main():
    from helpers import hostsim_path
    from phanotate_b200.engine import Engine
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rng = np.random.default_rng(seed)
    seqs = []
    for k in range(count):
        n = int(rng.integers(60, 9000))
        gc = rng.uniform(0.25, 0.75)
        p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
        s = rng.choice(np.frombuffer(b"acgt", dtype=np.uint8), size=n, p=p)
        m = rng.random(n) < 0.002
        s[m] = rng.choice(np.frombuffer(b"nrykmswbdhv", dtype=np.uint8), size=int(m.sum()))
        if k % 3 == 0:
            s = np.frombuffer(s.tobytes().upper(), dtype=np.uint8)
        seqs.append(s.tobytes())
    sim = Engine(0, lib_path=hostsim_path())
    res = sim.run(seqs)
    with Pool(os.cpu_count()) as pool:
        want = pool.map(oracle_rows, seqs, chunksize=4)
    bad, exc = [], 0
    for k in range(count):
        if isinstance(want[k], str):
            exc += 1
            continue
        if res.call_rows(k) != want[k] or int(res.contigs[k]["err"]) & ~16:
            bad.append(k)
    print(json.dumps({"seed": seed, "contigs": count, "bp": sum(len(s) for s in seqs), "oracle_exceptions": exc,
                      "contigs_with_ties": int((res.contigs["n_ties"] > 0).sum()), "calls": int(res.n_calls),
                      "mismatching_contigs": bad}))


if __name__ == "__main__":
    main()
