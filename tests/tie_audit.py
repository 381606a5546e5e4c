"""Audit of exact ties in the solve (run by hand; not collected by pytest):

    python tests/tie_audit.py [first] [count]

Contigs first..first+count-1 of the bench workload (BASELINE.json config 4) go through the HOST build of the stage
functions (tests/native/pb200_hostsim.so: the same sweep order and strict '<' as the CUDA solve) and through the oracle
(edge-order Bellman-Ford, oracle/phanotate_oracle.py).  Reports how many contigs saw an equal-distance relaxation
(`n_ties`) and whether any call table differs.  Test infrastructure: imports the oracle as the checker.
"""
import json
import os
import sys
import time
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)


def oracle_rows(seq):
    from oracle import phanotate_oracle as O
    return [tuple(r[:4]) for r in O.call_contig(seq.decode())[3]]


def main():
    from helpers import hostsim_path
    from phanotate_b200 import synth
    from phanotate_b200.engine import Engine
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    seqs = [synth.synth4_contig(first + k) for k in range(count)]
    t = time.time()
    sim = Engine(0, lib_path=hostsim_path())
    res = sim.run(seqs)
    t_sim = time.time() - t
    t = time.time()
    with Pool(os.cpu_count()) as pool:
        want = pool.map(oracle_rows, seqs, chunksize=1)
    t_or = time.time() - t
    tied = [k for k in range(count) if int(res.contigs[k]["n_ties"]) > 0]
    bad = [first + k for k in range(count) if res.call_rows(k) != want[k]]
    print(json.dumps({"contigs": [first, first + count - 1], "contigs_with_ties": len(tied),
                      "ties_total": int(sum(int(res.contigs[k]["n_ties"]) for k in tied)),
                      "contigs_with_ties_identical_to_oracle": sum(1 for k in tied if res.call_rows(k) == want[k]),
                      "mismatching_contigs": bad, "calls": int(res.n_calls), "host_build_s": round(t_sim, 1), "oracle_s": round(t_or, 1)}))


if __name__ == "__main__":
    main()
