"""Maximum-size check (run on the GPU box): one batch whose concatenated length exceeds 2^31 bases.

    python tools/big_batch.py [contigs=50000]

The batch repeats 16 unique 50-kb contigs of the bench workload cyclically, so every replica must produce exactly the
call rows of the first copy of its contig (a size-independent property: no oracle needed at this size), and global base
positions, node / ORF / edge offsets all pass 2^31 inside the run.
"""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
uniq, uoffs = synth.synth4_batch(16)
buf, offs = synth.tile_batch(uniq, uoffs, n)
e = Engine(0)
e.run_packed(buf, offs, fetch=False)                 # first run uploads the batch and sizes the device buffers
t = time.perf_counter()
res = e.run_packed(buf, offs, resident=True)         # timed: batch resident in HBM, tables fetched afterwards
wall = time.perf_counter() - t
ms = e.last_run_ms()
ref = [res.call_rows(k) for k in range(16)]
cs = res.contigs
bad = 0
for k in range(n):
    c = cs[k]
    a = res.calls[c["call_off"]:c["call_off"] + c["n_calls"]]
    b = res.calls[cs[k % 16]["call_off"]:cs[k % 16]["call_off"] + cs[k % 16]["n_calls"]]
    if len(a) != len(b) or not (np.array_equal(a["left"], b["left"]) and np.array_equal(a["right"], b["right"]) and
                                np.array_equal(a["strand"], b["strand"]) and np.array_equal(a["score"], b["score"])) \
            or not (a["contig"] == k).all() or int(c["err"]) != 0:
        bad += 1
print(json.dumps({"contigs": n, "bp": int(offs[-1]), "beyond_2^31": bool(offs[-1] > 2**31), "nodes": res.n_nodes, "orfs": res.n_orfs,
                  "overlap_edges": res.n_overlaps, "calls": res.n_calls, "replicas_differing_from_first_copy": bad,
                  "device_ms": round(ms, 2), "stage_ms": {k: round(v, 2) for k, v in sorted(res.stage_ms.items(), key=lambda kv: -kv[1])[:12]}, "Gbp_s_device": round(offs[-1] / ms / 1e6, 2), "wall_incl_table_fetch_s": round(wall, 2)}))
