"""BASELINE.json config 5: one 10-Mb contig (200 config-4 windows concatenated) on one GPU."""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth, _native as N
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seq = b"".join(synth.synth4_contig(10**6 + k) for k in range(nwin))
e = Engine(0)
out = {"bp": len(seq)}
for name, fl in (("run", 0),):
    e.run([seq], flags=fl)
    t = time.perf_counter(); r = e.run([seq], flags=fl); dt = time.perf_counter() - t
    out[name] = {"wall_ms": round(1e3 * dt, 2), "device_ms": round(e.last_run_ms(), 2), "solve_ms": round(r.stage_ms.get("solve", -1), 2),
                 "calls": r.n_calls, "err": int(r.contigs[0]["err"]), "nodes": r.n_nodes, "orfs": r.n_orfs, "overlaps": r.n_overlaps}
    out[name]["stage_ms"] = {k: round(v, 2) for k, v in sorted(r.stage_ms.items(), key=lambda kv: -kv[1])[:6]}
print(json.dumps(out))
