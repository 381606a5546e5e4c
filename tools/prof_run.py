"""One batch of N synthetic contigs through the engine (profiling target for ncu)."""
import sys, json
sys.path.insert(0, '.')
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
uniq, uoffs = synth.synth4_batch(16)
buf, offs = synth.tile_batch(uniq, uoffs, n)
for _ in range(reps):
    res = eng.run_packed(buf, offs, literal=(len(sys.argv) > 3))
print(json.dumps({"contigs": n, "bp": int(offs[-1]), "calls": res.n_calls, "orfs": res.n_orfs, "launches": res.launches, "lit": [res.n_literal_presolve, res.n_literal_postsolve, res.n_literal_overlaps],
                  "stage_ms": {k: round(v, 3) for k, v in res.stage_ms.items()}}))
