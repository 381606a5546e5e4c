"""Per-stage device ms of the bench workload (resident), median of a few runs; optional flags."""
import sys, json
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bases, offs = synth.synth4_batch(n, 50000)
e = Engine(0)
e.run_packed(bases, offs, fetch=False, flags=flags)
runs, tot = [], []
for _ in range(5):
    e.run_packed(bases, offs, fetch=False, resident=True, flags=flags)
    runs.append(e._stage_times()); tot.append(e.last_run_ms())
med = {k: float(np.median([r[k] for r in runs])) for k in runs[0]}
top = {k: round(v, 3) for k, v in sorted(med.items(), key=lambda kv: -kv[1])[:10]}
print(json.dumps({"flags": flags, "device_ms": round(float(np.median(tot)), 3), "Gbp_s": round(offs[-1] / np.median(tot) / 1e6, 3), "top": top, "calls": e.sizes()[6]}))
