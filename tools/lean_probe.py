"""Experiment: the whole bench batch through the chunked machinery with ONE chunk per contig (lean distance-only sweep +
parents/ties pulled by the check kernels) beside the default path: stage times."""
import sys, json
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
bases, offs = synth.synth4_batch(n, 50000)
e = Engine(0)
out = {}
ref = None
for name, geo in (("default", None), ("one_chunk_per_contig", (4096, 0, 0, 0))):
    if geo:
        e.set_chunking(*geo)
    e.run_packed(bases, offs, fetch=False)
    r = e.run_packed(bases, offs, resident=True)
    if ref is None:
        ref = r.calls.copy()
    st = r.stage_ms
    out[name] = {"device_ms": round(e.last_run_ms(), 2), "chunks": r.n_chunks, "fallbacks": r.n_chunk_fallbacks, "same_calls": bool(np.array_equal(ref, r.calls)),
                 "stages": {k: round(v, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.05}}
print(json.dumps(out))
