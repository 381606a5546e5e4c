"""One chunked run of the 10-Mb contig (profiling target for ncu: default geometry, two runs)."""
import sys, json
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
seq = np.frombuffer(synth.long_contig(200), dtype=np.uint8)
offs = np.array([0, len(seq)], dtype=np.int64)
e = Engine(0)
for _ in range(2):
    r = e.run_packed(seq, offs)
print(json.dumps({"bp": int(len(seq)), "calls": r.n_calls, "chunks": r.n_chunks, "fallbacks": r.n_chunk_fallbacks, "launches": r.launches,
                  "device_ms": round(e.last_run_ms(), 3)}))
