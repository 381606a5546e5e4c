"""Ordered stage list of one run: name, kernel ms, idle ms in front of it (pb200_stage_times / pb200_stage_gaps).
    python tools/stage_dump.py long 200 | phiX174 | lambda | T4 | batch 2000"""
import sys, json, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
what = sys.argv[1] if len(sys.argv) > 1 else "long"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
if what == "long":
    seqs = [synth.long_contig(n)]
elif what == "batch":
    seqs = [synth.synth4_contig(k) for k in range(n)]
else:
    from helpers import seq_of
    seqs = [seq_of(what).encode()]
e = Engine(0)
for _ in range(3):
    r = e.run(seqs)
names = (ctypes.c_char_p * 256)(); ms = (ctypes.c_float * 256)(); gaps = (ctypes.c_float * 256)()
k = e.lib.pb200_stage_times(e.ctx, names, ms, 256)
e.lib.pb200_stage_gaps(e.ctx, names, gaps, 256)
rows = [(names[i].decode(), round(float(ms[i]), 4), round(float(gaps[i]), 4)) for i in range(k)]
print(json.dumps({"what": what, "n": n, "device_ms": round(e.last_run_ms(), 3), "sum_stage_ms": round(sum(r[1] for r in rows), 3),
                  "sum_gap_ms": round(sum(r[2] for r in rows), 3), "launches": r.launches, "stages": rows}))
