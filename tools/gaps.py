"""Untimed device time between the stages of one run (pb200_stage_gaps)."""
import sys, ctypes, json
sys.path.insert(0, '.')
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
bases, offs = synth.synth4_batch(n, 50000)
e = Engine(0)
e.run_packed(bases, offs, fetch=False)
for _ in range(2):
    e.run_packed(bases, offs, fetch=False, resident=True)
names = (ctypes.c_char_p * 96)(); ms = (ctypes.c_float * 96)()
k = e.lib.pb200_stage_gaps(e.ctx, names, ms, 96)
gaps = [(names[i].decode(), round(float(ms[i]), 3)) for i in range(k)]
st = e._stage_times()
print(json.dumps({"device_ms": round(e.last_run_ms(), 3), "stages_ms": round(sum(st.values()), 3), "gaps_ms": round(sum(g for _, g in gaps), 3),
                  "largest_gaps_before": sorted(gaps, key=lambda x: -x[1])[:12]}))
