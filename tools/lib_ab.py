"""A/B of library builds on the bench batch: python tools/lib_ab.py build/libA.so build/libB.so ... (first = default lib)"""
import sys, json
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
bases, offs = synth.synth4_batch(10000, 50000)
ref = None
for path in [None] + sys.argv[1:]:
    e = Engine(0, lib_path=path)
    e.run_packed(bases, offs, fetch=False)
    ms, st = [], {}
    for _ in range(5):
        r = e.run_packed(bases, offs, resident=True)
        ms.append(e.last_run_ms())
        for k, v in r.stage_ms.items():
            st[k] = st.get(k, 0) + v / 5
    if ref is None:
        ref = r.calls.copy()
    print(json.dumps({"lib": path or "default", "device_ms": round(min(ms), 3), "solve": round(st["solve"], 3), "same_calls": bool(np.array_equal(ref, r.calls))}))
    e.close()
