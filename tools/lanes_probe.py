"""resident throughput of PipelinedEngine (several contexts on one GPU) against one context"""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine, PipelinedEngine
from phanotate_b200 import synth
bases, offs = synth.synth4_batch(10000, 50000)
out = {}
e = Engine(0); e.pin(bases)
e.run_packed(bases, offs, fetch=False)
t = time.perf_counter()
for _ in range(5): e.run_packed(bases, offs, resident=True, fetch=False)
out["one_context_ms"] = round((time.perf_counter() - t) / 5 * 1e3, 2)
e.close()
for lanes in (2, 3, 4, 6, 8):
    p = PipelinedEngine(0, lanes=lanes)
    p.run_packed(bases, offs)
    p.run_packed(bases, offs, resident=True, fetch=False)
    t = time.perf_counter()
    for _ in range(5): p.run_packed(bases, offs, resident=True, fetch=False)
    out["lanes_%d_resident_ms" % lanes] = round((time.perf_counter() - t) / 5 * 1e3, 2)
    t = time.perf_counter()
    for _ in range(5): p.run_packed(bases, offs)
    out["lanes_%d_e2e_ms" % lanes] = round((time.perf_counter() - t) / 5 * 1e3, 2)
    p.close()
print(json.dumps(out))
