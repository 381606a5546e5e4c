"""Where the end-to-end time goes: pipelined engine with 1..8 lanes on the bench workload."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine, PipelinedEngine, make_params
from phanotate_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
bases, offs = synth.synth4_batch(n, 50000)
params = make_params()
for lanes in (1, 2, 4, 8):
    pe = PipelinedEngine(0, lanes=lanes)
    pe.pin(bases); pe.pin(offs)
    pe.run_packed(bases, offs, params)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); r = pe.run_packed(bases, offs, params); ts.append(time.perf_counter() - t)
    tr = []
    for _ in range(4):
        t = time.perf_counter(); pe.run_packed(bases, offs, params, resident=True, fetch=False); tr.append(time.perf_counter() - t)
    print(json.dumps({"lanes": lanes, "ms": [round(1e3 * x, 2) for x in ts], "Gbp_s": round(offs[-1] / min(ts) / 1e9, 3), "calls": r.n_calls,
                      "resident_ms": [round(1e3 * x, 2) for x in tr]}))
    pe.unpin(bases); pe.unpin(offs)
    pe.close()
e = Engine(0)
e.pin(bases); e.pin(offs)
e.run_packed(bases, offs, params, fetch=False)
for _ in range(3):
    t = time.perf_counter(); e.run_packed(bases, offs, params, fetch=False); t1 = time.perf_counter() - t
    t = time.perf_counter(); e.run_packed(bases, offs, params, fetch=False, resident=True); t2 = time.perf_counter() - t
    print("single ctx: host->dev + kernels %.2f ms, resident %.2f ms, device %.2f ms" % (1e3 * t1, 1e3 * t2, e.last_run_ms()))
