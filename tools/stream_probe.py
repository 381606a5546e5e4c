"""e2e throughput with TWO batches in flight (two PipelinedEngines alternating) against one"""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from phanotate_b200.engine import Engine, PipelinedEngine
from phanotate_b200 import synth
bases, offs = synth.synth4_batch(10000, 50000)
K = 10
out = {}
A = PipelinedEngine(0, lanes=4); A.pin(bases); A.pin(offs)
A.run_packed(bases, offs); A.run_packed(bases, offs)
t = time.perf_counter()
for _ in range(K): r = A.run_packed(bases, offs)
out["one_in_flight_ms"] = round((time.perf_counter() - t) / K * 1e3, 2)
ref = r.calls.copy()
for lanes in (4, 3, 2):
    P = [PipelinedEngine(0, lanes=lanes) for _ in range(2)]
    for p in P: p.run_packed(bases, offs); p.run_packed(bases, offs)
    pool = ThreadPoolExecutor(2)
    t = time.perf_counter()
    futs = []
    same = True
    for i in range(K):
        futs.append(pool.submit(P[i % 2].run_packed, bases, offs))
        if i >= 1:
            r = futs[i - 1].result(); same = same and np.array_equal(r.calls, ref)
    r = futs[-1].result(); same = same and np.array_equal(r.calls, ref)
    out["two_in_flight_%d_lanes_ms" % lanes] = [round((time.perf_counter() - t) / K * 1e3, 2), same]
    for p in P: p.close()
    pool.shutdown()
print(json.dumps(out))
