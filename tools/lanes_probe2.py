import sys, json, time, os
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine, PipelinedEngine
from phanotate_b200 import synth
bases, offs = synth.synth4_batch(10000, 50000)
out = {}
e = Engine(0); e.pin(bases); e.pin(offs)
for lanes, w in ((4, "1,2,2,2"), (4, "1,3,4,4"), (4, "1,2,4,4"), (5, "1,2,4,4,4"), (4, "2,3,3,3"), (3, "1,3,3"), (4, "1,4,4,3"), (4, "1,3,4,2")):
    os.environ["PB200_LANE_WEIGHTS"] = w
    p = PipelinedEngine(0, lanes=lanes)
    p.run_packed(bases, offs); p.run_packed(bases, offs)
    ts = []
    for _ in range(6):
        t = time.perf_counter(); p.run_packed(bases, offs); ts.append(time.perf_counter() - t)
    out["%d:%s" % (lanes, w)] = [round(1e3 * min(ts), 2), round(1e3 * sum(ts) / len(ts), 2)]
    p.close()
print(json.dumps(out))
