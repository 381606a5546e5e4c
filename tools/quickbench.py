"""Small-batch timing probe (16 / 256 / 1024 contigs) for the GPU box."""
import sys, time, json
sys.path.insert(0,'.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth
eng = Engine(0)
uniq, uoffs = synth.synth4_batch(16)
for n in (16, 256, 1024):
    buf, offs = synth.tile_batch(uniq, uoffs, n)
    for rep in range(2):
        t=time.time(); res = eng.run_packed(buf, offs); dt=time.time()-t
    print(json.dumps({"contigs": n, "bp": int(offs[-1]), "wall_s": round(dt,4), "Gbp_s": round(offs[-1]/dt/1e9,4), "calls": res.n_calls, "nodes": res.n_nodes, "orfs": res.n_orfs, "ov": res.n_overlaps, "launches": res.launches, "stage_ms": {k: round(v,3) for k,v in res.stage_ms.items()}, "errs": int((res.contigs['err']!=0).sum()), "ties": int(res.contigs['n_ties'].sum())}))
