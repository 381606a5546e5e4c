"""BASELINE.json config 5: one 10-Mb contig (200 config-4 windows concatenated) on one GPU: chunked solve beside the
one-warp sweep, call tables compared."""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine
from phanotate_b200 import synth, _native as N
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 200
geos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(256, 768, 64, 4096)]
seq = synth.long_contig(nwin)
e = Engine(0)
out = {"bp": len(seq)}
ref = None
for name, fl, geo in [("one_sweep", N.SOLVE_NOCHUNK, None)] + [("chunked_%d_%d_%d_%d" % g, 0, g) for g in geos]:
    if geo:
        e.set_chunking(*geo)
    e.run([seq], flags=fl)
    best = None
    for _ in range(3):
        t = time.perf_counter(); r = e.run([seq], flags=fl); dt = time.perf_counter() - t
        if best is None or e.last_run_ms() < best[0]:
            best = (e.last_run_ms(), dt, r)
    ms, dt, r = best
    if ref is None:
        ref = r.calls.copy()
    st = dict(r.stage_ms)
    out[name] = {"wall_ms": round(1e3 * dt, 2), "device_ms": round(ms, 2), "Gbp_s": round(len(seq) / ms / 1e6, 3),
                 "solve_ms": round(st.get("solve", -1), 2),
                 "calls": r.n_calls, "err": int(r.contigs[0]["err"]), "nodes": r.n_nodes, "orfs": r.n_orfs, "overlaps": r.n_overlaps,
                 "chunks": r.n_chunks, "fallbacks": r.n_chunk_fallbacks, "launches": r.launches,
                 "calls_equal_one_sweep": bool(np.array_equal(ref, r.calls))}
    out[name]["stage_ms"] = {k: round(v, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]}
print(json.dumps(out))
