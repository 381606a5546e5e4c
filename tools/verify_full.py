"""Full-size check on the GPU: the default (certified) run against PB200_LITERAL on the bench workload
(N distinct synthetic 50-kb contigs): integer weights of every ORF and overlap edge, every call row, contig tables."""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200.engine import Engine, make_params
from phanotate_b200 import synth, _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
bases, offs = synth.synth4_batch(n, 50000)
e = Engine(0)
t = time.time(); fast = e.run_packed(bases, offs); tf = time.time() - t
raw_f = np.zeros((fast.n_orfs, 8), dtype=np.uint32)
e._ck(e.lib.pb200_get_orf_int_weights(e.ctx, raw_f.ctypes.data))
ov_f = fast.overlap_int_weights()
gs_f, gd_f = fast.gap_int_weights()
calls_f, contigs_f = fast.calls.copy(), fast.contigs.copy()
stats_f = (fast.n_literal_presolve, fast.n_literal_postsolve, fast.n_literal_overlaps)
t = time.time(); lit = e.run_packed(bases, offs, literal=True); tl = time.time() - t
raw_l = np.zeros((lit.n_orfs, 8), dtype=np.uint32)
e._ck(e.lib.pb200_get_orf_int_weights(e.ctx, raw_l.ctypes.data))
ov_l = lit.overlap_int_weights()
gs_l, gd_l = lit.gap_int_weights()
out = {"contigs": n, "bp": int(offs[-1]), "orfs": fast.n_orfs, "overlap_edges": fast.n_overlaps, "calls": fast.n_calls,
       "orf_int_weight_mismatches": int((raw_f != raw_l).any(axis=1).sum()),
       "overlap_int_weight_mismatches": int((ov_f != ov_l).sum()),
       "gap_int_weight_mismatches": int((gs_f != gs_l).sum() + (gd_f != gd_l).sum()),
       "call_rows_equal(contig,left,right,strand,score)": bool(all(np.array_equal(calls_f[c], lit.calls[c]) for c in ("contig", "left", "right", "strand", "score"))),
       "contig_tables_equal": bool(np.array_equal(contigs_f, lit.contigs)),
       "literal_orfs_presolve/postsolve/overlaps (certified run)": stats_f,
       "contig_errors": int((contigs_f["err"] != 0).sum()), "ties": int(contigs_f["n_ties"].sum()),
       "wall_s certified/literal": [round(tf, 3), round(tl, 3)]}
print(json.dumps(out))
