"""File to file: FASTA on disk -> vectorised ingest -> engine -> tabular text, timed per stage."""
import sys, time, json, io, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from phanotate_b200 import synth, fastio
from phanotate_b200.engine import PipelinedEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
path = os.path.join(tempfile.mkdtemp(), "batch.fasta")
with open(path, "w") as fh:
    for k in range(n):
        s = synth.synth4_contig(k).decode()
        fh.write(">contig%d synthetic\n" % k)
        fh.write("\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n")
pe = PipelinedEngine(0, lanes=4)
t0 = time.perf_counter(); names, bases, offs = fastio.read_fasta_packed(path, pe.engines[0].lib); t1 = time.perf_counter()
pe.run_packed(bases, offs)
t2 = time.perf_counter(); res = pe.run_packed(bases, offs); t3 = time.perf_counter()
text = fastio.tabular_text(res, names, pe.engines[0].lib); t4 = time.perf_counter()
out = io.StringIO(text.decode())
bp = int(offs[-1])
print(json.dumps({"contigs": n, "bp": bp, "file_MB": round(os.path.getsize(path) / 1e6, 1), "calls": res.n_calls,
                  "ingest_s": round(t1 - t0, 3), "engine_s (unpinned host buffer)": round(t3 - t2, 4), "tabular_s": round(t4 - t3, 3),
                  "text_MB": round(len(out.getvalue()) / 1e6, 1),
                  "Gbp_s": {"ingest": round(bp / (t1 - t0) / 1e9, 3), "engine": round(bp / (t3 - t2) / 1e9, 3),
                            "tabular": round(bp / (t4 - t3) / 1e9, 3), "file_to_text": round(bp / (t4 - t0 - (t2 - t1)) / 1e9, 3)}}))
