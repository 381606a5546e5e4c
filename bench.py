#!/usr/bin/env python3
"""Benchmark of the PHANOTATE hot path (contig Gbp/s through ORF-scan + graph-solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contigs C]

A step = one pass of the whole hot path over one batch of synthetic contigs: BASELINE.json config 4, `contigs` x 50 kb
phage-like windows (SURVEY.md 8d generator), default 10,000 contigs = 0.5 Gbp per GPU.  With N > 1 (launched under
torchrun, one rank per GPU; no PyTorch is imported -- the ranks meet through the library's own NCCL communicator) every
rank runs its own batch of that size (weak scaling, contigs are independent) and the per-rank call tables are gathered to
rank 0 over NCCL inside the timed region.

Prints ONE JSON line (rank 0).
  value  the batch already resident in HBM; timed with CUDA events on the library's stream around the K steps (runs +
         gathers), max over ranks -- the same definition at every N.
  e2e    THE HEADLINE (pinned host bases -> call tables on rank 0's host): the same through the public Python
         API (PipelinedEngine) with pinned HOST buffers, host->device copies of the bases, device->host copies of the call
         and contig tables and, for N > 1, the NCCL gather plus rank 0's copy of the other ranks' rows to its host, all
         inside the timed region; wall clock, max over ranks.
  cpu_baseline  the UNMODIFIED reference (baseline/_ref: get_orfs + get_graph, and an exact-integer edge-order
         Bellman-Ford standing in for the absent fastpathz) on a bounded sample of the same workload, one process per
         contig on all host cores; the oracle port on a larger sample beside it.
  single_genome / long_contig  (N = 1) BASELINE configs 1-3 and 5: phiX174, lambda, T4 as one-contig runs (GPU latency,
         reference seconds, calls compared) and the 10-Mb contig.
`--impl reference` times that same unmodified reference (or the oracle port when baseline/_ref is absent; --ref-kind).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from phanotate_b200 import synth  # noqa: E402

METRIC = "contig Gbp/s through ORF-scan+graph-solve"
UNIT = "Gbp/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
FASTA = {"phiX174": "phiX174.fasta", "lambda": "NC_001416.1.fasta", "T4": "NC_000866.1.fasta"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arms.  (1) the oracle port; (2) the unmodified reference out of baseline/_ref
def _oracle_one(seq: bytes):
    from oracle import phanotate_oracle as O
    rows = O.call_contig(seq.decode())[3]
    return [tuple(r[:4]) for r in rows]


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "phanotate_modules", "functions.py"))


def _ref_init():
    """worker start: the reference's phanotate_modules (baseline/_ref) goes in front of this repository's API mirror of
    the same name; its per-locus warning on stderr is silenced"""
    for k in [k for k in sys.modules if k == "phanotate_modules" or k.startswith("phanotate_modules.")]:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    import warnings
    warnings.filterwarnings("ignore")
    import phanotate_modules.functions as F
    assert os.path.realpath(F.__file__).startswith(os.path.realpath(REF_DIR)), F.__file__
    sys.stderr = open(os.devnull, "w")


class _Locus:
    """the duck type functions.get_orfs needs (functions.py:143-150, orfs.py:8-15), defaults of file_handling.py:50-62"""

    def __init__(self, seq):
        from decimal import Decimal
        self._seq = seq
        w = {"atg": Decimal("0.85"), "gtg": Decimal("0.10"), "ttg": Decimal("0.05")}
        m = max(w.values())
        self.start_codons = {k: v / m for k, v in w.items()}
        self.stop_codons = ["tag", "tga", "taa"]
        self.min_orf_len = 90

    def seq(self):
        return self._seq

    def length(self):
        return len(self._seq)

    def name(self):
        return "contig"


def _ref_one(seq: bytes):
    """phanotate.py:40-76 on one contig with the reference's own get_orfs / get_graph; the shortest path by an
    exact-integer Bellman-Ford over graph.iteredges() in order, strict '<' (the contract of the absent fastpathz:
    weights x1000 truncated, phanotate.py:55-59)"""
    from decimal import Decimal, ROUND_DOWN
    from phanotate_modules import functions
    from phanotate_modules.edges import Edge
    from phanotate_modules.nodes import Node  # noqa: F401  (eval of the node reprs)
    dna = seq.decode()
    orfs = functions.get_orfs(_Locus(dna))
    graph = functions.get_graph(orfs)
    idx, names, E = {}, [], []
    for e in graph.iteredges():
        a, b, w = str(e).split("\t")
        for n in (a, b):
            if n not in idx:
                idx[n] = len(names)
                names.append(n)
        E.append((idx[a], idx[b], int(Decimal(w).to_integral_value(rounding=ROUND_DOWN))))
    s, t = "Node('source','source',0,0)", "Node('target','target',0,%d)" % (len(dna) + 1)
    if s not in idx or t not in idx:
        return []
    dist, par = [None] * len(names), [-1] * len(names)
    dist[idx[s]] = 0
    changed = True
    while changed:
        changed = False
        for u, v, w in E:
            du = dist[u]
            if du is not None and (dist[v] is None or du + w < dist[v]):
                dist[v], par[v], changed = du + w, u, True
    if dist[idx[t]] is None:
        return []
    path, v = [], idx[t]
    while v != -1:
        path.append(names[v])
        v = par[v]
    path = path[::-1][1:]
    it = iter(path)
    rows = []
    for a, b in zip(it, it):
        left, right = eval(a), eval(b)
        w = graph.weight(Edge(left, right, 0))
        rows.append((left.position, right.position + 2, "+" if left.frame > 0 else "-", "%E" % w))
    return rows


class CpuArm:
    """a persistent pool of one process per host core running one of the two CPU implementations"""

    def __init__(self, kind):
        import multiprocessing as mp
        self.kind = kind
        self.cores = os.cpu_count() or 1
        if kind == "reference":
            self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_ref_init)
            self.fn = _ref_one
        else:
            self.pool = mp.get_context("spawn").Pool(self.cores)
            self.fn = _oracle_one

    def run(self, seqs):
        t = time.perf_counter()
        rows = self.pool.map(self.fn, seqs, chunksize=1)
        return rows, time.perf_counter() - t

    def close(self):
        self.pool.terminate()
        self.pool.join()


def describe(kind, n, length, dt, ncalls, cores):
    what = ("UNMODIFIED reference (baseline/_ref: phanotate_modules.functions.get_orfs + get_graph, exact-integer edge-order "
            "Bellman-Ford for the absent fastpathz)" if kind == "reference" else "oracle port (oracle/phanotate_oracle.py)")
    return "%d contigs x %d bp of the same synthetic workload, %s, one process per contig on %d cores, %.1f s wall, %d calls" % (
        n, length, what, cores, dt, ncalls)


def cpu_sample(kind, n_contigs, length, first=0):
    arm = CpuArm(kind)
    try:
        seqs = [synth.synth4_contig(first + k, length) for k in range(n_contigs)]
        arm.run([s[:3000] for s in seqs[:arm.cores]])             # warm the workers (imports)
        rows, dt = arm.run(seqs)
    finally:
        arm.close()
    ncalls = sum(len(r) for r in rows)
    return {"value": n_contigs * length / dt / 1e9, "unit": UNIT, "cores": arm.cores, "kind": kind,
            "sample": describe(kind, n_contigs, length, dt, ncalls, arm.cores)}, rows


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  Each step = one
    contig per core of the same workload (a 50-kb contig is the smallest unit the reference can be timed on: ~10 s)."""
    if rank != 0:
        return
    kind = args.ref_kind
    if kind == "auto":
        kind = "reference" if ref_available() else "port"
    if kind == "reference" and not ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing (pip install --target baseline/_ref "
                          "--no-deps /root/reference) -- use --ref-kind port"}), flush=True)
        return
    arm = CpuArm(kind)
    n = arm.cores if kind == "reference" else max(2 * arm.cores, 8)
    length = args.length
    try:
        arm.run([synth.synth4_contig(k, args.length)[:3000] for k in range(arm.cores)])      # imports
        if kind == "reference":
            # The whole run has to end within a few minutes (PB200_REF_BUDGET_S, default 200 s) whatever K and W are, and one
            # 50-kb contig costs the reference ~10 s of one core: a calibration pass (one 8-kb contig per core) gives its
            # rate, and when K + W full contigs per core would not fit, every step takes contigs of the same generator cut
            # to the length that does (>= 10 kb).  Shorter contigs favour the reference (its connect loop is quadratic in
            # the contig length, functions.py:360-438), so the number reported is an upper bound of its speed.
            budget = float(os.environ.get("PB200_REF_BUDGET_S", "200"))
            _, dt = arm.run([synth.synth4_contig(10 ** 5 + k, 8000) for k in range(n)])
            rate = 8000.0 / dt                                       # bp per second per core
            fit = int(budget / (args.warmup + args.steps) * rate)
            if fit < args.length:
                length = max(10000, fit // 5000 * 5000)
        vals, ncalls = [], 0
        for s in range(args.warmup + args.steps):
            # every step takes the NEXT n contigs of the workload (nothing is computed twice)
            seqs = [synth.synth4_contig(s * n + k, length) for k in range(n)]
            rows, dt = arm.run(seqs)
            if s >= args.warmup:
                vals.append(dt)
                ncalls += sum(len(r) for r in rows)
    finally:
        arm.close()
    sec = sum(vals)
    value = n * length * len(vals) / sec / 1e9
    cb = {"value": value, "unit": UNIT, "cores": arm.cores, "kind": kind,
          "sample": describe(kind, n, length, sec / max(len(vals), 1), ncalls // max(len(vals), 1), arm.cores) + " per step" +
                    ("" if length == args.length else " (contigs of the same generator cut to %d bp so that %d steps fit the time "
                     "budget: favours the reference, whose connect loop is quadratic in the contig length)" % (length, args.warmup + args.steps))}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(len(vals), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "decimal28+int",
            "data": "synthetic", "config": workload_config(args, world), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "synthetic %d phage contigs x %d bp per GPU (BASELINE.json config 4: donor lambda|T4|phiX "
                        "windows, 2%% point mutations, seed 20261017)" % (args.contigs, args.length),
            "contigs_per_gpu": args.contigs, "contig_bp": args.length, "parallelism": "contig-sharded x%d" % world,
            "l2": "inputs (%.0f MB per GPU) larger than the 126 MB L2" % (args.contigs * args.length / 1e6),
            "mode": "default pb200_run: certified integer edge weights and float scores, Decimal chain only where owed; "
                    "e2e through PipelinedEngine with %d lanes" % args.lanes}


def calls_text(res, k=0):
    return "".join("%d\t%d\t%s\t%s\n" % r for r in res.call_rows(k))


def single_genomes(eng, with_reference):
    """BASELINE configs 1-3: one genome = one batch of one contig.  GPU: best of 5 warm runs through Engine.run (host bytes in,
    call table on the host out): wall and device milliseconds.  CPU: the unmodified reference on the same genome (the three
    genomes side by side, one core each), calls compared with the GPU's."""
    out, seqs = {}, {}
    for name, f in FASTA.items():
        seqs[name] = synth.read_fasta_bytes(os.path.join(ROOT, "tests", "data", f))[0][1]
    ref_rows, ref_s = {}, {}
    if with_reference and ref_available():
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(3, initializer=_ref_init) as pool:
            pool.map(_ref_one, [s[:2000] for s in seqs.values()])
            t0 = time.perf_counter()
            jobs = {n: (pool.apply_async(_timed_ref, (s,))) for n, s in seqs.items()}
            for n, j in jobs.items():
                ref_rows[n], ref_s[n] = j.get()
    for name, s in seqs.items():
        eng.run([s])
        eng.run([s])
        best = None
        for _ in range(5):
            t = time.perf_counter()
            r = eng.run([s])
            dt = time.perf_counter() - t
            if best is None or dt < best[0]:
                best = (dt, eng.last_run_ms(), r)
        dt, dev, r = best
        row = {"bp": len(s), "gpu_wall_ms": round(1e3 * dt, 3), "gpu_device_ms": round(dev, 3), "calls": r.n_calls,
               "chunks": r.n_chunks, "launches": r.launches}
        if name in ref_rows:
            row["reference_s"] = round(ref_s[name], 2)
            row["speedup_wall"] = round(ref_s[name] / dt, 1)
            row["calls_identical_to_reference"] = r.call_rows(0) == [tuple(x) for x in ref_rows[name]]
        out[name] = row
    return out


def _timed_ref(seq):
    t = time.perf_counter()
    rows = _ref_one(seq)
    return rows, time.perf_counter() - t


def long_contig(eng):
    """BASELINE config 5: ONE contig of 10 Mb (synth.long_contig): chunked intra-contig solve; best of 3 warm runs; the call
    table's md5 against the golden made by the oracle (tests/golden/long_index.json)."""
    seq = np.frombuffer(synth.long_contig(200), dtype=np.uint8)
    offs = np.array([0, len(seq)], dtype=np.int64)
    eng.pin(seq)
    try:
        eng.run_packed(seq, offs, fetch=False)
        best = None
        for _ in range(3):
            t = time.perf_counter()
            r = eng.run_packed(seq, offs)
            dt = time.perf_counter() - t
            if best is None or dt < best[0]:
                best = (dt, eng.last_run_ms(), r)
    finally:
        eng.unpin(seq)
    dt, dev, r = best
    row = {"bp": int(len(seq)), "gpu_wall_ms": round(1e3 * dt, 3), "gpu_device_ms": round(dev, 3),
           "Gbp_s_wall": round(len(seq) / dt / 1e9, 3), "calls": r.n_calls, "nodes": r.n_nodes, "chunks": r.n_chunks,
           "chunk_fallbacks": r.n_chunk_fallbacks, "err": int(r.contigs[0]["err"]), "launches": r.launches,
           "solve_ms": round(r.stage_ms.get("solve", -1.0), 3)}
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "long_index.json")))["long200"]
        row["calls_md5_equals_oracle_golden"] = hashlib.md5(calls_text(r).encode()).hexdigest() == g["calls_md5"]
    except Exception as e:
        row["calls_md5_equals_oracle_golden"] = "golden missing: %s" % e
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the unmodified reference out of baseline/_ref, or the oracle port (auto: the former if present)")
    ap.add_argument("--contigs", type=int, default=10000)
    ap.add_argument("--length", type=int, default=50000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="contigs in the cpu_baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the extra configs")
    ap.add_argument("--no-extra", action="store_true", help="skip the single-genome / long-contig / strong-scaling rows")
    ap.add_argument("--lanes", type=int, default=4, help="contexts of the pipelined engine used for the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    saved_stdout = None
    if world > 1:
        # NCCL may print on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
    from phanotate_b200.engine import Engine, PipelinedEngine, make_params
    from phanotate_b200 import _native as N
    from phanotate_b200.dist import Comm, bind_near_gpu, shard_contigs
    # before the pinned buffers and the lane threads exist (PB200_NO_BIND=1: leave the process unbound)
    numa = {"gpu": local, "bound": False, "how": "PB200_NO_BIND"} if os.environ.get("PB200_NO_BIND") else bind_near_gpu(local, world)
    eng = Engine(local)
    comm = Comm(eng, rank, world) if world > 1 else None
    params = make_params()
    # this rank's batch: contigs rank*C .. rank*C + C-1 of the generator
    bases, offs = synth.synth4_batch(args.contigs, args.length, first=rank * args.contigs)
    total_bp = int(offs[-1])
    eng.pin(bases)
    eng.pin(offs)

    def barrier():
        if comm is not None:
            comm.barrier()

    def reduce(vals, op):
        return comm.allreduce(vals, op) if comm is not None else [float(v) for v in vals]

    def timed_resident(b, o, steps, warm):
        """`steps` resident runs (+ gather) between two marks on the library's stream -> (device ms of the region, stage
        sums, launches)"""
        eng.run_packed(b, o, params, fetch=False)                 # uploads the batch once
        for _ in range(warm):
            eng.run_packed(b, o, params, resident=True, fetch=False)
            if comm is not None:
                comm.gather_calls()
        barrier()
        stage, launches = {}, 0
        eng.lib.pb200_mark(eng.ctx, 0)
        for _ in range(steps):
            eng.run_packed(b, o, params, resident=True, fetch=False)
            for k, v in eng._stage_times().items():
                stage[k] = stage.get(k, 0.0) + v
            launches += int(eng.lib.pb200_launch_count(eng.ctx))
            if comm is not None:
                comm.gather_calls()
        eng.lib.pb200_mark(eng.ctx, 1)
        ms = float(eng.lib.pb200_elapsed_ms(eng.ctx, 0, 1))
        barrier()
        return ms, stage, launches

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, stage, launches = timed_resident(bases, offs, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    sizes = eng.sizes()
    ncalls = sizes[6]

    # ---- end to end through the public API with pinned host buffers (`e2e`): PipelinedEngine cuts the batch into
    #      `lanes` groups of contigs so that the copies of one group overlap the kernels of the others
    peng = PipelinedEngine(local, lanes=args.lanes)

    # input of the end-to-end leg: the batch as 4-bit letters (two bases per byte), the format a FASTA reader hands to the
    # engine when the host link matters (pb200_pack4; packing is parsing, outside the metric like the rest of the parser:
    # SURVEY.md 8d) -- half the bytes per step over PCIe.  The same leg from 1-byte letters is reported beside it.
    packed = peng.pack4(bases)
    peng.pin(packed)
    trace = []                                                     # PB200_E2E_TRACE: host seconds of every run / gather of this rank
    packed_b = packed.copy()                                       # the "next batch" of the double-buffered leg: its own host buffer
    peng.pin(packed_b)

    def e2e_step(ascii_input=False, k=0, nxt=False):
        """k: step number (the 4-bit leg alternates between two pinned host buffers); nxt: the NEXT step's letters are
        copied in while this step computes (PipelinedEngine.run_packed(prefetch=...): double buffering across steps --
        every step's copy still happens inside the timed region, behind the previous step's kernels instead of in front
        of its own)"""
        if ascii_input:
            res = peng.run_packed(bases, offs, params, compact=True)   # H2D letters+offsets, kernels, D2H calls+contig table
        else:
            cur, other = (packed, packed_b) if k % 2 == 0 else (packed_b, packed)
            t_a = time.perf_counter()
            res = peng.run_packed(cur, offs, params, packed4=True, prefetch=(other, offs) if nxt else None, compact=True)
            trace.append(["run", time.perf_counter() - t_a])
        moved = res.calls.nbytes + res.contigs.nbytes
        t_a = time.perf_counter()
        if comm is not None:
            # every rank has its rows on its host; the cross-rank gather goes device to device over NCCL straight from the
            # lanes' tables, and rank 0 copies the other ranks' rows to its host
            # (that copy runs on its own stream beside the next step's batch; the previous step's is collected here, the
            # last one before the clock stops)
            if rank == 0:
                comm.fetch_wait()
            counts, total = comm.gather_calls(peng.engines, compact=True)
            if rank == 0 and total > counts[0]:
                comm.fetch_begin(counts[0], total - counts[0])
                moved += (total - counts[0]) * N.CALL24.itemsize
            trace.append(["gather", time.perf_counter() - t_a])
        return res, moved

    def timed_resident_lanes(steps, warm):
        """the same resident batch through the PipelinedEngine (its lanes still hold their groups of contigs: no copy in, no
        copy out): the groups' kernels overlap on the GPU -- the tail of one group's solve runs beside the next group's scan.
        Timed from a mark on lane 0's idle stream to the last mark of any lane (or of the gather's stream)."""
        for _ in range(warm):
            peng.run_packed(bases, offs, params, resident=True, fetch=False)
            if comm is not None:
                comm.gather_calls(peng.engines)
        barrier()
        n_launch = 0
        for e in peng.engines:
            e.lib.pb200_mark(e.ctx, 0)
        for _ in range(steps):
            peng.run_packed(bases, offs, params, resident=True, fetch=False)
            n_launch += sum(int(e.lib.pb200_launch_count(e.ctx)) for e in peng.engines)
            if comm is not None:
                comm.gather_calls(peng.engines)
        ends = list(peng.engines) + ([eng] if comm is not None else [])
        for e in ends:
            e.lib.pb200_mark(e.ctx, 1)
        first = peng.engines[0]
        ms = max(float(first.lib.pb200_elapsed_between_ms(first.ctx, 0, e.ctx, 1)) for e in ends)
        barrier()
        return ms, n_launch

    res, _ = e2e_step()                                            # warm-up: sizes the lanes' device buffers
    res, _ = e2e_step()
    res, _ = e2e_step()
    if comm is not None and rank == 0:
        comm.fetch_wait()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        res, d2h = e2e_step()
    if comm is not None and rank == 0:
        comm.fetch_wait()                                          # the last step's rows of the other ranks are on the host
    wall_e2e_plain = time.perf_counter() - t1
    barrier()
    # ... the same K steps double-buffered across steps: step k's run queues step k+1's letters behind its own (the first
    # timed step starts with nothing prefetched, the last one prefetches nothing: K copies, all inside the timed region)
    for k in range(2):
        e2e_step(k=k, nxt=(k == 0))
    if comm is not None and rank == 0:
        comm.fetch_wait()
    barrier()
    t1 = time.perf_counter()
    for k in range(args.steps):
        res, d2h = e2e_step(k=k, nxt=(k + 1 < args.steps))
    if comm is not None and rank == 0:
        comm.fetch_wait()
    wall_e2e = time.perf_counter() - t1
    barrier()
    if os.environ.get("PB200_E2E_TRACE"):
        tr = trace[-2 * args.steps:] if comm is not None else trace[-args.steps:]
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"rank": rank, "wall_e2e": wall_e2e, "trace_ms": [[a, round(b * 1e3, 2)] for a, b in tr]},
                  open("gpurun_out/e2e_trace_rank%d.json" % rank, "w"))
    # ... and from 1-byte letters
    e2e_step(True)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(True)
    if comm is not None and rank == 0:
        comm.fetch_wait()
    wall_e2e_ascii = time.perf_counter() - t1
    barrier()
    errs = int((res.contigs["err"] != 0).sum())
    e2e_calls = res.n_calls
    lanes_ms, lanes_launches = timed_resident_lanes(args.steps, 2)

    # ---- every rank: a few of ITS contigs against the oracle port (the checker of this run's GPU result, never its source)
    nchk = 4 if world > 1 else 0
    mism = 0
    if nchk and not args.no_cpu:
        for k in range(nchk):
            kk = (k * 2477 + 13 * rank) % args.contigs
            want = _oracle_one(bytes(bases[offs[kk]:offs[kk + 1]]))
            mism += 0 if res.call_rows(kk) == want else 1

    # ---- strong scaling (extra): the SAME 10,000 contigs dealt to the ranks by dist.shard_contigs (LPT)
    strong = None
    if world > 1 and not args.no_extra:
        lens = [args.length] * args.contigs
        mine = shard_contigs(lens, world)[rank]
        sb, so = synth.synth4_batch(len(mine), args.length, first=int(mine[0])) if len(mine) and np.all(np.diff(mine) == 1) \
            else _pack([synth.synth4_contig(int(i), args.length) for i in mine])
        eng.pin(sb)
        sms, _, _ = timed_resident(sb, so, args.steps, 2)
        eng.unpin(sb)
        strong = {"contigs_total": args.contigs, "contigs_this_rank": int(len(mine)), "ms": sms}

    vals = reduce([dev_ms / 1e3, wall_e2e, strong["ms"] / 1e3 if strong else 0.0, lanes_ms / 1e3, wall_e2e_ascii, wall_e2e_plain], "max")
    dev_s, wall_e2e, strong_s, lanes_s, wall_e2e_ascii, wall_e2e_plain = vals
    sums = reduce([total_bp, ncalls, errs, mism, e2e_calls], "sum")
    job_bp, job_calls, errs, mism, job_e2e_calls = (int(round(x)) for x in sums)

    peng.unpin(packed)
    peng.unpin(packed_b)
    peng.close()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        value_one = job_bp * args.steps / dev_s / 1e9              # one context: one batch = one chain of kernels
        value_lanes = job_bp * args.steps / lanes_s / 1e9          # PipelinedEngine: the batch as `lanes` overlapping groups
        use_lanes = lanes_s > 0 and value_lanes > value_one
        value = value_lanes if use_lanes else value_one
        if use_lanes:
            dev_s, launches = lanes_s, lanes_launches
        e2e = job_bp * args.steps / wall_e2e / 1e9
        # roofline of the dominant kernel (largest share of the step), algorithmic bytes = 1 B/bp + 24 B/CDS
        peak, peak_src = measured_peak_gbs()
        dom = max(stage, key=stage.get) if stage else None
        roof = None
        if dom:
            per_launch_s = stage[dom] / args.steps / 1e3
            alg = total_bp * 1.0 + 24.0 * ncalls
            ach = alg / per_launch_s / 1e9
            kname = {"solve": "k_solve", "scan_tiles": "k_scan_tiles", "hold": "k_hold"}.get(dom, "k_" + dom)
            # DRAM bytes of that kernel per launch from the committed `ncu --set full` capture of this workload
            traffic, tsrc = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
                if tj.get("contigs") == args.contigs and tj.get("contig_bp") == args.length:
                    traffic = tj["kernels"].get(kname)
                    tsrc = "profiles/dram_traffic.json (ncu --set full at commit %s)" % tj.get("commit", "?")
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach,
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": tsrc,
                    "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_s * 1e3,
                    "share_of_step": stage[dom] / max(sum(stage.values()), 1e-9),
                    "note": "issue / dependent-latency bound, not HBM bound (DESIGN.md 4): frac is reported for the contract; "
                            "traffic = dram read+write bytes per launch from profiles/ (ncu --set full)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "decimal28+f64x2+int128", "data": "synthetic",
                "config": dict(workload_config(args, world), numa=numa), "clocks": clocks,
                "timing": "value: CUDA events around the K steps (runs + NCCL gathers), max over ranks, batch resident in HBM -- "
                          "through %s; e2e: wall clock around K PipelinedEngine runs from pinned host buffers, max over ranks" % (
                              "PipelinedEngine (%d overlapping groups of contigs, marks on the lanes' streams)" % args.lanes if use_lanes
                              else "one context (marks on its stream)"),
                "value_one_context": {"value": value_one, "unit": UNIT, "note": "one context = one chain of kernels per batch; the "
                                      "stage table and the roofline below are from this run"},
                "value_pipelined": {"value": value_lanes, "unit": UNIT, "lanes": args.lanes},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(packed.nbytes + offs.nbytes),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * wall_e2e / args.steps,
                        "headline": "pinned host bases -> call tables on rank 0's host (SURVEY.md 8d)",
                        "input": "4-bit letters, two bases per byte (pb200_pack4 / PB200_INPUT_PACKED4), expanded on the device",
                        "output": "call rows as pb200_call24 (contig, left, right, strand, f64 score: the columns Locus.tabular prints; SURVEY.md 8d's 24 B per CDS) + the contig table",
                        "double_buffering": "step k's run queues step k+1's letters (another pinned host buffer) behind its own "
                                            "(pb200_prefetch_async); K host->device copies and K device->host reads inside the timed region"},
                "e2e_no_prefetch": {"value": job_bp * args.steps / wall_e2e_plain / 1e9, "unit": UNIT,
                                    "ms_per_step": 1e3 * wall_e2e_plain / args.steps,
                                    "input": "4-bit letters; every step copies its own letters in first (round-2 mid-round definition)"},
                "e2e_ascii_input": {"value": job_bp * args.steps / wall_e2e_ascii / 1e9, "unit": UNIT,
                                    "h2d_bytes_per_step": int(bases.nbytes + offs.nbytes), "ms_per_step": 1e3 * wall_e2e_ascii / args.steps,
                                    "input": "one byte per base (any case, IUPAC)"},
                "gpu_launches": launches, "roofline": roof,
                "stage_ms_per_step": {k: round(v / args.steps, 3) for k, v in sorted(stage.items(), key=lambda x: -x[1])},
                "calls_per_step": job_calls, "contig_errors": errs,
                "tables": {"nodes": sizes[2], "orfs": sizes[3], "overlap_edges": sizes[4], "bridge_edges": sizes[5]}}
        if world > 1:
            line["comm"] = {"backend": "NCCL %s inside libpb200.so (csrc/comm.inc), no PyTorch" % eng.lib.pb200_comm_nccl_version(),
                            "collective": "all-gather of row counts + grouped send/recv of exactly the rows, to rank 0"}
            line["oracle_check"] = {"contigs_per_rank": nchk, "ranks": world, "mismatches": mism}
            if strong:
                line["strong_scaling"] = {"workload": "the SAME %d contigs x %d bp dealt to %d ranks (LPT, dist.shard_contigs)" % (
                    args.contigs, args.length, world), "value": args.contigs * args.length * args.steps / strong_s / 1e9, "unit": UNIT,
                    "ms_per_step": 1e3 * strong_s / args.steps}
        if world == 1 and not args.no_cpu:
            kind = "reference" if ref_available() else "port"
            n = args.cpu_sample or (os.cpu_count() or 1) * (1 if kind == "reference" else 2)
            cb, rows = cpu_sample(kind, n, args.length)
            same = sum(1 for k, want in enumerate(rows) if res.call_rows(k) == [tuple(w) for w in want])
            cb["gpu_calls_identical"] = "%d of %d sampled contigs" % (same, len(rows))
            line["cpu_baseline"] = cb
            if kind == "reference":
                pb, prow = cpu_sample("port", 4 * (os.cpu_count() or 1), args.length)
                same = sum(1 for k, want in enumerate(prow) if res.call_rows(k) == [tuple(w) for w in want])
                pb["gpu_calls_identical"] = "%d of %d sampled contigs" % (same, len(prow))
                line["cpu_baseline_port"] = pb
        if world == 1 and not args.no_extra:
            line["single_genome"] = single_genomes(eng, not args.no_cpu)
            line["long_contig"] = long_contig(eng)
        print(json.dumps(line), flush=True)
    eng.unpin(bases)
    eng.unpin(offs)
    if comm is not None:
        comm.close()
    eng.close()


def _pack(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(), offs


if __name__ == "__main__":
    main()
