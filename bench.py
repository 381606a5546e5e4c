#!/usr/bin/env python3
"""Benchmark of the PHANOTATE hot path (contig Gbp/s through ORF-scan + graph-solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contigs C]

A step = one pass of the whole hot path over one batch of synthetic contigs: BASELINE.json
config 4, `contigs` x 50 kb phage-like windows (SURVEY.md 8d generator), default 10,000 contigs
= 0.5 Gbp per GPU.  With N > 1 (launched under torchrun, one rank per GPU) every rank runs its own
batch of that size (weak scaling, contigs are independent) and the per-rank call tables are
gathered to rank 0 over NCCL inside the timed region.

Prints ONE JSON line (rank 0).  `value` is measured with the batch already resident in HBM;
`e2e` is the same through the public Python API with pinned HOST buffers, the host->device copy of
the bases and the device->host copy of the call table inside the timed region.
`--impl reference` times the CPU oracle port (oracle/phanotate_oracle.py, one process per contig
on all host cores) on a bounded sample of the same workload.
"""
import argparse
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
# CPU arm: the oracle port on all host cores
def _oracle_one(seq: bytes):
    from oracle import phanotate_oracle as O
    rows = O.call_contig(seq.decode())[3]
    return [tuple(r[:4]) for r in rows]


def cpu_sample(n_contigs: int, length: int, first: int = 0):
    """Time the oracle on n_contigs contigs of the workload with one process per contig on all cores."""
    from multiprocessing import Pool
    cores = os.cpu_count() or 1
    seqs = [synth.synth4_contig(first + k, length) for k in range(n_contigs)]
    with Pool(cores) as pool:
        pool.map(_oracle_one, seqs[:min(cores, len(seqs))])      # warm the workers (imports)
        t = time.perf_counter()
        rows = pool.map(_oracle_one, seqs, chunksize=1)
        dt = time.perf_counter() - t
    ncalls = sum(len(r) for r in rows)
    cpu_sample.rows = rows                       # kept for the parity check of the GPU result against the oracle
    bp = n_contigs * length
    return {"value": bp / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d contigs x %d bp of the same synthetic workload (oracle/phanotate_oracle.py, "
                      "multiprocessing, %.1f s wall, %d calls)" % (n_contigs, length, dt, ncalls)}, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = max(2 * (os.cpu_count() or 1), 8)
    vals, last = [], None
    for s in range(args.warmup + args.steps):
        cb, dt = cpu_sample(n, args.length, first=0)
        if s >= args.warmup:
            vals.append((n * args.length, dt))
        last = cb
    bp = sum(v[0] for v in vals)
    sec = sum(v[1] for v in vals)
    value = bp / sec / 1e9
    last["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(len(vals), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "decimal28+f64x2+int128",
            "data": "synthetic", "config": workload_config(args, world), "cpu_baseline": last,
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contigs", type=int, default=10000)
    ap.add_argument("--length", type=int, default=50000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="contigs in the cpu_baseline sample (0 = 2 x cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--lanes", type=int, default=4, help="contexts of the pipelined engine used for the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = torch = None
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from phanotate_b200.engine import Engine, PipelinedEngine, make_params
    from phanotate_b200 import _native as N
    from phanotate_b200.dist import bind_near_gpu
    numa = bind_near_gpu(local)                                   # before the pinned buffers and the lane threads exist
    eng = Engine(local)
    params = make_params()
    # this rank's batch: contigs rank*C .. rank*C + C-1 of the generator
    bases, offs = synth.synth4_batch(args.contigs, args.length, first=rank * args.contigs)
    total_bp = int(offs[-1])
    eng.pin(bases)
    eng.pin(offs)

    def barrier():
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    def gather_calls():
        """cross-contig gather of the call tables to rank 0 over NCCL (the only collective of the path)."""
        if dist is None:
            return
        from phanotate_b200.dist import DeviceCalls, gather_call_tables
        n = eng.sizes()[6]
        ptr = eng.lib.pb200_device_calls(eng.ctx)
        mine = torch.as_tensor(DeviceCalls(ptr, n), device="cuda") if n else torch.zeros(1, dtype=torch.uint8, device="cuda")
        gather_call_tables(mine, n, dist, rank, world)
        if rank == 0:
            torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    eng.run_packed(bases, offs, params, fetch=False)              # uploads the batch once
    for _ in range(args.warmup):
        eng.run_packed(bases, offs, params, resident=True, fetch=False)
        gather_calls()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, stage, launches = 0.0, {}, 0
    for _ in range(args.steps):
        eng.run_packed(bases, offs, params, resident=True, fetch=False)
        dev_ms += eng.last_run_ms()
        for k, v in eng._stage_times().items():
            stage[k] = stage.get(k, 0.0) + v
        launches += int(eng.lib.pb200_launch_count(eng.ctx))
        gather_calls()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    sizes = eng.sizes()
    ncalls = sizes[6]

    # ---- end to end through the public API with pinned host buffers (`e2e`): PipelinedEngine cuts the batch into
    #      `lanes` groups of contigs so that the copies of one group overlap the kernels of the others
    peng = PipelinedEngine(local, lanes=args.lanes)

    def gather_host_calls(res):
        """N > 1: every rank has its rows on its host (the run's device->host copy); the cross-rank gather to rank 0 goes
        device to device over NCCL, straight from the lanes' device tables (no second upload)."""
        if dist is None:
            return
        from phanotate_b200.dist import DeviceCalls, gather_call_tables
        parts, total = [], 0
        for e in peng.engines:
            n = e.sizes()[6]
            if n:
                parts.append(torch.as_tensor(DeviceCalls(e.lib.pb200_device_calls(e.ctx), n), device="cuda"))
                total += n
        assert total == res.n_calls
        mine = torch.cat(parts) if parts else torch.zeros(1, dtype=torch.uint8, device="cuda")
        gather_call_tables(mine, total, dist, rank, world)
        if rank == 0:
            torch.cuda.synchronize()

    res = peng.run_packed(bases, offs, params)                     # warm-up: sizes the lanes' device buffers
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        res = peng.run_packed(bases, offs, params)                 # H2D bases+offsets, kernels, D2H calls+contig table
        d2h = res.calls.nbytes + res.contigs.nbytes
        gather_host_calls(res)
    barrier()
    wall_e2e = time.perf_counter() - t1
    peng.close()
    errs = int((res.contigs["err"] != 0).sum())

    if dist is not None:
        tmax = torch.tensor([wall, wall_e2e, dev_ms / 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        wall, wall_e2e, dev_s = (float(x) for x in tmax.tolist())
        tot = torch.tensor([total_bp, ncalls, errs], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        job_bp, job_calls, errs = (int(x) for x in tot.tolist())
    else:
        dev_s = dev_ms / 1e3
        job_bp, job_calls = total_bp, ncalls

    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        # N=1: device time from CUDA events on the library's stream; N>1: max-over-ranks wall incl. the NCCL gather
        t_value = dev_s if world == 1 else wall
        value = job_bp * args.steps / t_value / 1e9
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
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
                if tj.get("contigs") == args.contigs and tj.get("contig_bp") == args.length:
                    traffic = tj["kernels"].get(kname)
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach,
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_s * 1e3,
                    "share_of_step": stage[dom] / max(sum(stage.values()), 1e-9),
                    "note": "issue / dependent-latency bound, not HBM bound (DESIGN.md 4): frac is reported for the contract; "
                            "traffic = dram read+write bytes per launch from profiles/ (ncu --set full)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_value / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "decimal28+f64x2+int128", "data": "synthetic",
                "config": dict(workload_config(args, world), numa=numa), "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(bases.nbytes + offs.nbytes),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * wall_e2e / args.steps},
                "gpu_launches": launches, "roofline": roof,
                "stage_ms_per_step": {k: round(v / args.steps, 3) for k, v in sorted(stage.items(), key=lambda x: -x[1])},
                "wall_ms_per_step": 1e3 * wall / args.steps, "calls_per_step": job_calls, "contig_errors": errs,
                "tables": {"nodes": sizes[2], "orfs": sizes[3], "overlap_edges": sizes[4], "bridge_edges": sizes[5]}}
        if world == 1 and not args.no_cpu:
            n = args.cpu_sample or 2 * (os.cpu_count() or 1)
            line["cpu_baseline"] = cpu_sample(n, args.length)[0]
            # the oracle's calls for the sampled contigs are the checker of this run's GPU result (never its source)
            same = sum(1 for k, want in enumerate(cpu_sample.rows) if res.call_rows(k) == [tuple(w) for w in want])
            line["cpu_baseline"]["gpu_calls_identical_to_oracle"] = "%d of %d sampled contigs" % (same, len(cpu_sample.rows))
        print(json.dumps(line), flush=True)
    eng.unpin(bases)
    eng.unpin(offs)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
