"""gpurun_out/<tag>_* -> profiles/<tag>_*: bench line, launch list + per-kernel summary, full-capture metric summary,
DRAM traffic of the two largest kernels (read by bench.py's roofline)."""
import csv, json, sys, collections, subprocess, os, shutil
tag = sys.argv[1]
G, P = "gpurun_out/", "profiles/"
shutil.copy(G + tag + "_bench.json", P + tag + "_bench_default_10000contigs.json")
shutil.copy(G + tag + "_launches.csv", P + tag + "_launches.csv")
rows = [r for r in csv.reader(open(G + tag + "_launches.csv")) if len(r) > 10]
h = rows[0]; iK, iV = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[iK].split('(')[0].replace("void ", ""), [0, 0.0]); a[0] += 1; a[1] += float(r[iV].replace(',', ''))
tot = sum(v[1] for v in agg.values())
with open(P + tag + "_launch_summary.csv", "w") as fh:
    fh.write("kernel,launches,total_ms,avg_ms,share_of_gpu_time\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write("%s,%d,%.4f,%.4f,%.4f\n" % (k, v[0], v[1] / 1e6, v[1] / 1e6 / v[0], v[1] / tot))
raw = subprocess.run(["ncu", "-i", G + tag + "_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh, uu = rr[0], rr[1]
keep = [k for k in hh if any(s in k for s in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe", "launch__occupancy_limit"))
        and not k.endswith(".per_second")]
traffic = {}
with open(P + tag + "_ncu_full_top_kernels.csv", "w") as fh:
    w = csv.writer(fh)
    w.writerow(["kernel"] + keep)
    w.writerow(["unit"] + [uu[hh.index(k)] for k in keep])
    for r in rr[2:]:
        d = dict(zip(hh, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]      # k_solve<32> -> k_solve
        w.writerow([name] + [d[k] for k in keep])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        rd = float(d["dram__bytes_read.sum"]) * scale[uu[hh.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * scale[uu[hh.index("dram__bytes_write.sum")]]
        traffic[name] = rd + wr
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"contigs": 10000, "contig_bp": 50000, "commit": commit, "source": tag + "_ncu_full_top_kernels.csv (ncu --set full, one launch each, bench workload)",
           "kernels": traffic}, open(P + "dram_traffic.json", "w"), indent=1)
if os.path.exists(G + tag + "_metrics_table.txt"):
    shutil.copy(G + tag + "_metrics_table.txt", P + tag + "_kernel_metrics_table.txt")
print(json.dumps(traffic))
b = json.load(open(G + tag + "_bench.json"))
print({k: b[k] for k in ("value", "ms_per_step", "gpu_launches")}, b["e2e"], b["roofline"]["kernel"], b["roofline"]["frac"])
