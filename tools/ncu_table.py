"""Pivot an ncu --csv (long format) metric log into one row per kernel launch."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iID, iM, iV = hdr.index('Kernel Name'), hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iID], r[iK]), {})[r[iM]] = r[iV].replace(',', '')
short = [('gpu__time_duration.sum', 'us', 1e-3), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1),
         ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'iss%', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 1),
         ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1), ('launch__registers_per_thread', 'regs', 1),
         ('smsp__inst_executed.sum', 'Minst', 1e-6), ('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'lclLdM', 1e-6),
         ('l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'lclStM', 1e-6),
         ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'longsb', 1),
         ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst', 1), ('dram__bytes_read.sum', 'rdMB', 1e-6), ('dram__bytes_write.sum', 'wrMB', 1e-6)]
print('%-28s' % 'kernel' + ''.join('%9s' % s for _, s, _ in short))
tot = 0
for (i, k), m in d.items():
    name = k.split('(')[0][:28]
    vals = []
    for key, s, sc in short:
        try:
            vals.append('%9.1f' % (float(m.get(key, 'nan')) * sc))
        except ValueError:
            vals.append('%9s' % '-')
    tot += float(m.get('gpu__time_duration.sum', 0)) * 1e-3
    print('%-28s' % name + ''.join(vals))
print('total us', tot)
