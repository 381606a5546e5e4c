#!/bin/bash
# per-kernel metric table of one batch (ncu, few metrics): tools/ncu_metrics.sh <contigs> <out.csv>
M=gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 200 --csv --log-file "$2" python tools/prof_run.py "$1" 1
