#!/bin/bash
# Round profile set (run on the GPU box): bench line, ncu launch list of the same command, full captures of the two
# largest kernels at the bench size, a per-kernel metric table of one batch, the long-contig run under ncu.
# Outputs under gpurun_out/<tag>_*; tools/summarise_profiles.py <tag> (here) turns them into profiles/<tag>_*.
TAG=${1:-r2_z}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_scan_tiles$|^k_solve$" -c 2 -o gpurun_out/${TAG}_full \
    python tools/prof_run.py 10000 1 > gpurun_out/${TAG}_full.log 2>&1
M=gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
M=$M,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
M=$M,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_metrics.csv \
    python tools/prof_run.py 10000 1 > gpurun_out/${TAG}_metrics.log 2>&1
python tools/ncu_table.py gpurun_out/${TAG}_metrics.csv > gpurun_out/${TAG}_metrics_table.txt
ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_long_metrics.csv \
    python tools/prof_long.py > gpurun_out/${TAG}_long_under_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/${TAG}_long_metrics.csv > gpurun_out/${TAG}_long_metrics_table.txt
python tools/long_contig.py 200 > gpurun_out/${TAG}_long_contig_10Mb.json 2>/dev/null
python tools/cli_throughput.py > gpurun_out/${TAG}_cli_file_to_text.json 2>/dev/null
tail -c 600 gpurun_out/${TAG}_bench.json
