#!/bin/bash
# Round profile set (run on the GPU box): bench line, ncu launch list of the same command, full captures of the two
# largest kernels at the bench size.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r1_t}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_scan_tiles$|^k_solve$" -c 2 -o gpurun_out/${TAG}_full \
    python tools/prof_run.py 10000 1 > gpurun_out/${TAG}_full.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench.json
