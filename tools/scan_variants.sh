#!/bin/bash
# compile and time k_scan_tiles with different occupancy targets (run on the GPU box)
for n in 2 3 4 5; do
  sed -i "s/__global__ void __launch_bounds__(ST_NT[^)]*) k_scan_tiles/__global__ void __launch_bounds__(ST_NT, $n) k_scan_tiles/" phanotate_b200/csrc/scan_tile.cuh
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o phanotate_b200/libpb200.so phanotate_b200/csrc/pb200.cu
  echo "blocks/SM target $n:"; python tools/prof_run.py 10000 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['stage_ms']['scan_tiles'])"
done
