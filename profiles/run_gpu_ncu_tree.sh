#!/bin/bash
# ncu --set full of the fused C5 rounds kernel.  bash profiles/run_gpu_ncu_tree.sh <tag>
TAG=${1:-r02n}
O=gpurun_out; mkdir -p $O
R=/tmp/ncu_$TAG; mkdir -p $R
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_warps -s 6 -c 1 -o $R/tree_c5 -f python profiles/c5_probe.py ncu > $O/${TAG}_ncu_c5.log 2>&1
ncu -i $R/tree_c5.ncu-rep --page raw --csv > $O/${TAG}_tree_c5_raw.csv 2>/dev/null
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source sass > $O/${TAG}_tree_c5_source_sass.csv 2>/dev/null
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source cuda > $O/${TAG}_tree_c5_source_cuda.csv 2>/dev/null
gzip -f $O/${TAG}_*_source_*.csv
tail -2 $O/${TAG}_ncu_c5.log
