#!/bin/bash
# One gpurun call: bench line, knob sweep, ncu launch list and ncu --set full captures of the hot kernels (exported to CSV
# on the box: gpurun_out/ is capped at 64 MiB, the .ncu-rep files are dropped when they would not fit).
# Usage (from the repo root on the GPU box): bash profiles/run_gpu_round.sh <tag>
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "$1 rc=$2 $(( $(date +%s)-T0 ))s" | tee -a $O/${TAG}_times.log; }
timeout 400 python bench.py --steps 6 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; stamp bench $?
for K in "AGZ_PIPELINE=0" "AGZ_PIPELINE=1" "AGZ_PIPELINE=1 AGZ_CONV_PAIRS=74" "AGZ_PIPELINE=1 AGZ_CONV_PAIRS=70" "AGZ_PIPELINE=1 AGZ_CONV_PAIRS=68" "AGZ_PIPELINE=0 AGZ_CONV_PAIRS=74"; do
  env $K timeout 200 python profiles/quick_c2.py 4 >> $O/${TAG}_sweep.jsonl 2>> $O/${TAG}_sweep.err
done; stamp sweep $?
timeout 300 python profiles/measure_configs.py C5 > $O/${TAG}_c5.jsonl 2>&1; stamp c5 $?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 340 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1; stamp launches $?
R=/tmp/ncu_$TAG; mkdir -p $R
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 3 -o $R/conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_conv.log 2>&1; stamp ncu_conv $?
AGZ_PIPELINE=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_warps|heads_tc" -s 200 -c 4 -o $R/tree_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_tree.log 2>&1; stamp ncu_tree $?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_warps -s 420 -c 2 -o $R/tree_c5 -f python profiles/c5_probe.py > $O/${TAG}_ncu_c5.log 2>&1; stamp ncu_c5 $?
for f in conv tree_c2 tree_c5; do
  ncu -i $R/$f.ncu-rep --page raw --csv > $O/${TAG}_${f}_raw.csv 2>/dev/null
done
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source sass > $O/${TAG}_tree_c5_source_sass.csv 2>/dev/null
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source cuda > $O/${TAG}_tree_c5_source_cuda.csv 2>/dev/null
ncu -i $R/conv.ncu-rep --page source --csv --print-source cuda > $O/${TAG}_conv_source_cuda.csv 2>/dev/null
gzip -f $O/${TAG}_*_source_*.csv
ls -la $R $O | tee -a $O/${TAG}_times.log
SZ=$(du -sm $O | cut -f1); RS=$(du -sm $R | cut -f1)
if [ $((SZ+RS)) -lt 55 ]; then cp $R/*.ncu-rep $O/; fi
stamp export 0
cat $O/${TAG}_bench.json | cut -c1-300; cat $O/${TAG}_sweep.jsonl
