#!/bin/bash
# One gpurun call: GPU parity tests, bench line, other BASELINE configs, ncu launch list and ncu --set full captures of the hot
# kernels (exported to CSV on the box: gpurun_out/ is capped at 64 MiB, the .ncu-rep files are dropped when they would not fit).
# Usage (from the repo root on the GPU box): bash profiles/run_gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "$1 rc=$2 $(( $(date +%s)-T0 ))s" | tee -a $O/${TAG}_times.log; }
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; stamp pytest $?
timeout 400 python bench.py --steps 6 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; stamp bench $?
timeout 600 python profiles/measure_configs.py C5 C4 C3 > $O/${TAG}_configs.jsonl 2>&1; stamp configs $?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 340 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1; stamp launches $?
R=/tmp/ncu_$TAG; mkdir -p $R
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 3 -o $R/conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_conv.log 2>&1; stamp ncu_conv $?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_warps|heads_tc" -s 100 -c 4 -o $R/tree_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_tree.log 2>&1; stamp ncu_tree $?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_warps -s 420 -c 2 -o $R/tree_c5 -f python profiles/c5_probe.py > $O/${TAG}_ncu_c5.log 2>&1; stamp ncu_c5 $?
for f in conv tree_c2 tree_c5; do
  ncu -i $R/$f.ncu-rep --page raw --csv > $O/${TAG}_${f}_raw.csv 2>/dev/null
done
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source sass > $O/${TAG}_tree_c5_source_sass.csv 2>/dev/null
gzip -f $O/${TAG}_*_source_*.csv
stamp export 0
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json | cut -c1-400; cat $O/${TAG}_configs.jsonl | cut -c1-700
