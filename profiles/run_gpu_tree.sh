#!/bin/bash
# Tree-kernel iteration on one B200: parity subset, C5 leg, ncu of the fused rounds kernel.  bash profiles/run_gpu_tree.sh <tag>
TAG=${1:-r02t}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "$1 rc=$2 $(( $(date +%s)-T0 ))s" | tee -a $O/${TAG}_times.log; }
timeout 600 python -m pytest tests/test_abi_mcts.py tests/test_abi_go.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; stamp pytest $?
tail -3 $O/${TAG}_pytest.log
timeout 200 python profiles/c5_probe.py > $O/${TAG}_c5.json 2>$O/${TAG}_c5.err; stamp c5 $?
cat $O/${TAG}_c5.json
R=/tmp/ncu_$TAG; mkdir -p $R
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_warps -s 6 -c 1 -o $R/tree_c5 -f python profiles/c5_probe.py ncu > $O/${TAG}_ncu_c5.log 2>&1; stamp ncu_c5 $?
ncu -i $R/tree_c5.ncu-rep --page raw --csv > $O/${TAG}_tree_c5_raw.csv 2>/dev/null
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source sass > $O/${TAG}_tree_c5_source_sass.csv 2>/dev/null
gzip -f $O/${TAG}_*_source_*.csv
tail -2 $O/${TAG}_ncu_c5.log
stamp export 0
