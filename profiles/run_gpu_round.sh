#!/bin/bash
# One gpurun call: GPU parity tests, bench line (with the C5 / C4 / C3 legs), ncu launch list and ncu --set full captures of the hot
# kernels (exported to CSV on the box: gpurun_out/ is capped at 64 MiB, the .ncu-rep files are dropped when they would not fit),
# compute-sanitizer logs.  Usage (from the repo root on the GPU box): bash profiles/run_gpu_round.sh <tag> [skip-ncu]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "$1 rc=$2 $(( $(date +%s)-T0 ))s" | tee -a $O/${TAG}_times.log; }
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; stamp pytest $?
tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; stamp bench $?
timeout 300 python bench.py --steps 6 --warmup 3 --no-legs --no-cpu-baseline --precision 2 > $O/${TAG}_bench_split.json 2> $O/${TAG}_bench_split.err; stamp bench_split $?
if [ "$2" != "skip-ncu" ]; then
# launch list of one steady-state step: skip the burn-in (112 steps x 50 rounds x 17 launches + gathers) by counting from the end is not
# possible with ncu, so the list is taken on a short-burn-in run (8 steps): same kernels, same shapes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --burnin 8 --no-cpu-baseline --no-legs > $O/${TAG}_launches.log 2>&1; stamp launches $?
R=/tmp/ncu_$TAG; mkdir -p $R
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 3 -o $R/conv -f python bench.py --steps 1 --warmup 3 --burnin 0 --no-cpu-baseline --no-legs > $O/${TAG}_ncu_conv.log 2>&1; stamp ncu_conv $?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_warps|heads_tc" -s 100 -c 4 -o $R/tree_c2 -f python bench.py --steps 1 --warmup 3 --burnin 0 --no-cpu-baseline --no-legs > $O/${TAG}_ncu_tree.log 2>&1; stamp ncu_tree $?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_warps -s 6 -c 1 -o $R/tree_c5 -f python profiles/c5_probe.py ncu > $O/${TAG}_ncu_c5.log 2>&1; stamp ncu_c5 $?
for f in conv tree_c2 tree_c5; do
  ncu -i $R/$f.ncu-rep --page raw --csv > $O/${TAG}_${f}_raw.csv 2>/dev/null
done
ncu -i $R/tree_c5.ncu-rep --page source --csv --print-source sass > $O/${TAG}_tree_c5_source_sass.csv 2>/dev/null
gzip -f $O/${TAG}_*_source_*.csv
stamp export 0
fi
timeout 120 python profiles/train_probe.py > $O/${TAG}_train.jsonl 2>&1; stamp train $?
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_probe.py nn > $O/${TAG}_sanitizer_memcheck.log 2>&1; stamp memcheck $?
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/sanitize_probe.py nn > $O/${TAG}_sanitizer_synccheck.log 2>&1; stamp synccheck $?
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python profiles/sanitize_probe.py nn > $O/${TAG}_sanitizer_initcheck.log 2>&1; stamp initcheck $?
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_probe.py > $O/${TAG}_sanitizer_racecheck.log 2>&1; stamp racecheck $?
tail -2 $O/${TAG}_sanitizer_memcheck.log $O/${TAG}_sanitizer_racecheck.log
cat $O/${TAG}_bench.json | cut -c1-300; cat $O/${TAG}_bench_split.json | cut -c1-200
