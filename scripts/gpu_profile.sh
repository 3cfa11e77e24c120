#!/bin/bash
# Run on the GPU box (gpurun): bench without ncu, then the ncu launch list of the same command and one
# `--set full` capture per hot kernel. Outputs land in gpurun_out/ (copy the summaries into profiles/).
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --pages-per-gpu 100000 --steps 3 --cpu-sample-pages 200 --latency-queries 20"
$BENCH > $O/bench_${TAG}_100k.json 2> $O/bench_${TAG}_100k.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/ncu_launch_${TAG}.log 2>&1
for W in large packed global pool_tokens pool_rows; do
  python scripts/prof_driver.py $W > $O/drv_$W.log 2>&1 || { echo "driver $W failed"; continue; }
  case $W in
    large|packed|global) K=maxsim_scan ;;
    pool_tokens) K=pool_tokens ;;
    pool_rows) K=pool_rows ;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $O/prof_${W}_${TAG} python scripts/prof_driver.py $W > $O/ncu_$W.log 2>&1
  ncu -i $O/prof_${W}_${TAG}.ncu-rep --page raw --csv > $O/prof_${W}_${TAG}_raw.csv 2>/dev/null
done
ls -la $O
