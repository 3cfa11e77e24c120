#!/bin/bash
# Run on the GPU box (gpurun): bench without ncu, then the ncu launch list of the same command and one
# `--set full` capture per hot kernel. Outputs land in gpurun_out/ (copy the summaries into profiles/).
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --pages-per-gpu 100000 --steps 3 --cpu-sample-pages 200 --latency-queries 20"
$BENCH > $O/bench_${TAG}_100k.json 2> $O/bench_${TAG}_100k.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/ncu_launch_${TAG}.log 2>&1
for W in large packed global large_batch packed_batch global_batch pool_tokens pool_fused; do
  python scripts/prof_driver.py $W > $O/drv_$W.log 2>&1 || { echo "driver $W failed"; continue; }
  case $W in
    large|packed|global|large_batch) K=maxsim_scan; S=2 ;;
    packed_batch|global_batch) K=maxsim_scan; S=5 ;;   # calls alternate sample pass / filtered pass: launch 6 = a full filtered pass
    pool_tokens|pool_fused) K=pool_tokens; S=2 ;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/prof_${W}_${TAG} python scripts/prof_driver.py $W > $O/ncu_$W.log 2>&1
  ncu -i $O/prof_${W}_${TAG}.ncu-rep --page raw --csv > $O/prof_${W}_${TAG}_raw.csv 2>/dev/null
  rm -f $O/prof_${W}_${TAG}.ncu-rep    # gpurun_out is capped at 64 MiB: keep the CSV exports only
done
ls -la $O
