#!/bin/bash
# Run on the GPU box (gpurun): bench without ncu, then the ncu launch list of the same command and one
# `--set full` capture per hot kernel. Outputs land in gpurun_out/ (copy the summaries into profiles/).
TAG=${1:-r1e}
O=gpurun_out
mkdir -p $O
# 1. the default bench command (what the driver runs), then its launch list under ncu (metrics-only pass)
BENCH="python bench.py"
$BENCH > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/ncu_launch_${TAG}.log 2>&1
# 2. one --set full capture per hot kernel: <name> <driver args> <kernel regex> <launches to skip>
while read -r W ARGS K S; do
  python scripts/prof_driver.py $ARGS > $O/drv_$W.log 2>&1 || { echo "driver $W failed"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/prof_${W}_${TAG} python scripts/prof_driver.py $ARGS > $O/ncu_$W.log 2>&1
  ncu -i $O/prof_${W}_${TAG}.ncu-rep --page raw --csv > $O/prof_${W}_${TAG}_raw.csv 2>/dev/null
  rm -f $O/prof_${W}_${TAG}.ncu-rep    # gpurun_out is capped at 64 MiB: keep the CSV exports only
done <<'LIST'
large_500k large:500000 maxsim_scan 2
large large maxsim_scan 2
packed packed maxsim_scan 2
global global maxsim_scan 2
large_batch large_batch maxsim_scan 2
large_batch8 large_batch8 maxsim_scan 2
packed_batch packed_batch maxsim_scan 5
global_batch global_batch maxsim_scan 5
cfg2_gather cfg2 maxsim_scan 10
cfg2_rerank cfg2 maxsim_scan 11
pool_tokens pool_tokens pool_tokens 2
pool_fused pool_fused pool_tokens 2
LIST
ls -la $O | head -50
