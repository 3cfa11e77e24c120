# final single-GPU pass of the round: parity tests, smoke, both bench arms, the ncu launch list of the bench command
cd $GRAFT_REPO_ROOT
TAG=${1:-r2f}
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log)
(timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err)
(timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"; head -c 1200 gpurun_out/${TAG}_bench_reference.json; echo; tail -3 gpurun_out/${TAG}_bench_reference.err)
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_bench_default.csv python bench.py > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/${TAG}_launches_bench_default.csv)
