# single-GPU kernel iteration: parity tests first, then the cfg2 batch timing and its launch list
cd $GRAFT_REPO_ROOT
TAG=${1:-r2c}
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log)
timeout 300 python scripts/prof_driver.py cfg2 > gpurun_out/${TAG}_cfg2.log 2>&1 && tail -2 gpurun_out/${TAG}_cfg2.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv python scripts/prof_driver.py cfg2 > gpurun_out/${TAG}_ncu_cfg2.log 2>&1
python - <<PY
import csv,sys
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches_cfg2.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:][-28:]:
    print(r[ki][:60], r[vi], r[ui])
PY
