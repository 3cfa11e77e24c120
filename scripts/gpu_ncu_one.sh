# one `ncu --set full` capture of a kernel of scripts/prof_driver.py: gpu_ncu_one.sh <tag> <driver what> <kernel regex> <skip>
cd $GRAFT_REPO_ROOT
TAG=$1; WHAT=$2; K=$3; S=$4
timeout 300 python scripts/prof_driver.py $WHAT > gpurun_out/${TAG}_drv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/${TAG} python scripts/prof_driver.py $WHAT > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}*; tail -3 gpurun_out/${TAG}_ncu.log
