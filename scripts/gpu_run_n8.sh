cd $GRAFT_REPO_ROOT
N=${1:-8}; TAG=${2:-r2g}
nvidia-smi -L | wc -l
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_n$N.json; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n$N.err | tail -5)
