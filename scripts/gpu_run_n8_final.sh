# 8-GPU run: the plain-C host on the three transports, then the bench line (BASELINE configs[3] weak scaling + cfg1 strong scaling)
cd $GRAFT_REPO_ROOT
N=${1:-8}; TAG=${2:-r2y}
nvidia-smi -L | wc -l
(timeout 300 tests/c_host/vrag_host sharded $N 160000 256 > gpurun_out/${TAG}_chost_fused.log 2>&1; echo "c host (peer memory, fused) rc=$?"; tail -1 gpurun_out/${TAG}_chost_fused.log)
(VRAG_P2P_FUSED=0 timeout 300 tests/c_host/vrag_host sharded $N 160000 256 > gpurun_out/${TAG}_chost_unfused.log 2>&1; echo "c host (peer memory, exchange kernels) rc=$?"; tail -1 gpurun_out/${TAG}_chost_unfused.log)
(VRAG_P2P=0 timeout 300 tests/c_host/vrag_host sharded $N 160000 256 > gpurun_out/${TAG}_chost_nccl.log 2>&1; echo "c host (NCCL) rc=$?"; tail -1 gpurun_out/${TAG}_chost_nccl.log)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/${TAG}_bench_n$N.json; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n$N.err | tail -5)
