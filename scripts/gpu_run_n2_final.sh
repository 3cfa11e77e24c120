# 2-GPU check of the collective path: the plain-C host (peer memory and NCCL), the torchrun parity script, the bench line
cd $GRAFT_REPO_ROOT
TAG=${1:-r2w}
nvidia-smi -L | wc -l
(timeout 300 tests/c_host/vrag_host sharded 2 40000 256 > gpurun_out/${TAG}_chost_p2p.log 2>&1; echo "c host (peer memory) rc=$?"; tail -3 gpurun_out/${TAG}_chost_p2p.log)
(VRAG_P2P_FUSED=0 timeout 300 tests/c_host/vrag_host sharded 2 40000 256 > gpurun_out/${TAG}_chost_unfused.log 2>&1; echo "c host (peer memory, exchange kernels) rc=$?"; tail -1 gpurun_out/${TAG}_chost_unfused.log)
(VRAG_P2P=0 timeout 300 tests/c_host/vrag_host sharded 2 40000 256 > gpurun_out/${TAG}_chost_nccl.log 2>&1; echo "c host (NCCL) rc=$?"; tail -3 gpurun_out/${TAG}_chost_nccl.log)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_sharded_gpu.py > gpurun_out/${TAG}_sharded.log 2>&1; echo "sharded rc=$?"; tail -4 gpurun_out/${TAG}_sharded.log)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_n2.json; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n2.err | tail -5)
