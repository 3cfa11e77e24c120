set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -15 gpurun_out/r2u_pytest.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2u_smoke.log)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_sharded_gpu.py > gpurun_out/r2u_sharded.log 2>&1; echo "sharded rc=$?"; tail -8 gpurun_out/r2u_sharded.log)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2u_bench_n2.json; tail -5 gpurun_out/r2u_bench_n2.err)
