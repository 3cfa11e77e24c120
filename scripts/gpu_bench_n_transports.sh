# bench line at N GPUs on both transports (peer memory, NCCL)
cd $GRAFT_REPO_ROOT
N=${1:-2}; TAG=${2:-r2t}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}_p2p.json 2> gpurun_out/${TAG}_bench_n${N}_p2p.err; echo "bench (peer memory) rc=$?"; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n${N}_p2p.err | tail -3)
(VRAG_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n${N}_nccl.json 2> gpurun_out/${TAG}_bench_n${N}_nccl.err; echo "bench (NCCL) rc=$?"; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n${N}_nccl.err | tail -3)
python - <<PY
import json
for t in ("p2p","nccl"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_n${N}_%s.json"%t).read().strip().splitlines()[-1])
        print(t, "value", round(d["value"]/1e6,2), "two_stage_strong p50", round(d["two_stage_strong"]["p50_ms"],3), d["two_stage_strong"].get("collective_us"))
        print(t, json.dumps(d.get("three_stage_batched_sharded")))
    except Exception as e:
        print(t, "failed", e)
PY
