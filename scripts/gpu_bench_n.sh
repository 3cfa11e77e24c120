# bench line at N GPUs (the command the driver runs for the scaling table)
cd $GRAFT_REPO_ROOT
N=${1:-8}; TAG=${2:-r2u}
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench rc=$?"; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${TAG}_bench_n${N}.err | tail -3)
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
print("value", round(d["value"]/1e6,2), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]/1e6,2), "parity", d.get("sharded_parity"))
print("two_stage_strong", json.dumps(d["two_stage_strong"]))
print("cfg2", json.dumps(d.get("three_stage_batched_sharded")))
print("cfg4", json.dumps(d.get("pooling_cfg4")))
PY
