"""Multi-GPU parity run (hardware): torchrun --nproc-per-node N tests/run_sharded_gpu.py

Runs bench.py's sharded_parity_check — merged exhaustive / two-stage / three-stage-batch / filtered / full-ranking
lists of the collective C-ABI searches through ShardedCorpusClient == the oracle on the whole corpus — and prints the
result of every rank. Exit code 0 iff every check passed on every rank. (The driver's `pytest -m gpu` runs on one GPU;
bench.py repeats this check at every N > 1 and reports it as `sharded_parity` in its JSON line.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import bench

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = bench.sharded_parity_check(local, rank, world)
    print(f"rank {rank}: {json.dumps(res)}", flush=True)
    ok = torch.tensor([1 if res["sharded_parity"] else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(ok.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
