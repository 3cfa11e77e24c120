"""torchrun --nproc-per-node N scripts/dev_sharded_batch_check.py : sharded batched three-stage search over N GPUs must
equal the single-GPU batched search over the whole corpus (rank 0 also holds the whole corpus). Prints timings."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus, pack_queries
from visual_rag_b200.distributed import ShardedSearcher, shard_page_range

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
rng = np.random.default_rng(3)
h = rng.integers(16, 33, size=n); w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
b, e = shard_page_range(n, rank, world)

def build(c, lo, hi):
    c.add_synthetic_store("initial", 0, page_offsets=off[lo:hi + 1] - off[lo], seed=1, row_seed_base=int(off[lo]))
    c.add_synthetic_store("experimental_pooling", 0, page_offsets=offp[lo:hi + 1] - offp[lo], seed=2, row_seed_base=int(offp[lo]))
    c.add_synthetic_store("global_pooling", hi - lo, fixed_rows=1, seed=3, row_seed_base=lo)

shard = GpuCorpus(lr, page_base=b)
build(shard, b, e)
qs = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(256)]
packed = pack_queries(qs)
stages = [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)]
s = ShardedSearcher(shard)
for _ in range(3):
    got = s.search_multistage_batch(stages, packed)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    got = s.search_multistage_batch(stages, packed)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
if rank == 0:
    full = GpuCorpus(lr, page_base=0)
    build(full, 0, n)
    want = full.search_multistage_batch(stages, packed, as_arrays=True)
    ok = all(np.array_equal(wi, gi) and np.allclose(ws, gs, rtol=1e-6) for (ws, wi, _), (gs, gi) in zip(want, got))
    for si, ((ws, wi, _), (gs, gi)) in enumerate(zip(want, got)):
        bad = np.argwhere(wi != gi)
        print(f"stage {si}: {len(bad)} id mismatches of {wi.size}; queries affected {len(set(bad[:,0].tolist()))}", flush=True)
        if len(bad):
            q, j = bad[0]
            print("   first:", q, j, wi[q, max(0,j-2):j+3], gi[q, max(0,j-2):j+3], ws[q, max(0,j-2):j+3], gs[q, max(0,j-2):j+3], flush=True)
            print("   set-equal rows:", sum(set(wi[r].tolist()) == set(gi[r].tolist()) for r in range(wi.shape[0])), "of", wi.shape[0])
    print(f"world={world} pages={n}: sharded batched three-stage, 256 queries: {dt*1e3:.2f} ms wall ({256/dt:.0f} QPS); "
          f"equals single-GPU result: {ok}; prefilter runs/fallbacks on rank0 shard: {shard._lib.vrag_launch_count(shard._h)} launches", flush=True)
    assert ok
dist.destroy_process_group()
