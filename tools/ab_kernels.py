"""A/B of two builds of the library on the same box (developer tool): run with VRAG_LIB=<path> to pick the build.
Times (CUDA events, 20 launches after warm-up) the latency-path kernels: pooled stage-1 scans, a small exhaustive scan,
the full-token scan, a rerank gather."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus, query_flags

c = GpuCorpus(0)
rng = np.random.default_rng(0)
qd = torch.from_numpy(rng.standard_normal((20, 128)).astype(np.float32)).cuda()
st = torch.cuda.current_stream().cuda_stream
out = {"lib": os.environ.get("VRAG_LIB", "in-tree")}

def timed(name, store, n_items, cand=None, pool=False, reps=20):
    sc = torch.empty((n_items,), dtype=torch.float32, device="cuda")
    fl = query_flags(True, pool)
    f = lambda: c.score_dev(store, qd.data_ptr(), 20, fl, cand.data_ptr() if cand is not None else 0, n_items, sc.data_ptr(), st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / reps * 1e3, 1)   # us

c.add_synthetic_store("p32", 1_000_000, fixed_rows=32, seed=1)
timed("pooled32_1M_us", "p32", 1_000_000)
timed("pooled32_1M_pooledq_us", "p32", 1_000_000, pool=True)
c.drop_store("p32")
c.add_synthetic_store("p13", 1_000_000, fixed_rows=13, seed=2)
timed("colsmol13_1M_us", "p13", 1_000_000)
c.drop_store("p13")
c.add_synthetic_store("g1", 4_000_000, fixed_rows=1, seed=3)
timed("global1_4M_pooledq_us", "g1", 4_000_000, pool=True)
c.drop_store("g1")
c.add_synthetic_store("s768", 10_000, fixed_rows=768, seed=4)
timed("cfg0_10k_x768_us", "s768", 10_000)
c.drop_store("s768")
c.add_synthetic_store("big", 200_000, fixed_rows=1030, seed=5)
timed("large_200k_us", "big", 200_000)
cand = torch.from_numpy(rng.permutation(200_000)[:256].astype(np.int64)).cuda()
timed("rerank_256_us", "big", 256, cand=cand)
print(json.dumps(out))
c.close()
