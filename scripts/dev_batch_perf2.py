"""cfg2 three-stage batch: per-stage device-time breakdown (prefix runs)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
rng = np.random.default_rng(0)
c = GpuCorpus(0)
NQ = 256
queries = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(NQ)]
n = 1_000_000
h = rng.integers(16, 33, size=n); w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
c.add_synthetic_store("initial", 0, page_offsets=off, seed=1)
offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
c.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=2)
c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
full = [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)]
for ns in (1, 2, 3):
    st = full[:ns]
    for _ in range(2):
        c.search_multistage_batch(st, queries)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        c.search_multistage_batch(st, queries)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"stages={ns}: device {c.last_timing_ms()[0]:.3f} ms, wall median {np.median(ts):.3f} ms", flush=True)
