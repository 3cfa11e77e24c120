"""Dev: where does the wall time of ThreeStageRetriever.search_server_side_batch go (cfg2 shape, 256 queries)?"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200")]
from visual_rag_b200.client import GpuCorpusClient
from visual_rag_b200.corpus import GpuCorpus
from visual_rag_b200.retrieval import ThreeStageRetriever

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(3)
h = rng.integers(16, 33, size=n)
w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
c = GpuCorpus(0)
c.add_synthetic_store("initial", 0, page_offsets=off, seed=1)
c.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=2)
c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
qs = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(256)]
r = ThreeStageRetriever(GpuCorpusClient(c, "b"), "b")
for _ in range(3):
    t0 = time.perf_counter()
    out = r.search_server_side_batch(query_embeddings=qs, top_k=100, stage1_k=1000, stage2_k=300)
    print("wall ms", 1e3 * (time.perf_counter() - t0), "device ms", c.last_timing_ms())
cProfile.run("r.search_server_side_batch(query_embeddings=qs, top_k=100, stage1_k=1000, stage2_k=300)", "/tmp/p.out")
pstats.Stats("/tmp/p.out").sort_stats("tottime").print_stats(12)
