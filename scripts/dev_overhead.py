"""Dev: wall time vs device time of single-query calls (where does the host-side overhead go?)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200")]
from visual_rag_b200.corpus import GpuCorpus
from visual_rag_b200.embedding import pooling as GP
rng = np.random.default_rng(0)
q = rng.standard_normal((20, 128)).astype(np.float32)
c = GpuCorpus(0)
def run(label, fn, n=300):
    for _ in range(20): fn()
    w, d0, d1 = [], [], []
    for _ in range(n):
        t = time.perf_counter(); fn(); w.append(1e3 * (time.perf_counter() - t))
        a, b = c.last_timing_ms(); d0.append(a); d1.append(b)
    print(f"{label:40s} wall p50 {np.percentile(w,50):.3f} ms | device whole call {np.percentile(d0,50):.3f} | scan kernel {np.percentile(d1,50):.3f}")
c.add_synthetic_store("initial", 10000, fixed_rows=768, seed=1)
c.pool_store("initial", [GP.spec_tile_mean(64)], ["mean_pooling"])
run("cfg0 exhaustive top-10 (10k x 768)", lambda: c.search("initial", q, 10))
run("cfg0 two-stage 256 -> 10", lambda: c.search_multistage([("mean_pooling", True, 256), ("initial", False, 10)], q))
run("score only, 256 candidates", lambda: c.score("initial", q, candidate_ids=np.arange(256)))
c.drop_store("initial"); c.drop_store("mean_pooling")
c.add_synthetic_store("initial", 500000, fixed_rows=1030, seed=1)
c.pool_store("initial", [GP.spec_seq_chunks(32)], ["mean_pooling"])
run("cfg1 shard two-stage (500k pages)", lambda: c.search_multistage([("mean_pooling", False, 256), ("initial", False, 10)], q))
run("stage-1 only top-256 (500k x 32)", lambda: c.search("mean_pooling", q, 256))
