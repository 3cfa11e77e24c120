"""Developer perf probe (gpurun): batched-query paths — cfg2 three-stage, batched two-stage, batched exhaustive."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus

rng = np.random.default_rng(0)
c = GpuCorpus(0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
NQ = 256
queries = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(NQ)]

def run(label, stages, qs, reps=3, seq_n=16):
    for _ in range(2):
        c.search_multistage_batch(stages, qs)
    t0 = time.perf_counter()
    for _ in range(reps):
        c.search_multistage_batch(stages, qs)
    wall = (time.perf_counter() - t0) / reps
    dev = c.last_timing_ms()
    for q in qs[:3]:
        c.search_multistage(stages, q)
    t0 = time.perf_counter()
    for q in qs[:seq_n]:
        c.search_multistage(stages, q)
    seq = (time.perf_counter() - t0) / seq_n
    print(f"{label}: batch of {len(qs)}: wall {wall*1e3:.2f} ms, device {dev[0]:.2f} ms (first batched kernel {dev[1]:.3f} ms) -> "
          f"{len(qs)/wall:.0f} QPS, {wall*1e3/len(qs):.3f} ms/query; sequential {seq*1e3:.3f} ms/query -> speed-up {seq/(wall/len(qs)):.1f}x", flush=True)

if what in ("all", "cfg2"):
    n = 1_000_000
    h = rng.integers(16, 33, size=n); w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
    T = h * w
    off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    c.add_synthetic_store("initial", 0, page_offsets=off, seed=1)
    offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
    c.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=2)
    c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
    print(f"cfg2 corpus: {n} pages, {off[-1]} tokens ({off[-1]*256/1e9:.1f} GB), pooled rows {offp[-1]}", flush=True)
    run("cfg2 three-stage 1000/300/100", [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)], queries)
    run("cfg2 two-stage tokens_vs_experimental 256/10", [("experimental_pooling", False, 256), ("initial", False, 10)], queries, reps=2)
    for nm in ("initial", "experimental_pooling", "global_pooling"):
        c.drop_store(nm)
if what in ("all", "cfg1"):
    n = 500_000
    c.add_synthetic_store("initial", n, fixed_rows=1030, seed=1)
    c.add_synthetic_store("mean_pooling", n, fixed_rows=32, seed=2)
    q20 = [rng.standard_normal((20, 128)).astype(np.float32) for _ in range(64)]
    run("cfg1 two-stage tokens_vs_standard_pooling 256/10 (500k pages)", [("mean_pooling", False, 256), ("initial", False, 10)], q20)
    run("cfg1 two-stage pooled_query_vs_standard_pooling 256/10", [("mean_pooling", True, 256), ("initial", False, 10)], q20)
    run("exhaustive batched (500k x 1030)", [("initial", False, 10)], q20[:16], reps=2, seq_n=4)
    gb = n * 1030 * 260 / 1e9
    print(f"   exhaustive: {gb:.1f} GB per corpus pass")
