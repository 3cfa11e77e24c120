"""Profiling driver (run under ncu on the GPU box): launches ONE kernel family repeatedly so that
`ncu -k regex:<name> -s <skip> -c 1` captures a steady-state launch.
  python scripts/prof_driver.py large|packed|global|rerank|pool_tokens|pool_rows [n_pages]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
from visual_rag_b200.embedding import pooling as GP

what = sys.argv[1]
if ":" in what:      # "large:500000" == "large 500000" (one token, for scripts/gpu_profile.sh)
    what, _n = what.split(":", 1)
    sys.argv[1:2] = [what, _n]
rng = np.random.default_rng(0)
q20 = rng.standard_normal((20, 128)).astype(np.float32)
c = GpuCorpus(0)
reps = 4
if what == "large":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    c.add_synthetic_store("initial", n, fixed_rows=1030, seed=1)
    for _ in range(reps):
        c.search("initial", q20, 10)
    print("large", n, c.last_timing_ms())
elif what == "packed":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    c.add_synthetic_store("mean_pooling", n, fixed_rows=32, seed=2)
    for _ in range(reps):
        c.search("mean_pooling", q20, 256)
    print("packed", n, c.last_timing_ms())
elif what == "global":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
    for _ in range(reps):
        c.search("global_pooling", q20, 1000, pool_query=True)
    print("global", n, c.last_timing_ms())
elif what == "global_batch":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
    qs = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(128)]
    for _ in range(reps):
        c.search_multistage_batch([("global_pooling", True, 1000)], qs)
    print("global_batch", n, c.last_timing_ms())
elif what == "large_batch":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    c.add_synthetic_store("initial", n, fixed_rows=1030, seed=1)
    qs = [rng.standard_normal((20, 128)).astype(np.float32) for _ in range(4)]
    for _ in range(reps):
        c.search_multistage_batch([("initial", False, 10)], qs)
    print("large_batch", n, c.last_timing_ms())
elif what == "large_batch8":
    # batched exhaustive scan, approximate first pass: 8 plain-fp16 queries share every document tile (kernel <128,32,0,0,2>)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    c.add_synthetic_store("initial", n, fixed_rows=1030, seed=1)
    qs = [rng.standard_normal((20, 128)).astype(np.float32) for _ in range(8)]
    for _ in range(reps):
        c.search_multistage_batch([("initial", False, 10)], qs)
    print("large_batch8", n, c.last_timing_ms())
elif what == "packed_batch":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
    c.add_synthetic_store("mean_pooling", n, fixed_rows=32, seed=2)
    qs = [rng.standard_normal((20, 128)).astype(np.float32) for _ in range(4)]
    for _ in range(reps):
        c.search_multistage_batch([("mean_pooling", False, 256)], qs)
    print("packed_batch", n, c.last_timing_ms())
elif what == "colsmol13":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    c.add_synthetic_store("mean_pooling", n, fixed_rows=13, seed=2)
    for _ in range(reps):
        c.search("mean_pooling", q20, 256)
    print("colsmol13", n, c.last_timing_ms())
elif what == "cfg2":
    # BASELINE configs[2]: the batched three-stage search (launch list: which kernels make up the 3.2 ms)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    h = rng.integers(16, 33, size=n)
    w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
    off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
    offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
    c.add_synthetic_store("initial", 0, page_offsets=off, seed=1)
    c.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=2)
    c.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=3)
    qs = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(256)]
    stages = [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)]
    for _ in range(3):
        c.search_multistage_batch(stages, qs, final_only=True)
    print("cfg2", n, c.last_timing_ms())
elif what == "rerank":
    c.add_synthetic_store("initial", 100_000, fixed_rows=1030, seed=1)
    cand = rng.permutation(100_000)[:256]
    for _ in range(reps):
        c.search("initial", q20, 10, candidate_ids=cand)
    print("rerank", c.last_timing_ms())
elif what == "pool_tokens":
    c.add_synthetic_store("vis", 200_000, fixed_rows=1024, seed=6)
    for _ in range(reps):
        ms = c.pool_store("vis", [GP.spec_adaptive_rows(32, 32, 32)], ["mean_pooling"])
    print("pool_tokens", ms)
elif what == "pool_fused":
    c.add_synthetic_store("vis", 200_000, fixed_rows=1024, seed=6)
    specs = [GP.spec_adaptive_rows(32, 32, 32)] + [GP.derived_from(x, 0) for x in (GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"),
             GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True))]
    for _ in range(reps):
        ms = c.pool_store("vis", specs, ["mean_pooling", "e1", "e2", "e3", "g"])
    print("pool_fused", ms)
elif what == "pool_qwen":
    # cfg4 ColQwen2.5: bulk re-pooling of variable-grid pages from `initial` (adaptive rows + gaussian x2 + triangular + global)
    from visual_rag_b200.embedding.repool import recompute_pooling_from_initial
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
    h = rng.integers(16, 33, size=n)
    w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
    off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
    c.add_synthetic_store("initial", 0, page_offsets=off, seed=1)
    for _ in range(reps):
        info = recompute_pooling_from_initial(c, grids=np.stack([h, w], axis=1))
    b = int(off[-1]) * 256 + (4 * int(np.minimum(h, 32).sum()) + n) * 256
    print("pool_qwen", n, info, "GB/s", b / (info["ms"] * 1e-3) / 1e9)
elif what == "pool_rows":
    c.add_synthetic_store("mean_pooling", 1_000_000, fixed_rows=32, seed=6)
    for _ in range(reps):
        ms = c.pool_store("mean_pooling", [GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"),
                                           GP.spec_global_mean(True)], ["e1", "e2", "e3", "g"])
    print("pool_rows", ms)
c.close()
