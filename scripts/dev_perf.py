"""Developer perf probe (run under gpurun): stage-1 scans over pooled stores, top-k, pooling throughput."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
from visual_rag_b200.embedding import pooling as GP

PEAK = 6549.8
rng = np.random.default_rng(0)
c = GpuCorpus(0)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    tot, ker = [], []
    for _ in range(n):
        fn()
        t = c.last_timing_ms()
        tot.append(t[0]); ker.append(t[1])
    return float(np.median(tot)), float(np.median(ker))

q20 = rng.standard_normal((20, 128)).astype(np.float32)
for name, n_pages, R, q, pool in (("mean_pooling R=32 Q=20", 1_000_000, 32, q20, False), ("R=32 pooled query", 1_000_000, 32, q20, True),
                                  ("legacy R=34", 1_000_000, 34, q20, False), ("colsmol R=13", 2_000_000, 13, q20, False),
                                  ("colsmol exp R=76", 500_000, 76, q20, False), ("global R=1 pooled", 4_000_000, 1, q20, True),
                                  ("R=16", 2_000_000, 16, q20, False), ("R=64", 500_000, 64, q20, False)):
    c.add_synthetic_store("p", n_pages, fixed_rows=R, seed=3)
    for k in (10, 256, 1000):
        tot, ker = timeit(lambda: c.search("p", q, k, pool_query=pool))
        gb = n_pages * R * 260 / 1e9
        print(f"{name:26s} k={k:4d}: total {tot:7.3f} ms  scan {ker:7.3f} ms  {gb/ker*1e3:6.0f} GB/s ({gb/ker*1e3/PEAK:5.1%})  topk+rest {tot-ker:6.3f} ms", flush=True)
    c.drop_store("p")
# variable <= 32 rows
lens = rng.integers(16, 33, size=1_000_000)
off = np.concatenate([[0], np.cumsum(lens)])
c.add_synthetic_store("p", 0, page_offsets=off, seed=4)
tot, ker = timeit(lambda: c.search("p", q20, 256))
gb = off[-1] * 260 / 1e9
print(f"variable 16..32 rows       k= 256: total {tot:7.3f} ms  scan {ker:7.3f} ms  {gb/ker*1e3:6.0f} GB/s ({gb/ker*1e3/PEAK:5.1%})", flush=True)
cand = rng.permutation(1_000_000)[:1000]
tot, ker = timeit(lambda: c.search("p", q20, 300, candidate_ids=cand))
print(f"1000 candidates (pooled)   k= 300: total {tot:7.3f} ms  scan {ker:7.3f} ms", flush=True)
c.drop_store("p")
# rerank
c.add_synthetic_store("initial", 100_000, fixed_rows=1030, seed=5)
for nc in (256, 300, 1000):
    cand = rng.permutation(100_000)[:nc]
    tot, ker = timeit(lambda: c.search("initial", q20, 10, candidate_ids=cand))
    print(f"rerank {nc:4d} x 1030 tok     k=  10: total {tot:7.3f} ms  scan {ker:7.3f} ms  {nc*1030*260/1e9/ker*1e3:6.0f} GB/s", flush=True)
# pooling throughput: ColPali cfg4
c.add_synthetic_store("vis", 200_000, fixed_rows=1024, seed=6)
for _ in range(3):
    ms1 = c.pool_store("vis", [GP.spec_adaptive_rows(32, 32, 32)], ["mean_pooling"])
    ms2 = c.pool_store("mean_pooling", [GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True)],
                       ["e1", "e2", "e3", "g"])
b1 = 200_000 * (1024 * 256 + 32 * 256) / 1e9
b2 = 200_000 * (32 + 34 + 32 + 32 + 1) * 256 / 1e9
print(f"pooling ColPali: tokens->32 rows {ms1:.3f} ms ({b1/ms1*1e3:.0f} GB/s, {b1/ms1*1e3/PEAK:.1%}); derived x4 {ms2:.3f} ms ({b2/ms2*1e3:.0f} GB/s); "
      f"{200_000/(ms1+ms2)*1e3/1e6:.2f} M pages/s", flush=True)
for _ in range(3):
    ms3 = c.pool_store("vis", [GP.spec_adaptive_rows(32, 32, 32)] + [GP.derived_from(x, 0) for x in (GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"),
                       GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True))], ["mean_pooling", "e1", "e2", "e3", "g"])
print(f"pooling ColPali FUSED single pass: {ms3:.3f} ms ({(b1+b2)/ms3*1e3:.0f} GB/s algorithmic, {(b1+b2)/ms3*1e3/PEAK:.1%}); {200_000/ms3*1e3/1e6:.2f} M pages/s", flush=True)
c.add_synthetic_store("smol", 200_000, fixed_rows=832, seed=7)
for _ in range(3):
    ms1 = c.pool_store("smol", [GP.spec_tile_mean(64)], ["mean_pooling"])
    ms1b = c.pool_store("smol", [GP.spec_colsmol_experimental(13, 64)], ["exp"])
    ms2 = c.pool_store("mean_pooling", [GP.spec_tile_4n(4, 3), GP.spec_global_mean(True)], ["e2d", "g"])
print(f"pooling ColSmol: tile mean {ms1:.3f} ms ({200_000*(832*256+13*256)/1e9/ms1*1e3:.0f} GB/s); experimental {ms1b:.3f} ms; 4n+global {ms2:.3f} ms; "
      f"{200_000/(ms1+ms1b+ms2)*1e3/1e6:.2f} M pages/s", flush=True)

g = np.tile(np.array([[4, 3]], dtype=np.int32), (200_000, 1))
for _ in range(3):
    ms4 = c.pool_store("smol", [GP.spec_tile_mean(64), GP.spec_colsmol_experimental(0, 64), GP.derived_from(GP.spec_tile_4n(0, 0), 0),
                                GP.derived_from(GP.spec_global_mean(True), 0)], ["mean_pooling", "exp", "e2d", "g"], grid_hw=g)
bs = 200_000 * (832 * 256 + (13 + 76 + 13 + 1) * 256) / 1e9
print(f"pooling ColSmol FUSED single pass: {ms4:.3f} ms ({bs/ms4*1e3:.0f} GB/s algorithmic, {bs/ms4*1e3/PEAK:.1%}); {200_000/ms4*1e3/1e6:.2f} M pages/s", flush=True)
