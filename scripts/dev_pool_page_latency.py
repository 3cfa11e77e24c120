"""Dev: latency of the per-page pooling calls (the drop-in for calling a pooling.py function on one numpy array) next
to the oracle port of the same function on the host."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200")]
from oracle import pooling_oracle as PO
from visual_rag_b200.embedding import pooling as GP
rng = np.random.default_rng(0)
tok = rng.standard_normal((1024, 128)).astype(np.float32)
smol = rng.standard_normal((832, 128)).astype(np.float32)
def bench(label, gpu, cpu, n=300):
    for _ in range(10): gpu()
    t = time.perf_counter()
    for _ in range(n): g = gpu()
    tg = 1e6 * (time.perf_counter() - t) / n
    t = time.perf_counter()
    for _ in range(30): c = cpu()
    tc = 1e6 * (time.perf_counter() - t) / 30
    same = np.array_equal(np.asarray(g), np.asarray(c))
    print(f"{label:46s} gpu call {tg:7.1f} us | cpu port {tc:7.1f} us | bit-identical {same}")
rows = GP.colpali_row_mean_pooling(tok, 32)
bench("colpali_row_mean_pooling 1024->32", lambda: GP.colpali_row_mean_pooling(tok, 32), lambda: PO.colpali_row_mean_pooling(tok, 32))
bench("tile_level_mean_pooling 832->13", lambda: GP.tile_level_mean_pooling(smol, 13), lambda: PO.tile_level_mean_pooling(smol, 13))
bench("adaptive_row_mean 24x30 -> 20", lambda: GP.adaptive_row_mean_pooling_from_grid(tok[:720], grid_h=24, grid_w=30, target_rows=20),
      lambda: PO.adaptive_row_mean_pooling_from_grid(tok[:720], grid_h=24, grid_w=30, target_rows=20))
bench("weighted_row_smoothing gaussian k=3 (32 rows)", lambda: GP.weighted_row_smoothing_same_length(rows, window_size=3, kernel="gaussian"),
      lambda: PO.weighted_row_smoothing_same_length(rows, window_size=3, kernel="gaussian"))
bench("colpali_experimental (legacy conv) k=3", lambda: GP.colpali_experimental_pooling_from_rows(rows, window_size=3),
      lambda: PO.colpali_experimental_pooling_from_rows(rows, window_size=3))
bench("global_mean_pooling 1024", lambda: GP.global_mean_pooling(tok), lambda: PO.global_mean_pooling(tok))
q = rng.standard_normal((20, 128)).astype(np.float32)
doc = rng.standard_normal((768, 128)).astype(np.float16).astype(np.float32)
from oracle import maxsim_oracle as MO
bench("compute_maxsim_score 20 x 768", lambda: GP.compute_maxsim_score(q, doc), lambda: MO.maxsim_score(q, doc))
docs = [rng.standard_normal((768, 128)).astype(np.float16).astype(np.float32) for _ in range(256)]
bench("compute_maxsim_batch 256 docs", lambda: GP.compute_maxsim_batch(q, docs), lambda: [MO.maxsim_score(q, d) for d in docs], n=20)
