"""Where a per-pair compute_maxsim_score call spends its time (run on the GPU box): the whole call, the resident-store
scoring alone, the upload alone, and the reference's numpy function."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
sys.path.insert(0, ROOT)
from visual_rag_b200.corpus import GpuCorpus
from visual_rag_b200.embedding import pooling as GP
from oracle import maxsim_oracle as MO

rng = np.random.default_rng(0)
q = rng.standard_normal((20, 128)).astype(np.float32)
doc = rng.standard_normal((768, 128)).astype(np.float32)
doc16 = doc.astype(np.float16)

def med(fn, n=300):
    for _ in range(20):
        fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return round(1e6 * float(np.median(ts)), 1)

out = {}
out["compute_maxsim_score_f32_us"] = med(lambda: GP.compute_maxsim_score(q, doc))
out["compute_maxsim_score_f16_us"] = med(lambda: GP.compute_maxsim_score(q, doc16))
out["numpy_us"] = med(lambda: MO.maxsim_score(q, doc))
with GpuCorpus(0) as c:
    out["score_pages_f32_us"] = med(lambda: c.score_pages(q, [doc]))
    out["score_pages_f16_us"] = med(lambda: c.score_pages(q, [doc16]))
    c.add_store("res", doc16, fixed_rows=768)
    out["score_resident_us"] = med(lambda: c.score("res", q))
    out["add_store_f16_us"] = med(lambda: c.add_store("tmp", doc16, fixed_rows=768))
    out["add_store_f32_us"] = med(lambda: c.add_store("tmp", doc, fixed_rows=768))
    out["search_resident_us"] = med(lambda: c.search("res", q, 1))
print(out)
