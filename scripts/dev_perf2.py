import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
rng = np.random.default_rng(0)
c = GpuCorpus(0)
q20 = rng.standard_normal((20, 128)).astype(np.float32)
lens = rng.integers(16, 33, size=1_000_000)
off = np.concatenate([[0], np.cumsum(lens)])
c.add_synthetic_store("p", 0, page_offsets=off, seed=4)
ref = c.score("p", q20)
for _ in range(3): c.search("p", q20, 256)
ker = np.median([ (c.search("p", q20, 256), c.last_timing_ms()[1])[1] for _ in range(10)])
gb = off[-1] * 260 / 1e9
print(f"VARSLOT={os.environ.get('VRAG_VARSLOT')} variable 16..32 rows scan {ker:.3f} ms {gb/ker*1e3:.0f} GB/s ({gb/ker*1e3/6549.8:.1%})", flush=True)
cand = rng.permutation(1_000_000)[:5000]
got = c.score("p", q20, candidate_ids=cand)
print("cand==dense", np.array_equal(got, ref[cand]))
