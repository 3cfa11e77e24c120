import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
variant = sys.argv[1]
rng = np.random.default_rng(0)
c = GpuCorpus(0)
n, R = 1001, 32
if variant == "r64": R = 64
if variant == "r128": R = 128
rows = rng.standard_normal((n * R, 128)).astype(np.float16)
c.add_store("p", rows, fixed_rows=R)
q = rng.standard_normal((20, 128)).astype(np.float32)
full = c.score("p", q, normalize=(variant != "nonorm"))
ncand = 4 if variant == "four" else (1 if variant == "one" else 77)
cand = rng.permutation(n)[:ncand]
if variant == "sorted": cand = np.sort(cand)
got = c.score("p", q, candidate_ids=cand, normalize=(variant != "nonorm"))
print(variant, np.abs(got - full[cand]).max())
