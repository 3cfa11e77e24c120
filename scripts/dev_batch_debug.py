import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus
rng = np.random.default_rng(107)
n = 700
def ragged(lo, hi):
    lens = rng.integers(lo, hi + 1, size=n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return rng.standard_normal((int(off[-1]), 128)).astype(np.float16), off
c = GpuCorpus(0)
r, o = ragged(150, 400); c.add_store("i", r, page_offsets=o)
r, o = ragged(16, 32); c.add_store("e", r, page_offsets=o)
c.add_store("g", rng.standard_normal((n, 128)).astype(np.float16), fixed_rows=1)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 7
queries = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(nq)]
stages = [("g", True, 200), ("e", False, 200), ("i", False, 200)]
got = c.search_multistage_batch(stages, queries)
for b, q in enumerate(queries):
    single = c.search_multistage(stages, q)
    for s in range(3):
        # compare as id -> score maps (k == n_prev so every candidate is reported)
        gm = dict(zip(got[b][s][1].tolist(), got[b][s][0].tolist()))
        sm = dict(zip(single[s][1].tolist(), single[s][0].tolist()))
        bad = [(i, gm.get(i), sm[i]) for i in sm if gm.get(i) != sm[i]]
        cand_pos = {pid: j for j, pid in enumerate(got[b][s - 1][1].tolist())} if s else {}
        print(f"q{b} rows={q.shape[0]} stage{s}: {len(bad)} mismatches", [(i, cand_pos.get(i), g_, s_) for i, g_, s_ in bad[:6]])
