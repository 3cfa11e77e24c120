"""Batched dense stage 1 over SMALL stores (the shards of a strongly scaled corpus): device time of a 256-query
pooled_query_vs_global batch (k=1000) with the fused top-k prefilter and with VRAG_PREFILTER=0 (score matrix + radix select)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus, pack_queries

rng = np.random.default_rng(0)
qs = pack_queries([rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(256)])
out = {}
with GpuCorpus(0) as c:
    for n in (125_000, 250_000, 500_000):
        c.add_synthetic_store("g", n, fixed_rows=1, seed=3)
        res = {}
        for mode in ("prefilter", "matrix"):
            if mode == "matrix":
                os.environ["VRAG_PREFILTER"] = "0"
            else:
                os.environ.pop("VRAG_PREFILTER", None)
            t = []
            for _ in range(6):
                r = c.search_multistage_batch([("g", True, 1000)], qs, as_arrays=True)
                t.append(c.last_timing_ms()[0])
            res[mode] = float(np.median(t[1:]))
            res[mode + "_ids"] = r[0][1]
        same = bool(np.array_equal(res.pop("prefilter_ids"), res.pop("matrix_ids")))
        out[str(n)] = dict(res, identical=same)
        c.drop_store("g")
os.environ.pop("VRAG_PREFILTER", None)
print(json.dumps(out))
