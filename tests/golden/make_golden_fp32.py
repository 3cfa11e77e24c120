"""Golden vectors on TRUE-fp32 inputs, produced by RUNNING THE REFERENCE (imported from /root/reference).

    python tests/golden/make_golden_fp32.py        (build container only)
Writes tests/golden/fp32_golden.npz + fp32_index.json: every pooling function on inputs that are not
fp16-representable, compute_maxsim_score / quick_test.search_exhaustive / search_two_stage on fp32 pages.
Inputs are regenerated from seeds (checksums in the index)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VRAG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import logging  # noqa: E402

logging.disable(logging.CRITICAL)

from benchmarks import quick_test  # noqa: E402
from visual_rag.embedding import pooling as P  # noqa: E402

import cases as CS  # noqa: E402


def main():
    out, index = {}, {"pooling": [], "maxsim": []}
    for c in CS.fp32_pooling_cases():
        x = CS.raw_rows(c["seed"], c["n"])
        res = getattr(P, c["fn"])(x, *c["args"], **CS.fix_kwargs(c["kwargs"]))
        out[c["key"]] = res
        index["pooling"].append({"key": c["key"], "in_crc": CS.checksum(x), "out_dtype": str(res.dtype), "out_shape": list(res.shape)})
    for c in CS.fp32_maxsim_cases():
        q = CS.query_rows(c["seed"], c["q"])
        d = CS.raw_rows(c["seed"] + 1, c["t"])
        out[c["key"]] = np.array([P.compute_maxsim_score(q, d), P.compute_maxsim_score(q, d, normalize=False)], dtype=np.float64)
        index["maxsim"].append({"key": c["key"], "in_crc": [CS.checksum(q), CS.checksum(d)]})
    q, docs = CS.fp32_corpus()
    ddict = {i: {"embedding": d, "pooled": P.tile_level_mean_pooling(d, 0, patches_per_tile=32)} for i, d in enumerate(docs)}
    out["corpus_scores"] = np.array(P.compute_maxsim_batch(q, docs), dtype=np.float64)
    ex = quick_test.search_exhaustive(q, ddict, top_k=10)
    out["exhaustive_ids"] = np.array([r["id"] for r in ex], dtype=np.int64)
    out["exhaustive_scores"] = np.array([r["score"] for r in ex], dtype=np.float64)
    ts = quick_test.search_two_stage(q, ddict, prefetch_k=40, top_k=10)
    out["two_stage_ids"] = np.array([r["id"] for r in ts], dtype=np.int64)
    out["two_stage_scores"] = np.array([r["score"] for r in ts], dtype=np.float64)
    index["corpus_in_crc"] = [CS.checksum(q), CS.checksum(np.concatenate(docs))]
    index["numpy"] = np.__version__
    np.savez_compressed(os.path.join(HERE, "fp32_golden.npz"), **out)
    with open(os.path.join(HERE, "fp32_index.json"), "w") as f:
        json.dump(index, f, indent=1)
    print("fp32 pooling cases", len(index["pooling"]), "maxsim", len(index["maxsim"]))


if __name__ == "__main__":
    main()
