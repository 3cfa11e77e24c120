"""Goldens for the bulk re-pooling path by RUNNING THE REFERENCE script's own functions:
scripts/qdrant_recompute_colqwen_pooling_from_initial.py::_infer_grid (64-105) and its per-point arithmetic
(292-327: adaptive row-mean pooling of the stored `initial` tokens with the inferred grid, gaussian / triangular
smoothing and the global mean, all in fp32).  Build container only:  python tests/golden/make_golden_repool.py
Writes tests/golden/repool_golden.npz + repool_index.json."""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VRAG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
from fake_qdrant import install_qdrant_stub  # noqa: E402

install_qdrant_stub()
import cases as CS  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_repool", os.path.join(REF, "scripts", "qdrant_recompute_colqwen_pooling_from_initial.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
from visual_rag.embedding.pooling import adaptive_row_mean_pooling_from_grid, weighted_row_smoothing_same_length  # noqa: E402


def main():
    grid_cases = CS.infer_grid_cases()
    grids = [list(mod._infer_grid(n, width=w, height=h)) for n, w, h in grid_cases]
    out = {}
    index = []
    for c in CS.repool_cases():
        emb = CS.unit_rows(c["seed"], c["n"], dtype=np.float16).astype(np.float32)   # what retrieve() returns from an fp16 store
        gh, gw = mod._infer_grid(c["n"], width=c["w"], height=c["h"])
        cap = c["cap"]
        mean_pool = adaptive_row_mean_pooling_from_grid(emb, grid_h=int(gh), grid_w=int(gw),
                                                        target_rows=(int(gh) if cap <= 0 else min(cap, int(gh))), output_dtype=np.float32)
        g = weighted_row_smoothing_same_length(mean_pool, window_size=3, kernel="gaussian", output_dtype=np.float32)
        t = weighted_row_smoothing_same_length(mean_pool, window_size=3, kernel="triangular", output_dtype=np.float32)
        glob = mean_pool.mean(axis=0).astype(np.float32)
        k = c["key"]
        out[k + "::mean_pooling"] = mean_pool
        out[k + "::experimental_pooling_gaussian"] = g
        out[k + "::experimental_pooling_triangular"] = t
        out[k + "::global_pooling"] = glob
        index.append({"key": k, "grid": [int(gh), int(gw)]})
    # saliency patch scores (visualization/saliency.py:53-79); without matplotlib the overlay is a no-op
    from PIL import Image
    from visual_rag.visualization.saliency import generate_saliency_map
    sal = []
    for c in CS.saliency_cases():
        q = CS.query_rows(c["qseed"], c["q"])
        d = CS.unit_rows(c["seed"], c["n"], dtype=np.float16).astype(np.float32)
        _, ps = generate_saliency_map(q, d, Image.new("RGB", (64, 64)), token_info=c["token_info"])
        out[c["key"] + "::patch_scores"] = np.asarray(ps, dtype=np.float32)
        sal.append(c["key"])
    np.savez_compressed(os.path.join(HERE, "repool_golden.npz"), **out)
    with open(os.path.join(HERE, "repool_index.json"), "w") as f:
        json.dump({"infer_grid": grids, "repool": index, "saliency": sal}, f)
    print("wrote", len(grids), "grid cases,", len(index), "repool cases")


if __name__ == "__main__":
    main()
