"""Generate golden input/output vectors by RUNNING THE REFERENCE (imported from /root/reference).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/pooling_golden.npz, maxsim_golden.npz and retrieval_golden.json.  Inputs are seeded;
the fixture stores both inputs and the reference's outputs so the tests never need the reference.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VRAG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from fake_qdrant import NumpyQdrant, install_qdrant_stub  # noqa: E402

install_qdrant_stub()

import logging  # noqa: E402

logging.disable(logging.CRITICAL)

from benchmarks import quick_test  # noqa: E402
from visual_rag.embedding import pooling as P  # noqa: E402
from visual_rag.embedding.visual_embedder import VisualEmbedder  # noqa: E402
from visual_rag.retrieval.three_stage import ThreeStageRetriever  # noqa: E402
from visual_rag.retrieval.two_stage import TwoStageRetriever  # noqa: E402


import cases as CS  # noqa: E402


def pooling_goldens():
    out, index = {}, []
    for c in CS.pooling_cases():
        x = CS.unit_rows(c["seed"], c["n"], dtype=np.dtype(c["dtype"]).type)
        fn = getattr(P, c["fn"])
        res = fn(x, *c["args"], **CS.fix_kwargs(c["kwargs"]))
        out[c["key"]] = res
        index.append({"key": c["key"], "in_crc": CS.checksum(x), "out_dtype": str(res.dtype), "out_shape": list(res.shape)})
    return out, index


def dispatch_goldens():
    """p9/p10/p11: model-aware dispatch driven in the pipeline's call order (pipeline.py:400-507)."""
    out, index = {}, []
    for c in CS.dispatch_cases():
        emb = VisualEmbedder(model_name=c["model"], device="cpu", output_dtype=np.dtype(c["out_dtype"]).type)
        visual = CS.unit_rows(c["seed"], c["n"])
        info = c["token_info"]
        ml = c["model"].lower()
        is_q = "colqwen2.5" in ml or "colqwen2_5" in ml
        is_s = "colsmol" in ml
        tv = c["cap"]
        if tv is not None:
            tv = None if int(tv) <= 0 else int(tv)
        mean_pool = emb.mean_pool_visual_embedding(visual, info, target_vectors=tv)
        named = {"mean_pooling": mean_pool}
        if is_q:
            g = emb.experimental_pool_visual_embedding(visual, info, target_vectors=tv, mean_pool=mean_pool, window_size=3, kernel="gaussian")
            tr = emb.experimental_pool_visual_embedding(visual, info, target_vectors=tv, mean_pool=mean_pool, window_size=3, kernel="triangular")
            named["experimental_pooling"] = g
            named["experimental_pooling_gaussian"] = g
            named["experimental_pooling_triangular"] = tr
        else:
            kernel = "legacy" if c["kernel"] == "auto" else c["kernel"]
            ks = c["windows"] if c["windows"] else [3]
            for k in ks:
                e = emb.experimental_pool_visual_embedding(visual, info, target_vectors=tv, mean_pool=mean_pool, window_size=int(k), kernel=kernel)
                named[f"experimental_pooling_{int(k)}"] = e
                if k == ks[0]:
                    named["experimental_pooling"] = e
        if is_s and c["twod"] and info.get("n_rows"):
            named["experimental_pooling_2d"] = P.colsmol_tile_4n_pooling_from_tiles(
                mean_pool, n_rows=info["n_rows"], n_cols=info["n_cols"], has_global=True, include_self=True, output_dtype=emb.output_dtype)
        named["global_pooling"] = emb.global_pool_from_mean_pool(mean_pool)
        for nm, arr in named.items():
            out[f"{c['key']}::{nm}"] = arr
        index.append({"key": c["key"], "in_crc": CS.checksum(visual), "names": sorted(named.keys())})
    return out, index


def maxsim_goldens():
    out, index = {}, []
    for c in CS.maxsim_cases():
        q = CS.query_rows(c["seed"], c["q"])
        d = CS.unit_rows(c["seed"] + 1, c["t"], scale=True)
        out[c["key"]] = np.array([P.compute_maxsim_score(q, d), P.compute_maxsim_score(q, d, normalize=False)], dtype=np.float64)
        index.append({"key": c["key"], "in_crc": [CS.checksum(q), CS.checksum(d)]})
    q, docs = CS.bench_corpus()
    out["batch_scores"] = np.array(P.compute_maxsim_batch(q, docs), dtype=np.float64)
    doc_dict = {i: {"embedding": d, "pooled": P.tile_level_mean_pooling(d, 4, patches_per_tile=64)} for i, d in enumerate(docs)}
    out["batch_pooled"] = np.stack([doc_dict[i]["pooled"] for i in range(len(docs))])
    ex = quick_test.search_exhaustive(q, doc_dict, top_k=10)
    out["exhaustive_ids"] = np.array([r["id"] for r in ex], dtype=np.int64)
    out["exhaustive_scores"] = np.array([r["score"] for r in ex], dtype=np.float64)
    ts = quick_test.search_two_stage(q, doc_dict, prefetch_k=30, top_k=10)
    out["two_stage_ids"] = np.array([r["id"] for r in ts], dtype=np.int64)
    out["two_stage_scores"] = np.array([r["score"] for r in ts], dtype=np.float64)
    out["two_stage_rank1"] = np.array([r["stage1_rank"] for r in ts], dtype=np.int64)
    index.append({"key": "bench", "in_crc": [CS.checksum(q), CS.checksum(np.stack(docs))]})
    return out, index


def retrieval_goldens():
    """Run the reference retriever classes against the in-memory client (scores by the reference's own
    compute_maxsim_score) and record their outputs. Derived pooled stores are built with the reference's
    pooling functions and saved (fp16) so the tests can load the very same corpus."""
    q, initial = CS.retrieval_corpus()
    n = len(initial)
    mean_pool = [P.tile_level_mean_pooling(d, 0, patches_per_tile=16, output_dtype=np.float16).astype(np.float32) for d in initial]
    experimental = [P.weighted_row_smoothing_same_length(m, window_size=3, kernel="gaussian", output_dtype=np.float16).astype(np.float32) for m in mean_pool]
    global_pool = [m.mean(axis=0).astype(np.float16).astype(np.float32) for m in mean_pool]
    vectors = {"initial": initial, "mean_pooling": mean_pool, "experimental_pooling": experimental, "global_pooling": global_pool}
    client = NumpyQdrant(vectors, P.compute_maxsim_score)
    arrays = {"offsets_pooled": np.concatenate([[0], np.cumsum([d.shape[0] for d in mean_pool])]).astype(np.int64),
              "mean_pooling": np.concatenate(mean_pool).astype(np.float16),
              "experimental_pooling": np.concatenate(experimental).astype(np.float16),
              "global_pooling": np.stack(global_pool).astype(np.float16)}
    results = {}
    two = TwoStageRetriever(client, "c")
    for mode in ("pooled_query_vs_tiles", "tokens_vs_tiles", "pooled_query_vs_global"):
        r = two.search(q, top_k=10, prefetch_k=40, stage1_mode=mode)
        results[f"two_stage_search::{mode}"] = [{"id": x["id"], "score_stage1": x["score_stage1"], "score_stage2": x["score_stage2"], "score_final": x["score_final"]} for x in r]
    r = two.search(q, top_k=10, prefetch_k=40, stage1_mode="tokens_vs_tiles", use_reranking=False)
    results["two_stage_search::norerank"] = [{"id": x["id"], "score_stage1": x["score_stage1"], "score_final": x["score_final"]} for x in r]
    for mode in ("pooled_query_vs_standard_pooling", "tokens_vs_standard_pooling", "pooled_query_vs_experimental_pooling",
                 "tokens_vs_experimental_pooling", "pooled_query_vs_global", "tokens_vs_tiles"):
        r = two.search_server_side(q, top_k=10, prefetch_k=40, stage1_mode=mode)
        results[f"two_stage_server::{mode}"] = [{"id": x["id"], "score_final": x["score_final"]} for x in r]
    for use_pooling in (False, True):
        r = two.search_single_stage(q, top_k=10, use_pooling=use_pooling)
        results[f"two_stage_single::{use_pooling}"] = [{"id": x["id"], "score_final": x["score_final"]} for x in r]
    three = ThreeStageRetriever(client, "c")
    r = three.search_server_side(query_embedding=q, top_k=10, stage1_k=80, stage2_k=30)
    results["three_stage"] = [{k: x[k] for k in ("id", "score_stage1", "score_stage2", "score_stage3", "score_final")} for x in r]
    from visual_rag.retrieval.single_stage import SingleStageRetriever

    single = SingleStageRetriever(client, "c")
    for strat in ("multi_vector", "tiles_maxsim", "pooled_tile", "pooled_global", "experimental_maxsim", "pooled_experimental"):
        r = single.search(q, top_k=10, strategy=strat)
        results[f"single::{strat}"] = [{"id": x["id"], "score": x["score"]} for x in r]
    crc = [CS.checksum(q), CS.checksum(np.concatenate(initial))]
    return arrays, results, crc


def main():
    pool_arrays, pool_index = pooling_goldens()
    disp_arrays, disp_index = dispatch_goldens()
    pool_arrays.update(disp_arrays)
    np.savez_compressed(os.path.join(HERE, "pooling_golden.npz"), **pool_arrays)
    ms_arrays, ms_index = maxsim_goldens()
    np.savez_compressed(os.path.join(HERE, "maxsim_golden.npz"), **ms_arrays)
    ret_arrays, ret_results, ret_crc = retrieval_goldens()
    np.savez_compressed(os.path.join(HERE, "retrieval_golden.npz"), **ret_arrays)
    with open(os.path.join(HERE, "golden_index.json"), "w") as f:
        json.dump({"pooling": pool_index, "dispatch": disp_index, "maxsim": ms_index, "retrieval": ret_results,
                   "retrieval_in_crc": ret_crc, "numpy": np.__version__}, f, indent=1)
    print("pooling cases", len(pool_index), "dispatch", len(disp_index), "maxsim", len(ms_index), "retrieval", len(ret_results))


if __name__ == "__main__":
    main()
