"""CPU oracle for the retrieval scoring path — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's client-side arithmetic, used as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
visual-rag-toolkit_b200/ imports this module; the product path is CUDA-only.

Parity pinning: tests/test_oracle_golden.py checks every function here against outputs of the
reference itself (tests/golden/*.npz, produced by tests/golden/make_golden.py which imports
/root/reference) and against the known answers of the reference's own tests (tests/test_pooling.py).

Each function cites the reference lines it restates (paths relative to the reference repo root).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

EPS = 1e-8  # visual_rag/embedding/pooling.py:498,500


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """x / (||x||_2 + 1e-8) row-wise, in the dtype of x — pooling.py:497-500."""
    return x / (np.linalg.norm(x, axis=1, keepdims=True) + EPS)


def maxsim_score(query: np.ndarray, doc: np.ndarray, normalize: bool = True) -> float:
    """sum_i max_j <q_i, d_j> — compute_maxsim_score, pooling.py:468-514."""
    q = l2_normalize_rows(query) if normalize else query
    d = l2_normalize_rows(doc) if normalize else doc
    sim = np.dot(q, d.T)                      # pooling.py:506
    return float(sim.max(axis=1).sum())       # pooling.py:509-512


def maxsim_batch(query: np.ndarray, docs: Sequence[np.ndarray], normalize: bool = True) -> List[float]:
    """compute_maxsim_batch, pooling.py:517-552 (query normalised once)."""
    q = l2_normalize_rows(query) if normalize else query
    out = []
    for doc in docs:
        d = l2_normalize_rows(doc) if normalize else doc
        out.append(float(np.dot(q, d.T).max(axis=1).sum()))
    return out


def pooled_query_score(query: np.ndarray, doc_rows: np.ndarray) -> float:
    """max_t cos(mean_q, d_t): stage 1 of quick_test.search_two_stage, benchmarks/quick_test.py:182-191
    (twin: benchmarks/run_vidore.py:226-235). Also the cosine/MAX_SIM restatement of the Qdrant
    pooled_query_vs_* modes (two_stage.py:141-155; qdrant_indexer.py:200-239)."""
    qb = query.mean(axis=0)
    qb = qb / (np.linalg.norm(qb) + EPS)
    dn = doc_rows / (np.linalg.norm(doc_rows, axis=1, keepdims=True) + EPS)
    return float(np.dot(dn, qb).max())


def stable_topk(scores: Sequence[float], k: int) -> List[int]:
    """Indices of the k best scores, descending, ties keep the original order — what
    `list.sort(key=score, reverse=True)[:k]` does (quick_test.py:165-166, two_stage.py:424-426)."""
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)
    return order[:k]


def search_exhaustive(query: np.ndarray, docs: Sequence[np.ndarray], top_k: int = 10) -> List[Tuple[int, float]]:
    """quick_test.search_exhaustive, benchmarks/quick_test.py:158-166: [(doc index, score)]."""
    scores = [maxsim_score(query, d) for d in docs]
    return [(i, scores[i]) for i in stable_topk(scores, top_k)]


def search_two_stage_pooled(
    query: np.ndarray,
    docs: Sequence[np.ndarray],
    pooled: Sequence[np.ndarray],
    prefetch_k: int = 20,
    top_k: int = 10,
) -> List[Tuple[int, float, int]]:
    """quick_test.search_two_stage, benchmarks/quick_test.py:169-206: pooled-query-vs-tiles prefetch,
    exact MaxSim rerank. Returns [(doc index, score, stage1_rank)]."""
    s1 = [pooled_query_score(query, p) for p in pooled]
    cand = stable_topk(s1, prefetch_k)
    s2 = [maxsim_score(query, docs[i]) for i in cand]
    order = stable_topk(s2, top_k)
    return [(cand[j], s2[j], j + 1) for j in order]


# ------------------------------------------------------------------------------------------------
# Stage semantics of the retriever classes (cosine + MAX_SIM restatement).

STAGE1_ALIASES = {  # two_stage.py:131-139
    "pooled_query_vs_tiles": "pooled_query_vs_standard_pooling",
    "tokens_vs_tiles": "tokens_vs_standard_pooling",
    "pooled_query_vs_experimental": "pooled_query_vs_experimental_pooling",
    "tokens_vs_experimental": "tokens_vs_experimental_pooling",
}


def stage1_plan(stage1_mode: str) -> Tuple[bool, str]:
    """mode -> (pool_query, store key in {"pooled","experimental","global"}) — two_stage.py:141-157."""
    mode = STAGE1_ALIASES.get(stage1_mode, stage1_mode)
    table = {
        "pooled_query_vs_standard_pooling": (True, "pooled"),
        "tokens_vs_standard_pooling": (False, "pooled"),
        "pooled_query_vs_experimental_pooling": (True, "experimental"),
        "tokens_vs_experimental_pooling": (False, "experimental"),
        "pooled_query_vs_global": (True, "global"),
    }
    if mode not in table:
        raise ValueError(f"Unknown stage1_mode: {stage1_mode}")
    return table[mode]


def stage_scores(query: np.ndarray, store: Sequence[np.ndarray], pool_query: bool,
                 cand: Optional[Sequence[int]] = None) -> List[float]:
    """Score pages of one named store (all, or the candidate subset) with one stage's semantics:
    tokens -> MaxSim; pooled query -> max_t cos(mean_q, d_t)."""
    idx = range(len(store)) if cand is None else cand
    q = query.mean(axis=0, keepdims=True) if pool_query else query
    return [maxsim_score(q, np.asarray(store[i], dtype=np.float32)) for i in idx]


def multistage(query: np.ndarray, stages: Sequence[Tuple[Sequence[np.ndarray], bool, int]]):
    """Generic restatement of the server-side pipelines: every stage scores only the survivors of the
    previous one and keeps its k best.
      two-stage  (two_stage.py:161-178):  [(pooled store, pool?, prefetch_k), (initial, False, top_k)]
      three-stage (three_stage.py:102-159): [(global, True, stage1_k), (experimental, False, min(stage2_k, n1)), (initial, False, top_k)]
    Returns per stage [(page index, score)]."""
    out = []
    cand: Optional[List[int]] = None
    for store, pool, k in stages:
        sc = stage_scores(query, store, pool, cand)
        ids = list(range(len(store))) if cand is None else list(cand)
        order = stable_topk(sc, k)
        res = [(ids[j], sc[j]) for j in order]
        out.append(res)
        cand = [r[0] for r in res]
    return out


def merge_shard_topk(lists: Sequence[Sequence[Tuple[int, float]]], k: int) -> List[Tuple[int, float]]:
    """Deterministic merge of per-shard top-k lists [(global id, score)]: score descending, ties -> lower id
    (the order a single-shard stable sort over pages in id order produces)."""
    flat = [x for lst in lists for x in lst]
    flat.sort(key=lambda t: (-t[1], t[0]))
    return flat[:k]


def saliency_patch_scores(query: np.ndarray, doc: np.ndarray) -> np.ndarray:
    """patch_scores of generate_saliency_map, visual_rag/visualization/saliency.py:69-79: max over query tokens of
    the cosine with each document token."""
    q = l2_normalize_rows(np.asarray(query, dtype=np.float32))
    d = l2_normalize_rows(np.asarray(doc, dtype=np.float32))
    return np.dot(q, d.T).max(axis=0)
