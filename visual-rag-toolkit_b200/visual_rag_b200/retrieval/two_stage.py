"""TwoStageRetriever on the GPU-resident corpus — mirror of visual_rag/retrieval/two_stage.py.

Same constructor, method names, keyword arguments, defaults and result-dict keys as the reference; the
`qdrant_client` argument is any object with query_points / retrieve (normally a GpuCorpusClient).  Where
the reference pulls prefetch_k full multi-vectors over the wire and scores them with numpy
(_stage2_rerank, two_stage.py:371-426), this class asks the client for an ID-restricted exact MaxSim, which
the GPU backend runs as one gather+rerank kernel.
"""

from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Union

import numpy as np

from . import models as qdrant_models
from ._common import no_gc, resolve_stage1, retry_call, to_numpy, wire
from .models import FieldCondition, Filter, HasIdCondition, MatchAny, MatchValue

logger = logging.getLogger(__name__)


class TwoStageRetriever:
    def __init__(
        self,
        qdrant_client,
        collection_name: str,
        full_vector_name: str = "initial",
        pooled_vector_name: str = "mean_pooling",
        experimental_vector_name: str = "experimental_pooling",
        global_vector_name: str = "global_pooling",
        request_timeout: int = 120,
        max_retries: int = 3,
        retry_sleep: float = 0.5,
    ):
        self.client = qdrant_client
        self.collection_name = collection_name
        self.full_vector_name = full_vector_name
        self.pooled_vector_name = pooled_vector_name
        self.experimental_vector_name = experimental_vector_name
        self.global_vector_name = global_vector_name
        self.request_timeout = int(request_timeout)
        self.max_retries = int(max_retries)
        self.retry_sleep = float(retry_sleep)

    def _retry_call(self, fn):
        return retry_call(fn, self.max_retries, self.retry_sleep)

    def _to_numpy(self, embedding) -> np.ndarray:
        return to_numpy(embedding)

    def _stage1_query(self, query_np: np.ndarray, stage1_mode: str):
        pool, name = resolve_stage1(stage1_mode, self.pooled_vector_name, self.experimental_vector_name,
                                    self.global_vector_name)
        vec = wire(self.client, query_np.mean(axis=0)) if pool else wire(self.client, query_np)   # two_stage.py:142-155
        return vec, name

    # ------------------------------------------------------------------ server-side (fused on the GPU)
    def search_server_side(
        self,
        query_embedding,
        top_k: int = 10,
        prefetch_k: Optional[int] = None,
        filter_obj=None,
        stage1_mode: str = "pooled_query_vs_standard_pooling",
    ) -> List[Dict[str, Any]]:
        """two_stage.py:102-191: one query_points call with prefetch=; all scoring on the device."""
        query_np = self._to_numpy(query_embedding)
        if prefetch_k is None:
            prefetch_k = max(100, top_k * 10)
        prefetch_query, prefetch_using = self._stage1_query(query_np, stage1_mode)
        rerank_query = wire(self.client, query_np)

        def _do_query():
            return self.client.query_points(
                collection_name=self.collection_name,
                query=rerank_query,
                using=self.full_vector_name,
                limit=top_k,
                query_filter=filter_obj,
                with_payload=True,
                search_params=qdrant_models.SearchParams(exact=True),
                prefetch=[qdrant_models.Prefetch(query=prefetch_query, using=prefetch_using, limit=prefetch_k)],
                timeout=self.request_timeout,
            ).points

        results = self._retry_call(_do_query)
        return [
            {"id": r.id, "score_stage1": None, "score_stage2": r.score, "score_final": r.score, "payload": r.payload}
            for r in results
        ]

    def search_server_side_batch(
        self,
        query_embeddings,
        top_k: int = 10,
        prefetch_k: Optional[int] = None,
        filter_obj=None,
        stage1_mode: str = "pooled_query_vs_standard_pooling",
    ) -> List[List[Dict[str, Any]]]:
        """`search_server_side` for a batch of queries: one native call on a GpuCorpusClient (stage 1 as a dense
        batched scan, the rerank of all queries as one launch); otherwise the per-query loop."""
        batch = getattr(self.client, "query_multistage_batch_final", None)
        if batch is None or filter_obj is not None:
            return [self.search_server_side(q, top_k=top_k, prefetch_k=prefetch_k, filter_obj=filter_obj,
                                            stage1_mode=stage1_mode) for q in query_embeddings]
        if prefetch_k is None:
            prefetch_k = max(100, top_k * 10)
        pool, prefetch_using = resolve_stage1(stage1_mode, self.pooled_vector_name, self.experimental_vector_name,
                                              self.global_vector_name)
        qs = [self._to_numpy(q) for q in query_embeddings]
        with no_gc():
            res = self._retry_call(lambda: batch(usings=[prefetch_using, self.full_vector_name],
                                                 limits=[int(prefetch_k), int(top_k)], queries=qs, pool_flags=[pool, False]))
            return [[{"id": pid, "score_stage1": None, "score_stage2": score, "score_final": score, "payload": payload}
                     for pid, score, payload in zip(per_query[0], per_query[1], per_query[3])] for per_query in res]

    # ------------------------------------------------------------------ client-side flow
    def search(
        self,
        query_embedding,
        top_k: int = 10,
        prefetch_k: Optional[int] = None,
        filter_obj=None,
        use_reranking: bool = True,
        return_embeddings: bool = False,
        stage1_mode: str = "pooled_query_vs_standard_pooling",
    ) -> List[Dict[str, Any]]:
        """two_stage.py:193-272: prefetch, then rerank iff use_reranking and more than top_k candidates."""
        query_np = self._to_numpy(query_embedding)
        if prefetch_k is None:
            prefetch_k = max(100, top_k * 10)
        logger.info(f"Stage 1: Prefetching {prefetch_k} candidates ({stage1_mode})")
        candidates = self._stage1_prefetch(query_np=query_np, top_k=prefetch_k, filter_obj=filter_obj,
                                           stage1_mode=stage1_mode)
        if not candidates:
            logger.warning("No candidates found in stage 1")
            return []
        if use_reranking and len(candidates) > top_k:
            results = self._stage2_rerank(query_np=query_np, candidates=candidates, top_k=top_k,
                                          return_embeddings=return_embeddings)
        else:
            results = candidates[:top_k]
            for r in results:
                r["score_final"] = r["score_stage1"]
        return results

    def search_single_stage(self, query_embedding, top_k: int = 10, filter_obj=None,
                            use_pooling: bool = False) -> List[Dict[str, Any]]:
        """two_stage.py:274-326."""
        query_np = self._to_numpy(query_embedding)
        if use_pooling:
            vector_name = self.pooled_vector_name
            query_vector = wire(self.client, query_np.mean(axis=0))
        else:
            vector_name = self.full_vector_name
            query_vector = wire(self.client, query_np)
        results = self.client.query_points(
            collection_name=self.collection_name, query=query_vector, using=vector_name, query_filter=filter_obj,
            limit=top_k, with_payload=True, with_vectors=False, timeout=120,
        ).points
        return [{"id": r.id, "score_stage1": r.score, "score_final": r.score, "payload": r.payload} for r in results]

    def _stage1_prefetch(self, query_np: np.ndarray, top_k: int, filter_obj=None,
                         stage1_mode: str = "pooled_query_vs_standard_pooling") -> List[Dict[str, Any]]:
        """two_stage.py:328-369 (accepts every stage1_mode name, not only the legacy ones)."""
        query_vector, vector_name = self._stage1_query(query_np, stage1_mode)

        def _do_query():
            return self.client.query_points(
                collection_name=self.collection_name, query=query_vector, using=vector_name,
                query_filter=filter_obj, limit=top_k, with_payload=True, with_vectors=False,
                timeout=self.request_timeout,
            ).points

        results = self._retry_call(_do_query)
        return [{"id": r.id, "score_stage1": r.score, "payload": r.payload} for r in results]

    def _stage2_rerank(self, query_np: np.ndarray, candidates: List[Dict[str, Any]], top_k: int,
                       return_embeddings: bool = False) -> List[Dict[str, Any]]:
        """two_stage.py:371-426. The exact MaxSim of every candidate is computed on the device (ID-restricted
        scan of the full-token store) instead of retrieve() + numpy; candidates the store does not hold fall
        back to their stage-1 score (408-411); the final order is a stable sort by score_final (424), i.e.
        ties keep stage-1 order."""
        candidate_ids = [c["id"] for c in candidates]

        def _do_rerank():
            return self.client.query_points(
                collection_name=self.collection_name, query=wire(self.client, query_np), using=self.full_vector_name,
                query_filter=Filter(must=[HasIdCondition(has_id=candidate_ids)]), limit=len(candidate_ids),
                with_payload=False, with_vectors=False, search_params=qdrant_models.SearchParams(exact=True),
                timeout=self.request_timeout,
            ).points

        scored = {p.id: p.score for p in self._retry_call(_do_rerank)}
        embeddings = {}
        if return_embeddings:
            def _do_retrieve():
                return self.client.retrieve(collection_name=self.collection_name, ids=candidate_ids,
                                            with_payload=False, with_vectors=[self.full_vector_name],
                                            timeout=self.request_timeout)

            for point in self._retry_call(_do_retrieve):
                if point.vector and self.full_vector_name in point.vector:
                    embeddings[point.id] = np.array(point.vector[self.full_vector_name], dtype=np.float32)
        reranked = []
        for candidate in candidates:
            score = scored.get(candidate["id"])
            if score is None:
                candidate["score_stage2"] = candidate["score_stage1"]
                candidate["score_final"] = candidate["score_stage1"]
            else:
                candidate["score_stage2"] = score
                candidate["score_final"] = score
                if return_embeddings and candidate["id"] in embeddings:
                    candidate["embedding"] = embeddings[candidate["id"]]
            reranked.append(candidate)
        reranked.sort(key=lambda x: x["score_final"], reverse=True)
        return reranked[:top_k]

    def build_filter(self, year: Optional[Any] = None, source: Optional[str] = None, district: Optional[str] = None,
                     filename: Optional[str] = None, has_text: Optional[bool] = None):
        """two_stage.py:436-480: single values -> MatchValue, lists -> MatchAny."""
        conditions = []
        if year is not None:
            if isinstance(year, list):
                conditions.append(FieldCondition(key="year", match=MatchAny(any=[int(y) if isinstance(y, str) else y for y in year])))
            else:
                conditions.append(FieldCondition(key="year", match=MatchValue(value=int(year) if isinstance(year, str) else year)))
        for key, val in (("source", source), ("district", district), ("filename", filename)):
            if val is not None:
                if isinstance(val, list):
                    conditions.append(FieldCondition(key=key, match=MatchAny(any=val)))
                else:
                    conditions.append(FieldCondition(key=key, match=MatchValue(value=val)))
        if has_text is not None:
            conditions.append(FieldCondition(key="has_text", match=MatchValue(value=has_text)))
        return Filter(must=conditions) if conditions else None
