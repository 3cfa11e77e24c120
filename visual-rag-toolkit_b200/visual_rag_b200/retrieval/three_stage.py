"""ThreeStageRetriever — mirror of visual_rag/retrieval/three_stage.py (global -> experimental -> exact
MaxSim, each stage restricted to the previous stage's ids)."""

from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np

from . import models as m
from ._common import no_gc, retry_call, to_numpy, wire

logger = logging.getLogger(__name__)


class ThreeStageRetriever:
    def __init__(
        self,
        qdrant_client,
        collection_name: str,
        *,
        full_vector_name: str = "initial",
        experimental_vector_name: str = "experimental_pooling",
        global_vector_name: str = "global_pooling",
        request_timeout: int = 120,
        max_retries: int = 3,
        retry_sleep: float = 0.5,
    ):
        self.client = qdrant_client
        self.collection_name = collection_name
        self.full_vector_name = full_vector_name
        self.experimental_vector_name = experimental_vector_name
        self.global_vector_name = global_vector_name
        self.request_timeout = int(request_timeout)
        self.max_retries = int(max_retries)
        self.retry_sleep = float(retry_sleep)

    def _retry_call(self, fn):
        return retry_call(fn, self.max_retries, self.retry_sleep)

    def _to_numpy(self, embedding) -> np.ndarray:
        return to_numpy(embedding)

    def _and_filter(self, base_filter, ids: List[Any]):
        has_id = m.HasIdCondition(has_id=list(ids))
        if base_filter is None:
            return m.Filter(must=[has_id])
        return m.Filter(must=[base_filter, has_id])

    def search_server_side(
        self,
        *,
        query_embedding,
        top_k: int = 100,
        stage1_k: Optional[int] = 1000,
        stage2_k: Optional[int] = 300,
        filter_obj=None,
        stage1_mode: Optional[str] = None,
    ) -> List[Dict[str, Any]]:
        """three_stage.py:83-173. `stage1_mode` is accepted and ignored and None stage sizes fall back to
        1000/300, so MultiVectorRetriever.search_embedded(mode="three_stage") works (it raises TypeError in
        the reference)."""
        stage1_k = 1000 if stage1_k is None else int(stage1_k)
        stage2_k = 300 if stage2_k is None else int(stage2_k)
        query_np = self._to_numpy(query_embedding)
        stage1_query = wire(self.client, query_np.mean(axis=0))
        tokens = wire(self.client, query_np)

        fused = getattr(self.client, "query_three_stage", None)
        if fused is not None:
            # GPU backend: the three ID-restricted scans run back to back with one host synchronisation.
            s1, s2, s3 = self._retry_call(lambda: fused(
                stage1_query=stage1_query, stage2_query=tokens, stage3_query=tokens,
                stage1_using=self.global_vector_name, stage2_using=self.experimental_vector_name,
                stage3_using=self.full_vector_name, stage1_k=stage1_k, stage2_k=stage2_k, top_k=int(top_k),
                query_filter=filter_obj))
        else:
            s1 = self._retry_call(lambda: self.client.query_points(
                collection_name=self.collection_name, query=stage1_query, using=self.global_vector_name,
                limit=stage1_k, query_filter=filter_obj, with_payload=False, with_vectors=False,
                timeout=self.request_timeout).points)
            if not s1:
                return []
            s1_ids = [p.id for p in s1]
            s2 = self._retry_call(lambda: self.client.query_points(
                collection_name=self.collection_name, query=tokens, using=self.experimental_vector_name,
                limit=int(min(stage2_k, len(s1_ids))), query_filter=self._and_filter(filter_obj, s1_ids),
                with_payload=False, with_vectors=False, timeout=self.request_timeout).points)
            if not s2:
                return []
            s2_ids = [p.id for p in s2]
            s3 = self._retry_call(lambda: self.client.query_points(
                collection_name=self.collection_name, query=tokens, using=self.full_vector_name, limit=int(top_k),
                query_filter=self._and_filter(filter_obj, s2_ids), with_payload=True, with_vectors=False,
                search_params=m.SearchParams(exact=True), timeout=self.request_timeout).points)
        if not s1 or not s2:
            return []
        s1_score = {str(p.id): float(p.score) for p in s1}
        s2_score = {str(p.id): float(p.score) for p in s2}
        out = []
        for p in s3:
            pid = str(p.id)
            out.append({
                "id": p.id,
                "score_stage1": s1_score.get(pid),
                "score_stage2": s2_score.get(pid),
                "score_stage3": float(p.score),
                "score_final": float(p.score),
                "payload": p.payload,
            })
        return out

    def search_server_side_batch(
        self,
        *,
        query_embeddings,
        top_k: int = 100,
        stage1_k: Optional[int] = 1000,
        stage2_k: Optional[int] = 300,
        filter_obj=None,
    ) -> List[List[Dict[str, Any]]]:
        """`search_server_side` for a batch of queries (BASELINE configs[2]). On a GpuCorpusClient all queries run
        in one native call — the global stage as a dense batched scan with the fused top-k prefilter, the two
        ID-restricted stages as one launch each; with a filter (or any other client) it is the per-query loop the
        reference's evaluation runs (run_qdrant_beir.py:378-402). Same result dicts as search_server_side."""
        batch = getattr(self.client, "query_multistage_batch_final", None)
        if batch is None or filter_obj is not None:
            return [self.search_server_side(query_embedding=q, top_k=top_k, stage1_k=stage1_k, stage2_k=stage2_k,
                                            filter_obj=filter_obj) for q in query_embeddings]
        stage1_k = 1000 if stage1_k is None else int(stage1_k)
        stage2_k = 300 if stage2_k is None else int(stage2_k)
        qs = [self._to_numpy(q) for q in query_embeddings]
        # stage 1 scans with the mean-pooled query (three_stage.py:96), pooled on the device from the same token rows
        with no_gc():
            res = self._retry_call(lambda: batch(
                usings=[self.global_vector_name, self.experimental_vector_name, self.full_vector_name],
                limits=[stage1_k, stage2_k, int(top_k)], queries=qs, pool_flags=[True, False, False]))
        # compact columnar results: the final points plus their stage-1 / stage-2 scores (looked up on the device)
        with no_gc():
            return [[{
                "id": pid,
                "score_stage1": s1,
                "score_stage2": s2,
                "score_stage3": score,
                "score_final": score,
                "payload": payload,
            } for pid, score, s1, s2, payload in zip(ids, scores, stage_scores[0], stage_scores[1], payloads)]
                for ids, scores, stage_scores, payloads in res]
