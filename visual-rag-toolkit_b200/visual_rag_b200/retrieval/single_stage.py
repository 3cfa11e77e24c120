"""SingleStageRetriever — mirror of visual_rag/retrieval/single_stage.py: one exhaustive scan of one named
vector store, six strategies."""

from __future__ import annotations

import logging
from typing import Any, Dict, List

import numpy as np

from ._common import to_numpy, wire

logger = logging.getLogger(__name__)


class SingleStageRetriever:
    def __init__(
        self,
        qdrant_client,
        collection_name: str,
        experimental_vector_name: str = "experimental_pooling",
        request_timeout: int = 120,
        max_retries: int = 3,
        retry_sleep: float = 1.0,
    ):
        self.client = qdrant_client
        self.collection_name = collection_name
        self.experimental_vector_name = str(experimental_vector_name)
        self.request_timeout = int(request_timeout)
        self.max_retries = max_retries
        self.retry_sleep = retry_sleep

    def _to_numpy(self, embedding) -> np.ndarray:
        return to_numpy(embedding)

    def search(self, query_embedding, top_k: int = 10, strategy: str = "multi_vector",
               filter_obj=None) -> List[Dict[str, Any]]:
        """single_stage.py:60-142: (query tokens | mean-pooled query) x (initial | mean_pooling | experimental |
        global_pooling)."""
        query_np = self._to_numpy(query_embedding)
        table = {
            "multi_vector": ("initial", False),
            "tiles_maxsim": ("mean_pooling", False),
            "pooled_tile": ("mean_pooling", True),
            "pooled_global": ("global_pooling", True),
            "experimental_maxsim": (self.experimental_vector_name, False),
            "pooled_experimental": (self.experimental_vector_name, True),
        }
        if strategy not in table:
            raise ValueError(f"Unknown strategy: {strategy}")
        vector_name, pool = table[strategy]
        query_vector = wire(self.client, query_np.mean(axis=0)) if pool else wire(self.client, query_np)
        results = self.client.query_points(
            collection_name=self.collection_name, query=query_vector, using=vector_name, query_filter=filter_obj,
            limit=top_k, with_payload=True, with_vectors=False, timeout=self.request_timeout,
        ).points
        return [{"id": r.id, "score": r.score, "score_final": r.score, "payload": r.payload} for r in results]
