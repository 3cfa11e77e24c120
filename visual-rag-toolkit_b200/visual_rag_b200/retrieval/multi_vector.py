"""MultiVectorRetriever façade — mirror of visual_rag/retrieval/multi_vector.py for the embedded-query
path (`search_embedded`). Model inference is out of scope: `search(query: str)` needs an `embedder` object
with `embed_query`, exactly like the reference's constructor argument (multi_vector.py:48,108)."""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from .single_stage import SingleStageRetriever
from .three_stage import ThreeStageRetriever
from .two_stage import TwoStageRetriever


class MultiVectorRetriever:
    def __init__(
        self,
        collection_name: str,
        model_name: str = "vidore/colSmol-500M",
        qdrant_url: Optional[str] = None,
        qdrant_api_key: Optional[str] = None,
        prefer_grpc: bool = False,
        request_timeout: int = 120,
        max_retries: int = 3,
        retry_sleep: float = 0.5,
        qdrant_client=None,
        embedder=None,
        experimental_vector_name: str = "experimental_pooling",
    ):
        if qdrant_client is None:
            raise ValueError(
                "visual_rag_b200 serves a GPU-resident corpus: pass qdrant_client=GpuCorpusClient(corpus) "
                "(network Qdrant clients are outside this backend)."
            )
        self.client = qdrant_client
        self.collection_name = collection_name
        self.model_name = model_name
        self.embedder = embedder
        kw = dict(qdrant_client=qdrant_client, collection_name=collection_name,
                  experimental_vector_name=str(experimental_vector_name), request_timeout=request_timeout,
                  max_retries=max_retries, retry_sleep=retry_sleep)
        self._two_stage = TwoStageRetriever(**kw)
        self._three_stage = ThreeStageRetriever(**kw)
        self._single_stage = SingleStageRetriever(**kw)

    def build_filter(self, year=None, source=None, district=None, filename=None, has_text=None):
        return self._two_stage.build_filter(year=year, source=source, district=district, filename=filename,
                                            has_text=has_text)

    def search(self, query: str, top_k: int = 10, mode: str = "single_full", prefetch_k: Optional[int] = None,
               stage1_mode: str = "pooled_query_vs_standard_pooling", filter_obj=None,
               return_embeddings: bool = False) -> List[Dict[str, Any]]:
        """multi_vector.py:152-177."""
        if self.embedder is None:
            raise ValueError("search(query: str) needs an `embedder` with embed_query(); use search_embedded() "
                             "with a precomputed query embedding")
        q = self.embedder.embed_query(query)
        try:
            import torch

            if isinstance(q, torch.Tensor):
                q = q.detach().cpu().float().numpy()
        except ImportError:  # pragma: no cover
            pass
        return self.search_embedded(query_embedding=np.asarray(q, dtype=np.float32), top_k=top_k, mode=mode,
                                    prefetch_k=prefetch_k, stage1_mode=stage1_mode, filter_obj=filter_obj,
                                    return_embeddings=return_embeddings)

    def search_embedded(self, *, query_embedding, top_k: int = 10, mode: str = "single_full",
                        prefetch_k: Optional[int] = None, stage1_mode: str = "pooled_query_vs_standard_pooling",
                        stage1_k: Optional[int] = None, stage2_k: Optional[int] = None, filter_obj=None,
                        return_embeddings: bool = False) -> List[Dict[str, Any]]:
        """multi_vector.py:179-247."""
        single = {
            "single_full": "multi_vector",
            "single_tiles": "tiles_maxsim",
            "single_pooled": "pooled_tile",
            "single_global": "pooled_global",
            "single_experimental_tokens": "experimental_maxsim",
            "single_experimental_pooled": "pooled_experimental",
        }
        if mode in single:
            return self._single_stage.search(query_embedding=query_embedding, top_k=top_k, filter_obj=filter_obj,
                                             strategy=single[mode])
        if mode == "two_stage":
            return self._two_stage.search_server_side(query_embedding=query_embedding, top_k=top_k,
                                                      prefetch_k=prefetch_k, filter_obj=filter_obj,
                                                      stage1_mode=stage1_mode)
        if mode == "three_stage":
            return self._three_stage.search_server_side(query_embedding=query_embedding, top_k=top_k,
                                                        stage1_k=stage1_k, stage2_k=stage2_k, filter_obj=filter_obj,
                                                        stage1_mode=stage1_mode)
        raise ValueError(f"Unknown mode: {mode}")

    def search_embedded_batch(self, *, query_embeddings, top_k: int = 10, mode: str = "two_stage",
                              prefetch_k: Optional[int] = None, stage1_mode: str = "pooled_query_vs_standard_pooling",
                              stage1_k: Optional[int] = None, stage2_k: Optional[int] = None,
                              filter_obj=None) -> List[List[Dict[str, Any]]]:
        """search_embedded for a batch of queries; two_stage / three_stage run as one native call on the GPU
        backend, every other mode is the per-query loop."""
        if mode == "two_stage":
            return self._two_stage.search_server_side_batch(query_embeddings, top_k=top_k, prefetch_k=prefetch_k,
                                                            filter_obj=filter_obj, stage1_mode=stage1_mode)
        if mode == "three_stage":
            return self._three_stage.search_server_side_batch(query_embeddings=query_embeddings, top_k=top_k,
                                                              stage1_k=stage1_k, stage2_k=stage2_k, filter_obj=filter_obj)
        return [self.search_embedded(query_embedding=q, top_k=top_k, mode=mode, prefetch_k=prefetch_k,
                                     stage1_mode=stage1_mode, stage1_k=stage1_k, stage2_k=stage2_k,
                                     filter_obj=filter_obj) for q in query_embeddings]
