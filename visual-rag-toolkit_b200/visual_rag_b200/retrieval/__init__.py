"""Retriever classes of the reference (visual_rag/retrieval/__init__.py:9-12) on the GPU-resident corpus."""

from .multi_vector import MultiVectorRetriever
from .single_stage import SingleStageRetriever
from .three_stage import ThreeStageRetriever
from .two_stage import TwoStageRetriever

__all__ = ["MultiVectorRetriever", "SingleStageRetriever", "ThreeStageRetriever", "TwoStageRetriever"]
