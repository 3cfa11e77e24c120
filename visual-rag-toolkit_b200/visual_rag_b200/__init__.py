"""visual_rag_b200 — B200-native drop-in for the retrieval-scoring and pooling path of visual-rag-toolkit.

Mirrors the reference package layout for this path only:
    visual_rag_b200.embedding.pooling   <- visual_rag/embedding/pooling.py
    visual_rag_b200.retrieval           <- visual_rag/retrieval/{two_stage,three_stage,single_stage,multi_vector}.py
    visual_rag_b200.corpus              <- the Qdrant collection (GPU-resident store + duck-typed client)
All arithmetic runs in hand-written sm_100a CUDA kernels behind libvrag_b200.so (include/vrag_b200.h).
"""

__version__ = "0.1.0"

from . import _native  # noqa: F401
from .corpus import GpuCorpus  # noqa: F401
