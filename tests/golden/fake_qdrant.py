"""Test doubles for the `qdrant_client` seam of the reference retrievers (test infrastructure).

`install_qdrant_stub()` registers minimal `qdrant_client{,.http,.http.models,.models}` modules so that
`visual_rag.retrieval` of the reference imports without the real client.
`NumpyQdrant` is an in-memory client that answers query_points / retrieve with the cosine + MAX_SIM
semantics restated in SURVEY.md §3.2, scoring with a caller-supplied maxsim function (the reference's
own compute_maxsim_score when generating goldens).
"""

from __future__ import annotations

import sys
import types
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np


class _Kw:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def install_qdrant_stub() -> None:
    if "qdrant_client" in sys.modules and not getattr(sys.modules["qdrant_client"], "_vrag_stub", False):
        return
    root = types.ModuleType("qdrant_client")
    root._vrag_stub = True
    http = types.ModuleType("qdrant_client.http")
    models = types.ModuleType("qdrant_client.http.models")
    models2 = types.ModuleType("qdrant_client.models")
    for name in ("SearchParams", "Prefetch", "Filter", "FieldCondition", "MatchAny", "MatchValue", "HasIdCondition",
                 "Distance", "VectorParams", "PointStruct", "MultiVectorConfig", "MultiVectorComparator"):
        cls = type(name, (_Kw,), {})
        setattr(models, name, cls)
        setattr(models2, name, cls)
    root.http = http
    http.models = models
    root.models = models2
    root.QdrantClient = type("QdrantClient", (), {})
    sys.modules["qdrant_client"] = root
    sys.modules["qdrant_client.http"] = http
    sys.modules["qdrant_client.http.models"] = models
    sys.modules["qdrant_client.models"] = models2


class _Point:
    def __init__(self, id, score=None, payload=None, vector=None):
        self.id = id
        self.score = score
        self.payload = payload
        self.vector = vector


class _Resp:
    def __init__(self, points):
        self.points = points


def _has_ids(query_filter) -> Optional[set]:
    """Extract the HasIdCondition id set of a (possibly nested) Filter, three_stage.py:75-81."""
    if query_filter is None:
        return None
    ids = None
    for cond in getattr(query_filter, "must", None) or []:
        if hasattr(cond, "has_id"):
            s = set(cond.has_id)
            ids = s if ids is None else ids & s
        elif hasattr(cond, "must"):
            sub = _has_ids(cond)
            if sub is not None:
                ids = sub if ids is None else ids & sub
    return ids


class NumpyQdrant:
    """vectors: {name: list of [rows,128] float arrays (or [128] for dense)}, one entry per page; ids = 0..n-1."""

    def __init__(self, vectors: Dict[str, Sequence[np.ndarray]], maxsim: Callable[[np.ndarray, np.ndarray], float],
                 payloads: Optional[List[dict]] = None):
        self.vectors = vectors
        self.maxsim = maxsim
        n = len(next(iter(vectors.values())))
        self.payloads = payloads or [{"page": i} for i in range(n)]
        self.calls: List[Dict[str, Any]] = []

    def _score_all(self, query, using, ids):
        q = np.array(query, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        out = []
        for i in ids:
            d = np.array(self.vectors[using][i], dtype=np.float32)
            if d.ndim == 1:
                d = d[None, :]
            out.append(self.maxsim(q, d))
        return out

    def query_points(self, collection_name, query, using, limit, query_filter=None, with_payload=True,
                     with_vectors=False, search_params=None, prefetch=None, timeout=None):
        self.calls.append({"using": using, "limit": limit, "prefetch": prefetch is not None})
        n = len(self.vectors[using])
        ids = list(range(n))
        allowed = _has_ids(query_filter)
        if allowed is not None:
            ids = [i for i in ids if i in allowed]
        if prefetch:
            pf = prefetch[0]
            s1 = self._score_all(pf.query, pf.using, ids)
            order = sorted(range(len(ids)), key=lambda j: s1[j], reverse=True)[: pf.limit]
            ids = [ids[j] for j in order]
        sc = self._score_all(query, using, ids)
        order = sorted(range(len(ids)), key=lambda j: sc[j], reverse=True)[:limit]
        return _Resp([_Point(ids[j], sc[j], self.payloads[ids[j]] if with_payload else None) for j in order])

    def retrieve(self, collection_name, ids, with_payload=False, with_vectors=None, timeout=None):
        names = with_vectors or []
        return [_Point(i, vector={nm: np.asarray(self.vectors[nm][i], dtype=np.float32).tolist() for nm in names})
                for i in ids]

    def get_collection(self, name):
        return _Kw(config=None)
