"""`GpuCorpusClient`: the duck-typed `qdrant_client` of the reference retrievers, backed by a GpuCorpus.

The reference retrievers only ever call `query_points`, `retrieve` and `get_collection` on their client
(two_stage.py:162-178,307-316,349-358,384-390; three_stage.py:103-157;
single_stage.py:123-132).  This class implements exactly those with Qdrant's COSINE + MAX_SIM semantics
(qdrant_indexer.py:200-239) computed by the sm_100a kernels, so the reference's own retriever classes — and
the mirrors in visual_rag_b200.retrieval — run unchanged on top of the GPU store.
"""

from __future__ import annotations

import functools
import threading
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from .corpus import GpuCorpus


class ScoredPoint:
    __slots__ = ("id", "score", "payload", "vector", "version")

    def __init__(self, id, score=None, payload=None, vector=None):
        self.id = id
        self.score = score
        self.payload = payload
        self.vector = vector
        self.version = 0

    def __repr__(self):  # pragma: no cover
        return f"ScoredPoint(id={self.id!r}, score={self.score!r})"


class QueryResponse:
    __slots__ = ("points",)

    def __init__(self, points):
        self.points = points


class _VectorInfo:
    def __init__(self, multivector: bool):
        self.multivector_config = object() if multivector else None
        self.size = 128


class _CollectionInfo:
    def __init__(self, vectors: Dict[str, _VectorInfo], points_count: int):
        params = type("Params", (), {"vectors": vectors})()
        self.config = type("Config", (), {"params": params})()
        self.points_count = points_count


def _match(cond, payload: dict) -> bool:
    """FieldCondition(key, match=MatchValue(value)|MatchAny(any)) on a payload dict (two_stage.py:449-480)."""
    val = payload.get(getattr(cond, "key", None)) if payload else None
    m = getattr(cond, "match", None)
    if m is None:
        return True
    if hasattr(m, "any") and getattr(m, "any") is not None:
        return val in list(m.any)
    if hasattr(m, "value"):
        return val == m.value
    return True


def _locked(fn):
    """Client calls and ingest (GpuIndexer.upload_batch) serialise on one lock: a batch upload appends to several named
    stores and then registers its ids, and a search must not observe the collection in between."""

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self._lock:
            return fn(self, *args, **kwargs)

    return wrapper


class GpuCorpusClient:
    """In-process, GPU-resident replacement of QdrantClient for the retrieval path.

    point_ids: external id of page i (any hashable: int or UUID string as produced by
    QdrantIndexer.generate_point_id, qdrant_indexer.py:602-613). Defaults to the global page index.
    """

    accepts_numpy = True   # queries may arrive as fp32 numpy arrays (no list round trip), see retrieval/_common.py::wire

    def __init__(self, corpus: GpuCorpus, collection_name: str = "gpu", point_ids: Optional[Sequence[Any]] = None,
                 payloads: Optional[Sequence[Optional[dict]]] = None):
        self.corpus = corpus
        self.collection_name = collection_name
        self._ids: Optional[List[Any]] = list(point_ids) if point_ids is not None else None
        self._index: Optional[Dict[Any, int]] = (
            {pid: i for i, pid in enumerate(self._ids)} if self._ids is not None else None
        )
        self._payloads = list(payloads) if payloads is not None else None
        self._lock = threading.RLock()
        self._columns: Dict[str, np.ndarray] = {}       # payload key -> per-page value column (built on first use)
        self._filter_cache: Dict[Any, np.ndarray] = {}  # filter signature -> candidate page ids

    # ------------------------------------------------------------------ id mapping
    @_locked
    def set_points(self, point_ids: Sequence[Any], payloads: Optional[Sequence[Optional[dict]]] = None) -> None:
        self._ids = list(point_ids)
        self._index = {pid: i for i, pid in enumerate(self._ids)}
        self._payloads = list(payloads) if payloads is not None else None
        self._columns = {}
        self._filter_cache = {}

    @_locked
    def append_points(self, point_ids: Sequence[Any], payloads: Optional[Sequence[Optional[dict]]] = None) -> None:
        """Register the pages a batch upload appended (incremental: O(batch), not O(collection))."""
        if self._ids is None:
            self._ids, self._index = [], {}
        if self._payloads is None:
            self._payloads = [None] * len(self._ids)
        n0 = len(self._ids)
        self._ids.extend(point_ids)
        for i, pid in enumerate(point_ids):
            self._index[pid] = n0 + i
        self._payloads.extend(payloads if payloads is not None else [None] * len(point_ids))
        self._columns = {}
        self._filter_cache = {}

    @_locked
    def set_payload(self, point_id, payload: Optional[dict]) -> None:
        """Replace the payload of an existing point (upsert of an id that is already in the collection)."""
        page = self._page(point_id)
        if page < 0:
            raise KeyError(point_id)
        if self._payloads is None:
            self._payloads = [None] * len(self._ids)
        self._payloads[page - self.corpus.page_base] = payload
        self._columns = {}
        self._filter_cache = {}

    def _pid(self, page: int):
        local = page - self.corpus.page_base
        return self._ids[local] if self._ids is not None else page

    def _page(self, pid) -> int:
        if self._index is not None:
            i = self._index.get(pid)
            if i is None and not isinstance(pid, str):
                i = self._index.get(str(pid))
            return -1 if i is None else i + self.corpus.page_base
        try:
            return int(pid)
        except (TypeError, ValueError):
            return -1

    def _payload(self, page: int):
        if self._payloads is None:
            return {}
        return self._payloads[page - self.corpus.page_base]

    # ------------------------------------------------------------------ filters
    def _candidates(self, query_filter, n_pages: int) -> Optional[np.ndarray]:
        """Filter -> sorted global page ids, or None for 'all pages'. HasIdCondition is the candidate
        restriction of three_stage.py:75-81; FieldConditions (build_filter, two_stage.py:436-480) are
        evaluated on the host payloads."""
        if query_filter is None:
            return None
        allowed: Optional[set] = None
        field_conds = []

        def walk(f):
            nonlocal allowed
            for cond in (getattr(f, "must", None) or []):
                if hasattr(cond, "has_id"):
                    s = {self._page(p) for p in cond.has_id}
                    s.discard(-1)
                    allowed = s if allowed is None else (allowed & s)
                elif hasattr(cond, "must") or hasattr(cond, "should") or hasattr(cond, "must_not"):
                    walk(cond)
                elif hasattr(cond, "key"):
                    field_conds.append(cond)

        walk(query_filter)
        if allowed is None and not field_conds:
            return None
        base = self.corpus.page_base
        # payload conditions are evaluated column-wise (one numpy pass per key) and the resulting page list is cached
        # per filter, so a benchmark that reuses one filter (run_qdrant_beir.py:1987-1997) pays for it once
        sig = None
        if field_conds:
            try:
                sig = (n_pages, tuple(sorted(self._cond_signature(c) for c in field_conds)),
                       None if allowed is None else frozenset(allowed))
                hit = self._filter_cache.get(sig)
                if hit is not None:
                    return hit
            except TypeError:   # unhashable match values: evaluate without caching
                sig = None
            mask = np.ones((n_pages,), dtype=bool)
            for cond in field_conds:
                mask &= self._cond_mask(cond, n_pages)
            if allowed is not None:
                sel = np.zeros((n_pages,), dtype=bool)
                idx = np.fromiter((p - base for p in allowed if 0 <= p - base < n_pages), dtype=np.int64)
                sel[idx] = True
                mask &= sel
            pages_arr = np.nonzero(mask)[0].astype(np.int64) + base
            if sig is not None:
                if len(self._filter_cache) > 64:
                    self._filter_cache.clear()
                self._filter_cache[sig] = pages_arr
            return pages_arr
        return np.asarray(sorted(allowed), dtype=np.int64)

    @staticmethod
    def _cond_signature(cond):
        m = getattr(cond, "match", None)
        if m is None:
            return (str(getattr(cond, "key", None)), "none", ())
        if hasattr(m, "any") and getattr(m, "any") is not None:
            return (str(cond.key), "any", tuple(sorted(map(repr, m.any))))
        return (str(cond.key), "value", (repr(getattr(m, "value", None)),))

    def _column(self, key: str, n_pages: int) -> np.ndarray:
        col = self._columns.get(key)
        if col is None or col.shape[0] != n_pages:
            col = np.empty((n_pages,), dtype=object)
            pl = self._payloads
            for i in range(n_pages):
                d = pl[i] if (pl is not None and i < len(pl)) else None
                col[i] = d.get(key) if d else None
            self._columns[key] = col
        return col

    def _cond_mask(self, cond, n_pages: int) -> np.ndarray:
        """FieldCondition(key, match=MatchValue|MatchAny) over all pages (two_stage.py:449-480)."""
        m = getattr(cond, "match", None)
        if m is None:
            return np.ones((n_pages,), dtype=bool)
        col = self._column(getattr(cond, "key", None), n_pages)
        if hasattr(m, "any") and getattr(m, "any") is not None:
            vals = list(m.any)
            out = np.zeros((n_pages,), dtype=bool)
            for v in vals:
                out |= (col == v)
            return out
        if hasattr(m, "value"):
            return np.asarray(col == m.value, dtype=bool)
        return np.ones((n_pages,), dtype=bool)

    # ------------------------------------------------------------------ the three client methods
    @staticmethod
    def _as_query(query) -> np.ndarray:
        q = np.asarray(query, dtype=np.float32)
        return q[None, :] if q.ndim == 1 else q

    @_locked
    def query_points(self, collection_name=None, query=None, using=None, limit=10, query_filter=None,
                     with_payload=True, with_vectors=False, search_params=None, prefetch=None, timeout=None,
                     **_ignored) -> QueryResponse:
        if using is None:
            raise ValueError("`using` (named vector) is required")
        limit = int(limit)
        n_pages = self.corpus.n_pages(using)
        cand = self._candidates(query_filter, n_pages)
        q = self._as_query(query)
        if prefetch:
            # Qdrant prefetch= (two_stage.py:170-176): stage-1 query and rerank query travel separately;
            # both stages run back to back on the device with a single host synchronisation.
            pf = prefetch[0] if isinstance(prefetch, (list, tuple)) else prefetch
            stages = self.corpus.search_multistage(
                [(pf.using, False, int(pf.limit)), (using, False, limit)], None,
                stage_queries=[self._as_query(pf.query), q], candidate_ids=cand)
            scores, ids = stages[-1]
        else:
            scores, ids = self.corpus.search(using, q, limit, candidate_ids=cand)
        keep = np.isfinite(scores)
        points = [
            ScoredPoint(self._pid(int(i)), float(s), self._payload(int(i)) if with_payload else None)
            for s, i in zip(scores[keep], ids[keep])
        ]
        if with_vectors:
            names = [using] if with_vectors is True else list(with_vectors)
            for p in points:
                page = self._page(p.id) - self.corpus.page_base
                p.vector = {nm: self.corpus.read_page(nm, page).astype(np.float32).tolist() for nm in names}
        return QueryResponse(points)

    @_locked
    def query_three_stage(self, *, stage1_query, stage2_query, stage3_query, stage1_using, stage2_using,
                          stage3_using, stage1_k: int, stage2_k: int, top_k: int, query_filter=None):
        """The three ID-restricted scans of ThreeStageRetriever.search_server_side (three_stage.py:102-159)
        fused on the device: returns the three point lists (stage 3 with payloads)."""
        cand = self._candidates(query_filter, self.corpus.n_pages(stage1_using))
        stages = self.corpus.search_multistage(
            [(stage1_using, False, int(stage1_k)), (stage2_using, False, int(stage2_k)), (stage3_using, False, int(top_k))],
            None, stage_queries=[self._as_query(stage1_query), self._as_query(stage2_query), self._as_query(stage3_query)],
            candidate_ids=cand)
        out = []
        for si, (scores, ids) in enumerate(stages):
            keep = np.isfinite(scores)
            out.append([ScoredPoint(self._pid(int(i)), float(s), self._payload(int(i)) if si == 2 else None)
                        for s, i in zip(scores[keep], ids[keep])])
        return out

    @_locked
    def query_multistage_batch(self, *, usings: Sequence[str], limits: Sequence[int],
                               stage_queries: Sequence[Sequence[Any]], with_payload: bool = True):
        """A batch of independent multi-stage searches in ONE native call (BASELINE configs[2]: 256 queries):
        stage s scans named vector usings[s] with that query's stage-s matrix, restricted to the survivors of
        stage s-1, and keeps limits[s] points. stage_queries[b][s] is what the reference would send as `query=`
        of stage s for query b (the mean-pooled vector or the token matrix, two_stage.py:142-159,
        three_stage.py:96-100). Returns per query the per-stage point lists (payloads on the last stage only)."""
        ns = len(usings)
        if len(limits) != ns:
            raise ValueError("usings and limits must have the same length")
        sq = [[self._as_query(x) for x in per_query] for per_query in stage_queries]
        res = self.corpus.search_multistage_batch([(usings[s], False, int(limits[s])) for s in range(ns)], None,
                                                  stage_queries=sq)
        out = []
        for per_query in res:
            stages = []
            for si, (scores, ids) in enumerate(per_query):
                keep = np.isfinite(scores)
                last = si == ns - 1
                stages.append([ScoredPoint(self._pid(int(i)), float(s), self._payload(int(i)) if (last and with_payload) else None)
                               for s, i in zip(scores[keep], ids[keep])])
            out.append(stages)
        return out

    @_locked
    def query_multistage_batch_final(self, *, usings: Sequence[str], limits: Sequence[int],
                                     stage_queries: Optional[Sequence[Sequence[Any]]] = None,
                                     queries: Optional[Sequence[Any]] = None,
                                     pool_flags: Optional[Sequence[bool]] = None, with_payload: bool = True):
        """`query_multistage_batch` with compact, columnar results: per query four parallel lists
        `(ids, scores, stage_scores, payloads)` for the points of the LAST stage — `stage_scores[s][j]` is the score
        point j had in stage s < last (None where absent). The long intermediate lists stay on the device and
        no per-point Python object is built here (a retriever turns the columns into its result dicts with one zip).
        Either `stage_queries[b][s]` (what the reference would send as `query=` of stage s for query b) or
        `queries[b]` (the token matrix) plus `pool_flags[s]` (stage s scans with the mean-pooled query,
        two_stage.py:142 / three_stage.py:96, pooled on the device) describes the batch."""
        ns = len(usings)
        if len(limits) != ns:
            raise ValueError("usings and limits must have the same length")
        if stage_queries is not None:
            sq = [[self._as_query(x) for x in per_query] for per_query in stage_queries]
            sc, ids, st, cnt = self.corpus.search_multistage_batch(
                [(usings[s], False, int(limits[s])) for s in range(ns)], None, stage_queries=sq, final_only=True)
        else:
            if queries is None:
                raise ValueError("either stage_queries or queries is required")
            pf = [False] * ns if pool_flags is None else [bool(x) for x in pool_flags]
            if len(pf) != ns:
                raise ValueError("pool_flags must have one entry per stage")
            sc, ids, st, cnt = self.corpus.search_multistage_batch(
                [(usings[s], pf[s], int(limits[s])) for s in range(ns)], queries, final_only=True)
        nq, kl = sc.shape
        valid = np.isfinite(sc) & (np.arange(kl, dtype=np.int64)[None, :] < np.asarray(cnt, dtype=np.int64)[:, None])
        row_full = valid.all(axis=1).tolist()
        # plain Python lists once (numpy scalar access per element would dominate the whole call); stage scores as one
        # column per stage
        st_cols = []
        for s_i in range(ns - 1):
            col = st[:, :, s_i]
            if np.isnan(col).any():
                col_o = col.astype(object)
                col_o[np.isnan(col)] = None
                st_cols.append(col_o.tolist())
            else:
                st_cols.append(col.tolist())
        sc_l, ids_l = sc.tolist(), ids.tolist()
        base, ext, pls = self.corpus.page_base, self._ids, self._payloads
        out = []
        for b in range(nq):
            pages, scores, stages = ids_l[b], sc_l[b], [col[b] for col in st_cols]
            if not row_full[b]:
                keep = np.nonzero(valid[b])[0].tolist()
                pages = [pages[j] for j in keep]
                scores = [scores[j] for j in keep]
                stages = [[col[j] for j in keep] for col in stages]
            pids = pages if ext is None else [ext[p - base] for p in pages]
            if not with_payload:
                payloads = [None] * len(pages)
            elif pls is None:
                payloads = [{} for _ in pages]
            else:
                payloads = [pls[p - base] for p in pages]
            out.append((pids, scores, stages, payloads))
        return out

    @_locked
    def retrieve(self, collection_name=None, ids=(), with_payload=False, with_vectors=None, timeout=None, **_ignored):
        out = []
        names = [] if not with_vectors else (list(with_vectors) if not isinstance(with_vectors, bool) else [])
        for pid in ids:
            page = self._page(pid)
            if page < 0:
                continue
            local = page - self.corpus.page_base
            vec = {nm: self.corpus.read_page(nm, local).astype(np.float32).tolist() for nm in names
                   if self.corpus.has_store(nm)}
            out.append(ScoredPoint(pid, None, self._payload(page) if with_payload else None, vec or None))
        return out

    @_locked
    def get_collection(self, collection_name=None):
        vectors = {}
        count = 0
        for nm in ("initial", "mean_pooling", "experimental_pooling", "global_pooling"):
            if self.corpus.has_store(nm):
                info = self.corpus.store_info(nm)
                vectors[nm] = _VectorInfo(multivector=(nm != "global_pooling"))
                count = info["n_pages"]
        return _CollectionInfo(vectors, count)
