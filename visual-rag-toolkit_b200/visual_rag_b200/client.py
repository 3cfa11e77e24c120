"""`GpuCorpusClient` / `ShardedCorpusClient`: the duck-typed `qdrant_client` of the reference retrievers, backed by a
GPU-resident corpus (one GPU, or page-sharded over the GPUs of a box).

The reference retrievers only ever call `query_points`, `retrieve` and `get_collection` on their client
(two_stage.py:162-178,307-316,349-358,384-390; three_stage.py:103-157;
single_stage.py:123-132).  These classes implement exactly those with Qdrant's COSINE + MAX_SIM semantics
(qdrant_indexer.py:200-239) computed by the sm_100a kernels, so the reference's own retriever classes — and
the mirrors in visual_rag_b200.retrieval — run unchanged on top of the GPU store.
"""

from __future__ import annotations

import bisect
import functools
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

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


def _locked(fn):
    """Client calls and ingest (GpuIndexer.upload_batch) serialise on one lock: a batch upload appends to several named
    stores and then registers its ids, and a search must not observe the collection in between."""

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self._lock:
            return fn(self, *args, **kwargs)

    return wrapper


# ---------------------------------------------------------------------------------------------------- filters
def _is_filter(obj) -> bool:
    return any(getattr(obj, a, None) is not None for a in ("must", "should", "must_not")) and not hasattr(obj, "key") \
        and not hasattr(obj, "has_id")


def _as_list(x) -> list:
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _cond_signature(cond) -> tuple:
    """Hashable structural signature of a filter / condition (cache key of its page mask). TypeError for unhashable
    match values."""
    if _is_filter(cond):
        return ("filter",
                tuple(_cond_signature(c) for c in _as_list(getattr(cond, "must", None))),
                tuple(_cond_signature(c) for c in _as_list(getattr(cond, "should", None))),
                tuple(_cond_signature(c) for c in _as_list(getattr(cond, "must_not", None))))
    if hasattr(cond, "has_id"):
        return ("has_id", frozenset(cond.has_id))
    key = str(getattr(cond, "key", None))
    m = getattr(cond, "match", None)
    rng = getattr(cond, "range", None)
    sig: tuple = ("field", key)
    if m is not None:
        for attr in ("value", "any", "except_"):
            v = getattr(m, attr, None)
            if v is not None:
                sig += (attr, tuple(sorted(map(repr, v))) if attr != "value" else repr(v))
    if rng is not None:
        sig += ("range",) + tuple(repr(getattr(rng, a, None)) for a in ("gt", "gte", "lt", "lte"))
    return sig


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
        self._filter_cache: Dict[Any, Any] = {}         # filter signature -> {mask, pages, fid (device bitmask, or None)}
        self._device_masks: List[dict] = []             # cache entries that currently own a device bitmask
        self._last_mask = None
        self._mask_entry = None

    # ------------------------------------------------------------------ id mapping (pages of THIS corpus handle)
    def _invalidate(self) -> None:
        self._columns = {}
        self._drop_device_masks()
        self._filter_cache = {}

    def _drop_device_masks(self) -> None:
        for entry in getattr(self, "_device_masks", []):
            if entry.get("fid") is not None:
                try:
                    self.corpus.destroy_filter(entry["fid"])
                except Exception:  # noqa: BLE001 - the handle may already be closed
                    pass
                entry["fid"] = None
        self._device_masks = []

    @_locked
    def set_points(self, point_ids: Sequence[Any], payloads: Optional[Sequence[Optional[dict]]] = None) -> None:
        self._ids = list(point_ids)
        self._index = {pid: i for i, pid in enumerate(self._ids)}
        self._payloads = list(payloads) if payloads is not None else None
        self._invalidate()

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
        self._invalidate()

    @_locked
    def set_payload(self, point_id, payload: Optional[dict]) -> None:
        """Replace the payload of an existing point of this handle (upsert of an id that is already in the collection)."""
        local = self.local_page(point_id)
        if local < 0:
            raise KeyError(point_id)
        if self._payloads is None:
            self._payloads = [None] * len(self._ids)
        self._payloads[local] = payload
        self._invalidate()

    @_locked
    def remove_points(self, point_ids: Sequence[Any]) -> None:
        """Forget deleted points: their pages stay (empty) in the stores, their ids leave the index."""
        if self._ids is None:       # implicit ids (page index): make them explicit first
            n = max([self.corpus.n_pages(nm) for nm in ("initial", "mean_pooling", "global_pooling") if self.corpus.has_store(nm)] or [0])
            self._ids = list(range(self.corpus.page_base, self.corpus.page_base + n))
            self._index = {pid: i for i, pid in enumerate(self._ids)}
        for pid in point_ids:
            i = self._index.pop(pid, None)
            if i is not None:
                self._ids[i] = None
                if self._payloads is not None:
                    self._payloads[i] = None
        self._invalidate()

    def _pid(self, page: int):
        local = page - self.corpus.page_base
        return self._ids[local] if self._ids is not None else page

    def _page(self, pid) -> int:
        """external point id -> global page id (-1 if unknown)."""
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

    def local_page(self, pid) -> int:
        """external point id -> index of its page in THIS corpus handle (-1 if this handle does not hold it). What the
        ingest path (GpuIndexer) uses: on a sharded client it never sees the points of other ranks."""
        if self._index is not None:
            i = self._index.get(pid)
            if i is None and not isinstance(pid, str):
                i = self._index.get(str(pid))
            return -1 if i is None else int(i)
        try:
            i = int(pid) - self.corpus.page_base
        except (TypeError, ValueError):
            return -1
        n = max([self.corpus.n_pages(nm) for nm in ("initial", "mean_pooling", "global_pooling") if self.corpus.has_store(nm)] or [0])
        return i if 0 <= i < n else -1

    def _pids_of(self, pages: List[int]) -> list:
        ext, base = self._ids, self.corpus.page_base
        return pages if ext is None else [ext[p - base] for p in pages]

    def _payloads_of(self, pages: List[int]) -> list:
        pls, base = self._payloads, self.corpus.page_base
        return [{} for _ in pages] if pls is None else [pls[p - base] for p in pages]

    def _read_pages(self, name: str, pages: Sequence[int]) -> Dict[int, np.ndarray]:
        """fp16 rows of the listed global pages (what `with_vectors=[name]` returns)."""
        base = self.corpus.page_base
        return {int(p): self.corpus.read_page(name, int(p) - base) for p in pages}

    def _n_points_total(self, n_local: int) -> int:
        return n_local

    # ------------------------------------------------------------------ filters
    # A payload filter that lets more than this many pages through (and more than 1/64 of the store) runs as a page bitmask
    # inside the scan (uploaded once per filter); below it the candidate-list gather reads less.
    MASK_MIN_PAGES = 4096

    def _restriction(self, query_filter, n_pages: int) -> dict:
        """Filter -> keyword arguments for the corpus search: {} (no restriction), {"candidate_ids": ids} or
        {"filter_id": id of a device-resident page bitmask}."""
        if query_filter is None:
            return {}
        self._last_mask = None
        cand = self._candidates(query_filter, n_pages)
        if cand is None:
            return {}
        hit = self._last_mask     # set by _candidates when the result is exactly a cached payload mask (no id restriction)
        if hit is not None and len(cand) > max(self.MASK_MIN_PAGES, n_pages // 64) and hasattr(self.corpus, "create_filter"):
            entry = hit
            if entry.get("fid") is None:
                if len(self._device_masks) >= 8:           # keep a handful of device masks alive
                    old = self._device_masks.pop(0)
                    if old.get("fid") is not None:
                        self.corpus.destroy_filter(old["fid"])
                        old["fid"] = None
                entry["fid"] = self.corpus.create_filter(entry["mask"])
                self._device_masks.append(entry)
            return {"filter_id": entry["fid"]}
        return {"candidate_ids": cand}

    def _candidates(self, query_filter, n_pages: int) -> Optional[np.ndarray]:
        """Filter -> ascending global page ids of THIS handle's pages that pass it, or None for 'all pages'.
        `must` is a conjunction, `should` a disjunction (at least one), `must_not` a negated disjunction — Qdrant's
        filter semantics; conditions: HasIdCondition (the candidate restriction of three_stage.py:75-81 and the rerank
        list of two_stage.py:380-390), FieldCondition with MatchValue / MatchAny / MatchExcept / Range (build_filter,
        two_stage.py:436-480; the per_dataset scope filter, run_qdrant_beir.py:1987-1997) and nested Filters. Anything
        else raises NotImplementedError — a clause is never silently dropped. Payload clauses are evaluated column-wise
        on the host payloads (one numpy pass per key) and the resulting page mask is cached per filter."""
        if query_filter is None:
            return None
        base = self.corpus.page_base
        # split the top-level conjunction into id restrictions (handled as small sorted sets: O(k), never O(n_pages))
        # and the payload part (a cached page mask)
        id_sets: List[set] = []
        rest_must: list = []

        def split(f):
            for cond in _as_list(getattr(f, "must", None)):
                if hasattr(cond, "has_id"):
                    s = {self._page(p) for p in cond.has_id}
                    s.discard(-1)
                    id_sets.append(s)
                elif _is_filter(cond) and getattr(cond, "should", None) is None and getattr(cond, "must_not", None) is None:
                    split(cond)
                else:
                    rest_must.append(cond)

        split(query_filter)
        should = _as_list(getattr(query_filter, "should", None))
        must_not = _as_list(getattr(query_filter, "must_not", None))
        has_rest = bool(rest_must or should or must_not)
        allowed: Optional[set] = None
        for s in id_sets:
            allowed = s if allowed is None else (allowed & s)
        if not has_rest:
            if allowed is None:
                return None
            return np.asarray(sorted(p for p in allowed if 0 <= p - base < n_pages), dtype=np.int64)
        mask, pages = self._payload_mask(rest_must, should, must_not, n_pages)
        if allowed is None:
            self._last_mask = self._mask_entry     # the cache entry of this payload mask (may carry a device filter id)
            return pages
        return np.asarray(sorted(p for p in allowed if 0 <= p - base < n_pages and mask[p - base]), dtype=np.int64)

    def _payload_mask(self, must, should, must_not, n_pages: int) -> Tuple[np.ndarray, np.ndarray]:
        """(bool mask over this handle's pages, ascending global ids of the pages that pass), cached per filter — a
        benchmark that reuses one filter (run_qdrant_beir.py:1987-1997) pays for it once."""
        sig = None
        try:
            sig = (n_pages, tuple(_cond_signature(c) for c in must), tuple(_cond_signature(c) for c in should),
                   tuple(_cond_signature(c) for c in must_not))
            hit = self._filter_cache.get(sig)
            if hit is not None:
                self._mask_entry = hit
                return hit["mask"], hit["pages"]
        except TypeError:   # unhashable match values: evaluate without caching
            sig = None
        mask = np.ones((n_pages,), dtype=bool)
        for cond in must:
            mask &= self._cond_mask(cond, n_pages)
        if should:
            any_of = np.zeros((n_pages,), dtype=bool)
            for cond in should:
                any_of |= self._cond_mask(cond, n_pages)
            mask &= any_of
        for cond in must_not:
            mask &= ~self._cond_mask(cond, n_pages)
        entry = {"mask": mask, "pages": np.nonzero(mask)[0].astype(np.int64) + self.corpus.page_base, "fid": None}
        self._mask_entry = entry
        if sig is not None:
            if len(self._filter_cache) > 64:
                self._drop_device_masks()
                self._filter_cache.clear()
            self._filter_cache[sig] = entry
        return entry["mask"], entry["pages"]

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
        """One condition over all pages of this handle -> bool mask."""
        if _is_filter(cond):
            m = np.ones((n_pages,), dtype=bool)
            for c in _as_list(getattr(cond, "must", None)):
                m &= self._cond_mask(c, n_pages)
            sh = _as_list(getattr(cond, "should", None))
            if sh:
                any_of = np.zeros((n_pages,), dtype=bool)
                for c in sh:
                    any_of |= self._cond_mask(c, n_pages)
                m &= any_of
            for c in _as_list(getattr(cond, "must_not", None)):
                m &= ~self._cond_mask(c, n_pages)
            return m
        if hasattr(cond, "has_id"):
            m = np.zeros((n_pages,), dtype=bool)
            base = self.corpus.page_base
            idx = [self._page(p) - base for p in cond.has_id]
            idx = [i for i in idx if 0 <= i < n_pages]
            m[idx] = True
            return m
        if not hasattr(cond, "key"):
            raise NotImplementedError(f"unsupported filter condition {type(cond).__name__}")
        col = self._column(getattr(cond, "key"), n_pages)
        match, rng = getattr(cond, "match", None), getattr(cond, "range", None)
        if match is None and rng is None:
            raise NotImplementedError(f"FieldCondition on '{cond.key}' has neither match nor range "
                                      "(only MatchValue / MatchAny / MatchExcept / Range are supported)")
        out = np.ones((n_pages,), dtype=bool)
        if match is not None:
            if getattr(match, "any", None) is not None:
                hit = np.zeros((n_pages,), dtype=bool)
                for v in list(match.any):
                    hit |= np.asarray(col == v, dtype=bool)
                out &= hit
            elif getattr(match, "except_", None) is not None:
                for v in list(match.except_):
                    out &= ~np.asarray(col == v, dtype=bool)
            elif hasattr(match, "value"):
                out &= np.asarray(col == match.value, dtype=bool)
            else:
                raise NotImplementedError(f"unsupported matcher {type(match).__name__} on '{cond.key}'")
        if rng is not None:
            num = np.array([x if isinstance(x, (int, float)) and not isinstance(x, bool) else np.nan for x in col], dtype=np.float64)
            with np.errstate(invalid="ignore"):
                for attr, op in (("gt", np.greater), ("gte", np.greater_equal), ("lt", np.less), ("lte", np.less_equal)):
                    bound = getattr(rng, attr, None)
                    if bound is not None:
                        out &= op(num, float(bound))
        return out

    # ------------------------------------------------------------------ the three client methods
    @staticmethod
    def _as_query(query) -> np.ndarray:
        q = np.asarray(query, dtype=np.float32)
        return q[None, :] if q.ndim == 1 else q

    def _points(self, scores, ids, with_payload) -> List[ScoredPoint]:
        keep = np.isfinite(scores) & (ids >= 0)
        pages = [int(i) for i in ids[keep]]
        pids = self._pids_of(pages)
        pls = self._payloads_of(pages) if with_payload else [None] * len(pages)
        return [ScoredPoint(pid, float(s), pl) for pid, s, pl in zip(pids, scores[keep], pls)]

    @_locked
    def query_points(self, collection_name=None, query=None, using=None, limit=10, query_filter=None,
                     with_payload=True, with_vectors=False, search_params=None, prefetch=None, timeout=None,
                     **_ignored) -> QueryResponse:
        if using is None:
            raise ValueError("`using` (named vector) is required")
        limit = int(limit)
        n_pages = self.corpus.n_pages(using)
        restrict = self._restriction(query_filter, n_pages)
        q = self._as_query(query)
        if prefetch:
            # Qdrant prefetch= (two_stage.py:170-176): stage-1 query and rerank query travel separately;
            # both stages run back to back on the device with a single host synchronisation.
            pf = prefetch[0] if isinstance(prefetch, (list, tuple)) else prefetch
            stages = self.corpus.search_multistage(
                [(pf.using, False, int(pf.limit)), (using, False, limit)], None,
                stage_queries=[self._as_query(pf.query), q], **restrict)
            scores, ids = stages[-1]
        else:
            scores, ids = self.corpus.search(using, q, limit, **restrict)
        points = self._points(scores, ids, with_payload)
        if with_vectors:
            names = [using] if with_vectors is True else list(with_vectors)
            pages = [self._page(p.id) for p in points]
            rows = {nm: self._read_pages(nm, pages) for nm in names}
            for p, page in zip(points, pages):
                p.vector = {nm: rows[nm][page].astype(np.float32).tolist() for nm in names}
        return QueryResponse(points)

    @_locked
    def query_three_stage(self, *, stage1_query, stage2_query, stage3_query, stage1_using, stage2_using,
                          stage3_using, stage1_k: int, stage2_k: int, top_k: int, query_filter=None):
        """The three ID-restricted scans of ThreeStageRetriever.search_server_side (three_stage.py:102-159)
        fused on the device: returns the three point lists (stage 3 with payloads)."""
        restrict = self._restriction(query_filter, self.corpus.n_pages(stage1_using))
        stages = self.corpus.search_multistage(
            [(stage1_using, False, int(stage1_k)), (stage2_using, False, int(stage2_k)), (stage3_using, False, int(top_k))],
            None, stage_queries=[self._as_query(stage1_query), self._as_query(stage2_query), self._as_query(stage3_query)],
            **restrict)
        return [self._points(scores, ids, si == 2) for si, (scores, ids) in enumerate(stages)]

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
        return [[self._points(scores, ids, with_payload and si == ns - 1) for si, (scores, ids) in enumerate(per_query)]
                for per_query in res]

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
            if not sq:
                return []
            sc, ids, st, cnt = self.corpus.search_multistage_batch(
                [(usings[s], False, int(limits[s])) for s in range(ns)], None, stage_queries=sq, final_only=True)
        else:
            if queries is None:
                raise ValueError("either stage_queries or queries is required")
            if len(queries) == 0:
                return []
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
        out = []
        for b in range(nq):
            pages, scores, stages = ids_l[b], sc_l[b], [col[b] for col in st_cols]
            if not row_full[b]:
                keep = np.nonzero(valid[b])[0].tolist()
                pages = [pages[j] for j in keep]
                scores = [scores[j] for j in keep]
                stages = [[col[j] for j in keep] for col in stages]
            pids = self._pids_of(pages)
            payloads = self._payloads_of(pages) if with_payload else [None] * len(pages)
            out.append((pids, scores, stages, payloads))
        return out

    @_locked
    def retrieve(self, collection_name=None, ids=(), with_payload=False, with_vectors=None, timeout=None, **_ignored):
        names = [] if not with_vectors else (list(with_vectors) if not isinstance(with_vectors, bool) else [])
        names = [nm for nm in names if self.corpus.has_store(nm)]
        known = [(pid, self._page(pid)) for pid in ids]
        known = [(pid, page) for pid, page in known if page >= 0]
        rows = {nm: self._read_pages(nm, [page for _, page in known]) for nm in names}
        out = []
        for pid, page in known:
            vec = {nm: rows[nm][page].astype(np.float32).tolist() for nm in names}
            out.append(ScoredPoint(pid, None, self._payloads_of([page])[0] if with_payload else None, vec or None))
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
        return _CollectionInfo(vectors, self._n_points_total(count))


class ShardedCorpusClient(GpuCorpusClient):
    """The same client over a corpus that is page-sharded across the GPUs of a box: one process per GPU
    (torchrun), every rank constructs the client over ITS shard (`corpus`, a GpuCorpus whose page_base is the first
    global page id of the rank's range) and all ranks make the same client calls in the same order (SPMD) — each call
    is one collective search in the library and every rank receives the same global result, so the retriever classes
    (`TwoStageRetriever`, `ThreeStageRetriever`, `SingleStageRetriever`, `MultiVectorRetriever`, and the reference's own)
    run unchanged at N > 1.

    point_ids / payloads describe THIS rank's pages (page order). The id / payload tables of all ranks are exchanged
    once here (and after every `sync_points`) through the torch.distributed group, so turning result pages into
    ScoredPoints never communicates; the vectors and the payload COLUMNS used by filters stay sharded with the pages:
    a filter is evaluated by every rank over its own pages only and enters the collective search as a rank-local
    candidate list. `with_vectors` / `retrieve(with_vectors=...)` fetch rows from the owning rank (one object
    all-gather per call: the reference's slow client-side rerank path, kept for completeness)."""

    def __init__(self, corpus, collection_name: str = "gpu", point_ids: Optional[Sequence[Any]] = None,
                 payloads: Optional[Sequence[Optional[dict]]] = None, group=None):
        super().__init__(corpus, collection_name, point_ids, payloads)
        import torch.distributed as dist

        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        if self._dist is not None and getattr(corpus, "world", 1) == 1 and self._dist.get_world_size(group) > 1:
            corpus.comm_init_torch(group)
        self.rank = getattr(corpus, "rank", 0)
        self.world = getattr(corpus, "world", 1)
        self._shards: List[Tuple[int, int, Optional[list], Optional[list]]] = []   # (base, n_pages, ids, payloads) by rank
        self._bases: List[int] = []
        self._gindex: Optional[Dict[Any, int]] = None
        self._plain = True
        self.sync_points()

    # ------------------------------------------------------------------ replicated id / payload tables
    def _gather(self, obj) -> list:
        if self._dist is None or self.world == 1:
            return [obj]
        out = [None] * self.world
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    @_locked
    def sync_points(self) -> None:
        """Collective: exchange (page range, ids, payloads) of every rank. Called by the constructor; call it again after
        the local tables changed (set_points / append_points on every rank)."""
        n_local = len(self._ids) if self._ids is not None else (
            len(self._payloads) if self._payloads is not None else self._local_pages())
        parts = self._gather((int(self.corpus.page_base), int(n_local), self._ids, self._payloads))
        parts.sort(key=lambda t: t[0])
        for (b0, n0, _, _), (b1, _, _, _) in zip(parts, parts[1:]):
            if b0 + n0 > b1:
                raise ValueError(f"page ranges of two ranks overlap: [{b0},{b0 + n0}) and [{b1},...)")
        self._shards = parts
        self._bases = [p[0] for p in parts]
        self._plain = all(p[2] is None for p in parts)
        if self._plain:
            self._gindex = None
        else:
            self._gindex = {}
            for base, n, ids, _ in parts:
                for i in range(n):
                    self._gindex[ids[i] if ids is not None else base + i] = base + i

    def _local_pages(self) -> int:
        for nm in ("initial", "mean_pooling", "experimental_pooling", "global_pooling"):
            if self.corpus.has_store(nm):
                return int(self.corpus.n_pages(nm))
        return 0

    def _locate(self, page: int) -> Tuple[int, int]:
        r = bisect.bisect_right(self._bases, page) - 1
        if r < 0 or page - self._bases[r] >= self._shards[r][1]:
            raise KeyError(f"page {page} belongs to no shard")
        return r, page - self._bases[r]

    def _pid(self, page: int):
        r, loc = self._locate(page)
        ids = self._shards[r][2]
        return ids[loc] if ids is not None else page

    def _page(self, pid) -> int:
        if self._gindex is not None:
            g = self._gindex.get(pid)
            if g is None and not isinstance(pid, str):
                g = self._gindex.get(str(pid))
            return -1 if g is None else g
        try:
            return int(pid)
        except (TypeError, ValueError):
            return -1

    def _payload(self, page: int):
        r, loc = self._locate(page)
        pls = self._shards[r][3]
        return {} if pls is None else pls[loc]

    def _pids_of(self, pages: List[int]) -> list:
        return pages if self._plain else [self._pid(p) for p in pages]

    def _payloads_of(self, pages: List[int]) -> list:
        return [self._payload(p) for p in pages]

    def _n_points_total(self, n_local: int) -> int:
        return int(sum(self._gather(int(n_local))))

    def _read_pages(self, name: str, pages: Sequence[int]) -> Dict[int, np.ndarray]:
        """Collective: every rank reads the listed pages it owns; one object all-gather hands everyone all of them."""
        base = self.corpus.page_base
        n = self.corpus.n_pages(name) if self.corpus.has_store(name) else 0
        mine = {int(p): self.corpus.read_page(name, int(p) - base) for p in pages if 0 <= int(p) - base < n}
        out: Dict[int, np.ndarray] = {}
        for part in self._gather(mine):
            out.update(part)
        return out

    # Ingest at N > 1: every rank ingests ITS OWN points (GpuIndexer(corpus, client=this client).upload_batch(points of
    # this rank) — route ids to ranks with `owner_rank_of_id`, so that an upsert of an id lands on the rank that holds it),
    # then all ranks call sync_points() together: the local table edits (append_points / set_payload / remove_points)
    # become visible to result building on every rank.


def owner_rank_of_id(point_id, world: int) -> int:
    """Stable rank assignment of a point id for sharded ingest (the same id always lands on the same rank)."""
    import hashlib

    h = hashlib.sha256(str(point_id).encode()).digest()
    return int.from_bytes(h[:8], "little") % max(int(world), 1)
