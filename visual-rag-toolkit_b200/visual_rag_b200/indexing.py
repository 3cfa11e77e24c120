"""`GpuIndexer`: the ingest side of the GPU-resident corpus — mirror of the parts of
visual_rag/indexing/qdrant_indexer.py::QdrantIndexer that the processing pipeline drives
(`create_collection` 154-267, `upload_batch` 341-507, `check_exists` 509-520, `generate_point_id` 602-613).

`ProcessingPipeline` (pipeline.py:620-629) hands `upload_batch` point dicts with `id`, `visual_embedding`,
`tile_pooled_embedding`, `experimental_pooled_embedding` (array or {name: array}), `global_pooled_embedding`
(optional) and `metadata`; this class appends them to the named stores of a GpuCorpus with the same store-dtype
cast (every array -> fp32 -> fp16, 423-441) and the same fallback global = mean(tile_pooled) (417-421), and
registers id / payload with the GpuCorpusClient so the retrievers can serve the pages immediately. Ids that already
exist are overwritten (upsert) whatever their new shape; named vectors are optional per point; a failed batch leaves
nothing behind (appends are rolled back) and reports 0 uploaded points, like the reference.
"""

from __future__ import annotations

import hashlib
import logging
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .client import GpuCorpusClient
from .corpus import GpuCorpus

logger = logging.getLogger(__name__)


class GpuIndexer:
    def __init__(self, corpus: GpuCorpus, collection_name: str = "gpu", client: Optional[GpuCorpusClient] = None,
                 vector_datatype: str = "float16"):
        if vector_datatype not in ("float16", "float32"):
            raise ValueError("vector_datatype must be 'float16' or 'float32'")   # qdrant_indexer.py:137-138
        # the HBM store dtype is always fp16 (tcgen05 has no fp32 operand path); fp32 collections are rounded on ingest
        self.vector_datatype = vector_datatype
        self.corpus = corpus
        self.collection_name = collection_name
        self.client = client if client is not None else GpuCorpusClient(corpus, collection_name, point_ids=[], payloads=[])
        if self.client._ids is None:
            self.client.set_points([], [])
        # At N > 1 (client = ShardedCorpusClient) every rank runs its own indexer over its own points; see the note at the
        # end of client.py (owner_rank_of_id, sync_points).
        self._extra_names: List[str] = []
        self._seen_names: List[str] = []      # every named vector an upload has written so far
        # the reference drives upload_batch from uploader threads (run_qdrant_beir.py:720-768) while queries may run: one
        # batch at a time validates, writes its pages to every named store and registers its ids; searches wait
        self._lock = self.client._lock      # shared with the client's query methods (re-entrant)

    # ------------------------------------------------------------------ collection
    def collection_exists(self) -> bool:
        return self.corpus.has_store("initial")

    def create_collection(self, force_recreate: bool = False, enable_quantization: bool = False,
                          indexing_threshold: int = 20000, full_scan_threshold: int = 0,
                          experimental_vector_names: Optional[Sequence[str]] = None) -> bool:
        """qdrant_indexer.py:154-267: named vectors initial / mean_pooling / experimental_pooling[...] /
        global_pooling. Stores are created lazily by the first upload; force_recreate drops existing ones."""
        names = ["initial", "mean_pooling", "experimental_pooling", "global_pooling"] + [str(n) for n in (experimental_vector_names or [])]
        self._extra_names = [str(n) for n in (experimental_vector_names or [])]
        if self.collection_exists():
            if not force_recreate:
                return False
            for nm in list(dict.fromkeys(names + self._seen_names)):
                if self.corpus.has_store(nm):
                    self.corpus.drop_store(nm)
            self._seen_names = []
            self.client.set_points([], [])
        return True

    @staticmethod
    def generate_point_id(filename: str, page_number: int) -> str:
        """qdrant_indexer.py:602-613."""
        content = f"{filename}:page:{page_number}"
        hash_obj = hashlib.sha256(content.encode())
        hex_str = hash_obj.hexdigest()[:32]
        return f"{hex_str[:8]}-{hex_str[8:12]}-{hex_str[12:16]}-{hex_str[16:20]}-{hex_str[20:32]}"

    def check_exists(self, chunk_id: str) -> bool:
        return self.client.local_page(chunk_id) >= 0

    def get_existing_ids(self, filename: Optional[str] = None) -> set:
        ids = self.client._ids or []
        if filename is None:
            return set(ids)
        pl = self.client._payloads or []
        return {i for i, p in zip(ids, pl) if p and p.get("filename") == filename}

    # ------------------------------------------------------------------ upload
    @staticmethod
    def _rows(val) -> np.ndarray:
        a = np.array(val, dtype=np.float32)          # every array goes through fp32 first (qdrant_indexer.py:423-441)
        return a.reshape(-1, a.shape[-1]) if a.ndim != 2 else a

    def upload_batch(self, points: List[Dict[str, Any]], max_retries: int = 3, delay_between_batches: float = 0.0,
                     wait: bool = True, stop_event=None) -> int:
        """qdrant_indexer.py:341-507 (client.upsert semantics). Returns the number of uploaded points — and, like the
        reference, 0 after logging when the batch cannot be written (malformed points, device errors): nothing of a
        failed batch stays in the collection.

        New ids are appended behind the existing pages; an id that is already in the collection is overwritten (vectors
        of every named store + payload) whatever its new shape — a page whose row count changed moves the store onto a
        page table until `compact()`. Named vectors are optional per point, as in Qdrant: the pipeline emits
        `experimental_pooling_2d` only for pages with a tile grid (pipeline.py:485-503); a point without a named vector
        owns an empty page in that store (it scores -inf there and is never returned from it), so page indices stay
        aligned across stores. Within one batch the last occurrence of an id wins."""
        if not points:
            return 0
        if stop_event is not None and getattr(stop_event, "is_set", lambda: False)():
            return 0
        with self._lock:
            try:
                return self._upload_locked(points)
            except Exception as e:  # noqa: BLE001 - the reference's contract: log and report 0 uploaded points
                logger.error(f"Upload failed: {e}")
                return 0

    def _known_names(self) -> List[str]:
        names = ["initial", "mean_pooling", "global_pooling", "experimental_pooling"] + list(self._extra_names) + list(self._seen_names)
        return [n for n in dict.fromkeys(names) if self.corpus.has_store(n)]

    def _upload_locked(self, points: List[Dict[str, Any]]) -> int:
        # ---- every point -> its named vectors (fp32 first, 423-441); the last occurrence of an id wins
        per_id: Dict[Any, Dict[str, np.ndarray]] = {}
        payload_of: Dict[Any, Any] = {}
        for p in points:
            tile = self._rows(p["tile_pooled_embedding"])
            glob = p.get("global_pooled_embedding")
            if glob is None:
                glob = tile.mean(axis=0)                                    # 417-421
            glob = np.array(glob, dtype=np.float32).reshape(1, -1)
            per_point = {"initial": self._rows(p["visual_embedding"]), "mean_pooling": tile, "global_pooling": glob}
            exp = p.get("experimental_pooled_embedding")
            if isinstance(exp, dict):
                for k, v in exp.items():
                    if v is not None:
                        per_point[str(k)] = self._rows(v)
            elif exp is not None:
                per_point["experimental_pooling"] = self._rows(exp)
            for k, v in per_point.items():
                if v.ndim != 2 or v.shape[1] != 128:
                    raise ValueError(f"point {p['id']!r}: named vector '{k}' has shape {v.shape}, expected [rows, 128]")
            per_id.pop(p["id"], None)            # re-insert so that dict order = order of the last occurrences
            per_id[p["id"]] = per_point
            payload_of[p["id"]] = p.get("metadata")
        new_ids = [i for i in per_id if not self.check_exists(i)]
        old_ids = [i for i in per_id if self.check_exists(i)]
        n_before = len(self.client._ids)
        existing = self._known_names()
        names = list(dict.fromkeys(existing + sorted({k for v in per_id.values() for k in v})))
        # ---- validate before anything is written: every store of the collection holds one page per point
        for name in existing:
            have = self.corpus.n_pages(name)
            if have != n_before:
                raise ValueError(f"named vector '{name}' holds {have} pages but the collection has {n_before} points")
        old_pages = [self.client.local_page(i) for i in old_ids]
        empty = np.zeros((0, 128), dtype=np.float32)

        def pack(ids, name):
            mats = [per_id[i].get(name, empty) for i in ids]       # a point without this named vector: an empty page
            off = np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]).astype(np.int64)
            rows = np.concatenate(mats, axis=0).astype(np.float16) if mats else np.zeros((0, 128), np.float16)   # store dtype cast
            return rows, off

        # ---- write. Appends first (they can be rolled back by truncating), then the in-place part of the upsert.
        appended: List[str] = []
        created: List[str] = []
        try:
            if new_ids:
                for name in names:
                    if not self.corpus.has_store(name):
                        created.append(name)
                        if n_before > 0:        # a named vector that appears late: the earlier points own empty pages in it
                            self.corpus.append_store(name, np.zeros((0, 128), np.float16), page_offsets=np.zeros((n_before + 1,), np.int64))
                    rows, off = pack(new_ids, name)
                    self.corpus.append_store(name, rows, page_offsets=off)
                    appended.append(name)
            if old_ids:
                for name in names:
                    if not self.corpus.has_store(name):
                        created.append(name)
                        self.corpus.append_store(name, np.zeros((0, 128), np.float16), page_offsets=np.zeros((n_before + 1,), np.int64))
                    rows, off = pack(old_ids, name)
                    self.corpus.replace_pages(name, old_pages, rows, off)
        except Exception:
            for name in created:
                if self.corpus.has_store(name):
                    self.corpus.drop_store(name)
            for name in appended:
                if name not in created and self.corpus.has_store(name):
                    self.corpus.truncate_store(name, n_before)
            raise
        for name in names:
            if name not in self._seen_names:
                self._seen_names.append(name)
        for i in old_ids:
            self.client.set_payload(i, payload_of[i])
        if new_ids:
            self.client.append_points(new_ids, [payload_of[i] for i in new_ids])
        return len(points)

    # ------------------------------------------------------------------ deletes / maintenance
    def delete_points(self, point_ids: Sequence[Any]) -> int:
        """qdrant `client.delete(points_selector=ids)`: the points disappear from every search and from check_exists; their
        pages keep their index (and are reclaimed by `compact`). Returns the number of deleted points."""
        with self._lock:
            pages = [self.client.local_page(i) for i in point_ids if self.check_exists(i)]
            if not pages:
                return 0
            for name in self._known_names():
                self.corpus.delete_pages(name, pages)
            self.client.remove_points([i for i in point_ids if self.check_exists(i)])
            return len(pages)

    def compact(self) -> None:
        """Return every named store to the dense layout (after shape-changing upserts / deletes)."""
        with self._lock:
            for name in self._known_names():
                self.corpus.compact_store(name)
