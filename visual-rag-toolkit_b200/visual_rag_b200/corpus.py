"""GPU-resident corpus store: the B200 stand-in for the Qdrant collection of the reference.

One `GpuCorpus` owns, on one GPU, the named vector stores the reference's collection schema defines
(visual_rag/indexing/qdrant_indexer.py:200-239): ``initial`` (all page tokens), ``mean_pooling``,
``experimental_pooling*`` (pooled multi-vectors) and ``global_pooling`` (one row per page) — fp16 rows in
HBM plus a per-row fp32 inverse norm, pages described by a row-offset array.  All scoring goes through
libvrag_b200 (hand-written sm_100a kernels); nothing in this module computes scores on the CPU.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

DIM = 128
MAX_K = 1 << 20      # kTopkHardMaxK of the library: the largest `limit` / `prefetch_k` / stage size of one search stage
MAX_K_BATCH = 4096   # kTopkMaxK: the largest stage size of the BATCHED native call (beyond it a batch runs query by query)


def _check_k(k: int, what: str = "k") -> int:
    """Stage sizes are validated here so that an oversized request is a plain ValueError (never retried by the
    retrievers' back-off loop) instead of a library error."""
    k = int(k)
    if k > MAX_K:
        raise ValueError(f"{what}={k} exceeds the supported maximum {MAX_K} results per search stage")
    return k


def _as_f32_query(query) -> np.ndarray:
    """Query -> contiguous fp32 [Q,128] numpy (same conversions as TwoStageRetriever._to_numpy,
    visual_rag/retrieval/two_stage.py:428-434)."""
    if isinstance(query, np.ndarray):
        if query.dtype == np.float32 and query.ndim == 2 and query.shape[1] == DIM and query.flags.c_contiguous:
            return query    # already in the native layout: borrowed for the duration of the call, never written
    elif not isinstance(query, (list, tuple)):
        try:
            import torch

            if isinstance(query, torch.Tensor):
                query = query.detach().cpu().float().numpy()
        except ImportError:  # pragma: no cover
            pass
    q = np.array(query, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if q.ndim != 2 or q.shape[1] != DIM:
        raise ValueError(f"query must be [num_tokens, {DIM}], got {q.shape}")
    return np.ascontiguousarray(q)


class PackedQueries:
    """A batch of query matrices packed once into the layout the native batch call consumes: all rows
    concatenated ([R,128] fp32, C-contiguous) + row offsets ([n+1] int32). Build it with `pack_queries` when the same
    batch is searched repeatedly or the embedder already emits one padded tensor — the per-call Python work of
    converting and concatenating hundreds of small arrays is otherwise a third of a batched search's wall time."""

    __slots__ = ("rows", "offsets")

    def __init__(self, rows: np.ndarray, offsets: np.ndarray):
        self.rows = np.ascontiguousarray(rows, dtype=np.float32)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if self.rows.ndim != 2 or self.rows.shape[1] != DIM:
            raise ValueError(f"rows must be [total_rows, {DIM}]")
        if self.offsets.ndim != 1 or self.offsets.size < 1 or int(self.offsets[-1]) != self.rows.shape[0] or int(self.offsets[0]) != 0:
            raise ValueError("offsets must be n+1 non-decreasing row offsets from 0 to total_rows")

    def __len__(self) -> int:
        return self.offsets.size - 1


def pack_queries(queries: Sequence) -> PackedQueries:
    mats = [_as_f32_query(x) for x in queries]
    if not mats:
        return PackedQueries(np.zeros((0, DIM), np.float32), np.zeros((1,), np.int32))
    return PackedQueries(np.concatenate(mats, axis=0), np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]))


def query_flags(normalize: bool = True, pool_query: bool = False, fp16_query: bool = False) -> int:
    """fp16_query: opt-in reduced-precision query operand (VRAG_Q_FP16, include/vrag_b200.h); default exact."""
    return ((N.VRAG_Q_NORMALIZE if normalize else 0) | (N.VRAG_Q_POOL if pool_query else 0)
            | (N.VRAG_Q_FP16 if fp16_query else 0))


class GpuCorpus:
    """Handle to one shard of the corpus on one GPU.

    page_base: global id of this shard's first page (global page id = page_base + local index).
    """

    def __init__(self, device: int = 0, page_base: int = 0):
        self._lib = N.load()
        h = C.c_void_p()
        N.check(self._lib.vrag_corpus_create(int(device), int(page_base), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.page_base = int(page_base)
        self.rank, self.world = 0, 1     # set by comm_init: this handle is one shard of a page-sharded corpus

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vrag_corpus_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    @staticmethod
    def comm_unique_id() -> bytes:
        """The 128-byte communicator id rank 0 creates (ncclGetUniqueId) and the host hands to every rank."""
        buf = C.create_string_buffer(128)
        N.check(N.load().vrag_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: Optional[bytes]) -> None:
        """Join the communicator of the page-sharded corpus (collective over all `world` ranks). Afterwards `search`,
        `search_multistage` and `search_multistage_batch` of this handle are collective calls that return the merged
        GLOBAL lists on every rank (include/vrag_b200.h, multi-GPU section)."""
        if world > 1 and (unique_id is None or len(unique_id) != 128):
            raise ValueError("unique_id must be the 128 bytes of GpuCorpus.comm_unique_id() from rank 0")
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        N.check(self._lib.vrag_comm_init(self._h, int(rank), int(world), buf))
        self.rank, self.world = int(rank), int(world)

    def comm_peer_memory(self) -> bool:
        """True when the handle's collectives (messages up to 4 MB per rank) run over mapped peer memory (NVLink / NVSwitch,
        one kernel per collective) rather than NCCL — agreed by all ranks at comm_init."""
        v = C.c_int(0)
        N.check(self._lib.vrag_comm_transport(self._h, C.byref(v)))
        return bool(v.value)

    def comm_init_torch(self, group=None) -> None:
        """comm_init with the id distributed through an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [self.comm_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(rank, world, box[0])

    def comm_destroy(self) -> None:
        N.check(self._lib.vrag_comm_destroy(self._h))
        self.rank, self.world = 0, 1

    def comm_timing_us(self) -> List[float]:
        """Device time (us) of each collective of the most recent host-facing search on this handle, in issue order."""
        buf = (C.c_float * 32)()
        n = C.c_int()
        N.check(self._lib.vrag_last_comm_timing(self._h, buf, 32, C.byref(n)))
        return [float(buf[i]) for i in range(min(n.value, 32))]

    def comm_offsets_us(self) -> List[float]:
        """Start of each of those collectives in microseconds after the start of the search: with `comm_timing_us` the device
        timeline of a collective search (local stage work | exchange | local stage work | ...)."""
        buf = (C.c_float * 32)()
        n = C.c_int()
        N.check(self._lib.vrag_last_comm_offsets(self._h, buf, 32, C.byref(n)))
        return [float(buf[i]) for i in range(min(n.value, 32))]

    # ------------------------------------------------------------------ stores
    def add_store(
        self,
        name: str,
        rows,
        page_offsets: Optional[Sequence[int]] = None,
        fixed_rows: int = 0,
    ) -> None:
        """Upload a named store. rows: [total_rows,128] numpy fp16/fp32 (host) or a torch CUDA tensor
        on this device. fp32 rows are cast to the fp16 store dtype (qdrant_indexer.py:423-441)."""
        on_device = 0
        keep = None
        try:
            import torch

            if isinstance(rows, torch.Tensor):
                if rows.is_cuda:
                    if rows.device.index != self.device:
                        raise ValueError("rows tensor is on a different device than the corpus")
                    if rows.dtype not in (torch.float16, torch.float32):
                        rows = rows.float()
                    keep = rows.contiguous()
                    torch.cuda.synchronize(rows.device)
                    dtype = N.VRAG_F16 if keep.dtype == torch.float16 else N.VRAG_F32
                    total = keep.shape[0]
                    ptr = C.c_void_p(keep.data_ptr())
                    on_device = 1
                else:
                    rows = rows.float().numpy() if rows.dtype == torch.bfloat16 else rows.numpy()
        except ImportError:  # pragma: no cover
            pass
        if not on_device:
            arr = np.asarray(rows)
            if arr.dtype not in (np.float16, np.float32):
                arr = arr.astype(np.float32)
            arr = np.ascontiguousarray(arr.reshape(-1, DIM))
            keep = arr
            dtype = N.VRAG_F16 if arr.dtype == np.float16 else N.VRAG_F32
            total = arr.shape[0]
            ptr = arr.ctypes.data_as(C.c_void_p)
        if fixed_rows > 0:
            if total % fixed_rows:
                raise ValueError("total rows is not a multiple of fixed_rows")
            n_pages = total // fixed_rows
            off_p = None
        else:
            off = np.ascontiguousarray(np.asarray(page_offsets, dtype=np.int64))
            if off.ndim != 1 or off.size < 1:
                raise ValueError("page_offsets must be a 1-D array of n_pages+1 row offsets")
            if int(off[-1]) != total:
                raise ValueError("page_offsets[-1] must equal the number of rows")
            n_pages = off.size - 1
            off_p = off.ctypes.data_as(C.POINTER(C.c_int64))
        N.check(
            self._lib.vrag_store_add(
                self._h, name.encode(), ptr, dtype, on_device, off_p, int(n_pages), int(fixed_rows)
            )
        )
        del keep

    def append_store(self, name: str, rows, page_offsets: Optional[Sequence[int]] = None, fixed_rows: int = 0) -> None:
        """Append pages behind the existing pages of `name` (created on first use) — the per-batch ingest of
        QdrantIndexer.upload_batch (qdrant_indexer.py:341-507). rows: host numpy fp16/fp32 [total_rows,128]."""
        arr = np.asarray(rows)
        if arr.dtype not in (np.float16, np.float32):
            arr = arr.astype(np.float32)
        arr = np.ascontiguousarray(arr.reshape(-1, DIM))
        dtype = N.VRAG_F16 if arr.dtype == np.float16 else N.VRAG_F32
        if fixed_rows > 0:
            if arr.shape[0] % fixed_rows:
                raise ValueError("total rows is not a multiple of fixed_rows")
            n_pages, off_p = arr.shape[0] // fixed_rows, None
        else:
            off = np.ascontiguousarray(np.asarray(page_offsets, dtype=np.int64))
            if off.ndim != 1 or off.size < 1 or int(off[-1]) != arr.shape[0]:
                raise ValueError("page_offsets must be n_pages+1 row offsets ending at the number of rows")
            n_pages, off_p = off.size - 1, off.ctypes.data_as(C.POINTER(C.c_int64))
        N.check(self._lib.vrag_store_append(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), dtype, 0, off_p,
                                            int(n_pages), int(fixed_rows)))

    def replace_pages(self, name: str, local_pages: Sequence[int], rows, page_offsets: Sequence[int]) -> None:
        """Overwrite existing pages of `name` (the upsert of points that already exist, qdrant_indexer.py:459-507).
        Pages that keep their row count are rewritten in place; a page whose row count changed puts the store on a page
        table (see include/vrag_b200.h) until `compact_store`."""
        arr = np.asarray(rows)
        if arr.dtype not in (np.float16, np.float32):
            arr = arr.astype(np.float32)
        arr = np.ascontiguousarray(arr.reshape(-1, DIM))
        dtype = N.VRAG_F16 if arr.dtype == np.float16 else N.VRAG_F32
        pages = np.ascontiguousarray(np.asarray(local_pages, dtype=np.int64))
        off = np.ascontiguousarray(np.asarray(page_offsets, dtype=np.int64))
        if off.ndim != 1 or off.size != pages.size + 1 or int(off[-1]) != arr.shape[0]:
            raise ValueError("page_offsets must be len(local_pages)+1 row offsets ending at the number of rows")
        N.check(self._lib.vrag_store_replace_pages(self._h, name.encode(), pages.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   int(pages.size), arr.ctypes.data_as(C.c_void_p), dtype, 0,
                                                   off.ctypes.data_as(C.POINTER(C.c_int64)), 0))

    def delete_pages(self, name: str, local_pages: Sequence[int]) -> None:
        """Deleted pages keep their index but own no rows: they score -inf and are never returned."""
        pages = np.ascontiguousarray(np.asarray(local_pages, dtype=np.int64))
        N.check(self._lib.vrag_store_delete_pages(self._h, name.encode(), pages.ctypes.data_as(C.POINTER(C.c_int64)), int(pages.size)))

    def truncate_store(self, name: str, n_pages: int) -> None:
        """Drop the last pages of `name`, keeping n_pages (rollback of a failed multi-store batch upload)."""
        N.check(self._lib.vrag_store_truncate(self._h, name.encode(), int(n_pages)))

    def compact_store(self, name: str) -> None:
        """Return a store with a page table to the dense layout (reclaims the rows of replaced / deleted pages)."""
        N.check(self._lib.vrag_store_compact(self._h, name.encode()))

    def add_synthetic_store(
        self,
        name: str,
        n_pages: int,
        fixed_rows: int = 0,
        page_offsets: Optional[Sequence[int]] = None,
        seed: int = 0,
        row_seed_base: int = 0,
    ) -> None:
        """Generate the seeded synthetic corpus of SURVEY.md §8(d) on the device (unit-norm gaussian
        rows rounded to fp16)."""
        off_p = None
        if fixed_rows <= 0:
            off = np.ascontiguousarray(np.asarray(page_offsets, dtype=np.int64))
            n_pages = off.size - 1
            off_p = off.ctypes.data_as(C.POINTER(C.c_int64))
        N.check(
            self._lib.vrag_store_add_synthetic(
                self._h, name.encode(), off_p, int(n_pages), int(fixed_rows), C.c_uint64(seed), int(row_seed_base)
            )
        )

    def drop_store(self, name: str) -> None:
        N.check(self._lib.vrag_store_drop(self._h, name.encode()))

    def store_info(self, name: str) -> Dict[str, int]:
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        N.check(self._lib.vrag_store_info(self._h, name.encode(), C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"n_pages": a.value, "total_rows": b.value, "fixed_rows": c.value, "max_rows": d.value}

    def has_store(self, name: str) -> bool:
        a = C.c_int64()
        return self._lib.vrag_store_info(self._h, name.encode(), C.byref(a), None, None, None) == 0

    def n_pages(self, name: str) -> int:
        return self.store_info(name)["n_pages"]

    def read_rows(self, name: str, row0: int, n_rows: int) -> np.ndarray:
        out = np.empty((int(n_rows), DIM), dtype=np.float16)
        N.check(
            self._lib.vrag_store_read_rows(self._h, name.encode(), int(row0), int(n_rows), out.ctypes.data_as(C.c_void_p))
        )
        return out

    def page_range(self, name: str, local_page: int) -> Tuple[int, int]:
        r0, n = C.c_int64(), C.c_int64()
        N.check(self._lib.vrag_store_page_range(self._h, name.encode(), int(local_page), C.byref(r0), C.byref(n)))
        return r0.value, n.value

    def page_rows(self, name: str) -> np.ndarray:
        """Row (token) count of every page of the store, int64 [n_pages] — one call."""
        n = self.n_pages(name)
        out = np.empty((n,), dtype=np.int64)
        N.check(self._lib.vrag_store_page_rows(self._h, name.encode(), 0, n, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def read_page(self, name: str, local_page: int) -> np.ndarray:
        """fp16 rows of one page ([rows,128]) — what qdrant `retrieve(with_vectors=[name])` returns
        (two_stage.py:383-400)."""
        r0, n = self.page_range(name, local_page)
        return self.read_rows(name, r0, n)

    # ------------------------------------------------------------------ bulk pooling on the device
    def pool_store(self, src: str, specs: Sequence, dst_names: Sequence[str], grid_hw=None) -> float:
        """Derive pooled stores from `src` without leaving HBM (pipeline.py:400-507 for a whole collection).
        specs: sequence of _native.PoolSpec (see visual_rag_b200.embedding.pooling.spec_*). grid_hw: optional
        [n_pages,2] int32 per-page (grid_h, grid_w) / (n_rows, n_cols). Returns the device time in ms."""
        n = len(specs)
        arr = (N.PoolSpec * n)(*specs)
        names = (C.c_char_p * n)(*[d.encode() for d in dst_names])
        g = None
        if grid_hw is not None:
            gh = np.ascontiguousarray(np.asarray(grid_hw, dtype=np.int32).reshape(-1, 2))
            if gh.shape[0] != self.n_pages(src):
                raise ValueError("grid_hw must have one (h, w) pair per page")
            g = gh.ctypes.data_as(C.POINTER(C.c_int32))
        N.check(self._lib.vrag_store_pool(self._h, src.encode(), n, arr, names, g))
        return self.last_timing_ms()[0]

    # ------------------------------------------------------------------ payload filters (page bitmasks)
    def create_filter(self, allowed) -> int:
        """Upload a payload filter as a bitmask over this handle's pages: allowed[p] truthy = page p passes. Returns a
        filter id for `search(..., filter_id=)` / `search_multistage(..., filter_id=)`; the mask stays on the device until
        destroy_filter (one upload per distinct filter, none per query)."""
        a = np.asarray(allowed, dtype=bool)
        n = int(a.shape[0])
        bits = np.packbits(a, bitorder="little")
        words = np.zeros(((n + 31) // 32) * 4, dtype=np.uint8)
        words[: bits.shape[0]] = bits
        w32 = np.ascontiguousarray(words.view(np.uint32)) if words.size else np.zeros((1,), np.uint32)
        fid = C.c_int()
        N.check(self._lib.vrag_filter_create(self._h, w32.ctypes.data_as(C.POINTER(C.c_uint32)), n, C.byref(fid)))
        return int(fid.value)

    def destroy_filter(self, filter_id: int) -> None:
        N.check(self._lib.vrag_filter_destroy(self._h, int(filter_id)))

    # ------------------------------------------------------------------ scoring
    @staticmethod
    def _cand_arg(candidate_ids):
        """candidate id list -> (array kept alive by the caller, pointer, count). On a sharded handle the list is
        rank-local and may be empty: the call is collective and still has to be made, so an empty list travels as a
        valid pointer with count 0."""
        if candidate_ids is None:
            return None, None, 0
        cand = np.ascontiguousarray(np.asarray(candidate_ids, dtype=np.int64))
        n = int(cand.size)
        if n == 0:
            cand = np.zeros((1,), dtype=np.int64)
        return cand, cand.ctypes.data_as(C.POINTER(C.c_int64)), n

    def score(
        self,
        name: str,
        query,
        normalize: bool = True,
        pool_query: bool = False,
        candidate_ids: Optional[Sequence[int]] = None,
        fp16_query: bool = False,
    ) -> np.ndarray:
        """MaxSim score of every page (or of the listed global page ids). fp32 [n].
        fp16_query (all scoring methods): opt-in VRAG_Q_FP16 — contract the query as plain fp16 instead of the exact
        hi/lo pair: scores move by < 1e-4 relative (inside the 1e-3 parity gate), scans of stores with > 128 rows per
        page do half the tensor work (+2-3 % single-query, +15-19 % for 4-query dense batches, both power-limited)."""
        q = _as_f32_query(query)
        if candidate_ids is None:
            n = self.n_pages(name)
            cand_p, n_cand = None, 0
        else:
            cand = np.ascontiguousarray(np.asarray(candidate_ids, dtype=np.int64))
            n = n_cand = cand.size
            cand_p = cand.ctypes.data_as(C.POINTER(C.c_int64))
        out = np.empty((n,), dtype=np.float32)
        N.check(
            self._lib.vrag_score(
                self._h, name.encode(), q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0],
                query_flags(normalize, pool_query, fp16_query), cand_p, n_cand, out.ctypes.data_as(C.POINTER(C.c_float)),
            )
        )
        return out

    def search(
        self,
        name: str,
        query,
        k: int,
        normalize: bool = True,
        pool_query: bool = False,
        candidate_ids: Optional[Sequence[int]] = None,
        fp16_query: bool = False,
        filter_id: Optional[int] = None,
    ) -> Tuple[np.ndarray, np.ndarray]:
        """Top-k pages by MaxSim: (scores fp32 [m], global page ids int64 [m]), m <= k, sorted by score
        descending, ties by lower id. filter_id: restrict the scan to the pages of a `create_filter` bitmask."""
        if filter_id is not None:
            return self.search_multistage([(name, pool_query, k)], query, normalize, fp16_query=fp16_query, filter_id=filter_id)[0]
        q = _as_f32_query(query)
        k = _check_k(k)
        if k < 1 or (candidate_ids is not None and len(candidate_ids) == 0 and self.world == 1):
            return np.empty((0,), np.float32), np.empty((0,), np.int64)
        cand, cand_p, n_cand = self._cand_arg(candidate_ids)
        scores = np.empty((k,), dtype=np.float32)
        ids = np.empty((k,), dtype=np.int64)
        cnt = C.c_int()
        N.check(
            self._lib.vrag_search(
                self._h, name.encode(), q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0],
                query_flags(normalize, pool_query, fp16_query), cand_p, n_cand, k,
                scores.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(cnt),
            )
        )
        m = cnt.value
        return scores[:m], ids[:m]

    def search_multistage(
        self,
        stages: Sequence[Tuple[str, bool, int]],
        query,
        normalize: bool = True,
        stage_queries: Optional[Sequence] = None,
        candidate_ids: Optional[Sequence[int]] = None,
        fp16_query: bool = False,
        filter_id: Optional[int] = None,
    ) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Fused multi-stage search: stages = [(store name, pool_query, k), ...]; stage s is restricted to
        the survivors of stage s-1 (stage 0 to `candidate_ids` if given, or to the pages of the `create_filter`
        bitmask `filter_id` — filtered-out pages are skipped inside the scan and come back with score -inf, i.e. after
        every page that passes). One host synchronisation.
        stage_queries: optional per-stage query matrices (else every stage uses `query`).
        Returns per-stage (scores, ids)."""
        ns = len(stages)
        for st in stages:
            _check_k(st[2], f"stage '{st[0]}': k")
        if stage_queries is not None:
            qs = [_as_f32_query(x) for x in stage_queries]
            if len(qs) != ns:
                raise ValueError("stage_queries must have one entry per stage")
            q = np.ascontiguousarray(np.concatenate(qs, axis=0))
            offs = np.concatenate([[0], np.cumsum([x.shape[0] for x in qs])]).astype(np.int32)
            off_p = offs.ctypes.data_as(C.POINTER(C.c_int))
        else:
            q = _as_f32_query(query)
            off_p = None
        if candidate_ids is not None and len(candidate_ids) == 0 and self.world == 1:
            return [(np.empty((0,), np.float32), np.empty((0,), np.int64)) for _ in stages]
        cand, cand_p, n_cand = self._cand_arg(candidate_ids)
        names = (C.c_char_p * ns)(*[s[0].encode() for s in stages])
        flags = (C.c_uint32 * ns)(*[query_flags(normalize, bool(s[1]), fp16_query) for s in stages])
        ks = (C.c_int * ns)(*[int(s[2]) for s in stages])
        total = int(sum(int(s[2]) for s in stages))
        scores = np.empty((total,), dtype=np.float32)
        ids = np.empty((total,), dtype=np.int64)
        counts = (C.c_int * ns)()
        if filter_id is not None:
            if candidate_ids is not None:
                raise ValueError("pass either candidate_ids or filter_id, not both")
            N.check(
                self._lib.vrag_search_multistage_filtered(
                    self._h, int(filter_id), ns, names, flags, ks, q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0], off_p,
                    scores.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int64)), counts,
                )
            )
        else:
            N.check(
                self._lib.vrag_search_multistage(
                    self._h, ns, names, flags, ks, q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0], off_p,
                    cand_p, n_cand,
                    scores.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int64)), counts,
                )
            )
        out = []
        off = 0
        for s in range(ns):
            m = counts[s]
            out.append((scores[off : off + m].copy(), ids[off : off + m].copy()))
            off += int(stages[s][2])
        return out

    def search_multistage_batch(
        self,
        stages: Sequence[Tuple[str, bool, int]],
        queries: Sequence,
        normalize: bool = True,
        stage_queries: Optional[Sequence[Sequence]] = None,
        as_arrays: bool = False,
        final_only: bool = False,
        fp16_query: bool = False,
    ):
        """`search_multistage` for a batch of independent queries in ONE native call / host synchronisation
        (BASELINE configs[2]: 256 queries). queries: sequence of [Q_b,128] matrices (ragged). stage_queries:
        optional, per query a sequence of one matrix per stage (e.g. the mean-pooled prefetch vector and the
        token matrix, two_stage.py:142,159). Returns, per query, the per-stage (scores, ids) lists; with
        as_arrays=True instead one (scores [nq,k_s], ids [nq,k_s], counts [nq]) triple per stage (no per-query
        Python work: rows are valid up to their count, the rest is (-inf, -1)). final_only=True returns just
        (scores [nq,k_last], ids [nq,k_last], stage_scores [nq,k_last,n_stages-1], counts [nq]): the last stage's lists and,
        for every final result, the score its page had in each earlier stage (NaN if absent) — what the retrievers' result
        dictionaries carry — so the long intermediate lists never leave the device."""
        ns = len(stages)
        for st in stages:
            _check_k(st[2], f"stage '{st[0]}': k")
        if stage_queries is not None:
            mats = []
            for sq in stage_queries:
                if len(sq) != ns:
                    raise ValueError("stage_queries must have one entry per stage for every query")
                mats.extend(_as_f32_query(x) for x in sq)
            nq = len(stage_queries)
            per_stage = 1
        elif isinstance(queries, PackedQueries):
            mats = None
            nq = len(queries)
            per_stage = 0
        else:
            mats = [_as_f32_query(x) for x in queries]
            nq = len(mats)
            per_stage = 0
        if nq > 0 and any(int(st[2]) > MAX_K_BATCH for st in stages):
            return self._batch_query_by_query(stages, queries, mats, nq, normalize, stage_queries, as_arrays, final_only, fp16_query)
        if nq == 0:
            if final_only:
                kl = int(stages[-1][2])
                return (np.empty((0, kl), np.float32), np.empty((0, kl), np.int64), np.empty((0, kl, max(ns - 1, 0)), np.float32),
                        np.empty((0,), np.int32))
            if as_arrays:
                return [(np.empty((0, int(k)), np.float32), np.empty((0, int(k)), np.int64), np.empty((0,), np.int32)) for _, _, k in stages]
            return []
        if mats is None:
            rows, offs = queries.rows, queries.offsets
        else:
            rows = np.ascontiguousarray(np.concatenate(mats, axis=0))
            offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]).astype(np.int32))
        names = (C.c_char_p * ns)(*[s[0].encode() for s in stages])
        flags = (C.c_uint32 * ns)(*[query_flags(normalize, bool(s[1]), fp16_query) for s in stages])
        ks = [int(s[2]) for s in stages]
        ks_c = (C.c_int * ns)(*ks)
        if final_only:
            kl = ks[-1]
            f_sc = np.empty((nq, kl), dtype=np.float32)
            f_id = np.empty((nq, kl), dtype=np.int64)
            f_st = np.empty((nq, kl, max(ns - 1, 1)), dtype=np.float32)
            f_cnt = np.zeros((nq,), dtype=np.int32)
            N.check(
                self._lib.vrag_search_multistage_batch_final(
                    self._h, ns, names, flags, ks_c, nq, rows.ctypes.data_as(C.POINTER(C.c_float)),
                    offs.ctypes.data_as(C.POINTER(C.c_int)), per_stage,
                    f_sc.ctypes.data_as(C.POINTER(C.c_float)), f_id.ctypes.data_as(C.POINTER(C.c_int64)),
                    f_st.ctypes.data_as(C.POINTER(C.c_float)), f_cnt.ctypes.data_as(C.POINTER(C.c_int)),
                )
            )
            return f_sc, f_id, f_st[:, :, : ns - 1], f_cnt
        total = int(sum(ks)) * nq
        scores = np.empty((total,), dtype=np.float32)
        ids = np.empty((total,), dtype=np.int64)
        counts = np.zeros((ns * nq,), dtype=np.int32)
        N.check(
            self._lib.vrag_search_multistage_batch(
                self._h, ns, names, flags, ks_c, nq, rows.ctypes.data_as(C.POINTER(C.c_float)),
                offs.ctypes.data_as(C.POINTER(C.c_int)), per_stage,
                scores.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int64)),
                counts.ctypes.data_as(C.POINTER(C.c_int)),
            )
        )
        if as_arrays:
            res, off = [], 0
            for s in range(ns):
                res.append((scores[off : off + nq * ks[s]].reshape(nq, ks[s]), ids[off : off + nq * ks[s]].reshape(nq, ks[s]),
                            counts[s * nq : (s + 1) * nq]))
                off += nq * ks[s]
            return res
        out: List[List[Tuple[np.ndarray, np.ndarray]]] = [[] for _ in range(nq)]
        off = 0
        for s in range(ns):
            sc = scores[off : off + nq * ks[s]].reshape(nq, ks[s])
            ii = ids[off : off + nq * ks[s]].reshape(nq, ks[s])
            for b in range(nq):
                m = int(counts[s * nq + b])
                out[b].append((sc[b, :m], ii[b, :m]))   # views into this call's own result arrays
            off += nq * ks[s]
        return out

    def _batch_query_by_query(self, stages, queries, mats, nq, normalize, stage_queries, as_arrays, final_only, fp16_query):
        """A batch with a stage size beyond MAX_K_BATCH (the batched kernels keep at most 4096 results per query and stage in
        shared memory): the same result formats from one `search_multistage` call per query."""
        ns = len(stages)
        ks = [int(st[2]) for st in stages]
        if stage_queries is not None:
            per = [self.search_multistage(stages, None, normalize, stage_queries=sq, fp16_query=fp16_query) for sq in stage_queries]
        else:
            if mats is None:   # PackedQueries
                o = queries.offsets
                mats = [queries.rows[int(o[b]) : int(o[b + 1])] for b in range(nq)]
            per = [self.search_multistage(stages, q, normalize, fp16_query=fp16_query) for q in mats]
        if not as_arrays and not final_only:
            return per
        arr = []
        for s in range(ns):
            sc = np.full((nq, ks[s]), -np.inf, dtype=np.float32)
            ii = np.full((nq, ks[s]), -1, dtype=np.int64)
            cnt = np.zeros((nq,), dtype=np.int32)
            for b in range(nq):
                m = len(per[b][s][1])
                sc[b, :m], ii[b, :m], cnt[b] = per[b][s][0], per[b][s][1], m
            arr.append((sc, ii, cnt))
        if not final_only:
            return arr
        f_sc, f_id, f_cnt = arr[-1]
        f_st = np.full((nq, ks[-1], max(ns - 1, 0)), np.nan, dtype=np.float32)
        for b in range(nq):
            for s in range(ns - 1):
                sc_s, id_s = per[b][s]
                order = np.argsort(id_s, kind="stable")
                pos = np.searchsorted(id_s[order], f_id[b, : f_cnt[b]])
                pos = np.minimum(pos, max(len(order) - 1, 0))
                if len(order):
                    hit = id_s[order][pos] == f_id[b, : f_cnt[b]]
                    f_st[b, : f_cnt[b], s] = np.where(hit, sc_s[order][pos], np.nan)
        return f_sc, f_id, f_st, f_cnt

    def saliency(self, name: str, query, page_id: int) -> np.ndarray:
        """patch_scores of generate_saliency_map (visualization/saliency.py:69-79) for one page (global page id):
        fp32 [tokens], max over query tokens of the cosine with each document token."""
        q = _as_f32_query(query)
        r0, n = self.page_range(name, int(page_id) - self.page_base)
        out = np.empty((max(n, 1),), dtype=np.float32)
        got = C.c_int64()
        N.check(self._lib.vrag_saliency(self._h, name.encode(), q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0],
                                        int(page_id), out.ctypes.data_as(C.POINTER(C.c_float)), out.shape[0], C.byref(got)))
        return out[: got.value]

    # ------------------------------------------------------------------ device-pointer variants (multi-GPU path)
    def score_pages(self, query, pages: Sequence[np.ndarray], normalize: bool = True) -> np.ndarray:
        """MaxSim of one query against documents in host memory (compute_maxsim_score / compute_maxsim_batch,
        pooling.py:468-552): one [rows, 128] fp16 or fp32 array per document, handed over by pointer — no concatenation;
        the library casts / copies them into pinned staging with its worker pool while the previous chunk is in flight.
        fp32 [len(pages)]; an empty document scores -inf."""
        q = _as_f32_query(query)
        mats = [np.ascontiguousarray(m) for m in pages]
        n = len(mats)
        out = np.empty((n,), dtype=np.float32)
        if n == 0:
            return out
        dt = np.float16 if all(m.dtype == np.float16 for m in mats) else np.float32
        mats = [m if m.dtype == dt else m.astype(dt) for m in mats]
        for m in mats:
            if m.ndim != 2 or m.shape[1] != 128:
                raise ValueError(f"documents must be [rows, 128] arrays, got {m.shape}")
        ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in mats])
        rows = (C.c_int64 * n)(*[int(m.shape[0]) for m in mats])
        N.check(
            self._lib.vrag_score_pages(
                self._h, q.ctypes.data_as(C.POINTER(C.c_float)), q.shape[0], query_flags(normalize, False, False), ptrs, rows, n,
                N.VRAG_F16 if dt == np.float16 else N.VRAG_F32, out.ctypes.data_as(C.POINTER(C.c_float)),
            )
        )
        return out

    def score_dev(self, name: str, query_dev_ptr: int, n_query_rows: int, flags: int, cand_dev_ptr: int,
                  n_cand: int, out_scores_dev_ptr: int, stream: int) -> None:
        N.check(
            self._lib.vrag_score_dev(
                self._h, name.encode(), C.c_void_p(query_dev_ptr), int(n_query_rows), int(flags),
                C.c_void_p(cand_dev_ptr) if cand_dev_ptr else None, int(n_cand), C.c_void_p(out_scores_dev_ptr),
                C.c_void_p(stream),
            )
        )

    def topk_dev(self, scores_dev_ptr: int, ids_dev_ptr: int, id_base: int, n: int, k: int,
                 out_scores_dev_ptr: int, out_ids_dev_ptr: int, stream: int) -> None:
        N.check(
            self._lib.vrag_topk_dev(
                self._h, C.c_void_p(scores_dev_ptr), C.c_void_p(ids_dev_ptr) if ids_dev_ptr else None, int(id_base),
                int(n), int(k), C.c_void_p(out_scores_dev_ptr), C.c_void_p(out_ids_dev_ptr), C.c_void_p(stream),
            )
        )

    def search_multistage_dev(self, stages: Sequence[Tuple[str, bool, int]], query_dev_ptr: int, n_query_rows: int,
                              out_scores_dev_ptr: int, out_ids_dev_ptr: int, stream: int, normalize: bool = True) -> None:
        """All stages on the device for a query that is already there; results stay on the device (stage s at
        [sum(k[:s]), sum(k[:s+1])) of the output arrays), nothing is synchronised. Collective on a sharded handle."""
        ns = len(stages)
        names = (C.c_char_p * ns)(*[s[0].encode() for s in stages])
        flags = (C.c_uint32 * ns)(*[query_flags(normalize, bool(s[1])) for s in stages])
        ks = (C.c_int * ns)(*[int(s[2]) for s in stages])
        N.check(self._lib.vrag_search_multistage_dev(self._h, ns, names, flags, ks, C.c_void_p(query_dev_ptr), int(n_query_rows),
                                                     None, C.c_void_p(out_scores_dev_ptr), C.c_void_p(out_ids_dev_ptr),
                                                     C.c_void_p(stream)))

    # ------------------------------------------------------------------ building blocks of a collective stage
    HIT_DTYPE = np.dtype([("score", np.float32), ("aux", np.uint32), ("id", np.int64)])   # vrag_hit_t, 16 bytes

    def stage_hits_dev(self, name: str, query_dev_ptr: int, n_query_rows: int, flags: int, cand_dev_ptr: int, n_cand: int,
                       k: int, out_hits_dev_ptr: int, stream: int) -> None:
        """Local scan + local top-k of this shard as k packed entries (the all-gather send buffer)."""
        N.check(self._lib.vrag_stage_hits_dev(self._h, name.encode(), C.c_void_p(query_dev_ptr), int(n_query_rows), int(flags),
                                              C.c_void_p(cand_dev_ptr) if cand_dev_ptr else None, int(n_cand), int(k),
                                              C.c_void_p(out_hits_dev_ptr), C.c_void_p(stream)))

    def allgather_topk(self, local_dev_ptr: int, n_lists: int, k: int, gathered_dev_ptr: int, stream: int) -> None:
        N.check(self._lib.vrag_allgather_topk(self._h, C.c_void_p(local_dev_ptr), int(n_lists), int(k),
                                              C.c_void_p(gathered_dev_ptr), C.c_void_p(stream)))

    def merge_hits_dev(self, gathered_dev_ptr: int, n_src: int, n_lists: int, k_src: int, k: int, out_scores_dev_ptr: int,
                       out_ids_dev_ptr: int, flag_dev_ptr: int, stream: int) -> None:
        """Merge n_src gathered lists ([src][list][k_src] packed entries) into the global top-k of every list."""
        N.check(self._lib.vrag_merge_hits_dev(self._h, C.c_void_p(gathered_dev_ptr), int(n_src), int(n_lists), int(k_src), int(k),
                                              C.c_void_p(out_scores_dev_ptr), C.c_void_p(out_ids_dev_ptr),
                                              C.c_void_p(flag_dev_ptr) if flag_dev_ptr else None, C.c_void_p(stream)))

    def allreduce_max_dev(self, scores_dev_ptr: int, n: int, stream: int) -> None:
        N.check(self._lib.vrag_allreduce_max_dev(self._h, C.c_void_p(scores_dev_ptr), int(n), C.c_void_p(stream)))

    # ------------------------------------------------------------------ device-level batched stages (sharded batched search)
    def batch_upload(self, n_stages: int, packed: "PackedQueries", per_stage: bool = False) -> None:
        N.check(self._lib.vrag_batch_upload(self._h, int(n_stages), len(packed) // (n_stages if per_stage else 1),
                                            packed.rows.ctypes.data_as(C.POINTER(C.c_float)),
                                            packed.offsets.ctypes.data_as(C.POINTER(C.c_int)), 1 if per_stage else 0))

    def batch_stage_dev(self, stage: int, name: str, flags: int, k: int, cand_dev_ptr: int, n_cand: int,
                        allow_prefilter: bool, out_scores_dev_ptr: int, out_ids_dev_ptr: int, stream: int) -> None:
        N.check(self._lib.vrag_batch_stage_dev(self._h, int(stage), name.encode(), int(flags), int(k),
                                               C.c_void_p(cand_dev_ptr) if cand_dev_ptr else None, int(n_cand),
                                               1 if allow_prefilter else 0, C.c_void_p(out_scores_dev_ptr),
                                               C.c_void_p(out_ids_dev_ptr) if out_ids_dev_ptr else None, C.c_void_p(stream)))

    def batch_prefilter_failed(self, stream: int) -> bool:
        f = C.c_int()
        N.check(self._lib.vrag_batch_prefilter_failed(self._h, C.c_void_p(stream), C.byref(f)))
        return bool(f.value)

    def topk_batch_dev(self, scores_dev_ptr: int, ids_dev_ptr: int, n: int, k: int, nq: int, out_scores_dev_ptr: int,
                       out_ids_dev_ptr: int, stream: int) -> None:
        N.check(self._lib.vrag_topk_batch_dev(self._h, C.c_void_p(scores_dev_ptr), C.c_void_p(ids_dev_ptr), int(n), int(k),
                                              int(nq), C.c_void_p(out_scores_dev_ptr), C.c_void_p(out_ids_dev_ptr),
                                              C.c_void_p(stream)))

    # ------------------------------------------------------------------ measurement
    def last_timing_ms(self) -> Tuple[float, float]:
        buf = (C.c_float * 2)()
        N.check(self._lib.vrag_last_timing(self._h, buf, 2))
        return float(buf[0]), float(buf[1])

    def launch_count(self) -> int:
        return int(self._lib.vrag_launch_count(self._h))
