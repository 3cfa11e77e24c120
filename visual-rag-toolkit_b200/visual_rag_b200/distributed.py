"""Page-sharded multi-GPU search (one process per GPU).

Every rank owns a contiguous page range of every named store and scans only its shard. The exchange itself lives in
the C ABI (include/vrag_b200.h, multi-GPU section): after `GpuCorpus.comm_init` the handle's search calls are collective
— per stage ONE message per rank, the local top-k as packed 16-byte (score, flags, id) entries written by the top-k kernel
straight into the send buffer and all-gathered, or one max-all-reduce of the candidate scores for a stage that is
restricted to the previous stage's survivors — and every rank receives the same merged global lists (score descending,
ties -> lower global page id). Multi-stage search keeps the reference semantics "global top-prefetch_k, then rerank"
(two_stage.py:161-178). No data-path collective touches the corpus itself.

This module holds the host-side helpers around that: the page partition and `ShardedSearcher`, a small convenience
wrapper that keeps a query resident on the device for back-to-back device-timed searches (bench.py's `value` loop).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .corpus import GpuCorpus, PackedQueries, _as_f32_query, pack_queries


def shard_page_range(n_pages_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous page range [begin, end) owned by `rank`: rank r owns
    [r*N/P, (r+1)*N/P) with integer floors, so ranges are disjoint, ordered by rank and cover [0, N)."""
    return (n_pages_total * rank) // world, (n_pages_total * (rank + 1)) // world


def owner_of_page(page: int, n_pages_total: int, world: int) -> int:
    """Rank whose range contains global page id `page` (inverse of shard_page_range)."""
    if page < 0 or page >= n_pages_total:
        return -1
    r = min(world - 1, (page * world) // max(n_pages_total, 1))
    while shard_page_range(n_pages_total, r, world)[0] > page:
        r -= 1
    while shard_page_range(n_pages_total, r, world)[1] <= page:
        r += 1
    return r


def init_sharded_corpus(corpus: GpuCorpus, group=None) -> GpuCorpus:
    """Join `corpus` (this rank's shard) to the communicator of the torch.distributed group it runs in: the 128-byte
    id of rank 0 travels through the group (any backend), the communicator itself belongs to the library."""
    if corpus.world == 1 and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size(group) > 1:
        corpus.comm_init_torch(group)
    return corpus


class ShardedSearcher:
    """Searches over a page-sharded corpus with the query kept on the device. `corpus` is this rank's GpuCorpus; when a
    torch.distributed group with more than one rank is initialised and the handle has no communicator yet, it joins
    one (init_sharded_corpus)."""

    def __init__(self, corpus: GpuCorpus, group=None, max_query_rows: int = 128):
        self.corpus = init_sharded_corpus(corpus, group)
        self.world, self.rank = corpus.world, corpus.rank
        self.device = torch.device("cuda", corpus.device)
        self._q_dev = torch.empty((max_query_rows, 128), dtype=torch.float32, device=self.device)
        self._q_pin = torch.empty((max_query_rows, 128), dtype=torch.float32).pin_memory()
        self._out = None

    def upload_query(self, query) -> int:
        """Host query -> device (pinned staging, async H2D on the current stream). Returns the row count."""
        q = _as_f32_query(query)
        n = q.shape[0]
        if n > self._q_pin.shape[0]:   # long queries (> 128 tokens are scored in row chunks by the library)
            self._q_dev = torch.empty((n, 128), dtype=torch.float32, device=self.device)
            self._q_pin = torch.empty((n, 128), dtype=torch.float32).pin_memory()
        self._q_pin[:n].copy_(torch.from_numpy(q))
        self._q_dev[:n].copy_(self._q_pin[:n], non_blocking=True)
        return n

    def search_multistage_device(self, stages: Sequence[Tuple[str, bool, int]], n_q: int,
                                 normalize: bool = True) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """All stages on the device for the uploaded query, enqueued on the current stream, no host synchronisation.
        Returns per stage (scores[k], global ids[k]) device tensors, identical on every rank; unused slots are (-inf, -1)."""
        total = sum(int(k) for _, _, k in stages)
        if self._out is None or self._out[0].numel() < total:
            self._out = (torch.empty((max(total, 1),), dtype=torch.float32, device=self.device),
                         torch.empty((max(total, 1),), dtype=torch.int64, device=self.device))
        sc, ids = self._out
        self.corpus.search_multistage_dev(stages, self._q_dev.data_ptr(), n_q, sc.data_ptr(), ids.data_ptr(),
                                          torch.cuda.current_stream(self.device).cuda_stream, normalize)
        out, off = [], 0
        for _, _, k in stages:
            out.append((sc[off:off + int(k)], ids[off:off + int(k)]))
            off += int(k)
        return out

    # ------------------------------------------------------------------ host-facing (collective) searches
    def search_multistage(self, stages: Sequence[Tuple[str, bool, int]], query,
                          normalize: bool = True) -> List[Tuple[np.ndarray, np.ndarray]]:
        return self.corpus.search_multistage(stages, query, normalize)

    def search(self, name: str, query, k: int, normalize: bool = True, pool_query: bool = False):
        return self.corpus.search(name, query, k, normalize, pool_query)

    def search_multistage_batch(self, stages: Sequence[Tuple[str, bool, int]], queries, normalize: bool = True):
        """A batch of independent multi-stage searches over the sharded corpus (BASELINE configs[2] on several GPUs) in
        one native call: per stage one collective for the WHOLE batch. Returns per stage (scores [nq,k], ids [nq,k])
        numpy arrays, identical on every rank; invalid slots are (-inf, -1)."""
        packed = queries if isinstance(queries, PackedQueries) else pack_queries(queries)
        if len(packed) == 0:
            return [(np.empty((0, int(k)), np.float32), np.empty((0, int(k)), np.int64)) for _, _, k in stages]
        res = self.corpus.search_multistage_batch(stages, packed, normalize, as_arrays=True)
        out = []
        for sc, ids, cnt in res:
            sc, ids = sc.copy(), ids.copy()
            dead = np.arange(sc.shape[1])[None, :] >= np.asarray(cnt)[:, None]
            sc[dead] = -np.inf
            ids[dead] = -1
            out.append((sc, ids))
        return out
