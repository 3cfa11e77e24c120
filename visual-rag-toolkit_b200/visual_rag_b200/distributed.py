"""Page-sharded multi-GPU search (one process per GPU, torch.distributed / NCCL over NVLink).

Every rank owns a contiguous page range of every named store and scans only its shard.
The only exchange on the path is the top-k merge: each rank contributes its local top-k (fp32 score, int64
global page id) to ONE all-gather per stage, after which every rank runs the same deterministic merge
(score descending, ties -> lower id) with the library's top-k kernel. Multi-stage search keeps the
reference semantics "global top-prefetch_k, then rerank": the merged stage-s list is the candidate list of
stage s+1 on every rank, and a rank scores only the candidates it owns (the others come back as -inf).
No data-path collective touches the corpus itself.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .corpus import GpuCorpus, PackedQueries, _as_f32_query, pack_queries, query_flags


def shard_page_range(n_pages_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous page range [begin, end) owned by `rank`: rank r owns
    [r*N/P, (r+1)*N/P) with integer floors, so ranges are disjoint, ordered by rank and cover [0, N)."""
    return (n_pages_total * rank) // world, (n_pages_total * (rank + 1)) // world


class ShardedSearcher:
    """`corpus` is a GpuCorpus (device pointers + CUDA stream). Anything exposing the same
    score_dev / topk_dev / n_pages / page_base / device surface works; device=None means host tensors
    (used by the gloo tests of the exchange logic)."""

    def __init__(self, corpus: GpuCorpus, group=None, max_query_rows: int = 128):
        self.corpus = corpus
        self.group = group
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.on_gpu = corpus.device is not None
        self.device = torch.device("cuda", corpus.device) if self.on_gpu else torch.device("cpu")
        self._q_dev = torch.empty((max_query_rows, 128), dtype=torch.float32, device=self.device)
        self._q_pin = torch.empty((max_query_rows, 128), dtype=torch.float32)
        if self.on_gpu:
            self._q_pin = self._q_pin.pin_memory()
        self._scores: Optional[torch.Tensor] = None
        self._bufs = {}
        self._h_out = None      # pinned (scores, ids) staging of the host-facing search

    # ------------------------------------------------------------------ helpers
    def _buf(self, key: str, n: int, dtype) -> torch.Tensor:
        t = self._bufs.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty((max(n, 1),), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t[:n]

    def _score_buf(self, n: int) -> torch.Tensor:
        if self._scores is None or self._scores.numel() < n:
            self._scores = torch.empty((max(n, 1),), dtype=torch.float32, device=self.device)
        return self._scores[:n]

    def upload_query(self, query) -> int:
        """Host query -> device (pinned staging, async H2D on the current stream). Returns the row count."""
        q = _as_f32_query(query)
        n = q.shape[0]
        if n > self._q_pin.shape[0]:   # long queries (> 128 tokens are scored in row chunks by the library)
            self._q_dev = torch.empty((n, 128), dtype=torch.float32, device=self.device)
            self._q_pin = torch.empty((n, 128), dtype=torch.float32, pin_memory=self._q_pin.is_pinned())
        self._q_pin[:n].copy_(torch.from_numpy(q))
        self._q_dev[:n].copy_(self._q_pin[:n], non_blocking=True)
        return n

    def _stage(self, name: str, n_q: int, flags: int, cand: Optional[torch.Tensor], k: int,
               tag: str, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One stage on the device: local scan -> local top-k -> all-gather -> merged global top-k.
        Returns (scores[k], global ids[k]) identical on every rank; invalid slots are (-inf, -1)."""
        c = self.corpus
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.on_gpu else 0
        n_items = int(cand.numel()) if cand is not None else c.n_pages(name)
        scores = self._score_buf(n_items)
        if n_items > 0:
            c.score_dev(name, self._q_dev.data_ptr(), n_q, flags, cand.data_ptr() if cand is not None else 0,
                        n_items, scores.data_ptr(), stream)
        if cand is not None and self.world > 1:
            # candidate stage: every candidate is owned by exactly one shard (-inf elsewhere) -> one max-all-reduce of
            # the score vector, then the same top-k on every rank: ties keep candidate order, exactly like one shard
            self.dist.all_reduce(scores, op=self.dist.ReduceOp.MAX, group=self.group)
            ms, mi = out if out is not None else (self._buf(tag + "_ms", k, torch.float32), self._buf(tag + "_mi", k, torch.int64))
            c.topk_dev(scores.data_ptr(), cand.data_ptr(), 0, n_items, k, ms.data_ptr(), mi.data_ptr(), stream)
            return ms, mi
        if self.world == 1 and out is not None:
            ls, li = out
        else:
            ls = self._buf(tag + "_ls", k, torch.float32)
            li = self._buf(tag + "_li", k, torch.int64)
        c.topk_dev(scores.data_ptr(), cand.data_ptr() if cand is not None else 0, c.page_base, n_items, k,
                   ls.data_ptr(), li.data_ptr(), stream)
        if self.world == 1:
            return ls, li
        gs = self._buf(tag + "_gs", k * self.world, torch.float32)
        gi = self._buf(tag + "_gi", k * self.world, torch.int64)
        self.dist.all_gather_into_tensor(gs, ls, group=self.group)
        self.dist.all_gather_into_tensor(gi, li, group=self.group)
        ms, mi = out if out is not None else (self._buf(tag + "_ms", k, torch.float32), self._buf(tag + "_mi", k, torch.int64))
        # merge: keys are (score, position in the gathered list); rank-major gather order + per-rank id order
        # make "lower position" == "lower global id" among equal scores of different ranks only if shards are
        # id-ordered by rank, which contiguous page ranges guarantee.
        c.topk_dev(gs.data_ptr(), gi.data_ptr(), 0, k * self.world, k, ms.data_ptr(), mi.data_ptr(), stream)
        return ms, mi

    # ------------------------------------------------------------------ public API
    def search_multistage_device(self, stages: Sequence[Tuple[str, bool, int]], n_q: int,
                                 normalize: bool = True, packed_out: bool = False) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """All stages on the device, no host synchronisation. The query must already be uploaded.
        packed_out: the stages' merged lists are views into ONE scores and ONE ids tensor (stage s at
        [sum(k[:s]), sum(k[:s+1]))), so the host needs two copies for the whole search instead of two per stage."""
        out = []
        cand = None
        total = sum(int(k) for _, _, k in stages)
        all_s = self._buf("out_s", total, torch.float32) if packed_out else None
        all_i = self._buf("out_i", total, torch.int64) if packed_out else None
        off = 0
        for s, (name, pool, k) in enumerate(stages):
            k = int(k)
            dst = (all_s[off:off + k], all_i[off:off + k]) if packed_out else None
            sc, ids = self._stage(name, n_q, query_flags(normalize, pool), cand, k, f"s{s}", dst)
            out.append((sc, ids))
            cand = ids
            off += k
        return out

    def search_multistage(self, stages: Sequence[Tuple[str, bool, int]], query,
                          normalize: bool = True) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Host-facing: uploads the query, runs the stages, reads back every stage's merged list."""
        n_q = self.upload_query(query)
        self.search_multistage_device(stages, n_q, normalize, packed_out=True)
        total = sum(int(k) for _, _, k in stages)
        if not self.on_gpu:
            s_all, i_all = self._bufs["out_s"][:total].numpy().copy(), self._bufs["out_i"][:total].numpy().copy()
        else:
            # two async copies into pinned staging and ONE synchronisation for all stages
            if self._h_out is None or self._h_out[0].numel() < total:
                self._h_out = (torch.empty((max(total, 1),), dtype=torch.float32).pin_memory(),
                               torch.empty((max(total, 1),), dtype=torch.int64).pin_memory())
            hs, hi = self._h_out[0][:total], self._h_out[1][:total]
            hs.copy_(self._bufs["out_s"][:total], non_blocking=True)
            hi.copy_(self._bufs["out_i"][:total], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            s_all, i_all = hs.numpy().copy(), hi.numpy().copy()
        res, off = [], 0
        for _, _, k in stages:
            s, i = s_all[off:off + int(k)], i_all[off:off + int(k)]
            keep = (i >= 0) & np.isfinite(s)
            res.append((s[keep], i[keep]))
            off += int(k)
        return res

    def search(self, name: str, query, k: int, normalize: bool = True, pool_query: bool = False):
        return self.search_multistage([(name, pool_query, k)], query, normalize)[0]

    # ------------------------------------------------------------------ batched queries
    def search_multistage_batch(self, stages: Sequence[Tuple[str, bool, int]], queries, normalize: bool = True):
        """A batch of independent multi-stage searches over the sharded corpus (BASELINE configs[2] on several GPUs):
        every rank uploads the same queries, scores each stage on its shard for ALL queries in one or two launches
        (dense stage 0 with the fused top-k prefilter, candidate stages with the operand-switching kernel), and the
        per-shard lists of the whole batch travel in ONE all-gather per stage ([n_queries, k] scores + ids) before the
        batched deterministic merge. Returns per stage (scores [nq,k], ids [nq,k]) numpy arrays, identical on every
        rank; invalid slots are (-inf, -1)."""
        packed = queries if isinstance(queries, PackedQueries) else pack_queries(queries)
        nq = len(packed)
        if nq == 0:
            return [(np.empty((0, int(k)), np.float32), np.empty((0, int(k)), np.int64)) for _, _, k in stages]
        c = self.corpus
        c.batch_upload(len(stages), packed)
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.on_gpu else 0
        for allow_prefilter in (True, False):
            dev = []
            cand = None
            for s, (name, pool, k) in enumerate(stages):
                k = int(k)
                ls = self._buf(f"b{s}_ls", nq * k, torch.float32)
                li = self._buf(f"b{s}_li", nq * k, torch.int64)
                if cand is not None and self.world > 1:
                    # candidate stage across shards: raw scores of every candidate (-inf where another shard owns the
                    # page), ONE max-all-reduce for the whole batch, then the batched top-k in candidate order
                    n_cand = int(cand.numel() // nq)
                    raw = self._buf(f"b{s}_raw", nq * n_cand, torch.float32)
                    c.batch_stage_dev(s, name, query_flags(normalize, pool), 0, cand.data_ptr(), n_cand, False,
                                      raw.data_ptr(), 0, stream)
                    self.dist.all_reduce(raw, op=self.dist.ReduceOp.MAX, group=self.group)
                    c.topk_batch_dev(raw.data_ptr(), cand.data_ptr(), n_cand, k, nq, ls.data_ptr(), li.data_ptr(), stream)
                    dev.append((ls, li, k))
                    cand = li
                    continue
                c.batch_stage_dev(s, name, query_flags(normalize, pool), k, cand.data_ptr() if cand is not None else 0,
                                  int(cand.numel() // nq) if cand is not None else 0, allow_prefilter, ls.data_ptr(),
                                  li.data_ptr(), stream)
                if self.world == 1:
                    ms, mi = ls, li
                else:
                    gs = self._buf(f"b{s}_gs", nq * k * self.world, torch.float32)
                    gi = self._buf(f"b{s}_gi", nq * k * self.world, torch.int64)
                    self.dist.all_gather_into_tensor(gs, ls, group=self.group)
                    self.dist.all_gather_into_tensor(gi, li, group=self.group)
                    # [world][nq][k] -> [nq][world*k]: rank-major inside every query row, so that "lower position" is
                    # "lower global id" among equal scores of different shards (contiguous page ranges per rank)
                    ps = self._buf(f"b{s}_ps", nq * k * self.world, torch.float32)
                    pi = self._buf(f"b{s}_pi", nq * k * self.world, torch.int64)
                    ps.view(nq, self.world, k).copy_(gs.view(self.world, nq, k).permute(1, 0, 2))
                    pi.view(nq, self.world, k).copy_(gi.view(self.world, nq, k).permute(1, 0, 2))
                    ms = self._buf(f"b{s}_ms", nq * k, torch.float32)
                    mi = self._buf(f"b{s}_mi", nq * k, torch.int64)
                    c.topk_batch_dev(ps.data_ptr(), pi.data_ptr(), k * self.world, k, nq, ms.data_ptr(), mi.data_ptr(), stream)
                dev.append((ms, mi, k))
                cand = mi
            failed = c.batch_prefilter_failed(stream) if allow_prefilter else False
            if self.world > 1 and allow_prefilter:
                flag = torch.tensor([1 if failed else 0], dtype=torch.int32, device=self.device)
                self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
                failed = bool(flag.item())
            if not failed:
                break
        return [(ms.view(nq, k).cpu().numpy(), mi.view(nq, k).cpu().numpy()) for ms, mi, k in dev]
