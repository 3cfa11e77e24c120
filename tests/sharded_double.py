"""TEST INFRASTRUCTURE: a CPU stand-in for a comm-initialised GpuCorpus (one shard of a page-sharded corpus).

Scores come from the oracle; the exchange between ranks follows the protocol of the C ABI's collective searches
(include/vrag_b200.h, multi-GPU section) over a torch.distributed gloo group:
  * a stage that scans the shard (or a rank-local candidate list): local top-k as (score, id) entries padded with
    (-inf, -1) -> ONE all-gather -> merge by (score descending, gathered position ascending = rank-major);
  * a stage restricted to the previous stage's merged list: local scores (-inf for foreign pages) -> ONE max-all-reduce
    -> top-k by (score descending, candidate position ascending).
It exposes the GpuCorpus surface the clients / retrievers use, so the product's host logic above the C ABI
(ShardedCorpusClient, the retriever classes) runs unchanged on world_size-2 CPU processes.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from oracle import maxsim_oracle as MO

NEG = np.float32(-np.inf)


def _topk_order(scores: np.ndarray, k: int) -> np.ndarray:
    """positions of the k best scores: descending, ties -> lower position (NaN last)."""
    key = np.where(np.isnan(scores), -np.inf, scores)
    return np.lexsort((np.arange(len(scores)), -key))[:k]


class OracleShardedCorpus:
    device = None

    def __init__(self, stores: Dict[str, List[np.ndarray]], page_base: int, group=None):
        self.stores = stores
        self.page_base = int(page_base)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.collectives = 0

    # ------------------------------------------------------------------ store surface
    def has_store(self, name):
        return name in self.stores

    def n_pages(self, name):
        return len(self.stores[name])

    def store_info(self, name):
        rows = [len(p) for p in self.stores[name]]
        return {"n_pages": len(rows), "total_rows": int(sum(rows)), "fixed_rows": 0, "max_rows": max(rows or [0])}

    def read_page(self, name, local_page):
        return np.asarray(self.stores[name][local_page], dtype=np.float16)

    def page_range(self, name, local_page):
        rows = [len(p) for p in self.stores[name]]
        return int(sum(rows[:local_page])), rows[local_page]

    # ------------------------------------------------------------------ exchange primitives
    def _allgather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        self.collectives += 1
        return out

    def _allreduce_max(self, arr: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return arr
        t = torch.from_numpy(arr.astype(np.float32).copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.collectives += 1
        return t.numpy()

    def _score(self, name, q, pool, global_ids) -> np.ndarray:
        store = self.stores[name]
        qq = q.mean(axis=0, keepdims=True) if pool else q
        out = np.full((len(global_ids),), NEG, dtype=np.float32)
        for j, g in enumerate(global_ids):
            i = int(g) - self.page_base
            if 0 <= i < len(store) and len(store[i]) > 0:
                out[j] = MO.maxsim_score(qq, np.asarray(store[i], dtype=np.float32))
        return out

    def _stage0(self, name, q, pool, k, cand):
        ids = np.arange(self.n_pages(name), dtype=np.int64) + self.page_base if cand is None else np.asarray(cand, dtype=np.int64)
        sc = self._score(name, q, pool, ids)
        order = _topk_order(sc, k)
        loc_s = np.full((k,), NEG, dtype=np.float32)
        loc_i = np.full((k,), -1, dtype=np.int64)
        loc_s[:len(order)] = sc[order]
        loc_i[:len(order)] = ids[order]
        parts = self._allgather((loc_s, loc_i))
        gs = np.concatenate([p[0] for p in parts])
        gi = np.concatenate([p[1] for p in parts])
        key = np.where(gi >= 0, gs, np.nan)                 # padding sorts behind every real entry (even -inf ones)
        valid_first = np.lexsort((np.arange(len(gs)), -np.where(np.isnan(key), -np.inf, key), gi < 0))[:k]
        out_s = np.where(gi[valid_first] >= 0, gs[valid_first], NEG).astype(np.float32)
        return out_s, gi[valid_first]

    def _stage_n(self, name, q, pool, k, cand_ids):
        sc = self._allreduce_max(self._score(name, q, pool, cand_ids))
        order = _topk_order(sc, k)
        out_s = np.full((k,), NEG, dtype=np.float32)
        out_i = np.full((k,), -1, dtype=np.int64)
        out_s[:len(order)] = sc[order]
        out_i[:len(order)] = np.asarray(cand_ids)[order]
        out_s[out_i < 0] = NEG
        return out_s, out_i

    # ------------------------------------------------------------------ GpuCorpus search surface (collective)
    def search_multistage(self, stages, query, normalize=True, stage_queries=None, candidate_ids=None, fp16_query=False):
        out = []
        cand = None
        for s, (name, pool, k) in enumerate(stages):
            q = np.asarray(stage_queries[s] if stage_queries is not None else query, dtype=np.float32)
            q = q[None, :] if q.ndim == 1 else q
            if s == 0:
                sc, ids = self._stage0(name, q, pool, int(k), candidate_ids)
            else:
                sc, ids = self._stage_n(name, q, pool, int(k), cand)
            cand = ids
            n = 0
            while n < len(ids) and ids[n] >= 0:
                n += 1
            out.append((sc[:n].copy(), ids[:n].copy()))
        return out

    def search(self, name, query, k, normalize=True, pool_query=False, candidate_ids=None, fp16_query=False):
        return self.search_multistage([(name, pool_query, k)], query, candidate_ids=candidate_ids)[0]

    def search_multistage_batch(self, stages, queries, normalize=True, stage_queries=None, as_arrays=False,
                                final_only=False, fp16_query=False):
        from visual_rag_b200.corpus import PackedQueries

        if stage_queries is not None:
            nq = len(stage_queries)
        elif isinstance(queries, PackedQueries):
            queries = [queries.rows[queries.offsets[b]:queries.offsets[b + 1]] for b in range(len(queries))]
            nq = len(queries)
        else:
            nq = len(queries)
        if nq == 0:
            return []
        per_query = [self.search_multistage(stages, None if stage_queries is not None else queries[b],
                                            stage_queries=None if stage_queries is None else stage_queries[b])
                     for b in range(nq)]
        ns = len(stages)
        if final_only:
            kl = int(stages[-1][2])
            sc = np.full((nq, kl), NEG, np.float32)
            ids = np.full((nq, kl), -1, np.int64)
            st = np.full((nq, kl, max(ns - 1, 1)), np.nan, np.float32)
            cnt = np.zeros((nq,), np.int32)
            for b, res in enumerate(per_query):
                fs, fi = res[-1]
                cnt[b] = len(fi)
                sc[b, :len(fi)] = fs
                ids[b, :len(fi)] = fi
                for s in range(ns - 1):
                    look = {int(i): float(x) for x, i in zip(*res[s])}
                    for j, i in enumerate(fi):
                        if int(i) in look:
                            st[b, j, s] = look[int(i)]
            return sc, ids, st[:, :, :ns - 1], cnt
        if as_arrays:
            out = []
            for s, (_, _, k) in enumerate(stages):
                sc = np.full((nq, int(k)), NEG, np.float32)
                ids = np.full((nq, int(k)), -1, np.int64)
                cnt = np.zeros((nq,), np.int32)
                for b, res in enumerate(per_query):
                    cnt[b] = len(res[s][1])
                    sc[b, :cnt[b]] = res[s][0]
                    ids[b, :cnt[b]] = res[s][1]
                out.append((sc, ids, cnt))
            return out
        return per_query
