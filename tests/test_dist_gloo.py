"""world_size-2 gloo test (CPU) of the sharded search exchange: per-shard top-k -> all-gather -> merge,
and the candidate hand-off between stages. The per-shard scoring here is a CPU test double driven by the
oracle (the product's shards are GpuCorpus objects); the exchange logic under test is the product's
visual_rag_b200.distributed.ShardedSearcher, unchanged."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths():
    for p in (ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)


class OracleShard:
    """CPU stand-in for GpuCorpus: same device-pointer-style surface, tensors addressed by data_ptr()."""

    device = None

    def __init__(self, stores, page_base):
        from oracle import maxsim_oracle as MO

        self.MO = MO
        self.stores = stores          # name -> list of [rows,128] fp32 arrays (this shard's pages)
        self.page_base = page_base
        self._t = {}

    def n_pages(self, name):
        return len(self.stores[name])

    def reg(self, *tensors):
        for t in tensors:
            self._t[t.data_ptr()] = t

    def _view(self, ptr, n, dtype):
        base = None
        for p, t in self._t.items():
            if p <= ptr < p + t.numel() * t.element_size() and t.dtype == dtype:
                base = t
                off = (ptr - p) // t.element_size()
                return base.view(-1)[off:off + n]
        raise KeyError(ptr)

    def score_dev(self, name, q_ptr, n_q, flags, cand_ptr, n_items, out_ptr, stream):
        q = self._view(q_ptr, n_q * 128, torch.float32).numpy().reshape(n_q, 128)
        pool = bool(flags & 2)
        out = self._view(out_ptr, n_items, torch.float32)
        store = self.stores[name]
        if cand_ptr:
            ids = self._view(cand_ptr, n_items, torch.int64).numpy() - self.page_base
        else:
            ids = np.arange(n_items)
        qq = q.mean(axis=0, keepdims=True) if pool else q
        for j, i in enumerate(ids):
            out[j] = self.MO.maxsim_score(qq, store[i]) if 0 <= i < len(store) else float("-inf")

    def topk_dev(self, scores_ptr, ids_ptr, id_base, n, k, out_s_ptr, out_i_ptr, stream):
        sc = self._view(scores_ptr, n, torch.float32).numpy() if n else np.zeros((0,), np.float32)
        ids = self._view(ids_ptr, n, torch.int64).numpy() if ids_ptr else np.arange(n) + id_base
        order = np.lexsort((np.arange(n), -sc))[:k]
        os_, oi = self._view(out_s_ptr, k, torch.float32), self._view(out_i_ptr, k, torch.int64)
        os_[:] = float("-inf")
        oi[:] = -1
        os_[:len(order)] = torch.from_numpy(sc[order].copy())
        oi[:len(order)] = torch.from_numpy(ids[order].copy())


    # ---- batched surface (device-level batched stages)
    def batch_upload(self, n_stages, packed, per_stage=False):
        self._batch = [packed.rows[packed.offsets[b]:packed.offsets[b + 1]] for b in range(len(packed))]

    def batch_stage_dev(self, stage, name, flags, k, cand_ptr, n_cand, allow_prefilter, out_s_ptr, out_i_ptr, stream):
        nq = len(self._batch)
        store = self.stores[name]
        pool = bool(flags & 2)
        cand = self._view(cand_ptr, nq * n_cand, torch.int64).numpy().reshape(nq, n_cand) if cand_ptr else None
        if k == 0:   # raw candidate scores
            raw = self._view(out_s_ptr, nq * n_cand, torch.float32).view(nq, n_cand)
            for b, q in enumerate(self._batch):
                qq = q.mean(axis=0, keepdims=True) if pool else q
                for j, i in enumerate(cand[b] - self.page_base):
                    raw[b, j] = self.MO.maxsim_score(qq, store[i]) if 0 <= i < len(store) else float("-inf")
            return
        os_, oi = self._view(out_s_ptr, nq * k, torch.float32).view(nq, k), self._view(out_i_ptr, nq * k, torch.int64).view(nq, k)
        for b, q in enumerate(self._batch):
            qq = q.mean(axis=0, keepdims=True) if pool else q
            ids = (cand[b] - self.page_base) if cand is not None else np.arange(len(store))
            sc = np.array([self.MO.maxsim_score(qq, store[i]) if 0 <= i < len(store) else float("-inf") for i in ids], np.float32)
            gid = cand[b] if cand is not None else np.arange(len(store)) + self.page_base
            order = np.lexsort((np.arange(len(sc)), -sc))[:k]
            os_[b] = float("-inf")
            oi[b] = -1
            os_[b, :len(order)] = torch.from_numpy(sc[order].copy())
            oi[b, :len(order)] = torch.from_numpy(np.asarray(gid)[order].copy())

    def batch_prefilter_failed(self, stream):
        return False

    def topk_batch_dev(self, scores_ptr, ids_ptr, n, k, nq, out_s_ptr, out_i_ptr, stream):
        sc = self._view(scores_ptr, nq * n, torch.float32).numpy().reshape(nq, n)
        ids = self._view(ids_ptr, nq * n, torch.int64).numpy().reshape(nq, n)
        os_, oi = self._view(out_s_ptr, nq * k, torch.float32).view(nq, k), self._view(out_i_ptr, nq * k, torch.int64).view(nq, k)
        for b in range(nq):
            order = np.lexsort((np.arange(n), -sc[b]))[:k]
            os_[b] = torch.from_numpy(sc[b][order].copy())
            oi[b] = torch.from_numpy(ids[b][order].copy())


def _worker(rank, world, port, ret):
    _paths()
    import cases as CS
    from oracle import maxsim_oracle as MO
    from visual_rag_b200.distributed import ShardedSearcher, shard_page_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 61
        q = CS.query_rows(31, 12)
        initial = [CS.unit_rows(1000 + i, 40 + (i * 7) % 50, scale=True) for i in range(n)]
        pooled = [d[: 8] for d in initial]
        glob = [d.mean(axis=0, keepdims=True) for d in initial]
        b, e = shard_page_range(n, rank, world)
        shard = OracleShard({"initial": initial[b:e], "mean_pooling": pooled[b:e], "global_pooling": glob[b:e]}, b)
        s = ShardedSearcher(shard)
        orig = s._buf

        def buf(key, m, dtype):
            t = orig(key, m, dtype)
            shard.reg(s._bufs[key])
            return t

        s._buf = buf
        orig_sb = s._score_buf

        def sbuf(m):
            t = orig_sb(m)
            shard.reg(s._scores)
            return t

        s._score_buf = sbuf
        shard.reg(s._q_dev)
        # exhaustive
        sc, ids = s.search("initial", q, 10)
        want = MO.search_exhaustive(q, initial, 10)
        assert ids.tolist() == [i for i, _ in want]
        np.testing.assert_allclose(sc, [x for _, x in want], rtol=1e-6)
        # two-stage with reference semantics: GLOBAL top-prefetch_k, then rerank
        st = s.search_multistage([("mean_pooling", False, 20), ("initial", False, 5)], q)
        ref = MO.multistage(q, [(pooled, False, 20), (initial, False, 5)])
        assert st[0][1].tolist() == [i for i, _ in ref[0]]
        assert st[1][1].tolist() == [i for i, _ in ref[1]]
        # three-stage, k larger than a shard
        st = s.search_multistage([("global_pooling", True, 50), ("mean_pooling", False, 45), ("initial", False, 7)], q)
        ref = MO.multistage(q, [(glob, True, 50), (pooled, False, 45), (initial, False, 7)])
        for a, r in zip(st, ref):
            assert a[1].tolist() == [i for i, _ in r]
        # k larger than the whole corpus
        sc, ids = s.search("initial", q, 100)
        assert len(ids) == n and sorted(ids.tolist()) == list(range(n))
        # batched three-stage: one all-gather per stage for the whole batch
        qs = [CS.query_rows(40 + j, 5 + 3 * j) for j in range(4)]
        stages = [("global_pooling", True, 50), ("mean_pooling", False, 45), ("initial", False, 7)]
        got = s.search_multistage_batch(stages, qs)
        for j, qq in enumerate(qs):
            ref = MO.multistage(qq, [(glob, True, 50), (pooled, False, 45), (initial, False, 7)])
            for (gsc, gid), r in zip(got, ref):
                keep = gid[j] >= 0
                assert gid[j][keep].tolist() == [i for i, _ in r]
                np.testing.assert_allclose(gsc[j][keep], [x for _, x in r], rtol=1e-6)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    _paths()
    from visual_rag_b200.distributed import shard_page_range

    for n in (0, 1, 7, 1000, 4_000_000):
        for world in (1, 2, 3, 8):
            rs = [shard_page_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1


@pytest.mark.timeout(300)
def test_sharded_search_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
