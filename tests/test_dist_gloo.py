"""world_size-2 / 3 gloo tests (CPU) of the N>1 host path: the product's ShardedCorpusClient and retriever classes run
SPMD over a page-sharded corpus and must return, on every rank, exactly what the oracle computes on the whole corpus.

The per-shard scoring and the exchange are a CPU test double (tests/sharded_double.py: oracle scores, the C ABI's
collective protocol over gloo); everything above it — id / payload tables across ranks, rank-local filter evaluation,
candidate hand-off between stages, result building — is the product code, unchanged. The CUDA exchange itself is checked
on hardware by bench.py at N > 1 (`sharded_parity`)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths():
    for p in (ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _corpus(n):
    import cases as CS

    initial = [CS.unit_rows(1000 + i, 40 + (i * 7) % 50, scale=True) for i in range(n)]
    initial[n - 2] = initial[3].copy()           # an exact cross-shard tie: page n-2 (last shard) == page 3 (first shard)
    pooled = [d[:8] for d in initial]
    exper = [d[4:14] for d in initial]
    glob = [d.mean(axis=0, keepdims=True) for d in initial]
    payloads = [{"page": i, "year": 2000 + i % 3, "source": "a" if i % 2 else "b", "dataset": f"d{i % 4}"} for i in range(n)]
    ids = [f"pt-{i:04d}" for i in range(n)]
    return {"initial": initial, "mean_pooling": pooled, "experimental_pooling": exper, "global_pooling": glob}, payloads, ids


def _same(got, want_pairs, ids, rtol=1e-6):
    assert [g["id"] for g in got] == [ids[i] for i, _ in want_pairs], ([g["id"] for g in got], [ids[i] for i, _ in want_pairs])
    np.testing.assert_allclose([g["score_final"] for g in got], [s for _, s in want_pairs], rtol=rtol)


def _worker(rank, world, port, ret):
    _paths()
    import cases as CS
    from oracle import maxsim_oracle as MO
    from sharded_double import OracleShardedCorpus
    from visual_rag_b200.client import ShardedCorpusClient
    from visual_rag_b200.distributed import shard_page_range
    from visual_rag_b200.retrieval import (MultiVectorRetriever, SingleStageRetriever, ThreeStageRetriever,
                                           TwoStageRetriever)
    from visual_rag_b200.retrieval import models as M

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 61
        full, payloads, ids = _corpus(n)
        b, e = shard_page_range(n, rank, world)
        shard = OracleShardedCorpus({k: v[b:e] for k, v in full.items()}, b)
        client = ShardedCorpusClient(shard, "c", point_ids=ids[b:e], payloads=payloads[b:e])
        assert client.world == world and client.get_collection("c").points_count == n
        q = CS.query_rows(31, 12)
        qm = q.mean(axis=0, keepdims=True)

        # ---- single stage, all strategies (a8) == exhaustive oracle over the WHOLE corpus, incl. the cross-shard tie
        single = SingleStageRetriever(client, "c")
        table = {"multi_vector": ("initial", q), "tiles_maxsim": ("mean_pooling", q), "pooled_tile": ("mean_pooling", qm),
                 "pooled_global": ("global_pooling", qm), "experimental_maxsim": ("experimental_pooling", q),
                 "pooled_experimental": ("experimental_pooling", qm)}
        for strat, (store, qq) in table.items():
            got = single.search(q, top_k=10, strategy=strat)
            _same(got, MO.search_exhaustive(qq, full[store], 10), ids)
            assert all(g["payload"] == payloads[ids.index(g["id"])] for g in got)      # payloads of foreign shards too
        all_ids = [g["id"] for g in single.search(q, top_k=n, strategy="multi_vector")]
        assert all_ids.index(ids[3]) + 1 == all_ids.index(ids[n - 2])                   # tie -> lower global page first
        assert len(single.search(q, top_k=500, strategy="multi_vector")) == n            # k larger than the corpus

        # ---- two stage, server side (a7): GLOBAL top-prefetch_k, then rerank
        two = TwoStageRetriever(client, "c")
        for mode, (store, pool) in {"pooled_query_vs_standard_pooling": ("mean_pooling", True),
                                    "tokens_vs_standard_pooling": ("mean_pooling", False),
                                    "tokens_vs_experimental_pooling": ("experimental_pooling", False),
                                    "pooled_query_vs_global": ("global_pooling", True)}.items():
            got = two.search_server_side(q, top_k=5, prefetch_k=20, stage1_mode=mode)
            ref = MO.multistage(q, [(full[store], pool, 20), (full["initial"], False, 5)])
            _same(got, ref[1], ids)
        # client-side flow (a5/a6) with embeddings fetched from the owning shards
        got = two.search(q, top_k=5, prefetch_k=20, stage1_mode="tokens_vs_standard_pooling", return_embeddings=True)
        ref = MO.multistage(q, [(full["mean_pooling"], False, 20), (full["initial"], False, 5)])
        _same(got, ref[1], ids)
        for g in got:
            np.testing.assert_array_equal(g["embedding"], full["initial"][ids.index(g["id"])].astype(np.float16).astype(np.float32))

        # ---- three stage (a9) + batch, k larger than a shard
        three = ThreeStageRetriever(client, "c")
        got = three.search_server_side(query_embedding=q, top_k=7, stage1_k=50, stage2_k=45)
        ref = MO.multistage(q, [(full["global_pooling"], True, 50), (full["experimental_pooling"], False, 45), (full["initial"], False, 7)])
        _same(got, ref[2], ids)
        s1 = dict(ref[0])
        assert all(abs(g["score_stage1"] - s1[ids.index(g["id"])]) < 1e-6 for g in got)
        qs = [CS.query_rows(40 + j, 5 + 3 * j) for j in range(4)]
        batch = three.search_server_side_batch(query_embeddings=qs, top_k=7, stage1_k=50, stage2_k=45)
        for qq, got in zip(qs, batch):
            ref = MO.multistage(qq, [(full["global_pooling"], True, 50), (full["experimental_pooling"], False, 45), (full["initial"], False, 7)])
            _same(got, ref[2], ids)
        mv = MultiVectorRetriever("c", qdrant_client=client)
        _same(mv.search_embedded(query_embedding=q, top_k=5, mode="two_stage", prefetch_k=20),
              MO.multistage(q, [(full["mean_pooling"], True, 20), (full["initial"], False, 5)])[1], ids)
        for qq, got in zip(qs, mv.search_embedded_batch(query_embeddings=qs, top_k=5, mode="two_stage", prefetch_k=20)):
            _same(got, MO.multistage(qq, [(full["mean_pooling"], True, 20), (full["initial"], False, 5)])[1], ids)

        # ---- payload filters (8f-3): every rank evaluates its own pages; ranking == oracle restricted to the filtered set
        def restricted(pred, k, stages):
            keep = [i for i in range(n) if pred(payloads[i])]
            sub = [([st[i] for i in keep], pool, kk) for st, pool, kk in stages]
            res = MO.multistage(q, sub)[-1]
            return [(keep[i], s) for i, s in res][:k]

        f_year = two.build_filter(year=2001)
        _same(single.search(q, top_k=8, strategy="multi_vector", filter_obj=f_year),
              restricted(lambda p: p["year"] == 2001, 8, [(full["initial"], False, 8)]), ids)
        _same(two.search_server_side(q, top_k=4, prefetch_k=12, filter_obj=f_year, stage1_mode="tokens_vs_standard_pooling"),
              restricted(lambda p: p["year"] == 2001, 4, [(full["mean_pooling"], False, 12), (full["initial"], False, 4)]), ids)
        f_mix = M.Filter(must=[M.FieldCondition(key="year", match=M.MatchAny(any=[2000, 2002]))],
                         must_not=[M.FieldCondition(key="source", match=M.MatchValue(value="a"))],
                         should=[M.FieldCondition(key="dataset", match=M.MatchValue(value="d0")),
                                 M.FieldCondition(key="page", range=M.Range(gte=30))])
        pred = lambda p: p["year"] in (2000, 2002) and p["source"] != "a" and (p["dataset"] == "d0" or p["page"] >= 30)  # noqa: E731
        _same(single.search(q, top_k=9, strategy="multi_vector", filter_obj=f_mix),
              restricted(pred, 9, [(full["initial"], False, 9)]), ids)
        _same(three.search_server_side(query_embedding=q, top_k=3, stage1_k=12, stage2_k=6, filter_obj=f_year),
              restricted(lambda p: p["year"] == 2001, 3, [(full["global_pooling"], True, 12), (full["experimental_pooling"], False, 6),
                                                          (full["initial"], False, 3)]), ids)
        f_none = two.build_filter(year=1900)                                     # nothing passes on any rank
        assert single.search(q, top_k=5, strategy="multi_vector", filter_obj=f_none) == []
        f_one_rank = M.Filter(must=[M.FieldCondition(key="page", range=M.Range(lt=5))])   # only rank 0 owns candidates
        _same(single.search(q, top_k=3, strategy="multi_vector", filter_obj=f_one_rank),
              restricted(lambda p: p["page"] < 5, 3, [(full["initial"], False, 3)]), ids)
        with pytest.raises(NotImplementedError):
            single.search(q, top_k=3, filter_obj=M.Filter(must=[M.FieldCondition(key="year")]))

        # ---- retrieve(): ids of any shard, vectors fetched from the owner
        pts = client.retrieve("c", ids=[ids[1], ids[n - 1], "missing"], with_payload=True, with_vectors=["mean_pooling"])
        assert [p.id for p in pts] == [ids[1], ids[n - 1]]
        np.testing.assert_array_equal(np.asarray(pts[1].vector["mean_pooling"], np.float32),
                                      full["mean_pooling"][n - 1].astype(np.float16).astype(np.float32))
        assert pts[0].payload == payloads[1]
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    _paths()
    from visual_rag_b200.distributed import owner_of_page, shard_page_range

    for n in (0, 1, 7, 1000, 4_000_000):
        for world in (1, 2, 3, 8):
            rs = [shard_page_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1
            for page in {0, n // 3, n // 2, n - 1} - {-1}:
                if 0 <= page < n:
                    r = owner_of_page(page, n, world)
                    assert rs[r][0] <= page < rs[r][1]
            assert owner_of_page(n, n, world) == -1


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_client_and_retrievers_gloo(world):
    port = 29500 + (os.getpid() % 2000) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
