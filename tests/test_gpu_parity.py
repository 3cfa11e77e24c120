"""Parity of the CUDA path (through the C ABI) with the oracle and the committed reference goldens.
Run on a B200: python -m pytest tests -m gpu.

Tolerances: the north star allows 1e-3 relative on scores; the kernels carry the fp32 query as an fp16 hi/lo
pair and accumulate in fp32, so the tests hold them to 2e-5 relative (+1e-6 absolute for near-zero scores).
Top-k ids must be identical (the seeded cases have no ties inside that tolerance)."""
import numpy as np
import pytest

import cases as CS
from oracle import maxsim_oracle as MO

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-5, 2e-6


@pytest.fixture(scope="module")
def corpus():
    from visual_rag_b200.corpus import GpuCorpus

    c = GpuCorpus(0)
    yield c
    c.close()


def rows16(seed, n, scale=True):
    return CS.unit_rows(seed, n, dtype=np.float16, scale=scale)


def close(got, want, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(got, np.float64), np.asarray(want, np.float64), rtol=rtol, atol=atol)


# ------------------------------------------------------------------ a1/a2: compute_maxsim_score / batch
@pytest.mark.parametrize("case", CS.maxsim_cases(), ids=lambda c: c["key"])
def test_maxsim_golden_cases(corpus, case, maxsim_golden):
    q = CS.query_rows(case["seed"], case["q"])
    d = CS.unit_rows(case["seed"] + 1, case["t"], dtype=np.float16, scale=True)
    corpus.add_store("g", d, fixed_rows=case["t"])
    want = maxsim_golden[case["key"]]
    close(corpus.score("g", q)[0], want[0])
    close(corpus.score("g", q, normalize=False)[0], want[1], rtol=1e-4)


def test_bench_corpus_exhaustive_and_two_stage(corpus, maxsim_golden):
    """a3/a4: quick_test.search_exhaustive / search_two_stage on the seeded in-memory corpus."""
    q, docs = CS.bench_corpus()
    rows = np.concatenate(docs).astype(np.float16)
    corpus.add_store("initial", rows, fixed_rows=docs[0].shape[0])
    close(corpus.score("initial", q), maxsim_golden["batch_scores"])
    s, ids = corpus.search("initial", q, 10)
    assert ids.tolist() == maxsim_golden["exhaustive_ids"].tolist()
    close(s, maxsim_golden["exhaustive_scores"])
    pooled = maxsim_golden["batch_pooled"]                      # fp32 tile means produced by the reference
    corpus.add_store("mean_pooling", pooled.reshape(-1, 128), fixed_rows=pooled.shape[1])
    (s1, id1), (s2, id2) = corpus.search_multistage([("mean_pooling", True, 30), ("initial", False, 10)], q)
    assert id2.tolist() == maxsim_golden["two_stage_ids"].tolist()
    close(s2, maxsim_golden["two_stage_scores"])
    rank1 = np.array([id1.tolist().index(i) + 1 for i in id2])
    # the store rounds the reference's fp32 pooled rows to fp16: stage-1 ranks may move by a tie-sized step only
    assert np.abs(rank1 - maxsim_golden["two_stage_rank1"]).max() <= 2


# ------------------------------------------------------------------ shape sweeps vs the oracle
@pytest.mark.parametrize("q_rows", [1, 8, 20, 33, 100])
@pytest.mark.parametrize("tokens", [129, 300, 1030])
def test_large_pages_fixed(corpus, q_rows, tokens):
    n = 37
    q = CS.query_rows(100 + q_rows, q_rows)
    rows = rows16(200 + tokens, n * tokens)
    corpus.add_store("s", rows, fixed_rows=tokens)
    want = [MO.maxsim_score(q, rows[i * tokens:(i + 1) * tokens].astype(np.float32)) for i in range(n)]
    close(corpus.score("s", q), want)


@pytest.mark.parametrize("q_rows", [1, 20, 64])
def test_large_pages_variable_and_candidates(corpus, q_rows):
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 700, size=61)
    lens[3] = 768
    lens[7] = 128
    lens[9] = 129
    off = np.concatenate([[0], np.cumsum(lens)])
    rows = rows16(300 + q_rows, int(off[-1]))
    q = CS.query_rows(400 + q_rows, q_rows)
    corpus.add_store("s", rows, page_offsets=off)
    want = np.array([MO.maxsim_score(q, rows[off[i]:off[i + 1]].astype(np.float32)) for i in range(len(lens))])
    close(corpus.score("s", q), want)
    cand = rng.permutation(len(lens))[:23]
    close(corpus.score("s", q, candidate_ids=cand), want[cand])
    want_nn = [MO.maxsim_score(q, rows[off[i]:off[i + 1]].astype(np.float32), normalize=False) for i in range(len(lens))]
    close(corpus.score("s", q, normalize=False), want_nn, rtol=1e-4)


@pytest.mark.parametrize("q_rows", [1, 20, 40])
@pytest.mark.parametrize("r", [1, 12, 13, 32, 34, 64, 76, 128])
def test_packed_pages_fixed(corpus, q_rows, r):
    """Pooled stores: 12/13 rows (ColSmol tiles), 32/34 (ColPali rows / legacy conv), 76 (ColSmol experimental),
    1 (global_pooling)."""
    rng = np.random.default_rng(r)
    n = 1001
    rows = rows16(500 + r, n * r)
    q = CS.query_rows(600 + q_rows, q_rows)
    corpus.add_store("p", rows, fixed_rows=r)
    want = np.array([MO.maxsim_score(q, rows[i * r:(i + 1) * r].astype(np.float32)) for i in range(n)])
    close(corpus.score("p", q), want)
    for ncand in (1, 3, 77):
        cand = rng.permutation(n)[:ncand]
        close(corpus.score("p", q, candidate_ids=cand), want[cand])
    wantp = [MO.pooled_query_score(q, rows[i * r:(i + 1) * r].astype(np.float32)) for i in range(n)]
    close(corpus.score("p", q, pool_query=True), wantp)


def test_packed_pages_variable(corpus):
    """ColQwen2.5-style pooled store: min(32, H_eff) rows per page."""
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 33, size=777)
    off = np.concatenate([[0], np.cumsum(lens)])
    rows = rows16(700, int(off[-1]))
    q = CS.query_rows(701, 20)
    corpus.add_store("p", rows, page_offsets=off)
    want = np.array([MO.maxsim_score(q, rows[off[i]:off[i + 1]].astype(np.float32)) for i in range(len(lens))])
    close(corpus.score("p", q), want)
    cand = rng.permutation(len(lens))[:50]
    close(corpus.score("p", q, candidate_ids=cand), want[cand])


# ------------------------------------------------------------------ top-k semantics
def test_topk_order_ties_and_padding(corpus):
    n, r = 3000, 4
    rows = rows16(800, n * r)
    rows[r * 17:r * 18] = rows[r * 5:r * 6]       # page 17 duplicates page 5 -> exact tie
    rows[r * 2999:r * 3000] = rows[r * 5:r * 6]
    q = rows[r * 5:r * 6].astype(np.float32)       # makes the tied pages the best ones
    corpus.add_store("t", rows, fixed_rows=r)
    sc = corpus.score("t", q)
    assert sc[5] == sc[17] == sc[2999]
    for k in (1, 2, 3, 10, 256, 1000, 3000, 4096):
        s, ids = corpus.search("t", q, k)
        order = np.lexsort((np.arange(n), -sc))[:k]   # score desc, ties -> lower id (Python's stable sort)
        assert ids.tolist() == order.tolist()
        assert np.array_equal(s, sc[order])
        assert len(ids) == min(k, n)
    assert corpus.search("t", q, 3)[1].tolist() == [5, 17, 2999]


def test_topk_multilevel_large(corpus):
    n = 300_000
    corpus.add_synthetic_store("g", n, fixed_rows=1, seed=1)
    q = CS.query_rows(900, 1)
    sc = corpus.score("g", q)
    d = corpus.read_rows("g", 0, n).astype(np.float32)
    want = (d / (np.linalg.norm(d, axis=1, keepdims=True) + 1e-8)) @ (q[0] / (np.linalg.norm(q[0]) + 1e-8))
    close(sc, want, rtol=1e-4, atol=1e-5)
    for k in (10, 1000, 4096):
        s, ids = corpus.search("g", q, k)
        order = np.lexsort((np.arange(n), -sc))[:k]
        assert ids.tolist() == order.tolist() and np.array_equal(s, sc[order])


def test_topk_radix_select_with_massive_ties(corpus):
    """n > 4096 goes through the radix select; duplicated pages give hundreds of exact ties at the threshold,
    which must be resolved by lower page id (Python's stable sort)."""
    base = rows16(950, 50, scale=False)
    rng = np.random.default_rng(1)
    pick = rng.integers(0, 50, size=20000)
    corpus.add_store("dup", base[pick], fixed_rows=1)
    q = CS.query_rows(951, 3)
    sc = corpus.score("dup", q)
    assert len(np.unique(sc)) <= 50
    for k in (1, 10, 256, 1000, 4096):
        s, ids = corpus.search("dup", q, k)
        order = np.lexsort((np.arange(len(sc)), -sc))[:k]
        assert ids.tolist() == order.tolist()
        assert np.array_equal(s, sc[order])
    allsame = np.repeat(base[:1], 9000, axis=0)
    corpus.add_store("same", allsame, fixed_rows=1)
    s, ids = corpus.search("same", q, 300)
    assert ids.tolist() == list(range(300))


def test_topk_beyond_the_shared_memory_sort(corpus):
    """k > 4096 (the reference accepts any prefetch_k / limit): radix select + global bitonic sort + emit. Exact order
    incl. ties, k between powers of two, k >= n, candidate lists, a multi-stage search whose prefetch is > 4096, and the
    batched call (which runs such a batch query by query)."""
    n = 300_000
    corpus.add_synthetic_store("gb", n, fixed_rows=1, seed=3)
    q = CS.query_rows(901, 1)
    sc = corpus.score("gb", q)
    full = np.lexsort((np.arange(n), -sc))
    for k in (4097, 5000, 8192, 8193, 20000, 70000):
        s, ids = corpus.search("gb", q, k)
        assert len(ids) == k and ids.tolist() == full[:k].tolist() and np.array_equal(s, sc[full[:k]])
    # massive ties at the threshold + k >= n
    base = rows16(952, 50, scale=False)
    pick = np.random.default_rng(2).integers(0, 50, size=20000)
    corpus.add_store("dupb", base[pick], fixed_rows=1)
    q3 = CS.query_rows(953, 3)
    sd = corpus.score("dupb", q3)
    order = np.lexsort((np.arange(len(sd)), -sd))
    for k in (4100, 9000, 19999, 20000, 25000, 40000):
        s, ids = corpus.search("dupb", q3, k)
        m = min(k, len(sd))
        assert len(ids) == m and ids.tolist() == order[:m].tolist() and np.array_equal(s, sd[order[:m]])
    # candidate list (ids in arbitrary order; ties -> lower POSITION in the list, as the reference's stable sort over the list)
    cand = np.random.default_rng(3).permutation(len(sd))[:12000]
    s, ids = corpus.search("dupb", q3, 6000, candidate_ids=cand)
    sdc = corpus.score("dupb", q3, candidate_ids=cand)     # the gather path's own scores (last-ulp differences to the dense scan)
    o = np.lexsort((np.arange(len(cand)), -sdc))[:6000]
    assert ids.tolist() == cand[o].tolist() and np.array_equal(s, sdc[o])
    # two-stage with a 6000-page prefetch == the oracle order on the same scores
    rows = rows16(954, 9000 * 6)
    corpus.add_store("bp", rows.reshape(9000, 6, 128).mean(axis=1).astype(np.float16), fixed_rows=1)
    corpus.add_store("bi", rows, fixed_rows=6)
    qq = CS.query_rows(955, 7)
    s1 = corpus.score("bp", qq)
    pre = np.lexsort((np.arange(9000), -s1))[:6000]
    s2 = corpus.score("bi", qq)
    fin = pre[np.lexsort((np.arange(6000), -s2[pre]))[:10]]
    res = corpus.search_multistage([("bp", False, 6000), ("bi", False, 10)], qq)
    assert res[0][1].tolist() == pre.tolist() and res[1][1].tolist() == fin.tolist()
    assert np.array_equal(res[1][0], s2[fin])
    b = corpus.search_multistage_batch([("bp", False, 6000), ("bi", False, 10)], [qq, q3], as_arrays=True)
    assert b[0][1].shape == (2, 6000) and b[1][1][0].tolist() == fin.tolist() and b[0][2].tolist() == [6000, 6000]
    f_sc, f_id, f_st, f_cnt = corpus.search_multistage_batch([("bp", False, 6000), ("bi", False, 10)], [qq, q3], final_only=True)
    assert f_id[0].tolist() == fin.tolist() and f_cnt.tolist() == [10, 10]
    assert np.array_equal(f_st[0, :, 0], s1[fin])
    for name in ("gb", "dupb", "bp", "bi"):
        corpus.drop_store(name)


# ------------------------------------------------------------------ edge cases
def test_edge_cases(corpus):
    from visual_rag_b200._native import VragError

    rows = rows16(1000, 10 * 40)
    q = CS.query_rows(1001, 5)
    corpus.add_store("e", rows, fixed_rows=40)
    s, ids = corpus.search("e", q, 50)                     # k > n
    assert len(ids) == 10 and sorted(ids.tolist()) == list(range(10))
    s, ids = corpus.search("e", q, 5, candidate_ids=[])    # empty candidate list
    assert len(ids) == 0
    sc = corpus.score("e", q, candidate_ids=[3, 99, -1, 3])  # ids outside the shard score -inf
    assert np.isneginf(sc[1]) and np.isneginf(sc[2]) and sc[0] == sc[3]
    off = np.array([0, 5, 5, 12], dtype=np.int64)          # an empty page
    corpus.add_store("e2", rows[:12], page_offsets=off)
    sc = corpus.score("e2", q)
    assert np.isneginf(sc[1]) and np.isfinite(sc[0]) and np.isfinite(sc[2])
    with pytest.raises(VragError, match="unknown vector store"):
        corpus.score("missing", q)
    with pytest.raises(VragError):
        corpus.search("e", CS.query_rows(1, 1025), 5)      # more query rows than the staging buffers hold (1024)
    with pytest.raises(ValueError):
        corpus.score("e", np.zeros((3, 64), np.float32))
    zero = np.zeros((2 * 8, 128), np.float16)               # all-zero rows: 0/(0+1e-8) = 0, as in numpy
    corpus.add_store("z", zero, fixed_rows=8)
    assert corpus.score("z", q).tolist() == [0.0, 0.0]


def test_page_base_shard_ids():
    from visual_rag_b200.corpus import GpuCorpus

    rows = rows16(1100, 20 * 150)
    q = CS.query_rows(1101, 20)
    with GpuCorpus(0, page_base=1000) as shard:
        shard.add_store("initial", rows, fixed_rows=150)
        s, ids = shard.search("initial", q, 5)
        assert ids.min() >= 1000 and ids.max() < 1020
        sc = shard.score("initial", q, candidate_ids=[1003, 3, 1019])
        assert np.isfinite(sc[0]) and np.isneginf(sc[1]) and np.isfinite(sc[2])


# ------------------------------------------------------------------ size-independent properties at scale
def test_properties_on_large_synthetic_corpus(corpus):
    n, t = 20_000, 1030
    corpus.add_synthetic_store("big", n, fixed_rows=t, seed=42)
    q = CS.query_rows(1200, 20)
    sc = corpus.score("big", q)
    assert np.all(np.isfinite(sc)) and sc.shape == (n,)
    # (1) oracle on a bounded random sample of pages read back from the device
    rng = np.random.default_rng(0)
    for p in rng.choice(n, size=12, replace=False):
        close(sc[p], MO.maxsim_score(q, corpus.read_page("big", int(p)).astype(np.float32)))
    # (2) gather path == scan path, and candidate order does not matter
    cand = rng.permutation(n)[:500]
    g = corpus.score("big", q, candidate_ids=cand)
    assert np.array_equal(g, sc[cand])
    # (3) top-k is sorted, consistent with the score array, and idempotent under restriction to itself
    s, ids = corpus.search("big", q, 100)
    assert np.all(np.diff(s) <= 0) and np.array_equal(s, sc[ids])
    s2, ids2 = corpus.search("big", q, 100, candidate_ids=ids)
    assert np.array_equal(ids2, ids) and np.array_equal(s2, s)
    # (4) MaxSim is additive over query tokens
    a, b = corpus.score("big", q[:7]), corpus.score("big", q[7:])
    close(a + b, sc, rtol=1e-5)
    # (5) a page scored against its own tokens reaches the self-similarity bound Q (unit rows)
    own = corpus.read_page("big", 123).astype(np.float32)[:16]
    close(corpus.score("big", own, candidate_ids=[123])[0], 16.0, rtol=1e-3)
    corpus.drop_store("big")


# ------------------------------------------------------------------ retriever classes on the GPU client
@pytest.fixture(scope="module")
def gpu_client(corpus, retrieval_golden):
    from visual_rag_b200.client import GpuCorpusClient

    q, initial = CS.retrieval_corpus()
    off_i = np.concatenate([[0], np.cumsum([d.shape[0] for d in initial])])
    corpus.add_store("initial", np.concatenate(initial).astype(np.float16), page_offsets=off_i)
    off = retrieval_golden["offsets_pooled"]
    corpus.add_store("mean_pooling", retrieval_golden["mean_pooling"], page_offsets=off)
    corpus.add_store("experimental_pooling", retrieval_golden["experimental_pooling"], page_offsets=off)
    corpus.add_store("global_pooling", retrieval_golden["global_pooling"], fixed_rows=1)
    n = len(initial)
    client = GpuCorpusClient(corpus, "c", point_ids=list(range(n)), payloads=[{"page": i, "year": 2000 + i % 3} for i in range(n)])
    return q, client


def _same(got, want, keys):
    assert [g["id"] for g in got] == [w["id"] for w in want]
    for k in keys:
        close([g[k] for g in got], [w[k] for w in want])


def test_retrievers_match_reference_goldens(gpu_client, golden_index):
    from visual_rag_b200.retrieval import (MultiVectorRetriever, SingleStageRetriever, ThreeStageRetriever,
                                           TwoStageRetriever)

    q, client = gpu_client
    want = golden_index["retrieval"]
    two = TwoStageRetriever(client, "c")
    for mode in ("pooled_query_vs_tiles", "tokens_vs_tiles", "pooled_query_vs_global"):
        _same(two.search(q, top_k=10, prefetch_k=40, stage1_mode=mode), want[f"two_stage_search::{mode}"],
              ("score_stage1", "score_stage2", "score_final"))
    _same(two.search(q, top_k=10, prefetch_k=40, stage1_mode="tokens_vs_tiles", use_reranking=False),
          want["two_stage_search::norerank"], ("score_stage1", "score_final"))
    for key, w in want.items():
        if key.startswith("two_stage_server::"):
            _same(two.search_server_side(q, top_k=10, prefetch_k=40, stage1_mode=key.split("::")[1]), w, ("score_final",))
    for use_pooling in (False, True):
        _same(two.search_single_stage(q, top_k=10, use_pooling=use_pooling), want[f"two_stage_single::{use_pooling}"],
              ("score_final",))
    three = ThreeStageRetriever(client, "c")
    _same(three.search_server_side(query_embedding=q, top_k=10, stage1_k=80, stage2_k=30), want["three_stage"],
          ("score_stage1", "score_stage2", "score_stage3", "score_final"))
    single = SingleStageRetriever(client, "c")
    for strat in ("multi_vector", "tiles_maxsim", "pooled_tile", "pooled_global", "experimental_maxsim", "pooled_experimental"):
        _same(single.search(q, top_k=10, strategy=strat), want[f"single::{strat}"], ("score",))
    mv = MultiVectorRetriever("c", qdrant_client=client)
    _same(mv.search_embedded(query_embedding=q, top_k=10, mode="three_stage", stage1_k=80, stage2_k=30),
          want["three_stage"], ("score_final",))
    res = two.search(q, top_k=3, prefetch_k=40, stage1_mode="tokens_vs_tiles", return_embeddings=True)
    assert res[0]["embedding"].shape[1] == 128 and res[0]["payload"]["page"] == res[0]["id"]


def test_retriever_batch_methods_equal_per_query_calls(gpu_client):
    """search_server_side_batch (one native call) == the per-query retrievers pinned to the reference goldens."""
    from visual_rag_b200.retrieval import MultiVectorRetriever, ThreeStageRetriever, TwoStageRetriever

    q, client = gpu_client
    rng = np.random.default_rng(9)
    queries = [q] + [rng.standard_normal((int(rng.integers(8, 30)), 128)).astype(np.float32) for _ in range(6)]
    three = ThreeStageRetriever(client, "c")
    got = three.search_server_side_batch(query_embeddings=queries, top_k=10, stage1_k=80, stage2_k=30)
    for qq, g in zip(queries, got):
        _same(g, three.search_server_side(query_embedding=qq, top_k=10, stage1_k=80, stage2_k=30),
              ("score_stage1", "score_stage2", "score_stage3", "score_final"))
    two = TwoStageRetriever(client, "c")
    for mode in ("pooled_query_vs_standard_pooling", "tokens_vs_standard_pooling", "pooled_query_vs_global",
                 "tokens_vs_experimental_pooling"):
        got = two.search_server_side_batch(queries, top_k=10, prefetch_k=40, stage1_mode=mode)
        for qq, g in zip(queries, got):
            _same(g, two.search_server_side(qq, top_k=10, prefetch_k=40, stage1_mode=mode), ("score_final",))
    mv = MultiVectorRetriever("c", qdrant_client=client)
    got = mv.search_embedded_batch(query_embeddings=queries[:2], top_k=5, mode="single_full")
    assert [r["id"] for r in got[0]] == [r["id"] for r in mv.search_embedded(query_embedding=queries[0], top_k=5)]
    f = two.build_filter(year=2001)   # filters fall back to the per-query path
    got = two.search_server_side_batch(queries[:2], top_k=5, prefetch_k=40, filter_obj=f)
    assert all(r["payload"]["year"] == 2001 for r in got[0]) and len(got) == 2


def test_payload_and_id_filters(gpu_client):
    from visual_rag_b200.retrieval import TwoStageRetriever
    from visual_rag_b200.retrieval.models import Filter, HasIdCondition

    q, client = gpu_client
    two = TwoStageRetriever(client, "c")
    f = two.build_filter(year=2001)
    res = two.search_server_side(q, top_k=10, prefetch_k=40, filter_obj=f, stage1_mode="tokens_vs_standard_pooling")
    assert len(res) == 10 and all(r["payload"]["year"] == 2001 for r in res)
    allowed = [4, 9, 77, 120]
    res = two.search_single_stage(q, top_k=10, filter_obj=Filter(must=[HasIdCondition(has_id=allowed)]))
    assert sorted(r["id"] for r in res) == allowed
    got = client.retrieve("c", ids=[9, 4], with_vectors=["initial", "global_pooling"])
    assert [p.id for p in got] == [9, 4] and len(got[0].vector["global_pooling"]) == 1


def test_reference_style_client_calls(gpu_client):
    """The exact call shapes of the reference retrievers (two_stage.py:349-358, 162-178) against the client."""
    from visual_rag_b200.retrieval.models import Prefetch, SearchParams

    q, client = gpu_client
    pts = client.query_points(collection_name="c", query=q.mean(axis=0).tolist(), using="mean_pooling",
                              query_filter=None, limit=7, with_payload=True, with_vectors=False, timeout=120).points
    assert len(pts) == 7 and pts[0].score >= pts[-1].score and isinstance(pts[0].score, float)
    pts2 = client.query_points(collection_name="c", query=q.tolist(), using="initial", limit=5, query_filter=None,
                               with_payload=True, search_params=SearchParams(exact=True),
                               prefetch=[Prefetch(query=q.mean(axis=0).tolist(), using="mean_pooling", limit=7)],
                               timeout=120).points
    assert len(pts2) == 5 and {p.id for p in pts2} <= {p.id for p in pts}
    assert client.get_collection("c").points_count == 150


def test_sharded_searcher_single_rank_matches_corpus_api(corpus):
    """The multi-GPU code path (device-pointer API + merge) at world size 1 equals the host API."""
    from visual_rag_b200.distributed import ShardedSearcher

    n = 4000
    corpus.add_synthetic_store("initial", n, fixed_rows=300, seed=5)
    corpus.add_synthetic_store("mean_pooling", n, fixed_rows=32, seed=6)
    q = CS.query_rows(1300, 20)
    s = ShardedSearcher(corpus)
    stages = [("mean_pooling", False, 256), ("initial", False, 10)]
    a = s.search_multistage(stages, q)
    b = corpus.search_multistage(stages, q)
    for (sa, ia), (sb, ib) in zip(a, b):
        assert ia.tolist() == ib.tolist() and np.array_equal(sa, sb)
    sa, ia = s.search("initial", q, 7)
    sb, ib = corpus.search("initial", q, 7)
    assert ia.tolist() == ib.tolist() and np.array_equal(sa, sb)
    q_long = CS.query_rows(1301, 200)                 # longer than the searcher's initial staging (128 rows): it grows
    a = s.search_multistage(stages, q_long)
    b = corpus.search_multistage(stages, q_long)
    for (sa, ia), (sb, ib) in zip(a, b):
        assert ia.tolist() == ib.tolist() and np.allclose(sa, sb, rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# batched queries (BASELINE configs[2]): one native call, stage >= 1 of all queries in one launch
def _ragged_store(rng, n_pages, lo, hi):
    lens = rng.integers(lo, hi + 1, size=n_pages)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rows = rng.standard_normal((int(off[-1]), 128)).astype(np.float16)
    return rows, off, [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(n_pages)]


def _same_ranking(ids, scores, want, rtol=RTOL, atol=ATOL):
    """ids identical to the oracle's, except swaps among oracle scores closer than the score tolerance."""
    want_ids = [i for i, _ in want]
    want_sc = np.asarray([x for _, x in want], np.float64)
    np.testing.assert_allclose(np.asarray(scores, np.float64), want_sc, rtol=rtol, atol=atol)
    got_ids = list(ids)
    if got_ids == want_ids:
        return
    by_id = dict(want)
    for j, (g, w) in enumerate(zip(got_ids, want_ids)):
        if g != w:
            assert g in by_id, f"position {j}: id {g} is not in the oracle list"
            assert abs(by_id[g] - by_id[w]) <= 2 * (atol + rtol * abs(by_id[w])), f"position {j}: {g} vs {w} is not a near-tie"


@pytest.mark.gpu
@pytest.mark.parametrize("n_queries", [1, 7, 40])
def test_batched_three_stage_matches_oracle_and_single_query_path(corpus, n_queries):
    rng = np.random.default_rng(100 + n_queries)
    n = 700
    init_rows, init_off, init_docs = _ragged_store(rng, n, 150, 400)
    exp_rows, exp_off, exp_docs = _ragged_store(rng, n, 16, 32)
    glob = rng.standard_normal((n, 128)).astype(np.float16)
    glob_docs = [glob[i:i + 1].astype(np.float32) for i in range(n)]
    corpus.add_store("b_initial", init_rows, page_offsets=init_off)
    corpus.add_store("b_exp", exp_rows, page_offsets=exp_off)
    corpus.add_store("b_glob", glob, fixed_rows=1)
    queries = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(n_queries)]
    stages = [("b_glob", True, 200), ("b_exp", False, 60), ("b_initial", False, 10)]
    got = corpus.search_multistage_batch(stages, queries)
    assert len(got) == n_queries
    for q, res in zip(queries, got):
        want = MO.multistage(q, [(glob_docs, True, 200), (exp_docs, False, 60), (init_docs, False, 10)])
        single = corpus.search_multistage(stages, q)
        for s in range(3):
            sc, ids = res[s]
            _same_ranking(ids.tolist(), sc, want[s])
            _same_ranking(ids.tolist(), sc, list(zip(single[s][1].tolist(), single[s][0].tolist())), rtol=1e-6, atol=1e-6)
    for nm in ("b_initial", "b_exp", "b_glob"):
        corpus.drop_store(nm)


@pytest.mark.gpu
def test_batched_two_stage_per_stage_queries_and_k_larger_than_corpus(corpus):
    rng = np.random.default_rng(77)
    n = 90
    init = rng.standard_normal((n * 300, 128)).astype(np.float16)
    pooled = rng.standard_normal((n * 32, 128)).astype(np.float16)
    corpus.add_store("b2_initial", init, fixed_rows=300)
    corpus.add_store("b2_pooled", pooled, fixed_rows=32)
    init_docs = [init[i * 300:(i + 1) * 300].astype(np.float32) for i in range(n)]
    pooled_docs = [pooled[i * 32:(i + 1) * 32].astype(np.float32) for i in range(n)]
    queries = [rng.standard_normal((int(rng.integers(5, 40)), 128)).astype(np.float32) for _ in range(9)]
    # the client sends the mean-pooled prefetch vector and the token matrix separately (two_stage.py:142,159)
    sq = [[q.mean(axis=0, keepdims=True), q] for q in queries]
    stages = [("b2_pooled", False, 128), ("b2_initial", False, 10)]      # prefetch_k > n pages
    got = corpus.search_multistage_batch(stages, None, stage_queries=sq)
    for q, res in zip(queries, got):
        want = MO.multistage(q, [(pooled_docs, True, 128), (init_docs, False, 10)])
        assert len(res[0][1]) == n
        for s in range(2):
            _same_ranking(res[s][1].tolist(), res[s][0], want[s])
    corpus.drop_store("b2_initial")
    corpus.drop_store("b2_pooled")


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["fixed1", "fixed8", "fixed32", "ragged_small", "fixed300", "ragged_large"])
@pytest.mark.parametrize("pool", [False, True])
def test_dense_batched_scan_equals_single_query_scan(corpus, layout, pool):
    """Dense batched stage 0 (several queries share every document tile) must reproduce the single-query kernels
    (same per-column arithmetic) for every store layout and both query kinds."""
    rng = np.random.default_rng(hash((layout, pool)) % (2 ** 31))
    n = 333
    if layout.startswith("fixed"):
        r = int(layout[5:])
        rows = rng.standard_normal((n * r, 128)).astype(np.float16)
        corpus.add_store("db", rows, fixed_rows=r)
        docs = [rows[i * r:(i + 1) * r].astype(np.float32) for i in range(n)]
    else:
        lo, hi = (5, 60) if layout == "ragged_small" else (100, 500)
        rows, off, docs = _ragged_store(rng, n, lo, hi)
        corpus.add_store("db", rows, page_offsets=off)
    queries = [rng.standard_normal((int(rng.integers(1, 33)), 128)).astype(np.float32) for _ in range(11)]
    got = corpus.search_multistage_batch([("db", pool, n)], queries)
    for b, q in enumerate(queries):
        sc, ids = got[b][0]
        s1, i1 = corpus.search("db", q, n, pool_query=pool)
        # same per-column arithmetic; only the order of the final sum over query tokens differs (<= a few ulp)
        _same_ranking(ids.tolist(), sc, list(zip(i1.tolist(), s1.tolist())), rtol=1e-6, atol=1e-6)
        if b < 3:
            want = MO.multistage(q, [(docs, pool, n)])[0]
            _same_ranking(ids.tolist(), sc, want)
    corpus.drop_store("db")


@pytest.mark.gpu
@pytest.mark.parametrize("layout,pool,k", [("fixed1", True, 1000), ("fixed32", False, 256), ("fixed32", True, 300),
                                           ("fixed300", False, 500), ("fixed1small", True, 1000), ("fixed32small", False, 256)])
def test_fused_topk_prefilter_is_exact(corpus, layout, pool, k, monkeypatch):
    """Large dense batched stage: the sample-threshold prefilter (scores never written to HBM) must return exactly
    the lists of the unfiltered path (VRAG_PREFILTER=0), which the tests above pin to the oracle."""
    rng = np.random.default_rng(5)
    small = layout.endswith("small")     # the shard of a strongly scaled corpus (1M pages over 8 GPUs): sample = n / 8
    r = int(layout[5:].replace("small", ""))
    n = 130_000 if small else (300_000 if r <= 32 else 270_000)
    corpus.add_synthetic_store("pf", n, fixed_rows=r, seed=11)
    queries = [rng.standard_normal((int(rng.integers(8, 33)), 128)).astype(np.float32) for _ in range(9)]
    stages = [("pf", pool, k)]
    got = corpus.search_multistage_batch(stages, queries)
    monkeypatch.setenv("VRAG_PREFILTER", "0")
    want = corpus.search_multistage_batch(stages, queries)
    monkeypatch.delenv("VRAG_PREFILTER")
    for g, w in zip(got, want):
        assert len(g[0][1]) == k
        assert g[0][1].tolist() == w[0][1].tolist()
        np.testing.assert_array_equal(g[0][0], w[0][0])
    # spot check against the oracle on the best page of the first query
    best = int(got[0][0][1][0])
    page = corpus.read_page("pf", best).astype(np.float32)
    q = queries[0].mean(axis=0, keepdims=True) if pool else queries[0]
    assert abs(MO.maxsim_score(q, page) - float(got[0][0][0][0])) <= RTOL * abs(float(got[0][0][0][0])) + ATOL
    corpus.drop_store("pf")


@pytest.mark.gpu
def test_fp16_query_flag_is_inside_the_parity_gate(corpus):
    """VRAG_Q_FP16 (opt-in): scores carry the fp16 rounding of the query — well inside the 1e-3 gate — and the exact
    default is unchanged."""
    from visual_rag_b200 import _native as N
    import ctypes as C

    rng = np.random.default_rng(12)
    n, t = 64, 700
    rows = rows16(31, n * t)
    corpus.add_store("fq", rows, fixed_rows=t)
    q = rng.standard_normal((20, 128)).astype(np.float32)
    exact = corpus.score("fq", q)
    out = np.empty((n,), np.float32)
    N.check(corpus._lib.vrag_score(corpus._h, b"fq", q.ctypes.data_as(C.POINTER(C.c_float)), 20,
                                   N.VRAG_Q_NORMALIZE | N.VRAG_Q_FP16, None, 0, out.ctypes.data_as(C.POINTER(C.c_float))))
    want = np.array(MO.maxsim_batch(q, [rows[i * t:(i + 1) * t].astype(np.float32) for i in range(n)]))
    np.testing.assert_allclose(exact, want, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out, want, rtol=5e-4)
    assert np.abs(out - want).max() > 0          # it really is the reduced-precision path
    # the same opt-in through the Python methods: single, multi-stage and batched calls
    np.testing.assert_array_equal(corpus.score("fq", q, fp16_query=True), out)
    s16, i16 = corpus.search("fq", q, 10, fp16_query=True)
    np.testing.assert_allclose(s16, np.sort(want)[::-1][:10], rtol=5e-4)
    ms = corpus.search_multistage([("fq", False, 20), ("fq", False, 5)], q, fp16_query=True)
    np.testing.assert_allclose(ms[1][0], np.sort(want)[::-1][:5], rtol=5e-4)
    qs = [q, rng.standard_normal((9, 128)).astype(np.float32), rng.standard_normal((31, 128)).astype(np.float32)]
    b16 = corpus.search_multistage_batch([("fq", False, 8)], qs, fp16_query=True)
    bex = corpus.search_multistage_batch([("fq", False, 8)], qs)
    for a, b in zip(b16, bex):
        np.testing.assert_allclose(a[0][0], b[0][0], rtol=5e-4)
        assert np.abs(a[0][0] - b[0][0]).max() > 0
    corpus.drop_store("fq")


@pytest.mark.gpu
def test_indexer_upload_batches_equals_bulk_store():
    """GpuIndexer.upload_batch (append path, qdrant_indexer.py:341-507) in several batches == one bulk add_store:
    same search results, fp32 -> fp16 store cast, global = mean(tile_pooled) fallback, ids/payloads served."""
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.indexing import GpuIndexer
    from visual_rag_b200.retrieval import ThreeStageRetriever, TwoStageRetriever

    rng = np.random.default_rng(21)
    n = 57
    pts = []
    for i in range(n):
        t = int(rng.integers(140, 420))
        r = int(rng.integers(10, 33))
        vis = rng.standard_normal((t, 128)).astype(np.float32)
        tile = rng.standard_normal((r, 128)).astype(np.float32)
        p = {"id": GpuIndexer.generate_point_id("doc.pdf", i), "visual_embedding": vis, "tile_pooled_embedding": tile,
             "experimental_pooled_embedding": {"experimental_pooling": rng.standard_normal((r, 128)).astype(np.float32)},
             "metadata": {"filename": "doc.pdf", "page_number": i, "year": 2020 + i % 2}}
        if i % 3:
            p["global_pooled_embedding"] = rng.standard_normal((128,)).astype(np.float32)
        pts.append(p)
    with GpuCorpus(0) as c1, GpuCorpus(0) as c2:
        idx = GpuIndexer(c1, "c")
        assert idx.create_collection(force_recreate=True)
        up = 0
        for lo, hi in ((0, 1), (1, 20), (20, 21), (21, 57)):
            up += idx.upload_batch(pts[lo:hi])
        assert up == n and idx.check_exists(pts[5]["id"]) and not idx.check_exists("nope")
        assert idx.upload_batch(pts[:1]) == 1     # re-sending an existing id is an (idempotent) in-place upsert
        # bulk twin
        def cat(key, f=lambda p: p):
            mats = [np.asarray(f(p), np.float32).reshape(-1, 128) for p in pts]
            return np.concatenate(mats).astype(np.float16), np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])])
        for name, f in (("initial", lambda p: p["visual_embedding"]), ("mean_pooling", lambda p: p["tile_pooled_embedding"]),
                        ("experimental_pooling", lambda p: p["experimental_pooled_embedding"]["experimental_pooling"]),
                        ("global_pooling", lambda p: p.get("global_pooled_embedding", p["tile_pooled_embedding"].mean(axis=0)))):
            rows, off = cat(name, f)
            c2.add_store(name, rows, page_offsets=off)
            assert np.array_equal(c1.read_rows(name, 0, rows.shape[0]).view(np.uint16), rows.view(np.uint16))
        client2 = GpuCorpusClient(c2, "c", point_ids=[p["id"] for p in pts], payloads=[p["metadata"] for p in pts])
        q = rng.standard_normal((17, 128)).astype(np.float32)
        a = TwoStageRetriever(idx.client, "c").search_server_side(q, top_k=5, prefetch_k=20, stage1_mode="tokens_vs_standard_pooling")
        b = TwoStageRetriever(client2, "c").search_server_side(q, top_k=5, prefetch_k=20, stage1_mode="tokens_vs_standard_pooling")
        assert [r["id"] for r in a] == [r["id"] for r in b] and [r["score_final"] for r in a] == [r["score_final"] for r in b]
        assert a[0]["payload"]["filename"] == "doc.pdf"
        a3 = ThreeStageRetriever(idx.client, "c").search_server_side(query_embedding=q, top_k=5, stage1_k=30, stage2_k=12)
        b3 = ThreeStageRetriever(client2, "c").search_server_side(query_embedding=q, top_k=5, stage1_k=30, stage2_k=12)
        assert [r["id"] for r in a3] == [r["id"] for r in b3]
        assert idx.get_existing_ids("doc.pdf") == {p["id"] for p in pts}


@pytest.mark.gpu
def test_saliency_kernel_matches_reference_goldens(corpus):
    """vrag_saliency (column-max twin of MaxSim) vs patch_scores produced by the reference's generate_saliency_map."""
    import os

    from visual_rag_b200.visualization import saliency_scores

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "repool_golden.npz"))
    cs = CS.saliency_cases()
    mats = [CS.unit_rows(c["seed"], c["n"], dtype=np.float16) for c in cs]
    off = np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])])
    corpus.add_store("sal", np.concatenate(mats), page_offsets=off)
    for p, c in enumerate(cs):
        q = CS.query_rows(c["qseed"], c["q"])
        res = saliency_scores(corpus, q, p, token_info=c["token_info"], vector_name="sal")
        want = gold[c["key"] + "::patch_scores"]
        np.testing.assert_allclose(res["patch_scores"], want, rtol=1e-5, atol=2e-6)
        assert res["patch_scores_norm"].min() >= 0.0 and res["patch_scores_norm"].max() <= 1.0
        if c["token_info"] and c["n"] >= c["token_info"]["n_rows"] * c["token_info"]["n_cols"] * 64:
            assert res["tile_scores"].shape == (c["token_info"]["n_rows"], c["token_info"]["n_cols"])
        # consistency with MaxSim: sum_q max_t >= max_t max_q ... and the best patch score is the best single cosine
        assert abs(res["patch_scores"].max() - MO.saliency_patch_scores(q, mats[p].astype(np.float32)).max()) < 1e-5
    with pytest.raises(Exception):
        corpus.saliency("sal", CS.query_rows(1, 5), 99)
    corpus.drop_store("sal")


@pytest.mark.gpu
def test_payload_filter_cache_and_vectorised_conditions(gpu_client):
    from visual_rag_b200.retrieval import TwoStageRetriever

    q, client = gpu_client
    two = TwoStageRetriever(client, "c")
    f = two.build_filter(year=[2000, 2002])
    n_cached = len(client._filter_cache)
    r1 = two.search_single_stage(q, top_k=20, filter_obj=f)
    r2 = two.search_single_stage(q, top_k=20, filter_obj=two.build_filter(year=[2002, 2000]))   # same filter, cached
    assert [r["id"] for r in r1] == [r["id"] for r in r2] and all(r["payload"]["year"] in (2000, 2002) for r in r1)
    assert len(client._filter_cache) == n_cached + 1
    none = two.search_single_stage(q, top_k=5, filter_obj=two.build_filter(year=1900))
    assert none == []


@pytest.mark.gpu
def test_cfg0_full_size_colsmol_top10_matches_reference_cpu_path(corpus):
    """BASELINE configs[0] at full size: 10,000 ColSmol-shaped pages x 768 tokens (1.97 GB fp16), 20-token queries,
    exact MaxSim top-10 vs the reference's client-side CPU path (oracle search_exhaustive on the same bits)."""
    n, t = 10_000, 768
    corpus.add_synthetic_store("cfg0", n, fixed_rows=t, seed=2024)
    rows = corpus.read_rows("cfg0", 0, n * t).astype(np.float32)
    docs = [rows[i * t:(i + 1) * t] for i in range(n)]
    rng = np.random.default_rng(2024)
    queries = [rng.standard_normal((20, 128)).astype(np.float32) for _ in range(6)]
    got = corpus.search_multistage_batch([("cfg0", False, 10)], queries)
    for q, g in zip(queries, got):
        want = MO.search_exhaustive(q, docs, 10)
        _same_ranking(g[0][1].tolist(), g[0][0], want)
        s1, i1 = corpus.search("cfg0", q, 10)
        assert i1.tolist() == g[0][1].tolist()
    corpus.drop_store("cfg0")


@pytest.mark.gpu
def test_packed_queries_and_array_results(corpus):
    from visual_rag_b200.corpus import pack_queries

    rng = np.random.default_rng(31)
    rows = rows16(32, 50 * 200)
    corpus.add_store("pq", rows, fixed_rows=200)
    qs = [rng.standard_normal((int(rng.integers(3, 28)), 128)).astype(np.float32) for _ in range(12)]
    a = corpus.search_multistage_batch([("pq", False, 7)], qs)
    (sc, ids, cnt), = corpus.search_multistage_batch([("pq", False, 7)], pack_queries(qs), as_arrays=True)
    assert sc.shape == (12, 7) and ids.shape == (12, 7) and cnt.tolist() == [7] * 12
    for b in range(12):
        assert ids[b].tolist() == a[b][0][1].tolist() and np.array_equal(sc[b], a[b][0][0])
    corpus.drop_store("pq")


@pytest.mark.gpu
def test_sharded_batched_search_single_rank_equals_corpus_batch(corpus):
    """ShardedSearcher.search_multistage_batch (device-level batched stages + merge) at world size 1 == the host batch API,
    including a prefiltered dense stage."""
    from visual_rag_b200.distributed import ShardedSearcher

    rng = np.random.default_rng(55)
    n = 270_000
    corpus.add_synthetic_store("sb_glob", n, fixed_rows=1, seed=5)
    lens = rng.integers(8, 33, size=n)
    corpus.add_synthetic_store("sb_exp", 0, page_offsets=np.concatenate([[0], np.cumsum(lens)]), seed=6)
    corpus.add_synthetic_store("sb_init", n, fixed_rows=130, seed=7)
    qs = [rng.standard_normal((int(rng.integers(6, 30)), 128)).astype(np.float32) for _ in range(10)]
    stages = [("sb_glob", True, 500), ("sb_exp", False, 100), ("sb_init", False, 10)]
    want = corpus.search_multistage_batch(stages, qs, as_arrays=True)
    got = ShardedSearcher(corpus).search_multistage_batch(stages, qs)
    for (ws, wi, _), (gs, gi) in zip(want, got):
        assert np.array_equal(wi, gi) and np.array_equal(ws, gs)
    for nm in ("sb_glob", "sb_exp", "sb_init"):
        corpus.drop_store(nm)


@pytest.mark.gpu
def test_final_only_batch_results(corpus):
    """final_only: last-stage lists + the earlier-stage scores of the final pages == the full per-stage lists."""
    rng = np.random.default_rng(71)
    n = 500
    g = rng.standard_normal((n, 128)).astype(np.float16)
    e = rng.standard_normal((n * 16, 128)).astype(np.float16)
    i = rng.standard_normal((n * 150, 128)).astype(np.float16)
    corpus.add_store("fo_g", g, fixed_rows=1)
    corpus.add_store("fo_e", e, fixed_rows=16)
    corpus.add_store("fo_i", i, fixed_rows=150)
    qs = [rng.standard_normal((int(rng.integers(4, 30)), 128)).astype(np.float32) for _ in range(9)]
    stages = [("fo_g", True, 120), ("fo_e", False, 40), ("fo_i", False, 7)]
    full = corpus.search_multistage_batch(stages, qs)
    sc, ids, st, cnt = corpus.search_multistage_batch(stages, qs, final_only=True)
    assert sc.shape == (9, 7) and st.shape == (9, 7, 2) and cnt.tolist() == [7] * 9
    for b in range(9):
        assert ids[b].tolist() == full[b][2][1].tolist() and np.array_equal(sc[b], full[b][2][0])
        for s in range(2):
            lut = dict(zip(full[b][s][1].tolist(), full[b][s][0].tolist()))
            assert [lut[p] for p in ids[b].tolist()] == st[b, :, s].tolist()
    for nm in ("fo_g", "fo_e", "fo_i"):
        corpus.drop_store(nm)


# ------------------------------------------------------------------ maximum sizes: queries longer than one operand image
@pytest.mark.gpu
@pytest.mark.parametrize("q_rows", [129, 200, 300, 1024])
def test_long_queries_are_scored_in_row_chunks(corpus, q_rows):
    """The reference accepts any number of query tokens (pooling.py:468-514); more than 128 rows are scored as balanced
    row chunks whose partial page scores add up — LARGE, PACKED and candidate-list scans, search and multistage."""
    rng = np.random.default_rng(q_rows)
    q = CS.query_rows(7000 + q_rows, q_rows)
    lens = rng.integers(1, 400, size=41)
    off = np.concatenate([[0], np.cumsum(lens)])
    rows = rows16(7100 + q_rows, int(off[-1]))
    docs = [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(len(lens))]
    corpus.add_store("lq", rows, page_offsets=off)
    want = np.array([MO.maxsim_score(q, d) for d in docs])
    close(corpus.score("lq", q), want)
    cand = rng.permutation(len(lens))[:17]
    close(corpus.score("lq", q, candidate_ids=cand), want[cand])
    pooled = rows16(7200 + q_rows, len(lens) * 32)
    pdocs = [pooled[i * 32:(i + 1) * 32].astype(np.float32) for i in range(len(lens))]
    corpus.add_store("lqp", pooled, fixed_rows=32)
    close(corpus.score("lqp", q), [MO.maxsim_score(q, d) for d in pdocs])
    got = corpus.search_multistage([("lqp", False, 20), ("lq", False, 5)], q)
    ref = MO.multistage(q, [(pdocs, False, 20), (docs, False, 5)])
    assert got[1][1].tolist() == [i for i, _ in ref[1]]
    close(got[1][0], [s for _, s in ref[1]])
    # pooled query of a long token matrix: one row after the mean, no chunking involved
    ref1 = MO.multistage(q, [(pdocs, True, 9)])
    s, ids = corpus.search("lqp", q, 9, pool_query=True)
    assert ids.tolist() == [i for i, _ in ref1[0]]
    for nm in ("lq", "lqp"):
        corpus.drop_store(nm)


@pytest.mark.gpu
def test_long_queries_in_batches_and_saliency(corpus):
    rng = np.random.default_rng(77)
    n, t, r = 50, 150, 16
    rows = rows16(7300, n * t)
    pooled = rows16(7301, n * r)
    docs = [rows[i * t:(i + 1) * t].astype(np.float32) for i in range(n)]
    pdocs = [pooled[i * r:(i + 1) * r].astype(np.float32) for i in range(n)]
    corpus.add_store("bq", rows, fixed_rows=t)
    corpus.add_store("bqp", pooled, fixed_rows=r)
    qs = [CS.query_rows(7400 + i, m) for i, m in enumerate((20, 260, 129, 7))]   # a batch mixing short and long queries
    stages = [("bqp", False, 12), ("bq", False, 4)]
    got = corpus.search_multistage_batch(stages, qs)
    for q, g in zip(qs, got):
        ref = MO.multistage(q, [(pdocs, False, 12), (docs, False, 4)])
        assert g[1][1].tolist() == [i for i, _ in ref[1]]
        close(g[1][0], [s for _, s in ref[1]])
    q = qs[1]
    close(corpus.saliency("bq", q, 3), MO.saliency_patch_scores(q, docs[3]), rtol=1e-5, atol=2e-6)
    for nm in ("bq", "bqp"):
        corpus.drop_store(nm)


@pytest.mark.gpu
def test_indexer_upsert_replaces_pages_in_place():
    """client.upsert semantics of QdrantIndexer.upload_batch (qdrant_indexer.py:459-507): a batch mixing ids that exist
    (same shapes, new vectors and payload) with new ids == a bulk store of the final state."""
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.indexing import GpuIndexer
    from visual_rag_b200.retrieval import TwoStageRetriever

    rng = np.random.default_rng(33)

    def point(i, t, r, ver):
        g = np.random.default_rng(1000 * ver + i)
        return {"id": GpuIndexer.generate_point_id("d.pdf", i), "visual_embedding": g.standard_normal((t, 128)).astype(np.float32),
                "tile_pooled_embedding": g.standard_normal((r, 128)).astype(np.float32),
                "experimental_pooled_embedding": g.standard_normal((r + 2, 128)).astype(np.float32),
                "metadata": {"filename": "d.pdf", "page_number": i, "version": ver}}

    shapes = [(int(rng.integers(130, 300)), int(rng.integers(8, 33))) for _ in range(30)]
    v1 = [point(i, *shapes[i], 1) for i in range(20)]
    with GpuCorpus(0) as c1, GpuCorpus(0) as c2:
        idx = GpuIndexer(c1, "c")
        idx.create_collection(force_recreate=True)
        assert idx.upload_batch(v1) == 20
        q = rng.standard_normal((21, 128)).astype(np.float32)
        two = TwoStageRetriever(idx.client, "c")
        two.search_server_side(q, top_k=5, prefetch_k=12)          # builds page tables / tensor maps before the upsert
        changed = [3, 0, 19, 7]
        batch = [point(i, *shapes[i], 2) for i in changed] + [point(i, *shapes[i], 2) for i in range(20, 30)]
        batch.insert(2, point(7, *shapes[7], 3))                   # duplicate id inside the batch: the last occurrence wins
        assert idx.upload_batch(batch) == len(batch)
        final = {p["id"]: p for p in v1}
        order = [p["id"] for p in v1]
        for p in batch:
            if p["id"] not in final:
                order.append(p["id"])
            final[p["id"]] = p
        assert final[GpuIndexer.generate_point_id("d.pdf", 7)]["metadata"]["version"] == 2
        pts = [final[i] for i in order]
        for name, f in (("initial", lambda p: p["visual_embedding"]), ("mean_pooling", lambda p: p["tile_pooled_embedding"]),
                        ("experimental_pooling", lambda p: p["experimental_pooled_embedding"]),
                        ("global_pooling", lambda p: p["tile_pooled_embedding"].mean(axis=0))):
            mats = [np.asarray(f(p), np.float32).reshape(-1, 128) for p in pts]
            rows = np.concatenate(mats).astype(np.float16)
            c2.add_store(name, rows, page_offsets=np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]))
            assert np.array_equal(c1.read_rows(name, 0, rows.shape[0]).view(np.uint16), rows.view(np.uint16)), name
        client2 = GpuCorpusClient(c2, "c", point_ids=order, payloads=[p["metadata"] for p in pts])
        for mode in ("tokens_vs_standard_pooling", "pooled_query_vs_global", "tokens_vs_experimental_pooling"):
            a = two.search_server_side(q, top_k=8, prefetch_k=15, stage1_mode=mode)
            b = TwoStageRetriever(client2, "c").search_server_side(q, top_k=8, prefetch_k=15, stage1_mode=mode)
            assert [(r["id"], r["score_final"], r["payload"]) for r in a] == [(r["id"], r["score_final"], r["payload"]) for r in b]
        assert {r["payload"]["version"] for r in two.search_server_side(q, top_k=30, prefetch_k=30)} == {1, 2}
        from visual_rag_b200._native import VragError
        with pytest.raises(VragError, match="out of range"):
            c1.replace_pages("initial", [99], np.zeros((3, 128), np.float16), [0, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["large", "pooled32", "fixed13", "mixed"])
def test_shape_changing_upserts_deletes_and_compaction_match_the_oracle(layout):
    """Page-table indirection (vrag_store_replace_pages with new shapes, vrag_store_delete_pages, vrag_store_compact):
    after every step the scans — exhaustive, candidate-restricted, batched, pooled-query — equal the oracle over the
    store's CURRENT pages; compaction changes no result and returns the store to the dense layout."""
    from visual_rag_b200.corpus import GpuCorpus

    rng = np.random.default_rng({"large": 1, "pooled32": 2, "fixed13": 3, "mixed": 4}[layout])
    n = 300
    if layout == "large":
        lens = rng.integers(129, 400, size=n)
    elif layout == "pooled32":
        lens = rng.integers(5, 33, size=n)
    elif layout == "fixed13":
        lens = np.full((n,), 13)
    else:
        lens = rng.integers(1, 260, size=n)
    pages = [rows16(int(rng.integers(1 << 30)), int(t)) for t in lens]
    q = CS.query_rows(77, 19)
    qs = [CS.query_rows(80 + j, 6 + 5 * j) for j in range(3)]

    def check(c, what):
        want = np.array([MO.maxsim_score(q, p.astype(np.float32)) if len(p) else -np.inf for p in pages], dtype=np.float64)
        got = c.score("s", q)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin), what
        close(got[fin], want[fin])
        k = 25
        s_, ids = c.search("s", q, k)
        order = np.lexsort((np.arange(len(want)), -want))[:k]
        _same_ranking(ids, s_, [(int(i), float(want[i])) for i in order])
        cand = rng.permutation(len(pages))[:40]
        gc = c.score("s", q, candidate_ids=cand)
        assert np.array_equal(np.isfinite(gc), fin[cand]) and np.allclose(gc[fin[cand]], want[cand][fin[cand]], rtol=RTOL, atol=ATOL), what
        wantp = np.array([MO.pooled_query_score(q, p.astype(np.float32)) if len(p) else -np.inf for p in pages])
        gp = c.score("s", q, pool_query=True)
        assert np.allclose(gp[fin], wantp[fin], rtol=RTOL, atol=ATOL), what
        res = c.search_multistage_batch([("s", False, 10)], qs)
        for qq, r in zip(qs, res):
            w = np.array([MO.maxsim_score(qq, p.astype(np.float32)) if len(p) else -np.inf for p in pages])
            _same_ranking(r[0][1], r[0][0], [(int(i), float(w[i])) for i in np.lexsort((np.arange(len(w)), -w))[:10]])

    with GpuCorpus(0) as c:
        off = np.concatenate([[0], np.cumsum([len(p) for p in pages])])
        if layout == "fixed13":
            c.add_store("s", np.concatenate(pages), fixed_rows=13)
        else:
            c.add_store("s", np.concatenate(pages), page_offsets=off)
        check(c, "initial")
        # ---- upsert with new shapes: some pages grow, some shrink, some keep their size
        hi = 400 if layout in ("large", "mixed") else (32 if layout == "pooled32" else 40)
        lo = 129 if layout == "large" else 1
        changed = rng.permutation(n)[:60].tolist()
        new_pages = [rows16(int(rng.integers(1 << 30)), int(rng.integers(lo, hi + 1))) for _ in changed]
        new_pages[0] = rows16(5, len(pages[changed[0]]))                      # same size: stays in place
        c.replace_pages("s", changed, np.concatenate(new_pages), np.concatenate([[0], np.cumsum([len(p) for p in new_pages])]))
        for pg, p in zip(changed, new_pages):
            pages[pg] = p
        assert c.store_info("s")["fixed_rows"] == 0
        check(c, "after shape-changing upsert")
        for pg in (changed[1], 0, n - 1):
            assert np.array_equal(c.read_page("s", pg).view(np.uint16), pages[pg].view(np.uint16))
        # ---- deletes
        gone = rng.permutation(n)[:25].tolist()
        c.delete_pages("s", gone)
        for pg in gone:
            pages[pg] = pages[pg][:0]
        check(c, "after deletes")
        # ---- appends behind a page table
        extra = [rows16(int(rng.integers(1 << 30)), int(rng.integers(lo, hi + 1))) for _ in range(17)]
        c.append_store("s", np.concatenate(extra), page_offsets=np.concatenate([[0], np.cumsum([len(p) for p in extra])]))
        pages.extend(extra)
        check(c, "after append")
        before = c.score("s", q)
        rows_before = c.store_info("s")["total_rows"]
        c.compact_store("s")
        info = c.store_info("s")
        assert info["total_rows"] == sum(len(p) for p in pages) <= rows_before and info["n_pages"] == len(pages)
        after = c.score("s", q)
        assert np.array_equal(np.isfinite(before), np.isfinite(after))
        assert np.allclose(before[np.isfinite(before)], after[np.isfinite(after)], rtol=1e-6)
        check(c, "after compaction")
        c.truncate_store("s", n)
        del pages[n:]
        check(c, "after truncate")


@pytest.mark.gpu
def test_indexer_upload_from_threads_and_shared_handle():
    """The reference ingests with uploader threads (run_qdrant_beir.py:720-768) while queries may run: upload_batch calls
    from several threads and searches on the same handle serialise; every id ends up on the page that holds its vectors."""
    import threading

    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.indexing import GpuIndexer
    from visual_rag_b200.retrieval import TwoStageRetriever

    def point(i):
        g = np.random.default_rng(500 + i)
        t, r = 130 + (i % 7) * 10, 8 + i % 5
        return {"id": f"p{i}", "visual_embedding": g.standard_normal((t, 128)).astype(np.float32),
                "tile_pooled_embedding": g.standard_normal((r, 128)).astype(np.float32),
                "experimental_pooled_embedding": g.standard_normal((r, 128)).astype(np.float32), "metadata": {"i": i}}

    pts = [point(i) for i in range(96)]
    with GpuCorpus(0) as c:
        idx = GpuIndexer(c, "c")
        idx.create_collection(force_recreate=True)
        idx.upload_batch(pts[:8])
        errors = []
        q = np.random.default_rng(1).standard_normal((15, 128)).astype(np.float32)

        def uploader(lo, hi):
            try:
                for a in range(lo, hi, 4):
                    idx.upload_batch(pts[a:a + 4])
            except Exception as e:   # noqa: BLE001
                errors.append(e)

        def searcher():
            try:
                two = TwoStageRetriever(idx.client, "c")
                for _ in range(30):
                    res = two.search_single_stage(q, top_k=3)
                    assert len(res) == 3
            except Exception as e:   # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=uploader, args=(8 + 22 * k, 8 + 22 * (k + 1))) for k in range(4)] + [threading.Thread(target=searcher)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        assert c.n_pages("initial") == 96 and idx.get_existing_ids() == {p["id"] for p in pts}
        for p in pts[::5]:
            got = idx.client.retrieve("c", ids=[p["id"]], with_payload=True, with_vectors=["initial", "mean_pooling"])[0]
            assert got.payload == p["metadata"]
            assert np.array_equal(np.asarray(got.vector["initial"], np.float32), p["visual_embedding"].astype(np.float16).astype(np.float32))
            assert np.array_equal(np.asarray(got.vector["mean_pooling"], np.float32), p["tile_pooled_embedding"].astype(np.float16).astype(np.float32))


@pytest.mark.gpu
def test_sampled_topk_equals_radix_select(corpus, monkeypatch):
    """Single-query searches over large score arrays take their top-k from a sampled threshold + one compaction pass; the
    lists must be bit-identical to the radix-select path (VRAG_SAMPLED_TOPK=0, pinned to the oracle by the tests above),
    including the cases where the estimate must fail and fall back: all scores tied, almost all pages empty."""
    def both(name, q, k, **kw):
        a = corpus.search(name, q, k, **kw)
        monkeypatch.setenv("VRAG_SAMPLED_TOPK", "0")
        b = corpus.search(name, q, k, **kw)
        monkeypatch.delenv("VRAG_SAMPLED_TOPK")
        assert a[1].tolist() == b[1].tolist() and a[0].tolist() == b[0].tolist(), (name, k)
        return a

    q = CS.query_rows(9100, 12)
    n = 300_000                                     # strided sample (n > 65536)
    corpus.add_store("tk", rows16(9101, n * 2), fixed_rows=2)
    for k in (1, 10, 256, 1000):
        s, ids = both("tk", q, k)
        assert len(ids) == k and np.all(np.diff(s) <= 0)
    both("tk", q, 64, pool_query=True)
    cand = np.random.default_rng(3).permutation(n)[:40_000]      # candidate list: ids map through the list
    s, ids = both("tk", q, 100, candidate_ids=cand)
    assert set(ids.tolist()) <= set(cand.tolist())
    st = corpus.search_multistage([("tk", False, 500), ("tk", False, 7)], q)
    monkeypatch.setenv("VRAG_SAMPLED_TOPK", "0")
    st0 = corpus.search_multistage([("tk", False, 500), ("tk", False, 7)], q)
    monkeypatch.delenv("VRAG_SAMPLED_TOPK")
    assert st[0][1].tolist() == st0[0][1].tolist() and st[1][1].tolist() == st0[1][1].tolist()
    corpus.add_store("tk_small", rows16(9102, 20_000 * 3), fixed_rows=3)    # n <= 65536: every score is "sampled"
    for k in (10, 300):
        both("tk_small", q, k)
    # all pages identical -> every score ties -> the estimate cannot separate k keys -> exact fallback, ties by lower id
    one = rows16(9103, 4)
    corpus.add_store("tk_ties", np.tile(one, (30_000, 1)), fixed_rows=4)
    s, ids = both("tk_ties", q, 50)
    assert ids.tolist() == list(range(50)) and len(set(s.tolist())) == 1
    # almost every page empty (-inf): fewer finite scores than k
    off = np.zeros(20_001, dtype=np.int64)
    off[-5:] = np.arange(1, 6) * 3
    corpus.add_store("tk_empty", rows16(9104, 15), page_offsets=off)
    s, ids = both("tk_empty", q, 10)
    assert sorted(ids[:5].tolist()) == [19_995, 19_996, 19_997, 19_998, 19_999][:5] or np.isfinite(s[:5]).all()
    for nm in ("tk", "tk_small", "tk_ties", "tk_empty"):
        corpus.drop_store(nm)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(24)))
def test_randomized_differential_vs_oracle(corpus, seed):
    """Seeded random layouts — page sizes from empty to > 1000 rows (or all small, or fixed), 1..150 query rows, random k,
    optional candidate lists with out-of-shard ids, pooled or token queries, with and without normalisation — scored by the
    kernels and by the oracle: scores within tolerance, top-k lists identical up to swaps between near-equal scores."""
    rng = np.random.default_rng(10_000 + seed)
    kind = seed % 4
    n = int(rng.integers(5, 400))
    if kind == 0:
        lens = rng.integers(0, 1100, size=n)          # LARGE, ragged, some empty
    elif kind == 1:
        lens = rng.integers(0, 33, size=n)            # pooled-store sized, ragged
    elif kind == 2:
        lens = np.full(n, int(rng.integers(1, 129)))  # fixed rows <= 128
    else:
        lens = rng.integers(1, 129, size=n)           # PACKED general path
    if lens.sum() == 0:
        lens[0] = 3
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rows = rows16(20_000 + seed, int(off[-1]), scale=bool(seed % 2))
    docs = [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(n)]
    if kind == 2:
        corpus.add_store("rd", rows, fixed_rows=int(lens[0]))
    else:
        corpus.add_store("rd", rows, page_offsets=off)
    for trial in range(3):
        qn = int(rng.choice([1, 2, 7, 20, 33, 64, 100, 150]))
        q = CS.query_rows(30_000 + seed * 10 + trial, qn)
        pool = bool(rng.integers(0, 2)) and qn > 1
        normalize = bool(rng.integers(0, 4))          # mostly cosine
        k = int(rng.integers(1, 2 * n))
        qq = q.mean(axis=0, keepdims=True) if pool else q
        want = np.array([MO.maxsim_score(qq, d, normalize=normalize) if len(d) else -np.inf for d in docs])
        got = corpus.score("rd", q, normalize=normalize, pool_query=pool)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        close(got[fin], want[fin], rtol=1e-4 if normalize else 3e-4, atol=1e-5 if normalize else 1e-3)
        cand = None
        if trial == 1:
            cand = np.concatenate([rng.permutation(n)[: max(1, n // 3)], [n + 5, -2]]).astype(np.int64)
            rng.shuffle(cand)
        s, ids = corpus.search("rd", q, k, normalize=normalize, pool_query=pool, candidate_ids=cand)
        # rank by the kernel's own scores: an exact-order check of the top-k logic. Empty pages and ids outside the shard
        # score -inf and sort last, in candidate order (the sharded merge relies on that; the client drops them)
        pool_idx = np.arange(n) if cand is None else cand
        inside = (pool_idx >= 0) & (pool_idx < n)
        ref_scores = np.where(inside, got[np.clip(pool_idx, 0, n - 1)], -np.inf)
        order = sorted(range(len(pool_idx)), key=lambda j: (-ref_scores[j], j))[:k]
        assert ids.tolist() == [int(pool_idx[j]) for j in order], (seed, trial, kind)
        assert s.tolist() == [float(ref_scores[j]) for j in order]
    corpus.drop_store("rd")


# ---------------------------------------------------------------------------------------------------------------
# the exchange building blocks of the C ABI on ONE GPU: two handles on device 0 act as the two shards of a corpus; their
# packed local lists are laid out as an all-gather would leave them ([rank][list][k]) and merged with vrag_merge_hits_dev;
# a candidate stage is completed with the element-wise max the all-reduce computes. (The collectives themselves run at
# N > 1 in bench.py's `sharded_parity`; vrag_allgather_topk / vrag_allreduce_max_dev are the identity on one rank.)
@pytest.mark.gpu
@pytest.mark.parametrize("k", [10, 256, 3000])
def test_packed_hit_exchange_blocks_emulate_two_shards(corpus, k):
    import torch

    from visual_rag_b200.corpus import GpuCorpus, query_flags

    n, t, split = 9000, 40, 3700
    rows = rows16(4100, n * t)
    rows[(n - 2) * t:(n - 1) * t] = rows[5 * t:6 * t]                 # exact tie across the two shards
    pooled = rows16(4101, n * 8)
    q = CS.query_rows(4102, 17)
    corpus.add_store("ex_i", rows, fixed_rows=t)
    corpus.add_store("ex_p", pooled, fixed_rows=8)
    want = corpus.search_multistage([("ex_p", False, k), ("ex_i", False, min(k, 50))], q)
    shards = [GpuCorpus(0, page_base=0), GpuCorpus(0, page_base=split)]
    try:
        shards[0].add_store("ex_i", rows[:split * t], fixed_rows=t)
        shards[0].add_store("ex_p", pooled[:split * 8], fixed_rows=8)
        shards[1].add_store("ex_i", rows[split * t:], fixed_rows=t)
        shards[1].add_store("ex_p", pooled[split * 8:], fixed_rows=8)
        dev = torch.device("cuda", 0)
        st = torch.cuda.current_stream(dev).cuda_stream
        qd = torch.from_numpy(q).to(dev)
        fl = query_flags(True, False)
        # ---- stage 1 (scan): local lists as packed entries -> "gathered" [rank][1][k] -> merge
        gathered = torch.zeros((2, k, 16), dtype=torch.uint8, device=dev)
        for r, sh in enumerate(shards):
            local = torch.zeros((k, 16), dtype=torch.uint8, device=dev)
            sh.stage_hits_dev("ex_p", qd.data_ptr(), q.shape[0], fl, 0, 0, k, local.data_ptr(), st)
            sh.allgather_topk(local.data_ptr(), 1, k, gathered[r].data_ptr(), st)      # one rank: a device copy
        ms = torch.empty((k,), dtype=torch.float32, device=dev)
        mi = torch.empty((k,), dtype=torch.int64, device=dev)
        flag = torch.zeros((1,), dtype=torch.int32, device=dev)
        shards[0].merge_hits_dev(gathered.data_ptr(), 2, 1, k, k, ms.data_ptr(), mi.data_ptr(), flag.data_ptr(), st)
        torch.cuda.synchronize()
        assert int(flag.item()) == 0
        assert mi.cpu().numpy().tolist() == want[0][1].tolist()
        assert np.array_equal(ms.cpu().numpy(), want[0][0])
        hits = np.frombuffer(gathered.cpu().numpy().tobytes(), dtype=GpuCorpus.HIT_DTYPE).reshape(2, k)
        assert (hits["id"][0] < split).all() and (hits["id"][1][hits["id"][1] >= 0] >= split).all()
        # ---- stage 2 (restricted to the merged list): -inf for foreign pages, max across shards, top-k in candidate order
        k2 = min(k, 50)
        sc = [torch.empty((k,), dtype=torch.float32, device=dev) for _ in shards]
        for sh, buf in zip(shards, sc):
            sh.score_dev("ex_i", qd.data_ptr(), q.shape[0], fl, mi.data_ptr(), k, buf.data_ptr(), st)
            sh.allreduce_max_dev(buf.data_ptr(), k, st)                                  # one rank: no-op
        torch.cuda.synchronize()
        assert torch.isinf(sc[0]).sum() + torch.isinf(sc[1]).sum() == k                 # every candidate has exactly one owner
        red = torch.maximum(sc[0], sc[1])
        o_s = torch.empty((k2,), dtype=torch.float32, device=dev)
        o_i = torch.empty((k2,), dtype=torch.int64, device=dev)
        shards[0].topk_dev(red.data_ptr(), mi.data_ptr(), 0, k, k2, o_s.data_ptr(), o_i.data_ptr(), st)
        torch.cuda.synchronize()
        assert o_i.cpu().numpy().tolist() == want[1][1].tolist()
        assert np.array_equal(o_s.cpu().numpy(), want[1][0])
    finally:
        for sh in shards:
            sh.close()
        corpus.drop_store("ex_i")
        corpus.drop_store("ex_p")


@pytest.mark.gpu
def test_merge_flags_a_missed_estimate_and_ignores_padding(corpus):
    """A kHitMiss bit in any gathered entry raises the flag on the merging rank; padding entries (id < 0) never reach the
    output ahead of real ones, even real -inf ones."""
    import torch

    from visual_rag_b200.corpus import GpuCorpus

    dev = torch.device("cuda", 0)
    h = np.zeros((3, 4), dtype=GpuCorpus.HIT_DTYPE)
    h["id"], h["score"] = -1, -np.inf
    h[0][:2] = [(0.5, 0, 7), (-np.inf, 0, 9)]            # a real page with score -inf (empty page) ...
    h[1][:1] = [(0.5, 0, 1000)]                          # ... a tie with rank 0's best: lower rank (= lower id) first
    h[2][:3] = [(0.9, 1, 2000), (0.1, 1, 2001), (0.0, 1, 2002)]   # rank 2's list came from a missed estimate
    g = torch.from_numpy(np.frombuffer(h.tobytes(), dtype=np.uint8).copy()).to(dev)
    ms = torch.empty((6,), dtype=torch.float32, device=dev)
    mi = torch.empty((6,), dtype=torch.int64, device=dev)
    flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    corpus.merge_hits_dev(g.data_ptr(), 3, 1, 4, 6, ms.data_ptr(), mi.data_ptr(), flag.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    assert int(flag.item()) == 1
    assert mi.cpu().tolist() == [2000, 7, 1000, 2001, 2002, 9]
    assert ms.cpu().tolist()[:5] == [np.float32(0.9), 0.5, 0.5, np.float32(0.1), 0.0] and np.isinf(ms.cpu().numpy()[5])


# ---------------------------------------------------------------------------------------------------------------
# payload filters: the RANKING of a filtered search equals the oracle restricted to the pages that pass (8f-3)
@pytest.mark.gpu
def test_filtered_searches_rank_like_the_oracle_on_the_filtered_set(retrieval_golden):
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.retrieval import SingleStageRetriever, ThreeStageRetriever, TwoStageRetriever
    from visual_rag_b200.retrieval import models as M

    q, initial = CS.retrieval_corpus()
    n = len(initial)
    off = retrieval_golden["offsets_pooled"]
    own = GpuCorpus(0)       # its own handle: other tests of this module overwrite the shared corpus' stores
    own.add_store("initial", np.concatenate(initial).astype(np.float16),
                  page_offsets=np.concatenate([[0], np.cumsum([d.shape[0] for d in initial])]))
    own.add_store("mean_pooling", retrieval_golden["mean_pooling"], page_offsets=off)
    own.add_store("experimental_pooling", retrieval_golden["experimental_pooling"], page_offsets=off)
    own.add_store("global_pooling", retrieval_golden["global_pooling"], fixed_rows=1)
    client = GpuCorpusClient(own, "c", payloads=[{"page": i, "year": 2000 + i % 3} for i in range(n)])
    split = lambda rows: [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(n)]  # noqa: E731
    pooled, exper = split(retrieval_golden["mean_pooling"]), split(retrieval_golden["experimental_pooling"])
    glob = [g.astype(np.float32)[None, :] for g in retrieval_golden["global_pooling"]]
    docs = [d.astype(np.float16).astype(np.float32) for d in initial]
    payload = lambda i: {"page": i, "year": 2000 + i % 3}   # noqa: E731  (the gpu_client fixture's payloads)

    def restricted(pred, stages):
        keep = [i for i in range(n) if pred(payload(i))]
        res = MO.multistage(q, [([st[i] for i in keep], pool, k) for st, pool, k in stages])[-1]
        return [(keep[i], s) for i, s in res]

    def same(got, want):
        assert [g["id"] for g in got] == [i for i, _ in want]
        close([g["score_final"] for g in got], [s for _, s in want])

    single, two, three = SingleStageRetriever(client, "c"), TwoStageRetriever(client, "c"), ThreeStageRetriever(client, "c")
    f = two.build_filter(year=2001)
    is01 = lambda p: p["year"] == 2001   # noqa: E731
    same(single.search(q, top_k=10, strategy="multi_vector", filter_obj=f), restricted(is01, [(docs, False, 10)]))
    same(two.search_server_side(q, top_k=10, prefetch_k=25, filter_obj=f, stage1_mode="tokens_vs_standard_pooling"),
         restricted(is01, [(pooled, False, 25), (docs, False, 10)]))
    same(three.search_server_side(query_embedding=q, top_k=5, stage1_k=30, stage2_k=12, filter_obj=f),
         restricted(is01, [(glob, True, 30), (exper, False, 12), (docs, False, 5)]))
    mix = M.Filter(must=[M.FieldCondition(key="year", match=M.MatchAny(any=[2000, 2002]))],
                   must_not=[M.FieldCondition(key="page", range=M.Range(lt=20))],
                   should=[M.FieldCondition(key="page", match=M.MatchExcept(**{"except": list(range(0, n, 2))})),
                           M.FieldCondition(key="page", range=M.Range(gte=100, lte=110))])
    pred = lambda p: p["year"] in (2000, 2002) and not p["page"] < 20 and (p["page"] % 2 == 1 or 100 <= p["page"] <= 110)  # noqa: E731
    same(single.search(q, top_k=12, strategy="multi_vector", filter_obj=mix), restricted(pred, [(docs, False, 12)]))
    with pytest.raises(NotImplementedError):
        single.search(q, top_k=3, filter_obj=M.Filter(must=[M.FieldCondition(key="year")]))
    own.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["large", "large_var", "pooled32", "fixed13", "ragged_small", "global1"])
def test_page_bitmask_filter_in_the_scan_equals_candidate_list_and_oracle(corpus, layout):
    """vrag_filter_create + vrag_search_multistage_filtered: the in-scan page bitmask gives exactly the lists of the
    candidate-list form of the same filter (and of the oracle restricted to the pages that pass), for every store layout;
    multi-stage: filtered-out pages never re-enter through a later stage, even when fewer pages pass than a stage keeps."""
    rng = np.random.default_rng({"large": 1, "large_var": 2, "pooled32": 3, "fixed13": 4, "ragged_small": 5, "global1": 6}[layout])
    n = 3000
    if layout == "large":
        lens = np.full((n,), 200)
    elif layout == "large_var":
        lens = rng.integers(129, 300, size=n)
    elif layout == "pooled32":
        lens = np.full((n,), 32)
    elif layout == "fixed13":
        lens = np.full((n,), 13)
    elif layout == "ragged_small":
        lens = rng.integers(1, 33, size=n)
    else:
        lens = np.full((n,), 1)
    off = np.concatenate([[0], np.cumsum(lens)])
    rows = rows16(900 + len(layout), int(off[-1]))
    fixed = int(lens[0]) if len(set(lens.tolist())) == 1 else 0
    if fixed:
        corpus.add_store("fm", rows, fixed_rows=fixed)
    else:
        corpus.add_store("fm", rows, page_offsets=off)
    corpus.add_synthetic_store("fm_pool", n, fixed_rows=8, seed=77)
    q = CS.query_rows(901, 18)
    for frac in (0.5, 0.05, 0.9):
        allowed = rng.random(n) < frac
        allowed[:3] = [True, False, True]
        fid = corpus.create_filter(allowed)
        ids_allowed = np.nonzero(allowed)[0]
        for pool in (False, True):
            s_m, i_m = corpus.search("fm", q, 40, pool_query=pool, filter_id=fid)
            s_c, i_c = corpus.search("fm", q, 40, pool_query=pool, candidate_ids=ids_allowed)
            keep = np.isfinite(s_m)
            # the masked scan and the gather sum a page's token maxima in different orders: ids equal, scores to the last ulps
            assert i_m[keep].tolist() == i_c.tolist()[:int(keep.sum())]
            np.testing.assert_allclose(s_m[keep], s_c[:int(keep.sum())], rtol=2e-6)
            assert allowed[i_m[keep]].all() and keep.sum() == min(40, len(ids_allowed))
        sub = [rows[off[i]:off[i + 1]].astype(np.float32) for i in ids_allowed]
        want = MO.search_exhaustive(q, sub, 10)
        s_m, i_m = corpus.search("fm", q, 10, filter_id=fid)
        _same_ranking(i_m, s_m, [(int(ids_allowed[i]), x) for i, x in want])
        # two-stage under the mask == two-stage over the candidate list
        st_m = corpus.search_multistage([("fm_pool", False, 64), ("fm", False, 10)], q, filter_id=fid)
        st_c = corpus.search_multistage([("fm_pool", False, 64), ("fm", False, 10)], q, candidate_ids=ids_allowed)
        assert st_m[1][1].tolist() == st_c[1][1].tolist()
        np.testing.assert_allclose(st_m[1][0], st_c[1][0], rtol=2e-6)
        corpus.destroy_filter(fid)
    # fewer pages pass than the stages keep: the padded (filtered-out) pages stay out of the later stage
    few = np.zeros(n, dtype=bool)
    few[[5, 17, 1234]] = True
    fid = corpus.create_filter(few)
    st = corpus.search_multistage([("fm_pool", False, 64), ("fm", False, 10)], q, filter_id=fid)
    fin = np.isfinite(st[1][0])
    assert sorted(st[1][1][fin].tolist()) == [5, 17, 1234]
    corpus.destroy_filter(fid)
    from visual_rag_b200._native import VragError
    with pytest.raises(VragError, match="unknown filter"):
        corpus.search("fm", q, 5, filter_id=fid)
    corpus.drop_store("fm")
    corpus.drop_store("fm_pool")


@pytest.mark.gpu
def test_client_picks_bitmask_or_candidate_list_by_selectivity():
    """GpuCorpusClient: a payload filter that lets many pages through runs as a device bitmask (created once, reused), a
    selective one as a candidate list; rankings equal the oracle on the filtered set either way."""
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.retrieval import SingleStageRetriever, TwoStageRetriever

    n, t = 20000, 12
    rows = rows16(31337, n * t)
    q = CS.query_rows(31338, 9)
    with GpuCorpus(0) as c:
        c.add_store("initial", rows, fixed_rows=t)
        client = GpuCorpusClient(c, "c", payloads=[{"year": 2000 + i % 2, "bucket": i % 100} for i in range(n)])
        single, two = SingleStageRetriever(client, "c"), TwoStageRetriever(client, "c")
        docs = rows.astype(np.float32).reshape(n, t, 128)
        for f, pred, expect_mask in ((two.build_filter(year=2001), lambda i: i % 2 == 1, True),
                                     (Filter_bucket(7), lambda i: i % 100 == 7, False)):
            keep = [i for i in range(n) if pred(i)]
            want = MO.search_exhaustive(q, [docs[i] for i in keep], 10)
            calls = []
            orig = c.create_filter
            c.create_filter = lambda m: (calls.append(1), orig(m))[1]
            for _ in range(3):
                got = single.search(q, top_k=10, strategy="multi_vector", filter_obj=f)
                _same_ranking([g["id"] for g in got], [g["score"] for g in got], [(keep[i], x) for i, x in want])
            c.create_filter = orig
            assert len(calls) == (1 if expect_mask else 0)      # one upload for three queries / none for the selective filter


def Filter_bucket(b):
    from visual_rag_b200.retrieval import models as M

    return M.Filter(must=[M.FieldCondition(key="bucket", match=M.MatchValue(value=b))])


# ---------------------------------------------------------------------------------------------------------------
# batched exhaustive scans of full-token stores: approximate first pass (8 plain-fp16 queries per document tile) + exact
# re-score of the candidates; a device-side guard keeps the result exact
@pytest.mark.gpu
@pytest.mark.parametrize("nq,k", [(8, 10), (13, 10), (5, 100), (32, 3)])
def test_approximate_first_pass_batches_are_exact(corpus, nq, k, monkeypatch):
    rng = np.random.default_rng(100 + nq)
    n = 6000
    lens = rng.integers(129, 420, size=n)
    off = np.concatenate([[0], np.cumsum(lens)])
    corpus.add_synthetic_store("ap", 0, page_offsets=off, seed=4242)
    qs = [rng.standard_normal((int(rng.integers(3, 33)), 128)).astype(np.float32) for _ in range(nq)]
    got = corpus.search_multistage_batch([("ap", False, k)], qs)                     # approximate pass + re-score (default)
    monkeypatch.setenv("VRAG_APPROX_PASS", "0")
    exact = corpus.search_multistage_batch([("ap", False, k)], qs)                   # 4 fp32-exact queries per tile
    monkeypatch.delenv("VRAG_APPROX_PASS")
    for b, q in enumerate(qs):
        assert got[b][0][1].tolist() == exact[b][0][1].tolist(), b
        np.testing.assert_allclose(got[b][0][0], exact[b][0][0], rtol=1e-6)
        s1, i1 = corpus.search("ap", q, k)                                           # single-query path
        assert got[b][0][1].tolist() == i1.tolist()
        np.testing.assert_allclose(got[b][0][0], s1, rtol=1e-6)
    # as stage 0 of a two-stage batch (the candidate hand-off sees the re-scored, exactly ordered list)
    corpus.add_synthetic_store("ap2", n, fixed_rows=140, seed=4243)
    st_a = corpus.search_multistage_batch([("ap", False, 40), ("ap2", False, 5)], qs, as_arrays=True)
    monkeypatch.setenv("VRAG_APPROX_PASS", "0")
    st_e = corpus.search_multistage_batch([("ap", False, 40), ("ap2", False, 5)], qs, as_arrays=True)
    monkeypatch.delenv("VRAG_APPROX_PASS")
    for (sa, ia, _), (se, ie, _) in zip(st_a, st_e):
        assert np.array_equal(ia, ie)
        np.testing.assert_allclose(sa, se, rtol=1e-6)
    corpus.drop_store("ap")
    corpus.drop_store("ap2")


@pytest.mark.gpu
def test_approximate_first_pass_guard_falls_back_on_near_ties(corpus):
    """Hundreds of identical pages tie at the top: the candidates of the fp16 pass cannot be proven to contain the top-k
    (no candidate beats the bound of the pages left out), the guard raises the flag and the batch is redone exactly —
    ties still resolve to the lower page id."""
    rng = np.random.default_rng(7)
    n, t = 5000, 150
    rows = rows16(5151, n * t).reshape(n, t, 128)
    rows[100:700] = rows[100]                                   # 600 identical pages
    qs = [rows[100, :20].astype(np.float32) + 0.01 * rng.standard_normal((20, 128)).astype(np.float32) for _ in range(8)]
    corpus.add_store("tie", rows.reshape(-1, 128), fixed_rows=t)
    got = corpus.search_multistage_batch([("tie", False, 10)], qs)
    for b in range(8):
        assert got[b][0][1].tolist() == list(range(100, 110)), got[b][0][1]
    corpus.drop_store("tie")
