"""Host-side logic of the retriever mirrors on CPU: they are driven through the in-memory client double
(tests/golden/fake_qdrant.py, scoring with the oracle) and must reproduce what the REFERENCE retriever classes
returned on the same corpus (tests/golden/golden_index.json)."""
import numpy as np
import pytest

import cases as CS
from fake_qdrant import NumpyQdrant
from oracle import maxsim_oracle as MO
from visual_rag_b200.retrieval import (MultiVectorRetriever, SingleStageRetriever, ThreeStageRetriever,
                                       TwoStageRetriever)
from visual_rag_b200.retrieval._common import resolve_stage1


def _split(rows, offsets):
    return [rows[offsets[i]:offsets[i + 1]].astype(np.float32) for i in range(len(offsets) - 1)]


@pytest.fixture(scope="module")
def setup(retrieval_golden, golden_index):
    q, initial = CS.retrieval_corpus()
    off = retrieval_golden["offsets_pooled"]
    vectors = {
        "initial": initial,
        "mean_pooling": _split(retrieval_golden["mean_pooling"], off),
        "experimental_pooling": _split(retrieval_golden["experimental_pooling"], off),
        "global_pooling": [g.astype(np.float32) for g in retrieval_golden["global_pooling"]],
    }
    return q, NumpyQdrant(vectors, MO.maxsim_score), golden_index["retrieval"]


def _ids(res):
    return [r["id"] for r in res]


def test_two_stage_search_matches_reference(setup):
    q, client, want = setup
    r = TwoStageRetriever(client, "c")
    for mode in ("pooled_query_vs_tiles", "tokens_vs_tiles", "pooled_query_vs_global"):
        got = r.search(q, top_k=10, prefetch_k=40, stage1_mode=mode)
        w = want[f"two_stage_search::{mode}"]
        assert _ids(got) == _ids(w)
        for g, x in zip(got, w):
            assert g["score_final"] == x["score_final"] and g["score_stage1"] == x["score_stage1"]
            assert g["score_stage2"] == x["score_stage2"] and "payload" in g
    got = r.search(q, top_k=10, prefetch_k=40, stage1_mode="tokens_vs_tiles", use_reranking=False)
    assert _ids(got) == _ids(want["two_stage_search::norerank"])
    assert all(g["score_final"] == g["score_stage1"] for g in got)


def test_two_stage_search_accepts_new_vocabulary(setup):
    """The reference's search() rejects its own default stage1_mode; the mirror accepts both."""
    q, client, want = setup
    r = TwoStageRetriever(client, "c")
    a = r.search(q, top_k=10, prefetch_k=40)  # default pooled_query_vs_standard_pooling
    assert _ids(a) == _ids(want["two_stage_search::pooled_query_vs_tiles"])
    b = r.search(q, top_k=10, prefetch_k=40, stage1_mode="tokens_vs_experimental_pooling")
    assert len(b) == 10
    with pytest.raises(ValueError, match="Unknown stage1_mode"):
        r.search(q, stage1_mode="bogus")
    with pytest.raises(ValueError, match="Unknown stage1_mode"):
        r.search_server_side(q, stage1_mode="bogus")


def test_two_stage_server_side_matches_reference(setup):
    q, client, want = setup
    r = TwoStageRetriever(client, "c")
    for key, w in want.items():
        if key.startswith("two_stage_server::"):
            got = r.search_server_side(q, top_k=10, prefetch_k=40, stage1_mode=key.split("::")[1])
            assert _ids(got) == _ids(w), key
            assert [g["score_final"] for g in got] == [x["score_final"] for x in w]
            assert all(g["score_stage1"] is None and g["score_stage2"] == g["score_final"] for g in got)
    for use_pooling in (False, True):
        got = r.search_single_stage(q, top_k=10, use_pooling=use_pooling)
        assert _ids(got) == _ids(want[f"two_stage_single::{use_pooling}"])


def test_default_prefetch_k(setup):
    q, client, _ = setup
    r = TwoStageRetriever(client, "c")
    client.calls.clear()
    r.search_server_side(q, top_k=3)
    r.search(q, top_k=20, stage1_mode="tokens_vs_tiles")
    limits = [c["limit"] for c in client.calls]
    assert limits[0] == 3          # outer limit; prefetch limit = max(100, 10*top_k) is inside Prefetch
    assert limits[1] == 200        # stage-1 prefetch of search(): max(100, 10*20)


def test_three_stage_matches_reference(setup):
    q, client, want = setup
    r = ThreeStageRetriever(client, "c")
    got = r.search_server_side(query_embedding=q, top_k=10, stage1_k=80, stage2_k=30)
    w = want["three_stage"]
    assert _ids(got) == _ids(w)
    for g, x in zip(got, w):
        for k in ("score_stage1", "score_stage2", "score_stage3", "score_final"):
            assert g[k] == x[k]
    # accepts-and-ignores stage1_mode, None sizes default to 1000/300 (reference raises here)
    got2 = r.search_server_side(query_embedding=q, top_k=10, stage1_k=None, stage2_k=None, stage1_mode="x")
    assert len(got2) == 10


def test_single_stage_matches_reference(setup):
    q, client, want = setup
    r = SingleStageRetriever(client, "c")
    for strat in ("multi_vector", "tiles_maxsim", "pooled_tile", "pooled_global", "experimental_maxsim",
                  "pooled_experimental"):
        got = r.search(q, top_k=10, strategy=strat)
        w = want[f"single::{strat}"]
        assert _ids(got) == _ids(w) and [g["score"] for g in got] == [x["score"] for x in w]
        assert all(g["score"] == g["score_final"] for g in got)
    with pytest.raises(ValueError, match="Unknown strategy"):
        r.search(q, strategy="nope")


def test_multi_vector_dispatch(setup):
    q, client, want = setup
    mv = MultiVectorRetriever("c", qdrant_client=client)
    assert _ids(mv.search_embedded(query_embedding=q, top_k=10, mode="single_full")) == _ids(want["single::multi_vector"])
    assert _ids(mv.search_embedded(query_embedding=q, top_k=10, mode="single_tiles")) == _ids(want["single::tiles_maxsim"])
    assert _ids(mv.search_embedded(query_embedding=q, top_k=10, mode="single_pooled")) == _ids(want["single::pooled_tile"])
    assert _ids(mv.search_embedded(query_embedding=q, top_k=10, mode="single_global")) == _ids(want["single::pooled_global"])
    got = mv.search_embedded(query_embedding=q, top_k=10, mode="two_stage", prefetch_k=40,
                             stage1_mode="tokens_vs_standard_pooling")
    assert _ids(got) == _ids(want["two_stage_server::tokens_vs_standard_pooling"])
    got = mv.search_embedded(query_embedding=q, top_k=10, mode="three_stage", stage1_k=80, stage2_k=30)
    assert _ids(got) == _ids(want["three_stage"])
    with pytest.raises(ValueError, match="Unknown mode"):
        mv.search_embedded(query_embedding=q, mode="nope")
    with pytest.raises(ValueError):
        MultiVectorRetriever("c")
    with pytest.raises(ValueError):
        mv.search("a text query")  # no embedder


def test_retry_and_torch_queries(setup):
    import torch

    q, client, want = setup

    class Flaky:
        def __init__(self, inner, fails):
            self.inner, self.fails = inner, fails

        def query_points(self, **kw):
            if self.fails > 0:
                self.fails -= 1
                raise RuntimeError("transient")
            return self.inner.query_points(**kw)

    r = TwoStageRetriever(Flaky(client, 2), "c", retry_sleep=0.0)
    got = r.search_server_side(torch.from_numpy(q).to(torch.bfloat16).float(), top_k=10, prefetch_k=40,
                               stage1_mode="tokens_vs_tiles")
    assert len(got) == 10
    r = TwoStageRetriever(Flaky(client, 5), "c", retry_sleep=0.0)
    with pytest.raises(RuntimeError, match="transient"):
        r.search_server_side(q)

    class Bad:
        calls = 0

        def query_points(self, **kw):
            Bad.calls += 1
            raise ValueError("k=5000 exceeds the supported maximum 4096 results per search stage")

    with pytest.raises(ValueError, match="exceeds"):          # deterministic errors are not retried (no back-off sleeps)
        TwoStageRetriever(Bad(), "c", retry_sleep=10.0).search_server_side(q)
    assert Bad.calls == 1


def test_build_filter_and_stage1_table():
    r = TwoStageRetriever(None, "c")
    assert r.build_filter() is None
    f = r.build_filter(year="2020", source=["a", "b"], has_text=True)
    keys = [c.key for c in f.must]
    assert keys == ["year", "source", "has_text"]
    assert f.must[0].match.value == 2020 and f.must[1].match.any == ["a", "b"]
    assert resolve_stage1("tokens_vs_tiles", "p", "e", "g") == (False, "p")
    assert resolve_stage1("pooled_query_vs_experimental", "p", "e", "g") == (True, "e")
    assert resolve_stage1("pooled_query_vs_global", "p", "e", "g") == (True, "g")


def test_infer_grid_mirror_matches_reference_goldens():
    """visual_rag_b200.embedding.repool.infer_grid (host logic) against outputs of the reference script's _infer_grid."""
    import json
    import os

    import cases as CS
    from visual_rag_b200.embedding.repool import infer_grid

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    idx = json.load(open(os.path.join(here, "repool_index.json")))
    for (n, w, h), want in zip(CS.infer_grid_cases(), idx["infer_grid"]):
        assert list(infer_grid(n, width=w, height=h)) == want
    import pytest
    with pytest.raises(ValueError):
        infer_grid(0, width=1, height=1)


class _ArrayCorpus:
    """Stand-in for GpuCorpus.search_multistage_batch(final_only=True): hands back prepared arrays and records the call."""

    page_base = 100

    def __init__(self, sc, ids, st, cnt):
        self.arrays = (sc, ids, st, cnt)
        self.calls = []

    def search_multistage_batch(self, stages, queries, normalize=True, stage_queries=None, as_arrays=False, final_only=False):
        self.calls.append((list(stages), queries, stage_queries, final_only))
        return self.arrays


def test_batched_retriever_results_from_columnar_client_lists():
    """Host logic of the batched retriever calls: short rows, -inf padding, NaN stage scores -> None, external ids,
    payload lookup, and the stage layout handed to the corpus (stage 1 pooled on the device, three_stage.py:96)."""
    from visual_rag_b200.client import GpuCorpusClient

    inf = np.float32(np.inf)
    sc = np.array([[3.0, 2.0, 1.0], [5.0, -inf, -inf]], dtype=np.float32)
    ids = np.array([[102, 100, 101], [101, -1, -1]], dtype=np.int64)
    st = np.array([[[0.3, 0.6], [0.2, np.nan], [0.1, 0.4]], [[0.9, 0.8], [np.nan, np.nan], [np.nan, np.nan]]], dtype=np.float32)
    cnt = np.array([3, 1], dtype=np.int32)
    corpus = _ArrayCorpus(sc, ids, st, cnt)
    client = GpuCorpusClient(corpus, "c", point_ids=["a", "b", "c"], payloads=[{"n": 0}, {"n": 1}, {"n": 2}])
    three = ThreeStageRetriever(client, "c")
    qs = [np.ones((4, 128), np.float32), np.ones((7, 128), np.float32)]
    res = three.search_server_side_batch(query_embeddings=qs, top_k=3, stage1_k=50, stage2_k=20)
    stages, queries, stage_queries, final_only = corpus.calls[-1]
    assert stages == [("global_pooling", True, 50), ("experimental_pooling", False, 20), ("initial", False, 3)]
    assert final_only and stage_queries is None and len(queries) == 2
    assert [r["id"] for r in res[0]] == ["c", "a", "b"] and [r["id"] for r in res[1]] == ["b"]
    assert res[0][1]["score_stage2"] is None and res[0][1]["score_stage1"] == pytest.approx(0.2)
    assert res[0][0] == {"id": "c", "score_stage1": pytest.approx(0.3), "score_stage2": pytest.approx(0.6),
                         "score_stage3": 3.0, "score_final": 3.0, "payload": {"n": 2}}
    assert res[1][0]["payload"] == {"n": 1} and res[1][0]["score_final"] == 5.0

    two = TwoStageRetriever(client, "c")
    res2 = two.search_server_side_batch(qs, top_k=3, prefetch_k=40, stage1_mode="pooled_query_vs_global")
    stages, _, _, _ = corpus.calls[-1]
    assert stages == [("global_pooling", True, 40), ("initial", False, 3)]
    assert res2[0][2] == {"id": "b", "score_stage1": None, "score_stage2": 1.0, "score_final": 1.0, "payload": {"n": 1}}
    stages = two.search_server_side_batch(qs, top_k=3, stage1_mode="tokens_vs_tiles") and corpus.calls[-1][0]
    assert stages == [("mean_pooling", False, 100), ("initial", False, 3)]

    # default ids / no payloads, explicit per-stage query matrices
    plain = GpuCorpusClient(corpus, "c")
    cols = plain.query_multistage_batch_final(usings=["mean_pooling", "initial"], limits=[40, 3],
                                              stage_queries=[[q.mean(axis=0), q] for q in qs], with_payload=False)
    assert cols[0][0] == [102, 100, 101] and cols[0][3] == [None, None, None] and cols[1][1] == [5.0]
    assert corpus.calls[-1][2] is not None and corpus.calls[-1][2][0][0].shape == (1, 128)
    with pytest.raises(ValueError):
        plain.query_multistage_batch_final(usings=["initial"], limits=[3])


class _ListCorpus:
    """Host-side stand-in for the store methods GpuIndexer drives (append / replace / page ranges)."""

    page_base = 0

    def __init__(self):
        self.stores = {}
        self.log = []
        self.fail_on = None

    def has_store(self, name):
        return name in self.stores

    def n_pages(self, name):
        return len(self.stores[name])

    def page_range(self, name, local_page):
        rows = [m.shape[0] for m in self.stores[name]]
        return int(sum(rows[:local_page])), int(rows[local_page])

    def append_store(self, name, rows, page_offsets=None, fixed_rows=0):
        if self.fail_on == ("append", name):
            raise RuntimeError("device error (injected)")
        off = np.asarray(page_offsets)
        self.stores.setdefault(name, []).extend(rows[off[i]:off[i + 1]] for i in range(len(off) - 1))
        self.log.append(("append", name, len(off) - 1))

    def replace_pages(self, name, local_pages, rows, page_offsets):
        if self.fail_on == ("replace", name):
            raise RuntimeError("device error (injected)")
        off = np.asarray(page_offsets)
        for j, pg in enumerate(local_pages):
            self.stores[name][pg] = rows[off[j]:off[j + 1]]        # any shape: the store keeps a page table
        self.log.append(("replace", name, list(local_pages)))

    def delete_pages(self, name, local_pages):
        for pg in local_pages:
            self.stores[name][pg] = self.stores[name][pg][:0]
        self.log.append(("delete", name, list(local_pages)))

    def truncate_store(self, name, n_pages):
        del self.stores[name][n_pages:]
        self.log.append(("truncate", name, n_pages))

    def compact_store(self, name):
        self.log.append(("compact", name))

    def drop_store(self, name):
        del self.stores[name]


def test_indexer_upsert_host_logic():
    """GpuIndexer.upload_batch without a GPU: client.upsert semantics (qdrant_indexer.py:459-507) — new ids appended, existing
    ids replaced with their payload (whatever their new shape), the last occurrence of an id inside a batch wins, named vectors
    are optional per point, a failed batch is rolled back and reports 0, deletes keep page indices stable."""
    from visual_rag_b200.indexing import GpuIndexer

    def point(i, t, r, ver, with_exp=True):
        g = np.random.default_rng(100 * ver + i)
        p = {"id": f"id{i}", "visual_embedding": g.standard_normal((t, 128)).astype(np.float32),
             "tile_pooled_embedding": g.standard_normal((r, 128)).astype(np.float32), "metadata": {"i": i, "ver": ver}}
        if with_exp:
            p["experimental_pooled_embedding"] = {"experimental_pooling": g.standard_normal((r, 128)).astype(np.float32)}
        return p

    c = _ListCorpus()
    idx = GpuIndexer(c, "c")
    assert idx.create_collection() and idx.upload_batch([]) == 0
    assert idx.upload_batch([point(i, 10 + i, 3, 1) for i in range(4)]) == 4
    assert c.n_pages("initial") == 4 and set(c.stores) == {"initial", "mean_pooling", "global_pooling", "experimental_pooling"}
    assert c.stores["initial"][2].dtype == np.float16                       # store dtype cast
    np.testing.assert_array_equal(c.stores["global_pooling"][1],           # global = mean(tile_pooled) fallback, 417-421
                                  point(1, 11, 3, 1)["tile_pooled_embedding"].mean(axis=0).reshape(1, -1).astype(np.float16))
    c.log.clear()
    batch = [point(2, 12, 3, 2), point(7, 30, 5, 2), point(2, 12, 3, 3), point(0, 10, 3, 2)]
    assert idx.upload_batch(batch) == 4
    assert ("replace", "initial", [2, 0]) in c.log and ("append", "initial", 1) in c.log
    assert idx.client._ids == ["id0", "id1", "id2", "id3", "id7"]
    assert [p["ver"] for p in idx.client._payloads] == [2, 1, 3, 1, 2]          # last occurrence of id2 (ver 3) won
    np.testing.assert_array_equal(c.stores["initial"][2], point(2, 12, 3, 3)["visual_embedding"].astype(np.float16))
    assert idx.check_exists("id7") and not idx.check_exists("id9") and idx.get_existing_ids() == set(idx.client._ids)
    # ---- an upsert may change a point's shape (client.upsert has no shape constraint): id1 grows from 11 to 99 tokens
    c.log.clear()
    assert idx.upload_batch([point(8, 9, 2, 4), point(1, 99, 3, 4)]) == 2
    assert c.stores["initial"][1].shape == (99, 128) and ("replace", "initial", [1]) in c.log
    assert idx.client._ids == ["id0", "id1", "id2", "id3", "id7", "id8"]
    # ---- named vectors are optional per point (pipeline.py:485-503): id10 has no experimental vector -> an empty page there;
    # a named vector that appears late (experimental_pooling_2d) gives the earlier points empty pages
    p11 = point(11, 9, 2, 4)
    p11["experimental_pooled_embedding"]["experimental_pooling_2d"] = np.ones((13, 128), np.float32)
    assert idx.upload_batch([point(9, 9, 2, 4), point(10, 9, 2, 4, with_exp=False), p11]) == 3
    n = len(idx.client._ids)
    assert n == 9 and all(c.n_pages(nm) == n for nm in c.stores)
    assert c.stores["experimental_pooling"][7].shape == (0, 128) and c.stores["experimental_pooling"][6].shape == (2, 128)
    assert [m.shape[0] for m in c.stores["experimental_pooling_2d"]] == [0] * 8 + [13]
    # ---- a failed batch leaves nothing behind and reports 0 (the reference logs and returns 0, qdrant_indexer.py:497-507)
    before = {nm: [m.copy() for m in v] for nm, v in c.stores.items()}
    c.fail_on = ("append", "mean_pooling")                              # the second store of the batch fails on the device
    assert idx.upload_batch([point(20, 5, 2, 5), point(21, 6, 2, 5)]) == 0
    c.fail_on = None
    assert len(idx.client._ids) == n and not idx.check_exists("id20")
    assert all(len(c.stores[nm]) == n and all(np.array_equal(a, b) for a, b in zip(before[nm], c.stores[nm])) for nm in before)
    assert idx.upload_batch([{"id": "bad", "visual_embedding": np.zeros((3, 64)), "tile_pooled_embedding": np.zeros((2, 64))}]) == 0
    assert idx.upload_batch([{"id": "bad2", "metadata": {}}]) == 0        # malformed point: logged, 0 uploaded, nothing written
    assert len(idx.client._ids) == n
    # ---- deletes: the pages stay (empty), the ids leave the collection
    assert idx.delete_points(["id2", "nope"]) == 1
    assert not idx.check_exists("id2") and c.stores["initial"][2].shape == (0, 128) and idx.client._ids[2] is None
    assert idx.upload_batch([point(2, 7, 2, 6)]) == 1 and idx.client._page("id2") == n    # re-inserted as a new page
    idx.compact()
    assert ("compact", "initial") in c.log
    assert idx.create_collection() is False and idx.create_collection(force_recreate=True) and not c.stores
    assert idx.client._ids == []


# ------------------------------------------------------------------ host-side ingest cast (qdrant_indexer.py:423-441)
def _cast_with_library(x: np.ndarray, threads: int, force_scalar: bool) -> np.ndarray:
    import ctypes as C

    from visual_rag_b200 import _native

    lib = _native.load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    rc = lib.vrag_host_f32_to_f16(x.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_uint16)), x.size,
                                  threads, int(force_scalar))
    assert rc == 0
    return out


@pytest.mark.parametrize("force_scalar", [False, True], ids=["f16c", "scalar"])
@pytest.mark.parametrize("threads", [1, 4])
def test_host_ingest_cast_is_numpy_astype_float16_bit_for_bit(force_scalar, threads):
    """The cast the library applies to fp32 rows arriving in host memory equals numpy's astype(float16) — what
    QdrantIndexer._build_qdrant_points stores — on every class of value: random bit patterns (all exponents), exact ties,
    fp16 denormals, the overflow boundary, signed zeros and infinities."""
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2**32, size=1 << 19, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals = vals[~np.isnan(vals)]
    # every fp16 value, the midpoints between neighbours (ties) and their fp32 neighbours
    h = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16).astype(np.float32)
    mid = (h[:-1] + h[1:]) * 0.5
    ties = np.concatenate([mid, np.nextafter(mid, np.float32(np.inf)), np.nextafter(mid, np.float32(-np.inf))])
    edge = np.array([0.0, -0.0, np.inf, -np.inf, 65504.0, 65519.99, 65520.0, 65536.0, 1e30, 2.0**-24, 2.0**-25,
                     np.nextafter(np.float32(2.0**-25), np.float32(1)), 2.0**-26, 6.1e-5, 5.96e-8, 1e-45], dtype=np.float32)
    x = np.concatenate([vals, h, -h, ties, -ties, edge, -edge,
                        rng.standard_normal(300_000).astype(np.float32)]).astype(np.float32)
    got = _cast_with_library(x, threads, force_scalar)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (x[bad[:5]], got[bad[:5]], want[bad[:5]])


def test_host_ingest_cast_handles_odd_lengths_and_empty():
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 15, 17, 1000003):
        x = rng.standard_normal(n).astype(np.float32) * 3
        for threads in (1, 8):
            got = _cast_with_library(x, threads, False)
            assert np.array_equal(got, x.astype(np.float16).view(np.uint16))


def test_infer_grids_vectorised_equals_per_page_inference():
    """repool.infer_grids (one infer_grid per distinct (tokens, width, height), scattered back) == infer_grid per page,
    with and without payload sizes, including pages whose payload lacks a usable size."""
    from visual_rag_b200.embedding.repool import _payload_size, infer_grid, infer_grids

    rng = np.random.default_rng(3)
    tokens = rng.integers(1, 769, size=500)
    payloads = []
    for i in range(500):
        kind = i % 4
        if kind == 0:
            payloads.append({"resized_width": int(rng.integers(200, 1200)), "resized_height": int(rng.integers(200, 1200))})
        elif kind == 1:
            payloads.append({"original_width": 800, "original_height": 600, "cropped_width": 640, "cropped_height": 640})
        elif kind == 2:
            payloads.append(None)
        else:
            payloads.append({"resized_width": "bad"})
    got = infer_grids(tokens, payloads)
    assert got.shape == (500, 2) and got.dtype == np.int32
    for p in range(500):
        w, h = _payload_size(payloads[p])
        assert tuple(got[p]) == infer_grid(int(tokens[p]), width=w, height=h)
        assert int(got[p, 0]) * int(got[p, 1]) == int(tokens[p])
    no_payload = infer_grids(tokens)
    for p in range(0, 500, 17):
        assert tuple(no_payload[p]) == infer_grid(int(tokens[p]))


def test_host_worker_pool_survives_fork():
    """A process that forks after the library's worker pool exists (multiprocessing's default start method) gets a fresh
    pool in the child instead of waiting for threads that were not copied."""
    import multiprocessing as mp

    x = np.random.default_rng(0).standard_normal(1 << 20).astype(np.float32)
    want = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(_cast_with_library(x, 8, False), want)      # the parent's pool exists now

    def child(q):
        q.put(bool(np.array_equal(_cast_with_library(x, 8, False), want)))

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=child, args=(q,))
    p.start()
    p.join(60)
    assert not p.is_alive(), "child hung in the worker pool"
    assert q.get(timeout=5) is True


def test_batch_with_a_stage_beyond_the_batched_limit_runs_query_by_query():
    """GpuCorpus.search_multistage_batch with a stage size > MAX_K_BATCH: the per-query fallback must reproduce the three
    result formats of the native batched call (lists, as_arrays, final_only with earlier-stage scores; NaN if absent).
    Host logic only: search_multistage is replaced by the oracle over small host stores."""
    from visual_rag_b200 import corpus as GC

    n = 40
    pooled = [CS.unit_rows(4000 + i, 3) for i in range(n)]
    full = [CS.unit_rows(5000 + i, 9) for i in range(n)]
    stores = {"p": pooled, "f": full}
    calls = []

    class Fake(GC.GpuCorpus):
        def __init__(self):   # no device
            pass

        def search_multistage(self, stages, query, normalize=True, stage_queries=None, candidate_ids=None, fp16_query=False,
                              filter_id=None):
            calls.append(1)
            qs = stage_queries if stage_queries is not None else [query] * len(stages)
            out, cand = [], None
            for (name, pool, k), q in zip(stages, qs):
                q = np.asarray(q, np.float32)
                if pool:
                    q = q.mean(axis=0, keepdims=True)
                ids = list(range(n)) if cand is None else cand
                sc = np.array([MO.maxsim_score(q, stores[name][i].astype(np.float32)) for i in ids], np.float32)
                order = np.lexsort((np.arange(len(ids)), -sc))[: min(k, len(ids))]
                out.append((sc[order], np.asarray(ids, np.int64)[order]))
                cand = [ids[j] for j in order]
            return out

    c = Fake()
    big = GC.MAX_K_BATCH + 1
    stages = [("p", False, big), ("f", False, 5)]
    queries = [CS.query_rows(6000 + b, 4 + b) for b in range(3)]
    lists = c.search_multistage_batch(stages, queries)
    assert len(calls) == 3 and len(lists) == 3
    for b, q in enumerate(queries):
        want = c.search_multistage(stages, q)
        assert lists[b][0][1].tolist() == want[0][1].tolist() and lists[b][1][1].tolist() == want[1][1].tolist()
    arr = c.search_multistage_batch(stages, GC.pack_queries(queries), as_arrays=True)
    assert arr[0][0].shape == (3, big) and arr[1][1].shape == (3, 5)
    assert arr[0][2].tolist() == [n, n, n] and arr[1][2].tolist() == [5, 5, 5]
    assert np.isneginf(arr[0][0][:, n:]).all() and (arr[0][1][:, n:] == -1).all()
    for b in range(3):
        assert arr[1][1][b].tolist() == lists[b][1][1].tolist()
        np.testing.assert_array_equal(arr[0][0][b, :n], lists[b][0][0])
    f_sc, f_id, f_st, f_cnt = c.search_multistage_batch(stages, queries, final_only=True)
    assert f_id.shape == (3, 5) and f_st.shape == (3, 5, 1) and f_cnt.tolist() == [5, 5, 5]
    for b in range(3):
        s1 = dict(zip(lists[b][0][1].tolist(), lists[b][0][0].tolist()))
        assert f_id[b].tolist() == lists[b][1][1].tolist()
        np.testing.assert_array_equal(f_st[b, :, 0], np.array([s1[i] for i in f_id[b]], np.float32))
    # a final page that is absent from an earlier list reports NaN for that stage (cannot happen in a chained search: forced)
    per_stage = [[q, q] for q in queries]
    f2 = c.search_multistage_batch(stages, None, stage_queries=per_stage, final_only=True)
    assert f2[1].tolist() == f_id.tolist()
    with pytest.raises(ValueError, match="exceeds"):
        c.search_multistage_batch([("p", False, GC.MAX_K + 1)], queries)
