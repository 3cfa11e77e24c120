"""The reference's OWN retriever classes, unmodified, on top of `GpuCorpusClient` (the drop-in seam).

The reference is imported from `/root/reference` where that exists (this container) and otherwise from `baseline/_ref`
— the verbatim install made by baseline/install_ref.py, which travels to the GPU box. Two variants: on the CPU the corpus
behind the client is a numpy double of `GpuCorpus` that scores with the oracle; the `gpu`-marked variant puts the same
corpus into a real `GpuCorpus` on cuda:0, so the reference's classes drive the CUDA kernels. What is exercised is the seam: the exact `query_points` / `retrieve` call shapes of
visual_rag/retrieval/{two_stage,three_stage,single_stage}.py — `prefetch=[Prefetch(...)]`, `Filter(must=[HasIdCondition])`,
`with_vectors=[name]`, list-typed queries — and the result objects the reference reads back (`.points[*].id/.score/
.payload`, `.vector[name]`). Results must equal the goldens the same classes produced on the in-memory Qdrant double."""
import os
import sys

import numpy as np
import pytest

import cases as CS
from fake_qdrant import install_qdrant_stub
from oracle import maxsim_oracle as MO

from conftest import ROOT

REF = next((p for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")) if os.path.isdir(os.path.join(p, "visual_rag"))), None)
pytestmark = pytest.mark.skipif(REF is None, reason="reference not present (neither /root/reference nor baseline/_ref)")


class OracleCorpus:
    """Duck-typed GpuCorpus (the methods GpuCorpusClient calls), scoring with the oracle. Test infrastructure."""

    page_base = 0

    def __init__(self, stores):
        self.stores = stores

    def n_pages(self, name):
        return len(self.stores[name])

    def has_store(self, name):
        return name in self.stores

    def store_info(self, name):
        return {"n_pages": len(self.stores[name])}

    def read_page(self, name, local_page):
        return self.stores[name][local_page].astype(np.float16)

    def search(self, name, query, k, normalize=True, pool_query=False, candidate_ids=None):
        return self.search_multistage([(name, pool_query, k)], query, candidate_ids=candidate_ids)[0]

    def search_multistage(self, stages, query, normalize=True, stage_queries=None, candidate_ids=None):
        cand = None if candidate_ids is None else [int(i) for i in candidate_ids]
        out = []
        for s, (name, pool, k) in enumerate(stages):
            q = np.asarray(stage_queries[s] if stage_queries is not None else query, dtype=np.float32)
            q = q[None, :] if q.ndim == 1 else q
            idx = list(range(len(self.stores[name]))) if cand is None else cand
            sc = MO.stage_scores(q, self.stores[name], bool(pool), idx)
            order = MO.stable_topk(sc, int(k))
            cand = [idx[j] for j in order]
            out.append((np.asarray([sc[j] for j in order], np.float32), np.asarray(cand, np.int64)))
        return out


def _reference_classes():
    install_qdrant_stub()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from visual_rag.retrieval.single_stage import SingleStageRetriever
    from visual_rag.retrieval.three_stage import ThreeStageRetriever
    from visual_rag.retrieval.two_stage import TwoStageRetriever

    return TwoStageRetriever, ThreeStageRetriever, SingleStageRetriever


def _stores(retrieval_golden):
    q, initial = CS.retrieval_corpus()
    off = retrieval_golden["offsets_pooled"]
    split = lambda rows: [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(len(off) - 1)]  # noqa: E731
    stores = {"initial": initial, "mean_pooling": split(retrieval_golden["mean_pooling"]),
              "experimental_pooling": split(retrieval_golden["experimental_pooling"]),
              "global_pooling": [g.astype(np.float32)[None, :] for g in retrieval_golden["global_pooling"]]}
    return q, stores, [{"page": i, "year": 2000 + i % 3} for i in range(len(initial))]


@pytest.fixture(scope="module")
def ref_setup_gpu(retrieval_golden, golden_index):
    """The same corpus in a real GpuCorpus on cuda:0 behind the product's GpuCorpusClient."""
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.corpus import GpuCorpus

    q, stores, payloads = _stores(retrieval_golden)
    corpus = GpuCorpus(0)
    for name, pages in stores.items():
        off = np.concatenate([[0], np.cumsum([len(p) for p in pages])])
        corpus.add_store(name, np.concatenate(pages).astype(np.float16), page_offsets=off)
    yield q, GpuCorpusClient(corpus, "c", payloads=payloads), golden_index["retrieval"], _reference_classes()
    corpus.close()


@pytest.fixture(scope="module")
def ref_setup(retrieval_golden, golden_index):
    from visual_rag_b200.client import GpuCorpusClient

    q, stores, payloads = _stores(retrieval_golden)
    client = GpuCorpusClient(OracleCorpus(stores), "c", payloads=payloads)
    return q, client, golden_index["retrieval"], _reference_classes()


def _ids(res):
    return [r["id"] for r in res]


def test_reference_two_stage_classes_run_on_the_client(ref_setup):
    q, client, gold, (Two, _, _) = ref_setup
    two = Two(client, "c")                                    # the reference class, imported from /root/reference
    assert type(two).__module__ == "visual_rag.retrieval.two_stage"
    for mode in ("pooled_query_vs_tiles", "tokens_vs_tiles", "pooled_query_vs_global"):
        res = two.search(q, top_k=10, prefetch_k=40, stage1_mode=mode)      # query_points + retrieve + numpy rerank
        g = gold[f"two_stage_search::{mode}"]
        assert _ids(res) == [x["id"] for x in g]
        np.testing.assert_allclose([r["score_final"] for r in res], [x["score_final"] for x in g], rtol=1e-3)
        assert res[0]["payload"]["page"] == res[0]["id"]
    for mode in ("pooled_query_vs_standard_pooling", "tokens_vs_standard_pooling", "pooled_query_vs_experimental_pooling",
                 "tokens_vs_experimental_pooling", "pooled_query_vs_global", "tokens_vs_tiles"):
        res = two.search_server_side(q, top_k=10, prefetch_k=40, stage1_mode=mode)   # query_points(prefetch=[Prefetch])
        g = gold[f"two_stage_server::{mode}"]
        assert _ids(res) == [x["id"] for x in g]
        np.testing.assert_allclose([r["score_final"] for r in res], [x["score_final"] for x in g], rtol=1e-3)
    for use_pooling in (False, True):
        res = two.search_single_stage(q, top_k=10, use_pooling=use_pooling)
        assert _ids(res) == [x["id"] for x in gold[f"two_stage_single::{use_pooling}"]]
    f = two.build_filter(year=2001)                           # FieldCondition / MatchValue from the (stubbed) qdrant models
    res = two.search_server_side(q, top_k=5, prefetch_k=20, filter_obj=f, stage1_mode="tokens_vs_tiles")
    assert len(res) == 5 and all(r["payload"]["year"] == 2001 for r in res)


def test_reference_three_and_single_stage_classes_run_on_the_client(ref_setup):
    q, client, gold, (_, Three, Single) = ref_setup
    three = Three(client, "c")
    assert type(three).__module__ == "visual_rag.retrieval.three_stage"
    res = three.search_server_side(query_embedding=q, top_k=10, stage1_k=80, stage2_k=30)   # HasIdCondition-restricted stages
    g = gold["three_stage"]
    assert _ids(res) == [x["id"] for x in g]
    for key in ("score_stage1", "score_stage2", "score_stage3"):
        np.testing.assert_allclose([r[key] for r in res], [x[key] for x in g], rtol=1e-3)
    single = Single(client, "c")
    for strat in ("multi_vector", "tiles_maxsim", "pooled_tile", "pooled_global", "experimental_maxsim", "pooled_experimental"):
        res = single.search(q, top_k=10, strategy=strat)
        assert _ids(res) == [x["id"] for x in gold[f"single::{strat}"]], strat


# ------------------------------------------------------------------ the same, on a real GpuCorpus (cuda:0)
@pytest.mark.gpu
def test_reference_two_stage_classes_drive_the_cuda_kernels(ref_setup_gpu):
    test_reference_two_stage_classes_run_on_the_client(ref_setup_gpu)
    q, client, gold, (Two, _, _) = ref_setup_gpu
    assert client.corpus.launch_count() > 0
    # the reference's client-side rerank with embeddings pulled through retrieve(with_vectors=[...])
    res = Two(client, "c").search(q, top_k=5, prefetch_k=20, stage1_mode="tokens_vs_tiles", return_embeddings=True)
    assert all(np.asarray(r["embedding"]).shape[1] == 128 for r in res)


@pytest.mark.gpu
def test_reference_three_and_single_stage_classes_drive_the_cuda_kernels(ref_setup_gpu):
    test_reference_three_and_single_stage_classes_run_on_the_client(ref_setup_gpu)
