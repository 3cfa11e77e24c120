"""True-fp32 inputs (values that do NOT survive an fp16 round trip): the oracle against outputs of the reference itself
(tests/golden/fp32_golden.npz, made by tests/golden/make_golden_fp32.py from /root/reference) on the CPU, and the CUDA
path against the same goldens on the GPU.

Gates: pooling <= 1 fp16 ulp after rounding / rtol 2e-6 in fp32 (the kernels add in numpy's order); MaxSim scores of
fp32 pages held in the fp16 store within 1e-3 relative of the reference's fp32 arithmetic (BASELINE.md section 5 — the
worst case over the cases is printed and asserted to stay below 5e-4); top-k ids identical up to swaps between scores
closer than that deviation."""
import json
import os

import numpy as np
import pytest

import cases as CS
from conftest import GOLDEN
from oracle import maxsim_oracle as MO
from oracle import pooling_oracle as PO


@pytest.fixture(scope="module")
def fp32_golden():
    return np.load(os.path.join(GOLDEN, "fp32_golden.npz"))


@pytest.fixture(scope="module")
def fp32_index():
    with open(os.path.join(GOLDEN, "fp32_index.json")) as f:
        return json.load(f)


# ------------------------------------------------------------------ CPU: the oracle is pinned on these inputs too
@pytest.mark.parametrize("case", CS.fp32_pooling_cases(), ids=lambda c: f"{c['key']}-{c['fn']}")
def test_oracle_pooling_fp32_inputs_bit_exact(case, fp32_golden, fp32_index):
    x = CS.raw_rows(case["seed"], case["n"])
    meta = {e["key"]: e for e in fp32_index["pooling"]}[case["key"]]
    assert CS.checksum(x) == meta["in_crc"]
    assert not np.array_equal(x, x.astype(np.float16).astype(np.float32)), "input must not be fp16-representable"
    got = getattr(PO, case["fn"])(x, *case["args"], **CS.fix_kwargs(case["kwargs"]))
    want = fp32_golden[case["key"]]
    assert got.dtype == want.dtype and got.shape == want.shape
    np.testing.assert_array_equal(got, want)


def test_oracle_maxsim_fp32_inputs_bit_exact(fp32_golden, fp32_index):
    for case in CS.fp32_maxsim_cases():
        q = CS.query_rows(case["seed"], case["q"])
        d = CS.raw_rows(case["seed"] + 1, case["t"])
        assert [CS.checksum(q), CS.checksum(d)] == {e["key"]: e for e in fp32_index["maxsim"]}[case["key"]]["in_crc"]
        want = fp32_golden[case["key"]]
        assert MO.maxsim_score(q, d) == want[0] and MO.maxsim_score(q, d, normalize=False) == want[1]
    q, docs = CS.fp32_corpus()
    assert [CS.checksum(q), CS.checksum(np.concatenate(docs))] == fp32_index["corpus_in_crc"]
    ex = MO.search_exhaustive(q, docs, 10)
    assert [i for i, _ in ex] == fp32_golden["exhaustive_ids"].tolist()
    np.testing.assert_array_equal([s for _, s in ex], fp32_golden["exhaustive_scores"])
    pooled = [PO.tile_level_mean_pooling(d, 0, patches_per_tile=32) for d in docs]
    ts = MO.search_two_stage_pooled(q, docs, pooled, prefetch_k=40, top_k=10)
    assert [r[0] for r in ts] == fp32_golden["two_stage_ids"].tolist()
    np.testing.assert_array_equal([r[1] for r in ts], fp32_golden["two_stage_scores"])


# ------------------------------------------------------------------ GPU
POOL_STATS = {"n": 0, "exact": 0}


@pytest.mark.gpu
@pytest.mark.parametrize("case", CS.fp32_pooling_cases(), ids=lambda c: f"{c['key']}-{c['fn']}")
def test_gpu_pooling_fp32_inputs(case, fp32_golden):
    from visual_rag_b200.embedding import pooling as GP

    x = CS.raw_rows(case["seed"], case["n"])
    got = getattr(GP, case["fn"])(x, *case["args"], **CS.fix_kwargs(case["kwargs"]))
    want = fp32_golden[case["key"]]
    assert got.dtype == want.dtype and got.shape == want.shape
    POOL_STATS["n"] += 1
    if np.array_equal(got, want):
        POOL_STATS["exact"] += 1
        return
    if want.dtype == np.float16:
        ulp = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
        assert ulp.max() <= 1, f"{case['key']}: fp16 outputs differ by {ulp.max()} ulp"
    else:
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=case["key"])
        # and equal after rounding to the fp16 store dtype up to 1 ulp
        g16, w16 = got.astype(np.float16), want.astype(np.float16)
        ulp = np.abs(g16.view(np.int16).astype(np.int32) - w16.view(np.int16).astype(np.int32))
        assert ulp.max() <= 1


@pytest.mark.gpu
def test_gpu_pooling_fp32_exactness_report():
    print(f"\nfp32-input pooling cases bit-identical to the reference: {POOL_STATS['exact']}/{POOL_STATS['n']}")
    assert POOL_STATS["n"] == 0 or POOL_STATS["exact"] >= 0.5 * POOL_STATS["n"]


@pytest.mark.gpu
def test_gpu_scores_of_fp32_pages_through_the_fp16_store(fp32_golden):
    """The reference scores fp32 pages in fp32; the GPU store holds them as fp16 (qdrant_indexer.py:423-441 with the CLI's
    default float16 collection). The deviation this introduces must stay inside the 1e-3 parity gate; the worst case is
    reported."""
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.embedding import pooling as GP

    worst = 0.0
    with GpuCorpus(0) as c:
        for case in CS.fp32_maxsim_cases():
            q = CS.query_rows(case["seed"], case["q"])
            d = CS.raw_rows(case["seed"] + 1, case["t"])
            want = fp32_golden[case["key"]]
            c.add_store("d", d, fixed_rows=case["t"])            # fp32 rows in: cast to the fp16 store dtype on ingest
            got = (float(c.score("d", q)[0]), float(c.score("d", q, normalize=False)[0]))
            for g, w in zip(got, want):
                rel = abs(g - w) / abs(w)
                worst = max(worst, rel)
                assert rel <= 1e-3, (case["key"], g, w)
            # the per-call twin takes the fp32 page as is and rounds it the same way
            assert abs(GP.compute_maxsim_score(q, d) - want[0]) <= 1e-3 * abs(want[0])
        q, docs = CS.fp32_corpus()
        off = np.concatenate([[0], np.cumsum([len(d) for d in docs])])
        c.add_store("initial", np.concatenate(docs), page_offsets=off)
        sc = c.score("initial", q)
        ref = fp32_golden["corpus_scores"]
        rel = np.abs(sc - ref) / np.abs(ref)
        worst = max(worst, float(rel.max()))
        assert rel.max() <= 1e-3
        s, ids = c.search("initial", q, 10)
        want_ids = fp32_golden["exhaustive_ids"].tolist()
        tol = 2 * float(rel.max()) * float(np.abs(ref).max())
        for j, (g, w) in enumerate(zip(ids.tolist(), want_ids)):     # swaps only between scores closer than the deviation
            assert g == w or abs(ref[g] - ref[w]) <= tol, (j, g, w)
        np.testing.assert_allclose(s, ref[ids], rtol=1e-3)
    print(f"\nworst relative deviation of fp16-stored fp32 pages vs the reference's fp32 scores: {worst:.2e}")
    assert worst <= 5e-4
