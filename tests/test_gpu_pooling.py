"""Parity of the pooling kernels (through the C ABI / the pooling.py mirror) with the reference goldens and
the oracle. Gate: allclose(rtol=1e-3) in fp32 and equal after fp16 rounding up to 1 ulp; the
kernels accumulate in numpy's order, so the tests hold them to 2e-6 relative / 1 fp16 ulp and report how many
cases are bit-identical."""
import numpy as np
import pytest

import cases as CS
from oracle import pooling_oracle as PO

pytestmark = pytest.mark.gpu

EXACT = {"n": 0, "exact": 0}


def assert_pooled(got, want, tag=""):
    assert got.dtype == want.dtype, tag
    assert got.shape == want.shape, tag
    EXACT["n"] += 1
    if np.array_equal(got, want):
        EXACT["exact"] += 1
        return
    if want.dtype == np.float16:
        ulp = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
        assert ulp.max() <= 1, f"{tag}: fp16 outputs differ by {ulp.max()} ulp"
    else:
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=tag)


@pytest.mark.parametrize("case", CS.pooling_cases(), ids=lambda c: f"{c['key']}-{c['fn']}")
def test_pooling_functions_match_reference(case, pooling_golden):
    from visual_rag_b200.embedding import pooling as GP

    x = CS.unit_rows(case["seed"], case["n"], dtype=np.dtype(case["dtype"]).type)
    got = getattr(GP, case["fn"])(x, *case["args"], **CS.fix_kwargs(case["kwargs"]))
    assert_pooled(got, pooling_golden[case["key"]], case["key"])


@pytest.mark.parametrize("case", CS.dispatch_cases(), ids=lambda c: f"{c['key']}-{c['model'].split('/')[-1]}")
def test_pool_page_matches_reference_pipeline(case, pooling_golden, golden_index):
    from visual_rag_b200.embedding import pooling as GP

    visual = CS.unit_rows(case["seed"], case["n"])
    named = GP.pool_page(case["model"], visual, case["token_info"], max_mean_pool_vectors=case["cap"],
                         pooling_windows=case["windows"], experimental_pooling_kernel=case["kernel"],
                         colsmol_experimental_2d=case["twod"], output_dtype=np.dtype(case["out_dtype"]).type)
    meta = {e["key"]: e for e in golden_index["dispatch"]}[case["key"]]
    assert sorted(named.keys()) == meta["names"]
    for nm in meta["names"]:
        assert_pooled(named[nm], pooling_golden[f"{case['key']}::{nm}"], f"{case['key']}::{nm}")


def test_exactness_report():
    print(f"\npooling cases bit-identical to the reference: {EXACT['exact']}/{EXACT['n']}")
    assert EXACT["n"] == 0 or EXACT["exact"] >= 0.5 * EXACT["n"]


def test_maxsim_function_mirrors(maxsim_golden):
    from visual_rag_b200.embedding import pooling as GP

    q, docs = CS.bench_corpus()
    got = GP.compute_maxsim_batch(q, docs[:40])
    np.testing.assert_allclose(got, maxsim_golden["batch_scores"][:40], rtol=2e-5)
    s = GP.compute_maxsim_score(q, docs[3])
    assert isinstance(s, float) and abs(s - maxsim_golden["batch_scores"][3]) <= 2e-5 * abs(s)
    x = CS.unit_rows(1, 10, scale=False)
    assert GP.compute_maxsim_score(x, x) >= 9.0                      # reference tests/test_pooling.py:163-174
    assert GP.compute_maxsim_score(np.eye(128, dtype=np.float32)[:2], np.eye(128, dtype=np.float32)[2:4]) < 0.1
    assert GP.compute_maxsim_batch(q, []) == []


def test_score_pages_pipelined_host_upload_matches_oracle_and_store_path():
    """vrag_score_pages (the engine of compute_maxsim_score / _batch): documents handed over one pointer each, cast and
    copied by the host worker pool through pinned staging chunks. Sized to span several chunks with ragged page lengths,
    an empty page in the middle, true-fp32 and fp16 documents; scores must equal (bit for bit) those of a store built
    from the concatenated rows, and the oracle on the fp16-cast documents within the score tolerance."""
    from oracle import maxsim_oracle as MO
    from visual_rag_b200.corpus import GpuCorpus

    rng = np.random.default_rng(11)
    lens = [int(v) for v in rng.integers(1, 1400, size=60)]
    lens[17] = 0
    lens += [20000, 3, 128, 129]                       # one page larger than a 2M-value staging chunk
    q = rng.standard_normal((23, 128)).astype(np.float32)
    for dt in (np.float32, np.float16):
        docs = [(rng.standard_normal((n, 128)) * 0.7).astype(dt) for n in lens]
        with GpuCorpus(0) as c:
            got = c.score_pages(q, docs)
            rows = np.concatenate(docs, axis=0)
            c.add_store("cat", rows, page_offsets=np.concatenate([[0], np.cumsum(lens)]))
            via_store = c.score("cat", q)
            # the rows the device holds are numpy's astype(float16) of the inputs
            back = c.read_rows("cat", 0, rows.shape[0])
            assert np.array_equal(back.view(np.uint16), rows.astype(np.float16).view(np.uint16))
        assert np.array_equal(got.view(np.uint32), via_store.view(np.uint32))
        assert got[17] == -np.inf
        for i in (0, 5, 33, 60, 63):
            want = MO.maxsim_score(q, docs[i].astype(np.float16).astype(np.float32))
            assert abs(got[i] - want) <= 2e-5 * abs(want), (i, got[i], want)


def test_page_rows_in_one_call_for_every_layout():
    from visual_rag_b200.corpus import GpuCorpus

    rng = np.random.default_rng(5)
    lens = [int(v) for v in rng.integers(0, 300, size=40)]
    rows = rng.standard_normal((sum(lens), 128)).astype(np.float16)
    with GpuCorpus(0) as c:
        c.add_store("var", rows, page_offsets=np.concatenate([[0], np.cumsum(lens)]))
        assert c.page_rows("var").tolist() == lens
        c.add_store("fix", rows[: 13 * 20], fixed_rows=13)
        assert c.page_rows("fix").tolist() == [13] * 20
        c.delete_pages("var", [3, 7])               # page table
        want = list(lens)
        want[3] = want[7] = 0
        assert c.page_rows("var").tolist() == want


def test_error_behaviour_matches_reference():
    from visual_rag_b200.embedding import pooling as GP

    x = np.zeros((10, 128), dtype=np.float32)
    with pytest.raises(ValueError, match="Expected 9 visual tokens"):
        GP.colpali_row_mean_pooling(x, grid_size=3)
    with pytest.raises(ValueError, match="Expected 9 visual tokens"):
        GP.adaptive_row_mean_pooling_from_grid(x, grid_h=3, grid_w=3)
    with pytest.raises(ValueError, match="target_rows must be > 0"):
        GP.adaptive_row_mean_pooling_from_grid(x, grid_h=5, grid_w=2, target_rows=0)
    with pytest.raises(ValueError, match="num_tiles must be > 0"):
        GP.colsmol_experimental_pooling(x, 0)
    with pytest.raises(ValueError, match="window_size must be odd"):
        GP.colpali_experimental_pooling_from_rows(x, window_size=4)
    with pytest.raises(ValueError, match="window_size must be >= 1"):
        GP.weighted_row_smoothing_same_length(x, window_size=0)
    with pytest.raises(ValueError, match="Unknown kernel"):
        GP.weighted_row_smoothing_same_length(x, kernel="box")
    with pytest.raises(ValueError, match="sigma must be > 0"):
        GP.weighted_row_smoothing_same_length(x, window_size=3, kernel="gaussian", sigma=0.0)
    with pytest.raises(ValueError, match="Expected at least 16 tile vectors"):
        GP.colsmol_tile_4n_pooling_from_tiles(x, n_rows=4, n_cols=4)
    import torch

    assert GP.tile_level_mean_pooling(torch.zeros(128, 128, dtype=torch.bfloat16), 2).dtype == np.float32
    assert GP.tile_level_mean_pooling(torch.zeros(128, 128, dtype=torch.float16), 2).dtype == np.float16
    assert GP.global_pool_from_mean_pool(np.zeros((0, 128), np.float32)).tolist() == [0.0] * 128


# ------------------------------------------------------------------ bulk pooling of whole stores on the device
def _pages(rows, offs):
    return [rows[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]


def _check_store(corpus, name, want_pages):
    info = corpus.store_info(name)
    assert info["n_pages"] == len(want_pages)
    for p in (0, 1, len(want_pages) // 2, len(want_pages) - 1):
        got = corpus.read_page(name, p)
        want = np.asarray(want_pages[p])
        if want.ndim == 1:
            want = want[None, :]
        assert_pooled(got, want.astype(np.float16), f"{name}[{p}]")


def test_store_pool_colpali():
    """cfg4 ColPali: [1024,128] -> row-mean 32 -> {legacy k=3 (34), gaussian (32), triangular (32), global (1)}."""
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.embedding import pooling as GP

    n = 40
    rows = CS.unit_rows(77, n * 1024, dtype=np.float16)
    with GpuCorpus(0) as c:
        c.add_store("initial", rows, fixed_rows=1024)
        c.pool_store("initial", [GP.spec_adaptive_rows(32, 32, 32)], ["mean_pooling"])
        c.pool_store("mean_pooling", [GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"),
                                      GP.spec_global_mean(True)],
                     ["experimental_pooling", "experimental_pooling_gaussian", "experimental_pooling_triangular", "global_pooling"])
        want = [PO.pool_page("vidore/colpali-v1.3", rows[i * 1024:(i + 1) * 1024].astype(np.float32), {}, output_dtype=np.float16)
                for i in range(n)]
        _check_store(c, "mean_pooling", [w["mean_pooling"] for w in want])
        _check_store(c, "experimental_pooling", [w["experimental_pooling"] for w in want])
        _check_store(c, "global_pooling", [w["global_pooling"] for w in want])
        mp = [w["mean_pooling"] for w in want]
        _check_store(c, "experimental_pooling_gaussian", [PO.weighted_row_smoothing_same_length(m, window_size=3, kernel="gaussian") for m in mp])
        _check_store(c, "experimental_pooling_triangular", [PO.weighted_row_smoothing_same_length(m, window_size=3, kernel="triangular") for m in mp])
        assert c.store_info("experimental_pooling")["fixed_rows"] == 34
        # pooled stores are searchable right away (inverse norms were rebuilt)
        q = CS.query_rows(5, 20)
        s, ids = c.search_multistage([("mean_pooling", False, 8), ("initial", False, 3)], q)[1]
        assert len(ids) == 3


def test_store_pool_colsmol_and_colqwen():
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.embedding import pooling as GP

    rng = np.random.default_rng(3)
    with GpuCorpus(0) as c:
        # ColSmol: variable tile counts (832 = 4x3+1 tiles, 768 = 12 tiles, 800 = partial tile)
        lens = np.array([832, 768, 800, 832, 64, 130, 832])
        offs = np.concatenate([[0], np.cumsum(lens)])
        rows = CS.unit_rows(78, int(offs[-1]), dtype=np.float16)
        c.add_store("initial", rows, page_offsets=offs)
        c.pool_store("initial", [GP.spec_tile_mean(64), GP.spec_colsmol_experimental(0, 64)], ["mean_pooling", "experimental_pooling"])
        pages = _pages(rows, offs)
        _check_store(c, "mean_pooling", [PO.tile_level_mean_pooling(p, 13) for p in pages])
        _check_store(c, "experimental_pooling", [PO.colsmol_experimental_pooling(p, -(-len(p) // 64)) for p in pages])
        grids = np.array([[4, 3], [3, 4], [4, 3], [4, 3], [1, 1], [1, 2], [2, 6]], dtype=np.int32)
        c.pool_store("mean_pooling", [GP.spec_tile_4n(0, 0, True, True), GP.spec_global_mean(True)],
                     ["experimental_pooling_2d", "global_pooling"], grid_hw=grids)
        mp = [PO.tile_level_mean_pooling(p, 13) for p in pages]
        _check_store(c, "experimental_pooling_2d",
                     [PO.colsmol_tile_4n_pooling_from_tiles(m, n_rows=int(g[0]), n_cols=int(g[1])) for m, g in zip(mp, grids)])
        _check_store(c, "global_pooling", [PO.global_pool_from_mean_pool(m, np.float16) for m in mp])
        # ColQwen2.5: dynamic grids, cap 32 (visual_embedder.py:786-801), gaussian + triangular + global
        gh = rng.integers(16, 48, size=25)
        gw = rng.integers(8, 33, size=25)
        lens = gh * gw
        offs = np.concatenate([[0], np.cumsum(lens)])
        rows = CS.unit_rows(79, int(offs[-1]), dtype=np.float16)
        c.add_store("initial", rows, page_offsets=offs)
        c.pool_store("initial", [GP.spec_adaptive_rows(0, 0, 32, clamp_to_h=True)], ["mean_pooling"],
                     grid_hw=np.stack([gh, gw], axis=1))
        pages = _pages(rows, offs)
        mp = [PO.adaptive_row_mean_pooling_from_grid(p, grid_h=int(h), grid_w=int(w), target_rows=min(32, int(h)))
              for p, h, w in zip(pages, gh, gw)]
        _check_store(c, "mean_pooling", mp)
        c.pool_store("mean_pooling", [GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True)],
                     ["experimental_pooling", "experimental_pooling_triangular", "global_pooling"])
        _check_store(c, "experimental_pooling", [PO.weighted_row_smoothing_same_length(m, window_size=3, kernel="gaussian") for m in mp])
        _check_store(c, "experimental_pooling_triangular", [PO.weighted_row_smoothing_same_length(m, window_size=3, kernel="triangular") for m in mp])
        _check_store(c, "global_pooling", [PO.global_pool_from_mean_pool(m, np.float16) for m in mp])


def _stores_equal(c, a, b):
    ia, ib = c.store_info(a), c.store_info(b)
    assert ia["n_pages"] == ib["n_pages"] and ia["total_rows"] == ib["total_rows"]
    ra, rb = c.read_rows(a, 0, ia["total_rows"]), c.read_rows(b, 0, ib["total_rows"])
    assert np.array_equal(ra.view(np.uint16), rb.view(np.uint16)), (a, b)
    for p in (0, ia["n_pages"] - 1):
        assert c.page_range(a, p) == c.page_range(b, p)


def test_store_pool_chained_specs_single_pass():
    """input_spec chaining: the derived stores computed inside the token pass (pooled rows kept in shared memory)
    are bit-identical to deriving them from the stored pooled rows in a second call (checked against the oracle
    above), for the ColPali, ColSmol and ColQwen2.5 pipelines."""
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.embedding import pooling as GP

    rng = np.random.default_rng(4)
    with GpuCorpus(0) as c:
        # ColPali
        c.add_store("initial", CS.unit_rows(81, 60 * 1024, dtype=np.float16), fixed_rows=1024)
        c.pool_store("initial", [GP.spec_adaptive_rows(32, 32, 32)], ["mp"])
        two = [GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True)]
        c.pool_store("mp", two, ["e", "eg", "et", "g"])
        one = [GP.spec_adaptive_rows(32, 32, 32)] + [GP.derived_from(s, 0) for s in
                                                     (GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"),
                                                      GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True))]
        c.pool_store("initial", one, ["mp1", "e1", "eg1", "et1", "g1"])
        for a, b in (("mp", "mp1"), ("e", "e1"), ("eg", "eg1"), ("et", "et1"), ("g", "g1")):
            _stores_equal(c, a, b)
        # ColSmol (variable tile counts, per-page tile grids)
        lens = np.array([832, 768, 800, 832, 64, 130, 832])
        offs = np.concatenate([[0], np.cumsum(lens)])
        grids = np.array([[4, 3], [3, 4], [4, 3], [4, 3], [1, 1], [1, 2], [2, 6]], dtype=np.int32)
        c.add_store("initial", CS.unit_rows(82, int(offs[-1]), dtype=np.float16), page_offsets=offs)
        c.pool_store("initial", [GP.spec_tile_mean(64), GP.spec_colsmol_experimental(0, 64)], ["mp", "e"])
        c.pool_store("mp", [GP.spec_tile_4n(0, 0, True, True), GP.spec_global_mean(True)], ["e2d", "g"], grid_hw=grids)
        c.pool_store("initial", [GP.spec_tile_mean(64), GP.spec_colsmol_experimental(0, 64),
                                 GP.derived_from(GP.spec_tile_4n(0, 0, True, True), 0), GP.derived_from(GP.spec_global_mean(True), 0)],
                     ["mp1", "e1", "e2d1", "g1"], grid_hw=grids)
        for a, b in (("mp", "mp1"), ("e", "e1"), ("e2d", "e2d1"), ("g", "g1")):
            _stores_equal(c, a, b)
        # ColQwen2.5 (dynamic grids, cap 32)
        gh = rng.integers(16, 48, size=31)
        gw = rng.integers(8, 33, size=31)
        offs = np.concatenate([[0], np.cumsum(gh * gw)])
        ghw = np.stack([gh, gw], axis=1)
        c.add_store("initial", CS.unit_rows(83, int(offs[-1]), dtype=np.float16), page_offsets=offs)
        c.pool_store("initial", [GP.spec_adaptive_rows(0, 0, 32, clamp_to_h=True)], ["mp"], grid_hw=ghw)
        c.pool_store("mp", [GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True)], ["e", "et", "g"])
        c.pool_store("initial", [GP.spec_adaptive_rows(0, 0, 32, clamp_to_h=True), GP.derived_from(GP.spec_smooth(3, "gaussian"), 0),
                                 GP.derived_from(GP.spec_smooth(3, "triangular"), 0), GP.derived_from(GP.spec_global_mean(True), 0)],
                     ["mp1", "e1", "et1", "g1"], grid_hw=ghw)
        for a, b in (("mp", "mp1"), ("e", "e1"), ("et", "et1"), ("g", "g1")):
            _stores_equal(c, a, b)
        with pytest.raises(Exception):
            c.pool_store("initial", [GP.derived_from(GP.spec_smooth(3, "gaussian"), 0)], ["x"])


def test_bulk_repool_from_initial_matches_reference_script():
    """recompute_pooling_from_initial (one device pass) vs the reference script's per-point arithmetic (goldens made by
    running its functions), rounded to the fp16 store dtype."""
    import json
    import os

    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.embedding.repool import recompute_pooling_from_initial

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = np.load(os.path.join(here, "repool_golden.npz"))
    for cap in (32, 0):
        cases = [c for c in CS.repool_cases() if c["cap"] == cap]
        mats = [CS.unit_rows(c["seed"], c["n"], dtype=np.float16) for c in cases]
        payloads = [({"resized_width": c["w"], "resized_height": c["h"]} if c["w"] else {}) for c in cases]
        off = np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])])
        with GpuCorpus(0) as c:
            c.add_store("initial", np.concatenate(mats), page_offsets=off)
            info = recompute_pooling_from_initial(c, payloads, max_mean_pool_vectors=cap)
            assert info["pages"] == len(cases)
            for nm, gk in (("mean_pooling", "mean_pooling"), ("experimental_pooling", "experimental_pooling_gaussian"),
                           ("experimental_pooling_gaussian", "experimental_pooling_gaussian"),
                           ("experimental_pooling_triangular", "experimental_pooling_triangular"), ("global_pooling", "global_pooling")):
                for p, cs in enumerate(cases):
                    want = gold[f"{cs['key']}::{gk}"]
                    if want.ndim == 1:
                        want = want[None, :]
                    assert_pooled(c.read_page(nm, p), want.astype(np.float16), f"{cs['key']}::{nm}")
