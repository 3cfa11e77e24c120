#!/usr/bin/env python
"""Benchmark of the hot path: exhaustive ColBERT MaxSim over a GPU-resident, page-sharded corpus.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's CPU algorithm on the host cores)

Workload (BASELINE.json configs[3] per-GPU shard; weak scaling): every GPU holds
`--pages-per-gpu` (default 500,000) ColPali-v1.3-shaped pages x 1030 tokens x 128-d fp16
(131.8 GB/GPU — 4M pages at N=8, i.e. cfg3; 1M pages at N=2, i.e. the cfg1 corpus) generated on the
device from a counter-based seeded generator. One step = one 20-token query scored against EVERY page
(exact MaxSim), exact top-10 per shard, NCCL all-gather merge of the per-shard lists.
  value  : pages/s, corpus and query resident in HBM, CUDA-event timed, max over ranks.
  e2e    : same metric through the reference-facing seam — SingleStageRetriever.search(strategy="multi_vector") on the
           (Sharded)GpuCorpusClient: numpy query in, result dicts out; H2D of the query and D2H of the result inside the
           timed region, wall-clock, max over ranks.
  extra  : two-stage (tokens_vs_standard_pooling, prefetch_k=256, top-10) QPS / p50 / p95 through the C ABI and through
           TwoStageRetriever.search_server_side, with the device time of every collective broken out; at N > 1 the
           `sharded_parity` block (exhaustive / two-stage / three-stage-batch / filtered lists over a host-generated corpus
           striped across the ranks == the oracle on the whole corpus, incl. a cross-shard exact tie), the cfg1
           strong-scaling run (1M pages TOTAL) and cfg4 pooling over all GPUs; at N = 1 the other BASELINE configs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "visual-rag-toolkit_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

TOKENS = 1030          # ColPali-v1.3: 1024 visual + 6 instruction tokens
VISUAL_TOKENS = 1024
POOLED_ROWS = 32       # mean_pooling rows per page (colpali_row_mean_pooling)
Q_TOKENS = 20
TOP_K = 10
PREFETCH_K = 256
SEED = 20260318


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return 1400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
def load_reference():
    """The UNMODIFIED reference, installed by baseline/install_ref.py into baseline/_ref (git-ignored, shipped to the GPU
    box): returns its `benchmarks.quick_test` module (search_exhaustive / search_two_stage, quick_test.py:158-206, which
    call visual_rag.embedding.pooling.compute_maxsim_score) or None when it is not installed."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "benchmarks")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import logging

        logging.disable(logging.INFO)
        from benchmarks import quick_test   # the reference's own module, nothing of this repository on that path

        return quick_test
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: reference import failed ({e!r}); falling back to the oracle port", file=sys.stderr)
        return None


def cpu_search_fn():
    """(search_exhaustive(query, docs_f32_list, k) -> [(index, score)], kind) — the reference's own function when
    baseline/_ref is there (kind "reference"), else the oracle port (kind "port")."""
    qt = load_reference()
    if qt is not None:
        def run(query, docs, k):
            res = qt.search_exhaustive(query, {i: {"embedding": d} for i, d in enumerate(docs)}, top_k=k)
            return [(r["id"], r["score"]) for r in res]

        return run, "reference"
    from oracle import maxsim_oracle as MO

    return (lambda query, docs, k: MO.search_exhaustive(query, docs, k)), "port"


def cpu_sample_pages_per_s(docs_f32, queries, min_seconds=10.0, max_passes=200):
    """The reference's client-side CPU path (quick_test.search_exhaustive, quick_test.py:158-166) timed on a bounded
    sample: one query after the other over the sampled pages until `min_seconds` of CPU work are done.
    Returns (pages/s, seconds, blas_threads, passes, kind)."""
    fn, kind = cpu_search_fn()
    t0 = time.perf_counter()
    passes = 0
    while True:
        fn(queries[passes % len(queries)], docs_f32, TOP_K)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or passes >= max_passes:
            break
    threads = 1
    try:
        from threadpoolctl import threadpool_info

        threads = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        pass
    return passes * len(docs_f32) / dt, dt, threads, passes, kind


def host_sample(n_pages, seed):
    """Seeded ColPali-shaped pages on the host (fp16-representable fp32), for the CPU arms."""
    rng = np.random.default_rng(seed)
    docs = []
    for _ in range(n_pages):
        x = rng.standard_normal((TOKENS, 128), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        docs.append(x.astype(np.float16).astype(np.float32))
    return docs


_REF_FN = None   # set in the parent before the pool forks, so that the workers inherit the imported reference


def _ref_worker(args):
    n_pages, seed, qseed, reps = args
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    docs = host_sample(n_pages, seed)
    q = np.random.default_rng(qseed).standard_normal((Q_TOKENS, 128)).astype(np.float32)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        _REF_FN(q, docs, TOP_K)
        times.append(time.perf_counter() - t0)
    return times


def workload_config(pages, world):
    """`config` of the JSON line — the SAME dict on both arms."""
    total_pages = pages * world
    return {
        "workload": f"exhaustive MaxSim top-{TOP_K} over {total_pages} ColPali-shaped pages "
                    f"({pages}/GPU x {TOKENS} tok x 128-d fp16 = {pages * TOKENS * 256 / 1e9:.1f} GB/GPU; "
                    f"BASELINE configs[3] shard), {Q_TOKENS}-token query, NCCL all-gather top-k merge",
        "pages_per_gpu": pages, "tokens_per_page": TOKENS, "query_tokens": Q_TOKENS, "top_k": TOP_K,
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path — benchmarks/quick_test.py::search_exhaustive
    (-> visual_rag.embedding.pooling.compute_maxsim_score), imported unmodified from baseline/_ref — on all host cores: the
    reference is a single-process per-page Python loop, so the page sample of a step is split over one process per
    core, each running that function on its slice; a step = one query over the whole sample. Falls back to the oracle
    port (kind "port") only when baseline/_ref is missing."""
    global _REF_FN
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    _REF_FN, kind = cpu_search_fn()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    pages_per_worker = max(8, args.ref_sample_pages // workers)
    total_pages = pages_per_worker * workers
    reps = args.warmup + args.steps
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        res = pool.map(_ref_worker, [(pages_per_worker, SEED + 1000 + w, SEED + 7, reps) for w in range(workers)])
    # a step finishes when the slowest worker has scored its slice
    step_times = [max(r[i] for r in res) for i in range(args.warmup, reps)]
    dt = sum(step_times)
    value = total_pages * args.steps / dt
    sample = (f"each step scores a bounded sample of {total_pages} pages of this workload, split over {workers} processes "
              f"(1 BLAS thread each); " + ("benchmarks/quick_test.py::search_exhaustive of the unmodified reference (baseline/_ref)"
                                           if kind == "reference" else "oracle/maxsim_oracle.py::search_exhaustive (port)"))
    line = {
        "impl": "reference", "metric": "exhaustive_maxsim_pages_per_s", "value": value, "unit": "pages/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.pages_per_gpu, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "pages/s", "cores": workers, "kind": kind, "sample": sample,
                         "pages_per_step": total_pages},
        "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
def sharded_parity_check(local_rank, rank, world):
    """N > 1, on hardware: a host-generated corpus (every rank generates the same seeded pages) is striped across the
    ranks in contiguous page ranges and searched through the reference-facing classes on a ShardedCorpusClient; every
    rank compares the merged GLOBAL lists with the oracle run over the WHOLE corpus: exhaustive top-10, the two-stage
    (256 -> 10) lists, a three-stage batch, a payload-filtered search, a full ranking (k = all pages: the long-merge
    path) and a cross-shard exact tie (two identical pages on the first and the last rank: lower global id first).
    Returns a dict for the JSON line; never raises (a failure is reported as sharded_parity: false with the reason)."""
    from oracle import maxsim_oracle as MO
    from visual_rag_b200.client import ShardedCorpusClient
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.distributed import shard_page_range
    from visual_rag_b200.retrieval import SingleStageRetriever, ThreeStageRetriever, TwoStageRetriever

    P = 4096
    out = {"sharded_parity": False, "pages": P, "checks": []}
    c2 = None
    try:
        rng = np.random.default_rng(SEED + 4242)
        lens = rng.integers(129, 261, size=P)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        rows = rng.standard_normal((int(off[-1]), 128), dtype=np.float32)
        rows /= np.linalg.norm(rows, axis=1, keepdims=True)
        rows = rows.astype(np.float16)
        tie_a, tie_b = 5, P - 2                      # page P-2 (last rank) := page 5 (rank 0), same row count
        la, lb = int(lens[tie_a]), int(lens[tie_b])
        m = min(la, lb)
        rows[off[tie_b]:off[tie_b] + m] = rows[off[tie_a]:off[tie_a] + m]
        if lb > m:                                     # longer copy: repeat rows (max over tokens is unchanged)
            rows[off[tie_b] + m:off[tie_b + 1]] = rows[off[tie_a]:off[tie_a] + (lb - m)]
        if la > m:
            rows[off[tie_a] + m:off[tie_a + 1]] = rows[off[tie_a]:off[tie_a] + (la - m)]
        docs = [rows[off[i]:off[i + 1]].astype(np.float32) for i in range(P)]
        pooled = [d[:32] for d in docs]
        glob = [d.mean(axis=0, keepdims=True).astype(np.float16).astype(np.float32) for d in docs]
        payloads = [{"page": i, "year": 2000 + i % 3} for i in range(P)]
        b, e = shard_page_range(P, rank, world)
        c2 = GpuCorpus(local_rank, page_base=b)
        c2.comm_init_torch()
        c2.add_store("initial", rows[off[b]:off[e]], page_offsets=off[b:e + 1] - off[b])
        c2.add_store("mean_pooling", np.concatenate(pooled[b:e]).astype(np.float16), fixed_rows=32)
        c2.add_store("global_pooling", np.concatenate(glob[b:e]).astype(np.float16), fixed_rows=1)
        client = ShardedCorpusClient(c2, "parity", payloads=payloads[b:e])
        single = SingleStageRetriever(client, "parity")
        two = TwoStageRetriever(client, "parity")
        three = ThreeStageRetriever(client, "parity", experimental_vector_name="mean_pooling")
        qs = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(6)]

        def same(got, want, what):
            gi, wi = [g["id"] for g in got], [i for i, _ in want]
            if gi != wi:
                raise AssertionError(f"{what}: ids differ {gi[:12]} vs {wi[:12]}")
            gs, ws = np.array([g["score_final"] for g in got]), np.array([x for _, x in want])
            if not np.allclose(gs, ws, rtol=2e-5, atol=2e-6):
                raise AssertionError(f"{what}: scores differ by {np.abs(gs - ws).max()}")
            out["checks"].append(what)

        for j, q in enumerate(qs[:3]):
            same(single.search(q, top_k=10, strategy="multi_vector"), MO.search_exhaustive(q, docs, 10), f"exhaustive_top10_q{j}")
            ref = MO.multistage(q, [(pooled, False, PREFETCH_K), (docs, False, TOP_K)])
            same(two.search_server_side(q, top_k=TOP_K, prefetch_k=PREFETCH_K, stage1_mode="tokens_vs_standard_pooling"), ref[1],
                 f"two_stage_256_10_q{j}")
        got = three.search_server_side_batch(query_embeddings=qs, top_k=10, stage1_k=100, stage2_k=30)
        for j, q in enumerate(qs):
            ref = MO.multistage(q, [(glob, True, 100), (pooled, False, 30), (docs, False, 10)])
            same(got[j], ref[2], f"three_stage_batch_q{j}")
        # longer candidate lists: both candidate stages take the owned-candidates path (count -> compact -> scan -> scatter)
        got = three.search_server_side_batch(query_embeddings=qs[:3], top_k=20, stage1_k=512, stage2_k=256)
        for j, q in enumerate(qs[:3]):
            ref = MO.multistage(q, [(glob, True, 512), (pooled, False, 256), (docs, False, 20)])
            same(got[j], ref[2], f"three_stage_batch_long_lists_q{j}")
        keep = [i for i in range(P) if payloads[i]["year"] == 2001]
        ref = MO.search_exhaustive(qs[0], [docs[i] for i in keep], 10)
        same(single.search(qs[0], top_k=10, strategy="multi_vector", filter_obj=two.build_filter(year=2001)),
             [(keep[i], x) for i, x in ref], "filtered_top10_year2001")
        full = single.search(qs[1], top_k=P, strategy="multi_vector")
        ids = [g["id"] for g in full]
        want = MO.search_exhaustive(qs[1], docs, P)
        if sorted(ids) != list(range(P)):
            raise AssertionError("full ranking is not a permutation of all pages")
        if ids.index(tie_a) + 1 != ids.index(tie_b):
            raise AssertionError(f"cross-shard tie: page {tie_a} at {ids.index(tie_a)}, page {tie_b} at {ids.index(tie_b)}")
        if ids[:200] != [i for i, _ in want[:200]]:
            raise AssertionError("full ranking: first 200 ids differ from the oracle")
        out["checks"].append("full_ranking_k4096_cross_shard_tie")
        # ---- shards of >= 8192 pages: the host-facing search takes the SAMPLED local top-k; its miss flag travels in the
        # exchanged entries. Store A: random pages (estimate holds). Store B: 90 % identical pages (massive ties: the
        # estimate misses on every rank and all ranks redo the search exactly, together).
        n_big, r_big = 12288, 8
        big = rng.standard_normal((n_big * world * r_big, 128), dtype=np.float32)
        big /= np.linalg.norm(big, axis=1, keepdims=True)
        big = big.astype(np.float16).reshape(n_big * world, r_big, 128)
        tied = big.copy()
        tied[np.arange(n_big * world) % 10 != 0] = big[1]
        lo, hi = rank * n_big, (rank + 1) * n_big
        c3 = GpuCorpus(local_rank, page_base=lo)
        try:
            c3.comm_init_torch()
            c3.add_store("a", big[lo:hi].reshape(-1, 128), fixed_rows=r_big)
            c3.add_store("b", tied[lo:hi].reshape(-1, 128), fixed_rows=r_big)
            q_tie = big[1, :4].astype(np.float32)    # the tied pages are this query's best matches: ~90 % of all scores tie at the top
            for name, arr, qq, what in (("a", big, qs[2], "sampled_topk_through_exchange"),
                                        ("b", tied, q_tie, "sampled_miss_flag_redo_all_ranks")):
                sc, ids = c3.search(name, qq, 50)
                want = MO.search_exhaustive(qq, [x.astype(np.float32) for x in arr], 50)
                if [int(i) for i in ids] != [i for i, _ in want]:
                    raise AssertionError(f"{what}: ids differ {ids[:8].tolist()} vs {[i for i, _ in want[:8]]}")
                if not np.allclose(sc, [x for _, x in want], rtol=2e-5, atol=2e-6):
                    raise AssertionError(f"{what}: scores differ")
                out["checks"].append(what)
        finally:
            c3.close()
        # ---- sharded ingest: every rank uploads the points routed to it (stable hash of the id), upserts one of them with
        # a new shape, then all ranks sync the id / payload tables; searches see the union, with external ids and payloads
        from visual_rag_b200.client import owner_rank_of_id
        from visual_rag_b200.indexing import GpuIndexer

        c4 = GpuCorpus(local_rank, page_base=rank << 32)       # ingest shards own disjoint, sparse id ranges
        try:
            c4.comm_init_torch()
            cl4 = ShardedCorpusClient(c4, "ingest", point_ids=[], payloads=[])
            idx = GpuIndexer(c4, "ingest", client=cl4)
            n_pts = 600

            def point(i, ver):
                g = np.random.default_rng(7000 + 10 * i + ver)
                t = int(g.integers(130, 200)) + 7 * ver
                v = g.standard_normal((t, 128)).astype(np.float32)
                v /= np.linalg.norm(v, axis=1, keepdims=True)
                return {"id": f"doc-{i:05d}", "visual_embedding": v.astype(np.float16).astype(np.float32),
                        "tile_pooled_embedding": v[:8].astype(np.float16).astype(np.float32), "metadata": {"i": i, "ver": ver}}

            mine = [i for i in range(n_pts) if owner_rank_of_id(f"doc-{i:05d}", world) == rank]
            for lo in range(0, len(mine), 64):
                assert idx.upload_batch([point(i, 0) for i in mine[lo:lo + 64]]) == len(mine[lo:lo + 64])
            changed = [i for i in mine if i % 50 == 0]
            if changed:
                assert idx.upload_batch([point(i, 1) for i in changed]) == len(changed)     # upsert with a new token count
            cl4.sync_points()
            final = [point(i, 1 if i % 50 == 0 else 0) for i in range(n_pts)]
            got = SingleStageRetriever(cl4, "ingest").search(qs[3], top_k=10, strategy="multi_vector")
            want = MO.search_exhaustive(qs[3], [p_["visual_embedding"] for p_ in final], 10)
            if [g["id"] for g in got] != [final[i]["id"] for i, _ in want]:
                raise AssertionError(f"sharded ingest: ids differ {[g['id'] for g in got][:5]} vs {[final[i]['id'] for i, _ in want][:5]}")
            if not np.allclose([g["score"] for g in got], [x for _, x in want], rtol=2e-5):
                raise AssertionError("sharded ingest: scores differ")
            if [g["payload"] for g in got] != [final[i]["metadata"] for i, _ in want]:
                raise AssertionError("sharded ingest: payloads differ")
            if cl4.get_collection("ingest").points_count != n_pts:
                raise AssertionError("sharded ingest: points_count")
            out["checks"].append("sharded_ingest_upsert_sync_search")
        finally:
            c4.close()
        out["sharded_parity"] = True
        out["comm_us_last_search"] = c2.comm_timing_us()
        # ---- (informational, does not gate sharded_parity) shards of >= 65536 pages: the dense batched stage takes the fused
        # top-k prefilter on every rank and sends its lists as packed hits; the merged lists must equal those of the score-matrix
        # path (VRAG_PREFILTER=0), bit for bit
        try:
            n5 = 70_000
            c5 = GpuCorpus(local_rank, page_base=rank * n5)
            try:
                c5.comm_init_torch()
                c5.add_synthetic_store("g", n5, fixed_rows=1, seed=SEED + 31, row_seed_base=rank * n5)
                qrng = np.random.default_rng(SEED + 32)
                qb = [qrng.standard_normal((int(qrng.integers(10, 31)), 128)).astype(np.float32) for _ in range(16)]
                got5 = c5.search_multistage_batch([("g", True, 500)], qb, as_arrays=True)
                old_env = os.environ.get("VRAG_PREFILTER")
                os.environ["VRAG_PREFILTER"] = "0"
                try:
                    want5 = c5.search_multistage_batch([("g", True, 500)], qb, as_arrays=True)
                finally:
                    if old_env is None:
                        os.environ.pop("VRAG_PREFILTER", None)
                    else:
                        os.environ["VRAG_PREFILTER"] = old_env
                out["sharded_prefilter_equals_matrix_path"] = bool(np.array_equal(got5[0][1], want5[0][1])
                                                                   and np.array_equal(got5[0][0], want5[0][0]))
            finally:
                c5.close()
        except Exception as e5:  # noqa: BLE001
            out["sharded_prefilter_equals_matrix_path"] = "error: " + repr(e5)[:200]
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:400]
    finally:
        if c2 is not None:
            try:
                c2.close()
            except Exception:
                pass
    return out


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    lib = os.path.join(ROOT, "visual-rag-toolkit_b200", "visual_rag_b200", "libvrag_b200.so")
    if not os.path.exists(lib):   # git-ignored build artefact: a bare checkout builds it once (rank 0), the others wait
        if local_rank == 0:
            import __graft_entry__

            __graft_entry__.build()
        if dist is not None:
            dist.barrier()
    from visual_rag_b200.client import GpuCorpusClient, ShardedCorpusClient
    from visual_rag_b200.corpus import GpuCorpus, query_flags
    from visual_rag_b200.distributed import ShardedSearcher
    from visual_rag_b200.embedding import pooling as GP
    from visual_rag_b200.retrieval import SingleStageRetriever, TwoStageRetriever

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    pages = args.pages_per_gpu
    corpus = GpuCorpus(local_rank, page_base=rank * pages)
    if world > 1:
        corpus.comm_init_torch()          # the library's own communicator (C ABI: vrag_comm_init); the id travels via torch
    t_gen0 = time.perf_counter()
    corpus.add_synthetic_store("initial", pages, fixed_rows=TOKENS, seed=SEED, row_seed_base=rank * pages * TOKENS)
    gen_s = time.perf_counter() - t_gen0
    # cfg1: mean_pooling = colpali_row_mean_pooling (p2) of the page's 1024 VISUAL tokens (32 x 32 grid -> 32 rows); the
    # 6 instruction tokens behind them are part of `initial` only. Derived on the device.
    mean_spec = GP.with_token_window(GP.spec_adaptive_rows(32, 32, POOLED_ROWS), 0, VISUAL_TOKENS)
    corpus.pool_store("initial", [mean_spec], ["mean_pooling"])            # first call: one-time kernel attribute setup
    pool_ms = corpus.pool_store("initial", [mean_spec], ["mean_pooling"])
    searcher = ShardedSearcher(corpus)
    client = ShardedCorpusClient(corpus, "bench") if world > 1 else GpuCorpusClient(corpus, "bench")
    single = SingleStageRetriever(client, "bench")
    two = TwoStageRetriever(client, "bench")
    rng = np.random.default_rng(SEED + 7)
    queries = [rng.standard_normal((Q_TOKENS, 128)).astype(np.float32) for _ in range(max(8, args.steps + args.warmup))]

    ex_stage = [("initial", False, TOP_K)]
    ts_stages = [("mean_pooling", False, PREFETCH_K), ("initial", False, TOP_K)]

    # ---------------- correctness spot check against the oracle (rank 0, a few pages read back) ----------------
    if rank == 0:
        from oracle import maxsim_oracle as MO

        sc = corpus.score("initial", queries[0], candidate_ids=[corpus.page_base + 3, corpus.page_base + pages - 1])
        for s_, p_ in zip(sc, (3, pages - 1)):
            want = MO.maxsim_score(queries[0], corpus.read_page("initial", p_).astype(np.float32))
            assert abs(s_ - want) <= 1e-3 * abs(want), ("parity spot check failed", s_, want)

    # ---------------- device-resident throughput (value) ----------------
    nq = searcher.upload_query(queries[0])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        searcher.search_multistage_device(ex_stage, nq)
    barrier()
    if rank == 0:
        sampler.rows.clear()       # keep only samples taken during the timed region
    launches0 = corpus.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        dev_res = searcher.search_multistage_device(ex_stage, nq)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = corpus.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_top = (dev_res[0][0].cpu().numpy(), dev_res[0][1].cpu().numpy())

    # ---------------- dominant kernel alone (roofline numerator), same stream, CUDA events ----------------
    scores_buf = torch.empty((pages,), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        corpus.score_dev("initial", searcher._q_dev.data_ptr(), nq, query_flags(True, False), 0, pages, scores_buf.data_ptr(), stream)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        corpus.score_dev("initial", searcher._q_dev.data_ptr(), nq, query_flags(True, False), 0, pages, scores_buf.data_ptr(), stream)
    k1.record()
    torch.cuda.synchronize()
    # each score_dev = query_prep (tiny) + the scan kernel; the prep kernel is ~2 us of the ~20 ms
    kern_ms = k0.elapsed_time(k1) / args.steps
    # same launch with the opt-in fp16 query operand (VRAG_Q_FP16): half the tensor work, scores within ~1e-4 relative
    fl16 = query_flags(True, False, True)
    corpus.score_dev("initial", searcher._q_dev.data_ptr(), nq, fl16, 0, pages, scores_buf.data_ptr(), stream)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        corpus.score_dev("initial", searcher._q_dev.data_ptr(), nq, fl16, 0, pages, scores_buf.data_ptr(), stream)
    k1.record()
    torch.cuda.synchronize()
    kern16_ms = k0.elapsed_time(k1) / args.steps

    # ---------------- end to end through the reference-facing seam ----------------
    # SingleStageRetriever.search(strategy="multi_vector") -> client.query_points -> C ABI (collective at N > 1):
    # numpy query in (H2D inside), result dicts out (D2H inside). The host-facing search may use the sampled top-k;
    # its list must equal the device-resident (exact radix-select) list of the same query.
    for i in range(args.warmup):
        single.search(queries[i % len(queries)], top_k=TOP_K, strategy="multi_vector")
    barrier()
    t0 = time.perf_counter()
    last = None
    e2e_dev_ms = 0.0
    for i in range(args.steps):
        last = single.search(queries[i % len(queries)], top_k=TOP_K, strategy="multi_vector")
        e2e_dev_ms += corpus.last_timing_ms()[0]       # device time of the call (scan + top-k + exchange), CUDA events
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    first = single.search(queries[0], top_k=TOP_K, strategy="multi_vector")
    host_equals_device = ([h["id"] for h in first] == [int(i) for i in dev_top[1]]
                          and np.allclose([h["score"] for h in first], dev_top[0], rtol=1e-6))

    # ---------------- two-stage latency (extra): C ABI and retriever class ----------------
    def two_stage_latency(n_lat):
        for i in range(5):
            corpus.search_multistage(ts_stages, queries[i % len(queries)])
        barrier()
        lat, comm = [], []
        t_all0 = time.perf_counter()
        for i in range(n_lat):
            t1 = time.perf_counter()
            corpus.search_multistage(ts_stages, queries[i % len(queries)])
            lat.append(1e3 * (time.perf_counter() - t1))
            if world > 1 and i % 16 == 0:
                comm.append(corpus.comm_timing_us())
        wall = time.perf_counter() - t_all0
        dev_total_ms = corpus.last_timing_ms()[0]
        barrier()
        for i in range(5):
            two.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K, stage1_mode="tokens_vs_standard_pooling")
        barrier()
        rlat = []
        for i in range(n_lat):
            t1 = time.perf_counter()
            hits = two.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K,
                                          stage1_mode="tokens_vs_standard_pooling")
            rlat.append(1e3 * (time.perf_counter() - t1))
        assert len(hits) == TOP_K
        barrier()
        wall, p50, p95, r50, r95, dev_total_ms = max_over_ranks([wall, np.percentile(lat, 50), np.percentile(lat, 95),
                                                                np.percentile(rlat, 50), np.percentile(rlat, 95), dev_total_ms])
        res = {"qps": n_lat / wall, "p50_ms": p50, "p95_ms": p95, "queries": n_lat, "retriever_p50_ms": r50, "retriever_p95_ms": r95,
               "device_ms_last_query": dev_total_ms}
        if comm:
            med = np.median(np.array(comm), axis=0)
            res["collective_us"] = {"stage1_allgather_topk256": float(med[0]), "stage2_allreduce_max_256": float(med[1]),
                                    "note": "device time (CUDA events) of each collective, median of sampled queries, rank 0"}
        return res

    ts = two_stage_latency(args.latency_queries)

    # ---------------- max over ranks ----------------
    dev_ms, e2e_s, kern_ms, kern16_ms = max_over_ranks([dev_ms, e2e_s, kern_ms, kern16_ms])

    # ---------------- N > 1: merged lists on hardware vs the oracle on the whole corpus (all ranks take part) --------
    parity = sharded_parity_check(local_rank, rank, world) if world > 1 else None
    if parity is not None:
        ok = max_over_ranks([0.0 if parity["sharded_parity"] else 1.0])[0] == 0.0
        parity["all_ranks"] = bool(ok)
        parity["sharded_parity"] = bool(parity["sharded_parity"] and ok)

    line = None
    if rank == 0:
        total_pages = pages * world
        peak, peak_src = measured_peaks()
        bytes_per_page = TOKENS * 128 * 2 + TOKENS * 4          # fp16 tokens + fp32 inv-norm side array
        achieved = pages * bytes_per_page / (kern_ms * 1e-3) / 1e9
        # CPU baseline on a bounded sample of the SAME corpus (pages read back from the device)
        n_cpu = args.cpu_sample_pages
        docs = [corpus.read_page("initial", p).astype(np.float32) for p in range(n_cpu)]
        # the timed CPU leg runs at N=1 only; at N>1 one pass still checks the GPU top-k against the CPU path
        cpu_pps, cpu_s, blas_threads, cpu_passes, cpu_kind = cpu_sample_pages_per_s(docs, queries, args.cpu_seconds if world == 1 else 0.0)
        fn, _ = cpu_search_fn()
        cpu_top = fn(queries[0], docs, TOP_K)
        line = {
            "metric": "exhaustive_maxsim_pages_per_s",
            "value": total_pages * args.steps / (dev_ms * 1e-3),
            "unit": "pages/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16",
            "dtype_detail": "fp16 tensor-core operands, fp32 accumulate; the fp32 query is carried as an fp16 hi/lo pair (fp32-exact)",
            "data": "synthetic",
            "config": workload_config(pages, world),
            "notes": {"l2": "input (>=26 GB per step) is far larger than the 126 MB L2; no flush needed", "corpus_generation_s": gen_s,
                      "exchange": ("C ABI (vrag_comm_init / one packed all-gather per scanning stage, one max-all-reduce per "
                                   "candidate stage); transport: " + ("NVLink peer memory (CUDA IPC windows; small messages as LL lines — flag in the data, no fence — "
                                   "with a single-query stage's exchange fused into its two top-k kernels; larger ones one kernel per "
                                   "collective: stores into every peer's window -> flag -> wait -> consume)" if corpus.comm_peer_memory() else
                                   "NCCL (resolved by the library at run time)")) if world > 1 else "single shard"},
            "hbm_gbs_algorithmic": total_pages * bytes_per_page * args.steps / (dev_ms * 1e-3) / 1e9,
            "e2e": {"value": total_pages * args.steps / e2e_s, "unit": "pages/s",
                    "h2d_bytes_per_step": Q_TOKENS * 128 * 4, "d2h_bytes_per_step": TOP_K * (4 + 8) + 36,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "SingleStageRetriever.search(q, top_k=10, strategy='multi_vector') on the "
                            + ("ShardedCorpusClient (collective, every rank)" if world > 1 else "GpuCorpusClient"),
                    "frac_of_value": (total_pages * args.steps / e2e_s) / (total_pages * args.steps / (dev_ms * 1e-3)),
                    "device_ms_per_step": e2e_dev_ms / args.steps,
                    "host_list_equals_device_list": bool(host_equals_device)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "vrag::maxsim_scan_kernel<QP=24 (MMA N=48), LARGE>", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBs_nominal": achieved / 8000.0,
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": pages * bytes_per_page,
                         "traffic": (args.ncu_traffic_ratio * pages * bytes_per_page) if args.ncu_traffic_ratio else None,
                         "traffic_source": "dram__bytes_read+write per launch / algorithmic bytes = 1.0019 in the ncu --set full capture of "
                                           "this same launch shape, 500k pages (profiles/r2_kernels_ncu_summary.md, column `large_500k`: "
                                           "134.134 GB read + 13.9 MB written vs 133.9 GB algorithmic); scaled by pages for other sizes"},
            "cpu_baseline": {"value": cpu_pps, "unit": "pages/s", "cores": blas_threads, "kind": cpu_kind,
                             "sample": f"first {n_cpu} pages of the same corpus read back from the device, {cpu_passes} queries one after "
                                       f"the other ({cpu_passes * n_cpu} page scorings, {cpu_s:.1f} s); "
                                       + ("benchmarks/quick_test.py::search_exhaustive of the unmodified reference (baseline/_ref)"
                                          if cpu_kind == "reference" else "oracle/maxsim_oracle.py::search_exhaustive")
                                       + f" (single process, numpy BLAS threads={blas_threads})"} if world == 1 else
                            {"value": None, "note": "timed at N=1 only", "kind": cpu_kind},
            "fp16_query_variant": {"flag": "VRAG_Q_FP16 (opt-in; default is the fp32-exact hi/lo query)", "kernel_ms": kern16_ms,
                                   "hbm_gbs": pages * bytes_per_page / (kern16_ms * 1e-3) / 1e9,
                                   "frac": pages * bytes_per_page / (kern16_ms * 1e-3) / 1e9 / peak},
            "two_stage": dict({"mode": "tokens_vs_standard_pooling", "prefetch_k": PREFETCH_K, "top_k": TOP_K,
                               "total_pages": total_pages, "pooled_rows_per_page": POOLED_ROWS,
                               "retriever_call": "TwoStageRetriever.search_server_side(q, top_k=10, prefetch_k=256, "
                                                 "stage1_mode='tokens_vs_standard_pooling') -> result dicts",
                               "note": "mean_pooling = colpali_row_mean_pooling of the 1024 visual tokens (32 rows/page), derived on the device"},
                              **ts),
            "pooling": {"kind": "colpali_row_mean_pooling (p2) over the 1024 visual tokens of each 1030-token page -> 32 rows, fp16 in / fp16 out",
                        "pages_per_s_per_gpu": pages / (pool_ms * 1e-3), "ms": pool_ms,
                        "hbm_gbs": pages * (VISUAL_TOKENS * 256 + POOLED_ROWS * 256) / (pool_ms * 1e-3) / 1e9},
            "last_top1": [float(last[0]["score"]), int(last[0]["id"])] if last else None,
        }
        # the first n_cpu pages live on rank 0: the GPU's candidate-restricted top-10 over them vs the CPU path
        gpu_top = corpus.score("initial", queries[0], candidate_ids=list(range(corpus.page_base, corpus.page_base + n_cpu)))
        order = np.lexsort((np.arange(n_cpu), -gpu_top))[:TOP_K]
        line["cpu_baseline"]["topk_matches_gpu"] = bool([int(i) for i in order] == [int(i) for i, _ in cpu_top])
        if parity is not None:
            line["sharded_parity"] = parity["sharded_parity"]
            line["sharded_parity_detail"] = parity
    if args.extras:
        try:
            if world == 1:
                extra = run_extras(corpus, args, measured_peaks()[0], kern_ms)
            else:
                extra = run_multi_extras(corpus, client, two, args, rank, world, barrier, max_over_ranks, queries, ts)
            if line is not None:
                line.update(extra)
        except Exception as e:  # extras must never cost the headline line
            if line is not None:
                line["extras_error"] = repr(e)
    if line is not None:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    corpus.close()
    return 0


def run_multi_extras(corpus, client, two, args, rank, world, barrier, max_over_ranks, queries, ts_weak):
    """N > 1, after the headline measurements (all ranks take part):
      two_stage_strong : BASELINE configs[1] as stated — 1M pages TOTAL (1M / N per GPU), tokens_vs_standard_pooling
                         256 -> 10, single query: p50 / p95 / QPS through the C ABI and the retriever class, the device
                         time of each collective, and the ideal stage-1 scan time of the shard for scale.
      pooling_cfg4     : BASELINE configs[4] — 1M ColPali-shaped pages over the N GPUs (no collective), all pooled
                         stores of the collection in one pass."""
    import torch

    from visual_rag_b200.embedding import pooling as GP

    out = {}
    peak = measured_peaks()[0]
    total = args.total_pages
    per = total // world
    pages_main = corpus.n_pages("initial")
    if per == pages_main:
        strong = dict(ts_weak)
    else:
        for nm in ("initial", "mean_pooling"):
            corpus.drop_store(nm)
        corpus.add_synthetic_store("initial", per, fixed_rows=TOKENS, seed=SEED + 21, row_seed_base=rank * per * TOKENS)
        corpus.pool_store("initial", [GP.with_token_window(GP.spec_adaptive_rows(32, 32, POOLED_ROWS), 0, VISUAL_TOKENS)], ["mean_pooling"])
        if hasattr(client, "sync_points"):
            client.sync_points()
        ts_stages = [("mean_pooling", False, PREFETCH_K), ("initial", False, TOP_K)]
        for i in range(5):
            corpus.search_multistage(ts_stages, queries[i % len(queries)])
        barrier()
        lat, comm = [], []
        t0 = time.perf_counter()
        for i in range(args.latency_queries):
            t1 = time.perf_counter()
            corpus.search_multistage(ts_stages, queries[i % len(queries)])
            lat.append(1e3 * (time.perf_counter() - t1))
            if i % 16 == 0:
                comm.append(corpus.comm_timing_us())
        wall = time.perf_counter() - t0
        dev_ms = corpus.last_timing_ms()[0]
        barrier()
        for i in range(5):
            two.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K, stage1_mode="tokens_vs_standard_pooling")
        barrier()
        rlat = []
        for i in range(args.latency_queries):
            t1 = time.perf_counter()
            two.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K, stage1_mode="tokens_vs_standard_pooling")
            rlat.append(1e3 * (time.perf_counter() - t1))
        barrier()
        wall, p50, p95, r50, r95, dev_ms = max_over_ranks([wall, np.percentile(lat, 50), np.percentile(lat, 95),
                                                          np.percentile(rlat, 50), np.percentile(rlat, 95), dev_ms])
        med = np.median(np.array(comm), axis=0)
        strong = {"qps": args.latency_queries / wall, "p50_ms": p50, "p95_ms": p95, "queries": args.latency_queries,
                  "retriever_p50_ms": r50, "retriever_p95_ms": r95, "device_ms_last_query": dev_ms,
                  "collective_us": {"stage1_allgather_topk256": float(med[0]), "stage2_allreduce_max_256": float(med[1])}}
    ideal_ms = per * POOLED_ROWS * (256 + 4) / (peak * 1e9) * 1e3 + PREFETCH_K / world * TOKENS * 260 / (peak * 1e9) * 1e3
    out["two_stage_strong"] = dict(strong, total_pages=per * world, pages_per_gpu=per, mode="tokens_vs_standard_pooling",
                                   prefetch_k=PREFETCH_K, top_k=TOP_K,
                                   ideal_scan_ms_at_measured_hbm_peak=ideal_ms,
                                   p50_over_ideal=strong["p50_ms"] / ideal_ms,
                                   workload="BASELINE configs[1]: 1M ColPali pages TOTAL, strong scaling")
    # ---- cfg4 over all GPUs
    for nm in ("initial", "mean_pooling"):
        if corpus.has_store(nm):
            corpus.drop_store(nm)
    npg = args.cfg4_total_pages // world
    corpus.add_synthetic_store("vis", npg, fixed_rows=1024, seed=SEED + 4, row_seed_base=rank * npg * 1024)
    specs = [GP.spec_adaptive_rows(32, 32, 32)] + [GP.derived_from(x, 0) for x in (
        GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True))]
    names = ["mean_pooling", "experimental_pooling", "experimental_pooling_gaussian", "experimental_pooling_triangular", "global_pooling"]
    for _ in range(3):
        barrier()
        ms = corpus.pool_store("vis", specs, names)
    ms = max_over_ranks([ms])[0]
    b = npg * (1024 * 256 + (32 + 34 + 32 + 32 + 1) * 256)
    out["pooling_cfg4"] = {"colpali": {"total_pages": npg * world, "pages_per_gpu": npg, "ms_max_over_ranks": ms,
                                       "pages_per_s": npg * world / (ms * 1e-3), "hbm_gbs_algorithmic_per_gpu": b / (ms * 1e-3) / 1e9,
                                       "frac_of_peak": b / (ms * 1e-3) / 1e9 / peak, "collective": "none (pages are independent)",
                                       "stores": "row-mean 32 + legacy k=3 (34) + gaussian (32) + triangular (32) + global (1), one pass"}}
    for nm in names + ["vis"]:
        corpus.drop_store(nm)
    # ---- cfg2 over all GPUs: BASELINE configs[2] (three-stage, 256 queries per call) on the page-sharded corpus. One
    # collective per stage for the WHOLE batch: stage 1 an all-gather of 256 x 1000 packed hits (4 MB per rank: the bulk
    # peer-memory kernel), stages 2 and 3 a max-all-reduce of the 256 x 1000 / 256 x 300 candidate scores.
    try:
        from visual_rag_b200.corpus import pack_queries

        rng = np.random.default_rng(SEED + 77 + rank)
        n = args.cfg2_pages // world
        h = rng.integers(16, 33, size=n)
        w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
        off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
        offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
        corpus.add_synthetic_store("initial", 0, page_offsets=off, seed=SEED + 1 + 10 * rank)
        corpus.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=SEED + 2 + 10 * rank)
        corpus.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=SEED + 3 + 10 * rank)
        qrng = np.random.default_rng(SEED + 78)     # the same queries on every rank
        nq = 256
        qs = pack_queries([qrng.standard_normal((int(qrng.integers(10, 31)), 128)).astype(np.float32) for _ in range(nq)])
        stages = [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)]
        for _ in range(2):
            corpus.search_multistage_batch(stages, qs, final_only=True)
        walls, devs, comms, begins = [], [], [], []
        for _ in range(5):
            barrier()
            t0 = time.perf_counter()
            fin = corpus.search_multistage_batch(stages, qs, final_only=True)
            walls.append(time.perf_counter() - t0)
            devs.append(corpus.last_timing_ms()[0])
            comms.append(corpus.comm_timing_us())
            begins.append(corpus.comm_offsets_us())
        # every rank must hold the same merged lists
        import torch.distributed as dist

        mine = torch.from_numpy(np.ascontiguousarray(fin[1])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(mine, ref))
        wall, dev = max_over_ranks([float(np.median(walls)), float(np.median(devs))])
        med = np.median(np.array(comms), axis=0)
        out["three_stage_batched_sharded"] = {
            "workload": f"cfg2 over {world} GPUs: {n * world} ColQwen2.5-shaped pages TOTAL ({n} per GPU), 256 queries per call, "
                        "stage1_k=1000, stage2_k=300, top_k=100; final lists + stage scores to the host",
            "batch_wall_ms": 1e3 * wall, "batch_device_ms": dev, "qps": nq / wall,
            "collective_us": {"stage1_allgather_256x1000_hits_4MB": float(med[0]), "stage2_allreduce_max_256x1000": float(med[1]),
                              "stage3_allreduce_max_256x300": float(med[2])},
            "lists_identical_on_all_ranks": same, "transport_peer_memory": bool(corpus.comm_peer_memory())}
        try:   # device timeline of a batch on rank 0 (us): local work before each collective, the collectives, the tail
            bg = np.median(np.array(begins), axis=0)
            tl = {"stage1_local_us": float(bg[0]), "stage1_exchange_us": float(med[0]),
                  "stage2_merge_and_local_us": float(bg[1] - bg[0] - med[0]), "stage2_exchange_us": float(med[1]),
                  "stage3_topk_and_local_us": float(bg[2] - bg[1] - med[1]), "stage3_exchange_us": float(med[2]),
                  "tail_topk_and_result_copy_us": float(1e3 * np.median(devs) - bg[2] - med[2]),
                  "note": "rank 0, median of 5 batches; exchange times include the wait for the slowest rank"}
            out["three_stage_batched_sharded"]["device_timeline"] = tl
        except Exception as e:
            out["three_stage_batched_sharded"]["device_timeline"] = {"error": repr(e)[:200]}
        for nm in ("initial", "experimental_pooling", "global_pooling"):
            corpus.drop_store(nm)
    except Exception as e:  # an extra must not take the headline line down with it
        out["three_stage_batched_sharded"] = {"error": repr(e)[:300]}
    torch.cuda.synchronize()
    return out


def run_extras(corpus, args, peak, kern_ms):
    """N=1 only, after the headline measurements: the other BASELINE configs on the same GPU.
      three_stage_batched : cfg2 — ColQwen2.5-shaped variable pages, 256 queries, global -> experimental -> MaxSim.
      exhaustive_batched  : 4 queries share every document tile (tensor-pipe-heavier variant of the headline scan).
      pooling_cfg4        : cfg4 — all pooled stores of a ColPali / ColSmol collection derived in one pass."""
    from visual_rag_b200.embedding import pooling as GP

    out = {}
    rng = np.random.default_rng(SEED + 99)
    # ---- batched exhaustive on the headline corpus
    q4 = [rng.standard_normal((Q_TOKENS, 128)).astype(np.float32) for _ in range(4)]
    for _ in range(2):
        corpus.search_multistage_batch([("initial", False, TOP_K)], q4)
    corpus.search_multistage_batch([("initial", False, TOP_K)], q4)
    ms = corpus.last_timing_ms()[0]
    pages = corpus.n_pages("initial")
    out["exhaustive_batched"] = {"queries_per_pass": 4, "ms_per_pass": ms, "page_scorings_per_s": 4 * pages / (ms * 1e-3),
                                 "speedup_vs_single_query": 4 * kern_ms / ms}
    corpus.search_multistage_batch([("initial", False, TOP_K)], q4, fp16_query=True)
    corpus.search_multistage_batch([("initial", False, TOP_K)], q4, fp16_query=True)
    ms16 = corpus.last_timing_ms()[0]
    out["exhaustive_batched"]["fp16_query_ms_per_pass"] = ms16     # opt-in VRAG_Q_FP16: half the tensor work (power-limited scan)
    out["exhaustive_batched"]["fp16_query_page_scorings_per_s"] = 4 * pages / (ms16 * 1e-3)
    # 8 and 32 queries per call: approximate first pass (8 plain-fp16 queries share every document tile) + exact re-score of
    # 128 candidates per query; the lists must equal the single-query (fp32-exact) searches
    flops_per_page_query = 2 * Q_TOKENS * TOKENS * 128
    for nqb in (8, 32):
        qb = [rng.standard_normal((Q_TOKENS, 128)).astype(np.float32) for _ in range(nqb)]
        for _ in range(2):
            res = corpus.search_multistage_batch([("initial", False, TOP_K)], qb)
        ms_b, kern_b = corpus.last_timing_ms()
        same = all(res[b][0][1].tolist() == corpus.search("initial", qb[b], TOP_K)[1].tolist() for b in range(0, nqb, 5))
        issued = (nqb + 7) // 8 * pages * 2 * 256 * TOKENS * 128            # MMA flops issued: N = 256 columns per tile row
        out["exhaustive_batched"][f"queries_{nqb}"] = {
            "ms_per_call": ms_b, "first_pass_kernel_ms": kern_b, "page_scorings_per_s": nqb * pages / (ms_b * 1e-3),
            "speedup_vs_single_query": nqb * kern_ms / ms_b, "lists_equal_single_query_path": bool(same),
            "useful_tflops": nqb * pages * flops_per_page_query / (ms_b * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "achieved": issued / (ms_b * 1e-3) / 1e12, "peak": measured_tensor_peak(),
                         "unit": "TFLOP/s", "frac": issued / (ms_b * 1e-3) / 1e12 / measured_tensor_peak(),
                         "note": "issued fp16 MMA flops (128 x 256 x 128 per tile and 8 queries) over bf16_tflops_sustained of "
                                 "MEASURED_PEAKS.json; useful flops are 20/32 of the issued ones (20-token queries in 32-column slots)"}}
    # ---- payload-filter bitmask inside the scan (8(f)-3): the headline scan restricted to the pages that pass
    out["filtered_scan"] = {}
    for frac in (0.5, 0.05):
        allowed = rng.random(pages) < frac
        fid = corpus.create_filter(allowed)
        for _ in range(2):
            corpus.search("initial", q4[0], TOP_K, filter_id=fid)
        ms_f = []
        for _ in range(5):
            s_f, i_f = corpus.search("initial", q4[0], TOP_K, filter_id=fid)
            ms_f.append(corpus.last_timing_ms()[1])
        corpus.destroy_filter(fid)
        ms_f = float(np.median(ms_f))
        out["filtered_scan"][f"selectivity_{frac}"] = {
            "pages_passing": int(allowed.sum()), "scan_kernel_ms": ms_f, "unfiltered_kernel_ms": kern_ms,
            "store_pages_per_s": pages / (ms_f * 1e-3), "rate_vs_unfiltered": kern_ms / ms_f,
            "all_results_pass": bool(allowed[i_f - corpus.page_base].all()),
            "note": "filtered-out pages are skipped inside the scan (no tile fetched or multiplied)"}
    for nm in ("initial", "mean_pooling"):
        corpus.drop_store(nm)
    # ---- cfg0: the reference's own CPU-runnable case, in full on both sides (ColSmol-shaped, exact top-10)
    out["cfg0_colsmol"] = run_cfg0(corpus, args, rng)
    out["per_call_twins"] = run_per_call_twins(rng)
    # ---- cfg2
    n = args.cfg2_pages
    h = rng.integers(16, 33, size=n)
    w = np.minimum(rng.integers(16, 33, size=n), 768 // h)
    off = np.concatenate([[0], np.cumsum(h * w)]).astype(np.int64)
    offp = np.concatenate([[0], np.cumsum(np.minimum(h, 32))]).astype(np.int64)
    corpus.add_synthetic_store("initial", 0, page_offsets=off, seed=SEED + 1)
    corpus.add_synthetic_store("experimental_pooling", 0, page_offsets=offp, seed=SEED + 2)
    corpus.add_synthetic_store("global_pooling", n, fixed_rows=1, seed=SEED + 3)
    nq = 256
    queries = [rng.standard_normal((int(rng.integers(10, 31)), 128)).astype(np.float32) for _ in range(nq)]
    stages = [("global_pooling", True, 1000), ("experimental_pooling", False, 300), ("initial", False, 100)]
    for _ in range(2):
        corpus.search_multistage_batch(stages, queries)
    walls = []
    for _ in range(5):
        t0 = time.perf_counter()
        arr = corpus.search_multistage_batch(stages, queries, as_arrays=True)
        walls.append(time.perf_counter() - t0)
    dev_ms = corpus.last_timing_ms()[0]
    from visual_rag_b200.corpus import pack_queries

    packed = pack_queries(queries)
    walls_packed = []
    for _ in range(5):
        t0 = time.perf_counter()
        arr = corpus.search_multistage_batch(stages, packed, as_arrays=True)
        walls_packed.append(time.perf_counter() - t0)
    walls_final = []
    for _ in range(5):
        t0 = time.perf_counter()
        fin = corpus.search_multistage_batch(stages, packed, final_only=True)
        walls_final.append(time.perf_counter() - t0)
    # the reference-facing call: ThreeStageRetriever.search_server_side_batch -> list of result dicts per query
    from visual_rag_b200.client import GpuCorpusClient
    from visual_rag_b200.retrieval import ThreeStageRetriever

    retr = ThreeStageRetriever(GpuCorpusClient(corpus, "bench"), "bench")
    retr.search_server_side_batch(query_embeddings=queries, top_k=100, stage1_k=1000, stage2_k=300)
    t0 = time.perf_counter()
    hits = retr.search_server_side_batch(query_embeddings=queries, top_k=100, stage1_k=1000, stage2_k=300)
    retr_wall = time.perf_counter() - t0
    res = corpus.search_multistage_batch(stages, queries)
    t0 = time.perf_counter()
    for q in queries[:32]:
        single = corpus.search_multistage(stages, q)
    seq_ms = 1e3 * (time.perf_counter() - t0) / 32
    same = bool(np.array_equal(single[2][1], res[31][2][1]))
    out["three_stage_batched"] = {
        "workload": f"cfg2: {n} ColQwen2.5-shaped pages (H,W in [16,32], T=H*W<=768, {int(off[-1])} tokens = {off[-1] * 256 / 1e9:.1f} GB), "
                    f"pooled rows min(H,32), global 1 row; {nq} queries with 10..30 tokens; stage1_k=1000, stage2_k=300, top_k=100",
        "batch_wall_ms": 1e3 * float(np.median(walls)), "batch_device_ms": dev_ms, "qps": nq / float(np.median(walls)),
        "batch_wall_ms_prepacked_queries": 1e3 * float(np.median(walls_packed)), "qps_prepacked_queries": nq / float(np.median(walls_packed)),
        "batch_wall_ms_final_only": 1e3 * float(np.median(walls_final)), "qps_final_only": nq / float(np.median(walls_final)),
        "retriever_batch_wall_ms": 1e3 * retr_wall, "retriever_batch_qps": nq / retr_wall,
        "retriever_results_match": bool([h["id"] for h in hits[31]] == [int(i) for i in res[31][2][1]] if True else False),
        "ms_per_query_batched": 1e3 * float(np.median(walls)) / nq, "ms_per_query_sequential_api": seq_ms,
        "last_query_matches_single_query_path": same}
    # ---- cfg4, ColQwen2.5: bulk re-pooling of the same collection from its stored `initial` vectors (8(f)-2): per-page patch
    # grids, adaptive row means capped at 32, gaussian + triangular smoothing and the global vector in ONE pass over the tokens
    from visual_rag_b200.embedding.repool import infer_grids, recompute_pooling_from_initial

    for nm in ("experimental_pooling", "global_pooling"):
        corpus.drop_store(nm)
    grids = np.stack([h, w], axis=1).astype(np.int32)
    try:   # 143 GB of tokens + 4 x 6 GB of pooled stores: if the device cannot hold them beside the scratch buffers, say so
        t1 = time.perf_counter()
        inferred = infer_grids(corpus.page_rows("initial"))          # what the re-pooling script does without payload sizes
        infer_s = time.perf_counter() - t1
        for _ in range(2):
            info = recompute_pooling_from_initial(corpus, grids=grids)
        pooled_rows = int(np.minimum(h, 32).sum())
        bq = int(off[-1]) * 256 + (4 * pooled_rows + n) * 256
        qwen = {"pages": n, "tokens": int(off[-1]), "ms": info["ms"], "pages_per_s": n / (info["ms"] * 1e-3),
                "hbm_gbs_algorithmic": bq / (info["ms"] * 1e-3) / 1e9, "frac_of_peak": bq / (info["ms"] * 1e-3) / 1e9 / peak,
                "grid_inference_host_s": infer_s, "grids_valid": bool((inferred[:, 0].astype(np.int64) * inferred[:, 1] == h * w).all()),
                "stores": "adaptive row means (<= 32) + experimental (gaussian) + gaussian + triangular + global, one pass; "
                          "recompute_pooling_from_initial (scripts/qdrant_recompute_colqwen_pooling_from_initial.py)"}
    except Exception as e:  # noqa: BLE001
        qwen = {"skipped": repr(e)[:300]}
    for nm in ("initial", "mean_pooling", "experimental_pooling", "experimental_pooling_gaussian", "experimental_pooling_triangular",
               "global_pooling"):
        try:
            corpus.drop_store(nm)
        except Exception:  # noqa: BLE001  (a store the re-pooling never got to create)
            pass
    # ---- cfg4
    npg = args.cfg4_pages
    corpus.add_synthetic_store("vis", npg, fixed_rows=1024, seed=SEED + 4)
    specs = [GP.spec_adaptive_rows(32, 32, 32)] + [GP.derived_from(x, 0) for x in (
        GP.spec_legacy_conv(3), GP.spec_smooth(3, "gaussian"), GP.spec_smooth(3, "triangular"), GP.spec_global_mean(True))]
    names = ["mean_pooling", "experimental_pooling", "experimental_pooling_gaussian", "experimental_pooling_triangular", "global_pooling"]
    for _ in range(3):
        ms = corpus.pool_store("vis", specs, names)
    b = npg * (1024 * 256 + (32 + 34 + 32 + 32 + 1) * 256)
    out["pooling_cfg4"] = {"colpali": {"pages": npg, "ms": ms, "pages_per_s": npg / (ms * 1e-3), "hbm_gbs_algorithmic": b / (ms * 1e-3) / 1e9,
                                       "frac_of_peak": b / (ms * 1e-3) / 1e9 / peak,
                                       "stores": "row-mean 32 + legacy k=3 (34) + gaussian (32) + triangular (32) + global (1), one pass"}}
    for nm in names + ["vis"]:
        corpus.drop_store(nm)
    corpus.add_synthetic_store("smol", npg, fixed_rows=832, seed=SEED + 5)
    g = np.tile(np.array([[4, 3]], dtype=np.int32), (npg, 1))
    specs = [GP.spec_tile_mean(64), GP.spec_colsmol_experimental(0, 64), GP.derived_from(GP.spec_tile_4n(0, 0), 0),
             GP.derived_from(GP.spec_global_mean(True), 0)]
    names = ["mean_pooling", "experimental_pooling", "experimental_pooling_2d", "global_pooling"]
    for _ in range(3):
        ms = corpus.pool_store("smol", specs, names, grid_hw=g)
    b = npg * (832 * 256 + (13 + 76 + 13 + 1) * 256)
    out["pooling_cfg4"]["colqwen"] = qwen
    out["pooling_cfg4"]["colsmol"] = {"pages": npg, "ms": ms, "pages_per_s": npg / (ms * 1e-3),
                                      "hbm_gbs_algorithmic": b / (ms * 1e-3) / 1e9, "frac_of_peak": b / (ms * 1e-3) / 1e9 / peak,
                                      "stores": "tile mean 13 + experimental 76 + 4-neighbour 13 + global 1, one pass"}
    return out


def run_per_call_twins(rng):
    """compute_maxsim_score / compute_maxsim_batch called per document (list) with HOST arrays, as the reference's client-side
    callers do (two_stage.py:398-426): this build (pipelined host cast + upload + one scan) against the reference's own
    numpy functions on the same inputs and host."""
    from oracle import maxsim_oracle as MO
    from visual_rag_b200.embedding import pooling as GP

    q = rng.standard_normal((Q_TOKENS, 128)).astype(np.float32)
    docs = [rng.standard_normal((768, 128)).astype(np.float32) for _ in range(256)]
    ref_fn_b, ref_fn_s, kind = MO.maxsim_batch, MO.maxsim_score, "port (oracle/maxsim_oracle.py)"
    if load_reference() is not None:
        from visual_rag.embedding import pooling as RP   # baseline/_ref, unmodified

        ref_fn_b, ref_fn_s, kind = RP.compute_maxsim_batch, RP.compute_maxsim_score, "reference (visual_rag/embedding/pooling.py, baseline/_ref)"

    def best(fn, n):
        ts = []
        for _ in range(n):
            t1 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t1)
        return float(np.median(ts)), r

    GP.compute_maxsim_batch(q, docs[:8])   # scratch handle, pinned staging, worker pool
    g_b, got = best(lambda: GP.compute_maxsim_batch(q, docs), 7)
    c_b, want = best(lambda: ref_fn_b(q, docs), 3)
    g_s, _ = best(lambda: GP.compute_maxsim_score(q, docs[0]), 200)
    c_s, _ = best(lambda: ref_fn_s(q, docs[0]), 200)
    rel = max(abs(a - b) / abs(b) for a, b in zip(got, want))
    return {"workload": "compute_maxsim_batch(q, 256 fp32 documents x 768 tokens = 100.7 MB of host memory) and compute_maxsim_score(q, one of them); 20-token query",
            "batch_ms": 1e3 * g_b, "batch_cpu_ms": 1e3 * c_b, "single_us": 1e6 * g_s, "single_cpu_us": 1e6 * c_s, "cpu_kind": kind,
            "max_rel_score_diff_vs_cpu": rel, "note": "the documents are true fp32 (not fp16-representable): the deviation is the fp16 store cast"}


def run_cfg0(corpus, args, rng):
    """BASELINE configs[0]: 10k pages x 768 tokens x 128-d fp16, 20-token queries, exact MaxSim top-10 — GPU through the
    host API against the oracle port of benchmarks/quick_test.py::search_exhaustive / search_two_stage on the SAME pages
    (read back from the device), a few queries on the CPU side (1.7 s each), 100 on the GPU side."""
    from oracle import maxsim_oracle as MO
    from visual_rag_b200.embedding import pooling as GP

    n0, t0 = args.cfg0_pages, 768
    corpus.add_synthetic_store("initial", n0, fixed_rows=t0, seed=SEED + 10)
    corpus.pool_store("initial", [GP.spec_tile_mean(64)], ["mean_pooling"])      # 12 tile means per page
    qs = [rng.standard_normal((Q_TOKENS, 128)).astype(np.float32) for _ in range(100)]
    ts = [("mean_pooling", True, 256), ("initial", False, TOP_K)]                 # pooled_query_vs_tiles -> MaxSim rerank
    for q in qs[:5]:
        corpus.search("initial", q, TOP_K)
        corpus.search_multistage(ts, q)
    lat_ex, lat_ts = [], []
    for q in qs:
        t1 = time.perf_counter()
        ex = corpus.search("initial", q, TOP_K)
        lat_ex.append(1e3 * (time.perf_counter() - t1))
    dev_ms = corpus.last_timing_ms()[1]
    for q in qs:
        t1 = time.perf_counter()
        corpus.search_multistage(ts, q)
        lat_ts.append(1e3 * (time.perf_counter() - t1))
    # CPU side on the same bits
    raw = corpus.read_rows("initial", 0, n0 * t0).astype(np.float32)
    docs = [raw[i * t0:(i + 1) * t0] for i in range(n0)]
    praw = corpus.read_rows("mean_pooling", 0, n0 * 12).astype(np.float32)
    pooled = [praw[i * 12:(i + 1) * 12] for i in range(n0)]
    n_cpu = max(1, args.cfg0_cpu_queries)
    same_ex = same_ts = True
    qt = load_reference()
    if qt is not None:   # the reference's own functions on its own document dictionary (quick_test.py:158-206)
        ddict = {i: {"embedding": docs[i], "pooled": pooled[i]} for i in range(n0)}
        t1 = time.perf_counter()
        cpu_ex = [[(r["id"], r["score"]) for r in qt.search_exhaustive(q, ddict, top_k=TOP_K)] for q in qs[:n_cpu]]
        cpu_ex_s = (time.perf_counter() - t1) / n_cpu
        t1 = time.perf_counter()
        cpu_ts = [[(r["id"], r["score"], r["stage1_rank"]) for r in qt.search_two_stage(q, ddict, prefetch_k=256, top_k=TOP_K)]
                  for q in qs[:n_cpu]]
        cpu_ts_s = (time.perf_counter() - t1) / n_cpu
        cpu_kind = "reference (benchmarks/quick_test.py::search_exhaustive / search_two_stage, unmodified, baseline/_ref), single process"
    else:
        t1 = time.perf_counter()
        cpu_ex = [MO.search_exhaustive(q, docs, TOP_K) for q in qs[:n_cpu]]
        cpu_ex_s = (time.perf_counter() - t1) / n_cpu
        t1 = time.perf_counter()
        cpu_ts = [MO.search_two_stage_pooled(q, docs, pooled, 256, TOP_K) for q in qs[:n_cpu]]
        cpu_ts_s = (time.perf_counter() - t1) / n_cpu
        cpu_kind = "port (oracle/maxsim_oracle.py::search_exhaustive / search_two_stage_pooled = benchmarks/quick_test.py:158-206), single process"
    max_rel = 0.0
    for q, ce, ct in zip(qs, cpu_ex, cpu_ts):
        s, ids = corpus.search("initial", q, TOP_K)
        same_ex &= [int(i) for i in ids] == [i for i, _ in ce]
        max_rel = max(max_rel, max(abs(float(a) - b) / abs(b) for a, (_, b) in zip(s, ce)))
        g = corpus.search_multistage(ts, q)
        same_ts &= [int(i) for i in g[1][1]] == [i for i, _, _ in ct]
    for nm in ("initial", "mean_pooling"):
        corpus.drop_store(nm)
    bytes_pp = t0 * 256 + t0 * 4
    return {
        "workload": f"cfg0: {n0} ColSmol-shaped pages x {t0} tok x 128-d fp16 ({n0 * t0 * 256 / 1e9:.2f} GB), {Q_TOKENS}-token queries, exact top-{TOP_K}",
        "gpu_exhaustive_p50_ms": float(np.percentile(lat_ex, 50)), "gpu_exhaustive_qps": 1e3 / float(np.mean(lat_ex)),
        "gpu_exhaustive_pages_per_s": n0 * 1e3 / float(np.mean(lat_ex)), "gpu_scan_kernel_ms": dev_ms,
        "gpu_scan_hbm_gbs": n0 * bytes_pp / (dev_ms * 1e-3) / 1e9,
        "gpu_two_stage_p50_ms": float(np.percentile(lat_ts, 50)), "gpu_two_stage_qps": 1e3 / float(np.mean(lat_ts)),
        "cpu_exhaustive_s_per_query": cpu_ex_s, "cpu_exhaustive_pages_per_s": n0 / cpu_ex_s,
        "cpu_two_stage_s_per_query": cpu_ts_s, "cpu_queries": n_cpu,
        "cpu_kind": cpu_kind,
        "top10_ids_identical_exhaustive": bool(same_ex), "top10_ids_identical_two_stage": bool(same_ts),
        "max_rel_score_diff": max_rel,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages-per-gpu", type=int, default=500_000)
    ap.add_argument("--cpu-sample-pages", type=int, default=3000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of the cpu_baseline leg")
    ap.add_argument("--extras", type=int, default=1, help="0: skip the cfg2 / cfg4 / batched extra sections")
    ap.add_argument("--ref-sample-pages", type=int, default=16384)
    ap.add_argument("--latency-queries", type=int, default=200)
    ap.add_argument("--cfg0-pages", type=int, default=10_000)
    ap.add_argument("--cfg0-cpu-queries", type=int, default=2)
    ap.add_argument("--cfg2-pages", type=int, default=1_000_000)
    ap.add_argument("--cfg4-pages", type=int, default=400_000)
    ap.add_argument("--total-pages", type=int, default=1_000_000, help="N>1: corpus size of the cfg1 strong-scaling two-stage run")
    ap.add_argument("--cfg4-total-pages", type=int, default=1_000_000, help="N>1: pages pooled over all GPUs (cfg4)")
    ap.add_argument("--ncu-traffic-ratio", type=float, default=1.0019,
                    help="DRAM bytes / algorithmic bytes of the scan kernel in the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
