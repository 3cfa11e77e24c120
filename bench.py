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
  e2e    : same metric through the public host API (numpy query in, top-10 (score,id) out) — pinned H2D of the
           query and D2H of the result inside the timed region, wall-clock, max over ranks.
  extra  : two-stage (tokens_vs_standard_pooling, prefetch_k=256, top-10) QPS / p50 / p95 through the same API.
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
def cpu_sample_pages_per_s(docs_f32, queries, min_seconds=10.0, max_passes=200):
    """The oracle's search_exhaustive (= the reference's client-side CPU path, quick_test.py:158-166) timed on
    a bounded sample: one query after the other over the sampled pages until `min_seconds` of CPU work are done.
    Returns (pages/s, seconds, blas_threads, passes)."""
    from oracle import maxsim_oracle as MO

    t0 = time.perf_counter()
    passes = 0
    while True:
        MO.search_exhaustive(queries[passes % len(queries)], docs_f32, TOP_K)
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
    return passes * len(docs_f32) / dt, dt, threads, passes


def host_sample(n_pages, seed):
    """Seeded ColPali-shaped pages on the host (fp16-representable fp32), for the CPU arms."""
    rng = np.random.default_rng(seed)
    docs = []
    for _ in range(n_pages):
        x = rng.standard_normal((TOKENS, 128), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        docs.append(x.astype(np.float16).astype(np.float32))
    return docs


def _ref_worker(args):
    n_pages, seed, qseed, reps = args
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from oracle import maxsim_oracle as MO

    docs = host_sample(n_pages, seed)
    q = np.random.default_rng(qseed).standard_normal((Q_TOKENS, 128)).astype(np.float32)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        MO.search_exhaustive(q, docs, TOP_K)
        times.append(time.perf_counter() - t0)
    return times


def workload_config(pages, world):
    """`config` of the JSON line — the same dict on both arms (the reference arm adds what it sampled)."""
    total_pages = pages * world
    return {
        "workload": f"exhaustive MaxSim top-{TOP_K} over {total_pages} ColPali-shaped pages "
                    f"({pages}/GPU x {TOKENS} tok x 128-d fp16 = {pages * TOKENS * 256 / 1e9:.1f} GB/GPU; "
                    f"BASELINE configs[3] shard), {Q_TOKENS}-token query, NCCL all-gather top-k merge",
        "pages_per_gpu": pages, "tokens_per_page": TOKENS, "query_tokens": Q_TOKENS, "top_k": TOP_K,
    }


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port of quick_test.search_exhaustive ->
    compute_maxsim_score) on all host cores: the page sample is split over one process per core, each
    running the reference's per-page Python loop; a step = one query over the whole sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

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
    line = {
        "impl": "reference", "metric": "exhaustive_maxsim_pages_per_s", "value": value, "unit": "pages/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.pages_per_gpu, max(1, args.gpus)),
                       reference_sample=f"each step scores a bounded sample of {total_pages} pages of this workload on the host cores; "
                                        "pages/s is the rate over the sample",
                       pages_per_step=total_pages),
        "cpu_baseline": {"value": value, "unit": "pages/s", "cores": workers, "kind": "port",
                         "sample": f"{total_pages} pages/step split over {workers} processes (1 BLAS thread each), "
                                   "oracle/maxsim_oracle.py::search_exhaustive"},
        "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
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
    from visual_rag_b200.corpus import GpuCorpus
    from visual_rag_b200.distributed import ShardedSearcher

    pages = args.pages_per_gpu
    corpus = GpuCorpus(local_rank, page_base=rank * pages)
    t_gen0 = time.perf_counter()
    corpus.add_synthetic_store("initial", pages, fixed_rows=TOKENS, seed=SEED, row_seed_base=rank * pages * TOKENS)
    gen_s = time.perf_counter() - t_gen0
    # mean_pooling is DERIVED on the device with the pooling kernels: 1030 tokens is not a square grid, so the
    # reference takes its sequence-chunk path (visual_embedder.py:824-835) -> 32 rows per page.
    from visual_rag_b200.embedding import pooling as GP

    pool_ms = corpus.pool_store("initial", [GP.spec_seq_chunks(POOLED_ROWS)], ["mean_pooling"])
    searcher = ShardedSearcher(corpus)
    rng = np.random.default_rng(SEED + 7)
    queries = [rng.standard_normal((Q_TOKENS, 128)).astype(np.float32) for _ in range(max(8, args.steps + args.warmup))]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ex_stage = [("initial", False, TOP_K)]
    ts_stages = [("mean_pooling", False, PREFETCH_K), ("initial", False, TOP_K)]

    # ---------------- correctness spot check against the oracle (rank 0, a few pages read back) ----------------
    if rank == 0:
        from oracle import maxsim_oracle as MO

        sc = corpus.score("initial", queries[0], candidate_ids=[corpus.page_base + 3, corpus.page_base + pages - 1])
        for s, p in zip(sc, (3, pages - 1)):
            want = MO.maxsim_score(queries[0], corpus.read_page("initial", p).astype(np.float32))
            assert abs(s - want) <= 1e-3 * abs(want), ("parity spot check failed", s, want)

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
        searcher.search_multistage_device(ex_stage, nq)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = corpus.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- dominant kernel alone (roofline numerator), same stream, CUDA events ----------------
    scores_buf = torch.empty((pages,), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    from visual_rag_b200.corpus import query_flags

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

    # ---------------- end to end through the host API ----------------
    for i in range(args.warmup):
        searcher.search("initial", queries[i % len(queries)], TOP_K)
    barrier()
    t0 = time.perf_counter()
    last = None
    for i in range(args.steps):
        last = searcher.search("initial", queries[i % len(queries)], TOP_K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---------------- two-stage latency (extra) ----------------
    n_lat = args.latency_queries
    for i in range(5):
        searcher.search_multistage(ts_stages, queries[i % len(queries)])
    barrier()
    lat = []
    t_all0 = time.perf_counter()
    for i in range(n_lat):
        t1 = time.perf_counter()
        searcher.search_multistage(ts_stages, queries[i % len(queries)])
        lat.append(1e3 * (time.perf_counter() - t1))
    ts_wall = time.perf_counter() - t_all0
    barrier()

    # the same two-stage search through the reference-facing class (TwoStageRetriever.search_server_side on the
    # GpuCorpusClient: result dicts with payloads), single shard only — the sharded path has no per-shard client
    retr_lat = None
    if world == 1:
        from visual_rag_b200.client import GpuCorpusClient
        from visual_rag_b200.retrieval import TwoStageRetriever

        retr = TwoStageRetriever(GpuCorpusClient(corpus, "bench"), "bench")
        for i in range(5):
            retr.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K, stage1_mode="tokens_vs_standard_pooling")
        retr_lat = []
        for i in range(n_lat):
            t1 = time.perf_counter()
            hits = retr.search_server_side(queries[i % len(queries)], top_k=TOP_K, prefetch_k=PREFETCH_K,
                                           stage1_mode="tokens_vs_standard_pooling")
            retr_lat.append(1e3 * (time.perf_counter() - t1))
        assert len(hits) == TOP_K

    # ---------------- max over ranks ----------------
    vals = torch.tensor([dev_ms, e2e_s, kern_ms, ts_wall, float(np.percentile(lat, 50)), float(np.percentile(lat, 95))],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, kern_ms, ts_wall, p50, p95 = [float(v) for v in vals.cpu()]

    if rank == 0:
        total_pages = pages * world
        peak, peak_src = measured_peaks()
        bytes_per_page = TOKENS * 128 * 2 + TOKENS * 4          # fp16 tokens + fp32 inv-norm side array
        achieved = pages * bytes_per_page / (kern_ms * 1e-3) / 1e9
        # CPU baseline on a bounded sample of the SAME corpus (pages read back from the device)
        n_cpu = args.cpu_sample_pages
        docs = [corpus.read_page("initial", p).astype(np.float32) for p in range(n_cpu)]
        # the timed CPU leg runs at N=1 only; at N>1 one pass still checks the GPU top-k against the oracle
        cpu_pps, cpu_s, blas_threads, cpu_passes = cpu_sample_pages_per_s(docs, queries, args.cpu_seconds if world == 1 else 0.0)
        gpu_top = corpus.search("initial", queries[0], TOP_K, candidate_ids=list(range(corpus.page_base, corpus.page_base + n_cpu)))
        from oracle import maxsim_oracle as MO

        cpu_top = MO.search_exhaustive(queries[0], docs, TOP_K)
        parity_ok = [int(i) for i in gpu_top[1]] == [corpus.page_base + i for i, _ in cpu_top]
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
            "config": dict(workload_config(pages, world),
                           l2="input (>=26 GB per step) is far larger than the 126 MB L2; no flush needed",
                           corpus_generation_s=gen_s),
            "hbm_gbs_algorithmic": total_pages * bytes_per_page * args.steps / (dev_ms * 1e-3) / 1e9,
            "e2e": {"value": total_pages * args.steps / e2e_s, "unit": "pages/s",
                    "h2d_bytes_per_step": Q_TOKENS * 128 * 4, "d2h_bytes_per_step": TOP_K * (4 + 8),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "vrag::maxsim_scan_kernel<QP=24 (MMA N=48), LARGE>", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBs_nominal": achieved / 8000.0,
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": pages * bytes_per_page,
                         "traffic": (args.ncu_traffic_ratio * pages * bytes_per_page) if args.ncu_traffic_ratio else None,
                         "traffic_source": "dram__bytes_read+write per launch / algorithmic bytes = 1.0104 in the ncu --set full capture of "
                                           "this same launch shape, 500k pages (profiles/r1e_kernels_ncu_summary.md, column `large_500k`: "
                                           "135.285 GB read + 9.7 MB written vs 133.9 GB algorithmic); scaled by pages for other sizes"},
            "cpu_baseline": {"value": cpu_pps, "unit": "pages/s", "cores": blas_threads, "kind": "port",
                             "sample": f"first {n_cpu} pages of the same corpus read back from the device, {cpu_passes} queries one after "
                                       f"the other ({cpu_passes * n_cpu} page scorings, {cpu_s:.1f} s); oracle/maxsim_oracle.py::search_exhaustive "
                                       f"(single process, numpy BLAS threads={blas_threads})",
                             "topk_matches_gpu": bool(parity_ok)} if world == 1 else
                            {"value": None, "note": "timed at N=1 only", "topk_matches_gpu": bool(parity_ok)},
            "fp16_query_variant": {"flag": "VRAG_Q_FP16 (opt-in; default is the fp32-exact hi/lo query)", "kernel_ms": kern16_ms,
                                   "hbm_gbs": pages * bytes_per_page / (kern16_ms * 1e-3) / 1e9,
                                   "frac": pages * bytes_per_page / (kern16_ms * 1e-3) / 1e9 / peak},
            "two_stage": {"mode": "tokens_vs_standard_pooling", "prefetch_k": PREFETCH_K, "top_k": TOP_K,
                          "qps": n_lat / ts_wall, "p50_ms": p50, "p95_ms": p95, "queries": n_lat,
                          "pooled_rows_per_page": POOLED_ROWS,
                          "retriever_p50_ms": float(np.percentile(retr_lat, 50)) if retr_lat else None,
                          "retriever_p95_ms": float(np.percentile(retr_lat, 95)) if retr_lat else None,
                          "retriever_call": "TwoStageRetriever.search_server_side(q, top_k=10, prefetch_k=256, "
                                            "stage1_mode='tokens_vs_standard_pooling') -> result dicts (N=1 only)",
                          "note": "mean_pooling derived on the device from `initial` (sequence-chunk mean pooling, 32 rows/page)"},
            "pooling": {"kind": "seq_chunks 1030 -> 32 rows/page (visual_embedder.py:824-835), fp16 in / fp16 out",
                        "pages_per_s_per_gpu": pages / (pool_ms * 1e-3), "ms": pool_ms,
                        "hbm_gbs": pages * (TOKENS * 256 + POOLED_ROWS * 256) / (pool_ms * 1e-3) / 1e9},
            "last_top1": [float(last[0][0]), int(last[1][0])] if last is not None and len(last[0]) else None,
        }
        if world == 1 and args.extras:
            try:
                line.update(run_extras(corpus, args, peak, kern_ms))
            except Exception as e:  # extras must never cost the headline line
                line["extras_error"] = repr(e)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    corpus.close()
    return 0


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
    for nm in ("initial", "mean_pooling"):
        corpus.drop_store(nm)
    # ---- cfg0: the reference's own CPU-runnable case, in full on both sides (ColSmol-shaped, exact top-10)
    out["cfg0_colsmol"] = run_cfg0(corpus, args, rng)
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
    for nm in ("initial", "experimental_pooling", "global_pooling"):
        corpus.drop_store(nm)
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
    out["pooling_cfg4"]["colsmol"] = {"pages": npg, "ms": ms, "pages_per_s": npg / (ms * 1e-3),
                                      "hbm_gbs_algorithmic": b / (ms * 1e-3) / 1e9, "frac_of_peak": b / (ms * 1e-3) / 1e9 / peak,
                                      "stores": "tile mean 13 + experimental 76 + 4-neighbour 13 + global 1, one pass"}
    return out


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
    t1 = time.perf_counter()
    cpu_ex = [MO.search_exhaustive(q, docs, TOP_K) for q in qs[:n_cpu]]
    cpu_ex_s = (time.perf_counter() - t1) / n_cpu
    t1 = time.perf_counter()
    cpu_ts = [MO.search_two_stage_pooled(q, docs, pooled, 256, TOP_K) for q in qs[:n_cpu]]
    cpu_ts_s = (time.perf_counter() - t1) / n_cpu
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
        "cpu_kind": "port (oracle/maxsim_oracle.py::search_exhaustive / search_two_stage_pooled = benchmarks/quick_test.py:158-206), single process",
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
    ap.add_argument("--ncu-traffic-ratio", type=float, default=1.0104,
                    help="DRAM bytes / algorithmic bytes of the scan kernel in the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
