"""Deterministic case definitions shared by make_golden.py (which runs the reference on them) and the
tests (which run the oracle / the CUDA path on the same inputs).  Inputs are regenerated from seeds; the
golden index stores a checksum of every input so a drifting RNG stream is detected, not silently accepted.
"""

from __future__ import annotations

import zlib

import numpy as np


def unit_rows(seed: int, n: int, d: int = 128, dtype=np.float32, scale: bool = False) -> np.ndarray:
    """Seeded gaussian rows, L2-normalised, rounded to fp16 precision (so fp32 and fp16 variants of a case
    hold the same values), optionally rescaled per row to exercise the normalisation path."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if scale:
        x *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    return x.astype(np.float16).astype(dtype)


def query_rows(seed: int, n: int, d: int = 128) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def checksum(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def pooling_case_specs():
    """[(fn name, n_rows, kwargs, positional-after-embedding)] — one entry per pooling call."""
    specs = []
    for t, nt in ((832, 13), (800, 13), (768, 12), (100, 5), (64, 1), (65, 13)):
        specs.append(("tile_level_mean_pooling", t, {"patches_per_tile": 64}, [nt]))
    specs.append(("tile_level_mean_pooling", 832, {"patches_per_tile": 64, "output_dtype": "float16"}, [13]))
    specs.append(("tile_level_mean_pooling", 96, {"patches_per_tile": 16}, [6]))
    for g in (32, 4):
        specs.append(("colpali_row_mean_pooling", g * g, {"grid_size": g}, []))
    for (h, w, r) in ((7, 3, 4), (7, 3, 9), (24, 20, 24), (40, 19, 32), (1, 5, 4), (32, 32, 32), (27, 28, 32),
                      (33, 23, 32), (16, 48, 16), (50, 10, 7)):
        specs.append(("adaptive_row_mean_pooling_from_grid", h * w, {"grid_h": h, "grid_w": w, "target_rows": r}, []))
    for t, nt in ((832, 13), (800, 13), (100, 13), (768, 12), (130, 3)):
        specs.append(("colsmol_experimental_pooling", t, {"patches_per_tile": 64}, [nt]))
    for n in (1, 2, 3, 4, 32):
        for k in (1, 3, 5):
            specs.append(("colpali_experimental_pooling_from_rows", n, {"window_size": k}, []))
    for n in (1, 2, 5, 32, 24):
        for k in (1, 2, 3, 4, 5):
            for kern in ("gaussian", "triangular", "uniform"):
                specs.append(("weighted_row_smoothing_same_length", n, {"window_size": k, "kernel": kern}, []))
    specs.append(("weighted_row_smoothing_same_length", 32, {"window_size": 5, "kernel": "gaussian", "sigma": 1.3}, []))
    for (nr, nc, extra, hg, inc) in ((4, 3, 1, True, True), (1, 1, 1, True, True), (2, 5, 0, False, True),
                                     (3, 4, 1, True, False), (4, 3, 2, True, True), (1, 6, 1, True, True)):
        specs.append(("colsmol_tile_4n_pooling_from_tiles", nr * nc + extra,
                      {"n_rows": nr, "n_cols": nc, "has_global": hg, "include_self": inc}, []))
    for t in (832, 1030, 7):
        specs.append(("global_mean_pooling", t, {}, []))
    return specs


def pooling_cases():
    """Expand specs over input dtypes: yields dict(key, fn, n, dtype, seed, kwargs, args)."""
    out = []
    for dt in ("float32", "float16"):
        for i, (fn, n, kwargs, args) in enumerate(pooling_case_specs()):
            out.append({"key": f"p_{dt[-2:]}_{i:03d}", "fn": fn, "n": n, "dtype": dt, "seed": 1000 + i,
                        "kwargs": kwargs, "args": args})
    return out


def fix_kwargs(kwargs):
    kw = dict(kwargs)
    if "output_dtype" in kw and isinstance(kw["output_dtype"], str):
        kw["output_dtype"] = np.dtype(kw["output_dtype"]).type
    return kw


DISPATCH_SPECS = [
    # model, tokens, token_info, cap (max_mean_pool_vectors), pooling_windows, kernel, colsmol 2d
    ("vidore/colSmol-500M", 832, {"n_rows": 4, "n_cols": 3}, 32, [3], "auto", True),
    ("vidore/colSmol-500M", 768, {"n_rows": 3, "n_cols": 4}, 32, [3], "auto", False),
    ("vidore/colSmol-500M", 832, {}, 32, None, "auto", False),
    ("vidore/colpali-v1.3", 1024, {}, 32, [3, 5], "auto", False),
    ("vidore/colpali-v1.3", 1024, {}, 32, [3], "gaussian", False),
    ("vidore/colpali-v1.3", 1024, {}, 16, [3], "triangular", False),
    ("vidore/colpali-v1.3", 1000, {}, 32, [3], "auto", False),
    ("vidore/colqwen2.5-v0.2", 24 * 28, {"grid_h_eff": 24, "grid_w_eff": 28}, 32, None, "auto", False),
    ("vidore/colqwen2.5-v0.2", 40 * 19, {"grid_h_eff": 40, "grid_w_eff": 19}, 32, None, "auto", False),
    ("vidore/colqwen2.5-v0.2", 40 * 19, {"grid_h_eff": 40, "grid_w_eff": 19}, 0, None, "auto", False),
    ("vidore/colqwen2.5-v0.2", 625, {}, None, None, "auto", False),
    ("vidore/colqwen2.5-v0.2", 700, {}, 32, None, "auto", False),
]


def dispatch_cases():
    out = []
    for dt in ("float32", "float16"):
        for i, (model, t, info, cap, windows, kern, twod) in enumerate(DISPATCH_SPECS):
            out.append({"key": f"d_{dt[-2:]}_{i:02d}", "model": model, "n": t, "token_info": info, "cap": cap,
                        "windows": windows, "kernel": kern, "twod": twod, "out_dtype": dt, "seed": 5000 + i})
    return out


MAXSIM_SPECS = [(20, 768), (20, 1030), (1, 32), (5, 100), (33, 13), (10, 1), (100, 300)]


def maxsim_cases():
    return [{"key": f"m_{i:02d}", "q": q, "t": t, "seed": 7000 + i} for i, (q, t) in enumerate(MAXSIM_SPECS)]


def bench_corpus(seed: int = 8000, n_docs: int = 120, tokens: int = 256):
    """In-memory corpus for the quick_test searches: docs fp32 (fp16-representable), 64-token tiles."""
    docs = [unit_rows(seed + 1 + i, tokens, scale=True) for i in range(n_docs)]
    q = query_rows(seed, 20)
    return q, docs


def retrieval_corpus(seed: int = 9000, n: int = 150):
    """Variable-length pages for the retriever-class goldens."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(40, 200, size=n)
    initial = [unit_rows(seed + 1 + i, int(lens[i]), scale=True) for i in range(n)]
    q = query_rows(seed + 100000, 20)
    return q, initial


# ------------------------------------------------------------------ bulk re-pooling
def infer_grid_cases():
    """(num_tokens, width, height) triples for _infer_grid: primes, squares, ColQwen-like grids, missing sizes."""
    out = []
    for n in (1, 2, 7, 12, 64, 97, 256, 300, 391, 512, 640, 729, 736, 748, 750, 768, 1024):
        for w, h in ((None, None), (1240, 1754), (1754, 1240), (800, 800), (3000, 500), (0, 100), (1, 5000)):
            out.append((n, w, h))
    return out


def repool_cases():
    rng = np.random.default_rng(4242)
    out = []
    for i in range(14):
        gh, gw = int(rng.integers(8, 40)), int(rng.integers(6, 33))
        while gh * gw > 1100:
            gh -= 1
        w, h = (gw * 28 + int(rng.integers(0, 20)), gh * 28 + int(rng.integers(0, 20))) if i % 4 else (None, None)
        out.append({"key": f"repool{i}", "seed": 9000 + i, "n": gh * gw, "w": w, "h": h, "cap": 32 if i % 3 else 0})
    return out


def saliency_cases():
    return [{"key": f"sal{i}", "seed": 9100 + i, "qseed": 9200 + i, "n": n, "q": q, "token_info": ti}
            for i, (n, q, ti) in enumerate([(832, 20, {"n_rows": 4, "n_cols": 3}), (1030, 13, None), (300, 1, None),
                                            (768, 32, {"n_rows": 3, "n_cols": 4}), (70, 25, {"n_rows": 1, "n_cols": 1})])]


# ------------------------------------------------------------------ true-fp32 inputs (NOT fp16-representable)
def raw_rows(seed: int, n: int, d: int = 128, scale: bool = True) -> np.ndarray:
    """Seeded gaussian rows, L2-normalised in fp32 and optionally rescaled — full fp32 mantissas, i.e. values that do
    NOT survive an fp16 round trip (what an embedder emits in fp32; the reference's compute_maxsim_score callers,
    two_stage.py:398-400, see such arrays). Pins (a) the pooling arithmetic on inputs whose partial sums are inexact
    in fp32 and (b) the deviation the fp16 store dtype introduces against the reference's fp32 arithmetic."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if scale:
        x *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    return x


def fp32_pooling_cases():
    """Every pooling call of pooling_case_specs() once more, on true-fp32 inputs."""
    return [{"key": f"p_raw_{i:03d}", "fn": fn, "n": n, "seed": 11000 + i, "kwargs": kwargs, "args": args}
            for i, (fn, n, kwargs, args) in enumerate(pooling_case_specs())]


FP32_MAXSIM_SPECS = [(20, 768), (20, 1030), (1, 32), (13, 100), (33, 13), (64, 300), (20, 2048)]


def fp32_maxsim_cases():
    return [{"key": f"m_raw_{i:02d}", "q": q, "t": t, "seed": 12000 + i} for i, (q, t) in enumerate(FP32_MAXSIM_SPECS)]


def fp32_corpus(seed: int = 13000, n_docs: int = 200):
    """Variable-length true-fp32 pages + query for the exhaustive / two-stage searches on fp32 inputs."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(60, 400, size=n_docs)
    docs = [raw_rows(seed + 1 + i, int(lens[i])) for i in range(n_docs)]
    return query_rows(seed + 5000, 20), docs
