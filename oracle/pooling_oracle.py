"""CPU oracle for the indexing-time pooling path — TEST INFRASTRUCTURE ONLY.

numpy restatement of visual_rag/embedding/pooling.py (p1-p8), the model-aware dispatch in
visual_rag/embedding/visual_embedder.py:735-923 (p9-p10) and the per-page orchestration of
visual_rag/indexing/pipeline.py:400-507 (p11).  Pinned against the reference's own outputs by
tests/test_oracle_golden.py (tests/golden/pooling_golden.npz) and against the known answers in the
reference's tests/test_pooling.py.  Never imported by the product package.
"""

from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------------ helpers
def _as_f32(x) -> np.ndarray:
    """Input conversion every pooling function applies first (e.g. pooling.py:68-74): torch bf16 ->
    float -> numpy; everything else -> np.float32 copy."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return x.detach().cpu().float().numpy().astype(np.float32)
    except ImportError:  # pragma: no cover
        pass
    return np.array(x, dtype=np.float32)


def infer_output_dtype(x, output_dtype=None):
    """pooling.py:19-32: explicit dtype wins; fp16 in -> fp16 out; anything else -> fp32."""
    if output_dtype is not None:
        return output_dtype
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return np.float16 if x.dtype == torch.float16 else np.float32
    except ImportError:  # pragma: no cover
        pass
    if isinstance(x, np.ndarray) and x.dtype == np.float16:
        return np.float16
    return np.float32


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


# ------------------------------------------------------------------------------------------------ p1
def tile_level_mean_pooling(embedding, num_tiles: int, patches_per_tile: int = 64, output_dtype=None) -> np.ndarray:
    """pooling.py:35-98. The tile count that comes out is always ceil(T / patches_per_tile): a matching
    `num_tiles` is a no-op and a mismatching one is overridden (lines 79-84); the last tile may be partial."""
    out_dtype = infer_output_dtype(embedding, output_dtype)
    emb = _as_f32(embedding)
    t = emb.shape[0]
    ppt = int(patches_per_tile)
    if t != int(num_tiles) * ppt:
        num_tiles = _ceil_div(t, ppt)
    means = []
    for tile in range(int(num_tiles)):
        lo = tile * ppt
        if lo >= t:
            break
        means.append(emb[lo : min(lo + ppt, t)].mean(axis=0))
    return np.array(means, dtype=out_dtype)


# ------------------------------------------------------------------------------------------------ p2
def colpali_row_mean_pooling(embedding, grid_size: int = 32, output_dtype=None) -> np.ndarray:
    """pooling.py:101-124: [g*g, D] -> [g, D], mean over the columns of each grid row."""
    out_dtype = infer_output_dtype(embedding, output_dtype)
    emb = _as_f32(embedding)
    g = int(grid_size)
    if emb.shape[0] != g * g:
        raise ValueError(f"Expected {g * g} visual tokens for grid_size={grid_size}, got {emb.shape[0]}")
    return emb.reshape(g, g, emb.shape[1]).mean(axis=1).astype(out_dtype)


# ------------------------------------------------------------------------------------------------ p3
def adaptive_bins(h: int, target_rows: int) -> List[tuple]:
    """Row bins of adaptive_row_mean_pooling_from_grid, pooling.py:176-182: edges = linspace(0, h, R+1),
    bin i = [floor(e_i), ceil(e_{i+1})) clamped so that it is non-empty and inside [0, h). Bins overlap
    when h / R is fractional."""
    edges = np.linspace(0, h, target_rows + 1)
    bins = []
    for i in range(target_rows):
        lo = int(np.floor(edges[i]))
        hi = int(np.ceil(edges[i + 1]))
        lo = max(0, min(lo, h - 1))
        hi = max(lo + 1, min(hi, h))
        bins.append((lo, hi))
    return bins


def adaptive_row_mean_pooling_from_grid(embedding, *, grid_h: int, grid_w: int, target_rows: int = 32,
                                        output_dtype=None) -> np.ndarray:
    """pooling.py:127-185."""
    out_dtype = infer_output_dtype(embedding, output_dtype)
    emb = _as_f32(embedding)
    gh, gw = int(grid_h), int(grid_w)
    if emb.shape[0] != gh * gw:
        raise ValueError(f"Expected {gh * gw} visual tokens for grid_h×grid_w={grid_h}×{grid_w}, got {emb.shape[0]}")
    rows = emb.reshape(gh, gw, emb.shape[1]).mean(axis=1)
    r = int(target_rows)
    if r <= 0:
        raise ValueError("target_rows must be > 0")
    if gh == r:
        return rows.astype(out_dtype)
    if gh == 1:
        return np.repeat(rows, repeats=r, axis=0).astype(out_dtype)
    out = np.zeros((r, emb.shape[1]), dtype=np.float32)
    for i, (lo, hi) in enumerate(adaptive_bins(gh, r)):
        out[i] = rows[lo:hi].mean(axis=0)
    return out.astype(out_dtype)


# ------------------------------------------------------------------------------------------------ p4
def colsmol_experimental_pooling(embedding, num_tiles: int, patches_per_tile: int = 64, output_dtype=None) -> np.ndarray:
    """pooling.py:188-232: means of the first num_tiles-1 tiles followed by the RAW rows of the last tile.
    num_tiles is only overridden (to ceil(T/ppt)) when the last tile would start past the end (209-219)."""
    out_dtype = infer_output_dtype(embedding, output_dtype)
    emb = _as_f32(embedding)
    t, dim = emb.shape
    nt, ppt = int(num_tiles), int(patches_per_tile)
    if nt <= 0:
        raise ValueError("num_tiles must be > 0")
    if ppt <= 0:
        raise ValueError("patches_per_tile must be > 0")
    last = (nt - 1) * ppt
    if last >= t:
        nt = _ceil_div(t, ppt)
        if nt <= 0:
            raise ValueError(
                f"Not enough tokens for num_tiles={num_tiles}, patches_per_tile={patches_per_tile}: got {t}"
            )
        last = (nt - 1) * ppt
    head = emb[:last]
    tail = emb[last : min(last + ppt, t)]
    if head.size:
        head_means = head.reshape(-1, ppt, dim).mean(axis=1)
    else:
        head_means = np.zeros((0, dim), dtype=out_dtype)
    return np.concatenate([head_means.astype(out_dtype), tail.astype(out_dtype)], axis=0)


# ------------------------------------------------------------------------------------------------ p5
def colpali_experimental_pooling_from_rows(row_vectors, *, window_size: int = 3, output_dtype=None) -> np.ndarray:
    """pooling.py:235-286 ("legacy conv"): N rows -> N + 2r rows, r = window//2; output i is the uniform mean of
    rows[max(0, i-2r) : min(N-1, i)+1]."""
    out_dtype = infer_output_dtype(row_vectors, output_dtype)
    rows = _as_f32(row_vectors)
    n, dim = rows.shape
    if n < 1:
        raise ValueError("row_vectors must be non-empty")
    k = int(window_size)
    if k < 1:
        raise ValueError("window_size must be >= 1")
    if k % 2 == 0:
        raise ValueError("window_size must be odd")
    if k == 1 or n == 1:
        return rows.astype(out_dtype)
    r = k // 2
    if k == 3 and n == 2:  # pooling.py:277-279
        return np.stack([rows[0], rows.mean(axis=0), rows[1]], axis=0).astype(out_dtype)
    out = np.zeros((n + 2 * r, dim), dtype=np.float32)
    for i in range(n + 2 * r):
        lo = max(0, i - 2 * r)
        hi = min(n - 1, i)
        out[i] = rows[lo : hi + 1].mean(axis=0)
    return out.astype(out_dtype)


# ------------------------------------------------------------------------------------------------ p6
def smoothing_weights(window_size: int, kernel: str, sigma: Optional[float] = None) -> np.ndarray:
    """Normalised fp32 tap weights of weighted_row_smoothing_same_length, pooling.py:329-355."""
    k = int(window_size)
    center = (k - 1) / 2.0
    dist = np.abs(np.arange(k, dtype=np.float32) - center)
    if kernel == "uniform":
        w = np.ones((k,), dtype=np.float32)
    elif kernel == "triangular":
        w = np.clip((center + 1.0) - dist, 0.0, None).astype(np.float32)
    else:
        if sigma is None:
            s = max(0.5, float(center) / 2.0)
        else:
            s = float(sigma)
            if s <= 0:
                raise ValueError("sigma must be > 0")
        w = np.exp(-0.5 * (dist / s) ** 2).astype(np.float32)
    return w, float(w.sum())


def weighted_row_smoothing_same_length(row_vectors, *, window_size: int = 3, kernel: str = "gaussian",
                                       sigma: Optional[float] = None, output_dtype=None) -> np.ndarray:
    """pooling.py:289-375: N -> N weighted window, taps at i - k//2 + t, renormalised by the in-range
    weight mass at the borders."""
    out_dtype = infer_output_dtype(row_vectors, output_dtype)
    rows = _as_f32(row_vectors)
    n, dim = rows.shape
    if n < 1:
        raise ValueError("row_vectors must be non-empty")
    k = int(window_size)
    if k < 1:
        raise ValueError("window_size must be >= 1")
    if k == 1 or n == 1:
        return rows.astype(out_dtype)
    kernel = str(kernel).lower().strip()
    if kernel not in ("uniform", "triangular", "gaussian"):
        raise ValueError(f"Unknown kernel={kernel}. Choose uniform|triangular|gaussian.")
    w, w_sum = smoothing_weights(k, kernel, sigma)
    if w_sum <= 0:
        return rows.astype(out_dtype)
    w = w / w_sum
    left = k // 2
    acc = np.zeros((n, dim), dtype=np.float32)
    mass = np.zeros((n,), dtype=np.float64)
    idx = np.arange(n)
    for t in range(k):  # same tap order as the reference's inner loop, so fp32 sums round identically
        src = idx - left + t
        ok = (src >= 0) & (src < n)
        wt = float(w[t])
        acc[ok] += np.float32(wt) * rows[src[ok]]
        mass[ok] += wt
    out = rows.copy()
    pos = mass > 0
    out[pos] = acc[pos] / mass[pos].astype(np.float32)[:, None]
    return out.astype(out_dtype)


# ------------------------------------------------------------------------------------------------ p7
def colsmol_tile_4n_pooling_from_tiles(tile_vectors, *, n_rows: int, n_cols: int, has_global: bool = True,
                                       include_self: bool = True, output_dtype=None) -> np.ndarray:
    """pooling.py:378-436: every grid tile -> unweighted mean of itself and its existing 4-neighbours
    (order self, up, down, left, right); the trailing global tile is copied through."""
    out_dtype = infer_output_dtype(tile_vectors, output_dtype)
    tiles = _as_f32(tile_vectors)
    nr, nc = int(n_rows), int(n_cols)
    if nr <= 0 or nc <= 0:
        raise ValueError("n_rows and n_cols must be > 0")
    g = nr * nc
    if tiles.shape[0] < g:
        raise ValueError(
            f"Expected at least {g} tile vectors for n_rows×n_cols={n_rows}×{n_cols}, got {tiles.shape[0]}"
        )
    grid = tiles[:g].reshape(nr, nc, -1)
    if not include_self and nr == 1 and nc == 1:
        raise ValueError("need at least one array to stack")
    acc = grid.copy() if include_self else np.zeros_like(grid)
    cnt = np.full((nr, nc, 1), 1 if include_self else 0, dtype=np.int64)
    acc[1:] += grid[:-1]
    cnt[1:] += 1
    acc[:-1] += grid[1:]
    cnt[:-1] += 1
    acc[:, 1:] += grid[:, :-1]
    cnt[:, 1:] += 1
    acc[:, :-1] += grid[:, 1:]
    cnt[:, :-1] += 1
    out = [(acc / cnt.astype(np.float32)).reshape(g, -1)]
    if has_global and tiles.shape[0] > g:
        out.append(tiles[g : g + 1])
    return np.concatenate(out, axis=0).astype(out_dtype)


# ------------------------------------------------------------------------------------------------ p8
def global_mean_pooling(embedding, output_dtype=None) -> np.ndarray:
    """pooling.py:439-465. NOTE: no fp32 upcast of fp16 numpy input before the mean (line 463)."""
    out_dtype = infer_output_dtype(embedding, output_dtype)
    try:
        import torch

        if isinstance(embedding, torch.Tensor):
            emb = embedding.cpu().float().numpy() if embedding.dtype == torch.bfloat16 else embedding.cpu().numpy()
            return emb.mean(axis=0).astype(out_dtype)
    except ImportError:  # pragma: no cover
        pass
    return np.array(embedding).mean(axis=0).astype(out_dtype)


def global_pool_from_mean_pool(mean_pool: np.ndarray, output_dtype=np.float32) -> np.ndarray:
    """VisualEmbedder.global_pool_from_mean_pool, visual_embedder.py:837-840."""
    if mean_pool.size == 0:
        return np.zeros((128,), dtype=output_dtype)
    return mean_pool.mean(axis=0).astype(output_dtype)


# ------------------------------------------------------------------------------------------------ p9
def _model_flags(model_name: str):
    m = (model_name or "").lower()
    return "colsmol" in m, ("colqwen2.5" in m or "colqwen2_5" in m)


def sequence_chunk_mean_pooling(visual: np.ndarray, target: int, output_dtype=np.float32) -> np.ndarray:
    """Last-resort pooling of the token SEQUENCE into `target` overlapping chunks, visual_embedder.py:824-835."""
    n = int(visual.shape[0])
    out = np.zeros((target, int(visual.shape[1])), dtype=np.float32)
    for i, (lo, hi) in enumerate(adaptive_bins(n, target)):
        out[i] = visual[lo:hi].mean(axis=0)
    return out.astype(output_dtype)


def mean_pool_visual_embedding(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
                               target_vectors: Optional[int] = 32, output_dtype=np.float32) -> np.ndarray:
    """VisualEmbedder.mean_pool_visual_embedding, visual_embedder.py:735-835."""
    is_colsmol, is_colqwen25 = _model_flags(model_name)
    if target_vectors is None:
        cap = None
    else:
        try:
            tv = int(target_vectors)
        except Exception:
            tv = 32
        cap = None if tv <= 0 else tv
    if not is_colqwen25 and cap is None:
        cap = 32
    visual = _as_f32(visual_embedding)
    info = token_info or {}
    if is_colsmol:
        nr, nc = info.get("n_rows"), info.get("n_cols")
        num_tiles = int(nr) * int(nc) + 1 if nr and nc else 13
        return tile_level_mean_pooling(visual, num_tiles=num_tiles, patches_per_tile=64, output_dtype=output_dtype)
    t = int(visual.shape[0])
    if is_colqwen25:
        gh, gw = info.get("grid_h_eff"), info.get("grid_w_eff")
        if gh and gw and int(gh) * int(gw) == t:
            rows = int(gh) if cap is None else min(int(cap), int(gh))
            return adaptive_row_mean_pooling_from_grid(visual, grid_h=int(gh), grid_w=int(gw), target_rows=rows,
                                                       output_dtype=output_dtype)
    g = int(round(float(t) ** 0.5))
    if g * g == t:
        eff = int(g) if (is_colqwen25 and cap is None) else int(cap)
        if g == eff:
            return colpali_row_mean_pooling(visual, grid_size=eff, output_dtype=output_dtype)
        return adaptive_row_mean_pooling_from_grid(visual, grid_h=g, grid_w=g, target_rows=eff, output_dtype=output_dtype)
    return sequence_chunk_mean_pooling(visual, int(cap or 32), output_dtype)


# ------------------------------------------------------------------------------------------------ p10
def experimental_pool_visual_embedding(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
                                       target_vectors: Optional[int] = 32, mean_pool: Optional[np.ndarray] = None,
                                       window_size: Optional[int] = None, kernel: Optional[str] = None,
                                       output_dtype=np.float32) -> np.ndarray:
    """VisualEmbedder.experimental_pool_visual_embedding, visual_embedder.py:842-923."""
    is_colsmol, is_colqwen25 = _model_flags(model_name)
    visual = _as_f32(visual_embedding)
    info = token_info or {}
    if is_colsmol:
        if mean_pool is not None and getattr(mean_pool, "shape", None) is not None and int(mean_pool.shape[0]) > 0:
            num_tiles = int(mean_pool.shape[0])
        else:
            num_tiles = info.get("num_tiles")
            if num_tiles is None:
                nvt = info.get("num_visual_tokens")
                if nvt is None:
                    nvt = int(visual.shape[0])
                num_tiles = _ceil_div(int(nvt), 64)
            num_tiles = int(num_tiles)
        return colsmol_experimental_pooling(visual, num_tiles=num_tiles, patches_per_tile=64, output_dtype=output_dtype)
    rows = mean_pool if mean_pool is not None else mean_pool_visual_embedding(
        model_name, visual, token_info, target_vectors=target_vectors, output_dtype=output_dtype)
    k = (kernel or ("gaussian" if is_colqwen25 else "legacy")).lower().strip()
    if k in ("legacy", "legacy_conv", "conv"):
        window = int(window_size) if window_size is not None else (5 if is_colqwen25 else 3)
        return colpali_experimental_pooling_from_rows(rows, window_size=window, output_dtype=output_dtype)
    window = int(window_size) if window_size is not None else 3
    kern = "gaussian" if k == "gaussian" else ("triangular" if k == "triangular" else "uniform")
    return weighted_row_smoothing_same_length(rows, window_size=window, kernel=kern, output_dtype=output_dtype)


# ------------------------------------------------------------------------------------------------ p11
def pool_page(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
              max_mean_pool_vectors: Optional[int] = 32, pooling_windows: Optional[Sequence[int]] = None,
              experimental_pooling_kernel: str = "auto", colsmol_experimental_2d: bool = False,
              output_dtype=np.float32) -> Dict[str, np.ndarray]:
    """The pooling part of ProcessingPipeline._process_single_page, pipeline.py:400-507: returns the named
    vectors of one page: mean_pooling, experimental_pooling[...], global_pooling."""
    is_colsmol, is_colqwen25 = _model_flags(model_name)
    tv = max_mean_pool_vectors
    if tv is not None:
        try:
            tv_i = int(tv)
            tv = None if tv_i <= 0 else tv_i
        except Exception:
            tv = 32
    mean_pool = mean_pool_visual_embedding(model_name, visual_embedding, token_info, target_vectors=tv,
                                           output_dtype=output_dtype)
    out: Dict[str, np.ndarray] = {"mean_pooling": mean_pool}
    if is_colqwen25:
        g = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, target_vectors=tv,
                                               mean_pool=mean_pool, window_size=3, kernel="gaussian",
                                               output_dtype=output_dtype)
        t = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, target_vectors=tv,
                                               mean_pool=mean_pool, window_size=3, kernel="triangular",
                                               output_dtype=output_dtype)
        out["experimental_pooling"] = g
        out["experimental_pooling_gaussian"] = g
        out["experimental_pooling_triangular"] = t
    else:
        karg = str(experimental_pooling_kernel or "auto").lower().strip()
        kern = "legacy" if karg == "auto" else karg
        ks: List[int] = []
        for k in (pooling_windows if pooling_windows else [3]):
            try:
                ki = int(k)
            except Exception:
                continue
            if ki > 0 and ki not in ks:
                ks.append(ki)
        if not ks:
            ks = [3]
        for k in ks:
            e = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, target_vectors=tv,
                                                   mean_pool=mean_pool, window_size=k, kernel=kern,
                                                   output_dtype=output_dtype)
            out[f"experimental_pooling_{k}"] = e
            if k == ks[0]:
                out["experimental_pooling"] = e
    if is_colsmol and colsmol_experimental_2d:
        nr, nc = (token_info or {}).get("n_rows"), (token_info or {}).get("n_cols")
        if nr and nc:
            try:
                out["experimental_pooling_2d"] = colsmol_tile_4n_pooling_from_tiles(
                    mean_pool, n_rows=int(nr), n_cols=int(nc), has_global=True, include_self=True,
                    output_dtype=output_dtype)
            except Exception:
                pass
    out["global_pooling"] = global_pool_from_mean_pool(mean_pool, output_dtype)
    return out


# ------------------------------------------------------------------------------------------------
# Bulk re-pooling from the stored `initial` vectors.
def infer_grid(num_tokens: int, *, width: Optional[int] = None, height: Optional[int] = None):
    """(grid_h, grid_w) with grid_h*grid_w == num_tokens whose aspect ratio w/h is closest (log distance) to
    width/height — scripts/qdrant_recompute_colqwen_pooling_from_initial.py:64-105. Ties keep the first pair in
    the enumeration order h = 1..isqrt(n), orientation (h, w) before (w, h)."""
    import math

    n = int(num_tokens)
    if n <= 0:
        raise ValueError("num_tokens must be > 0")
    aspect = float(width) / float(height) if (width and height and int(width) > 0 and int(height) > 0) else 1.0
    best, best_score = None, float("inf")
    for h in range(1, int(math.isqrt(n)) + 1):
        if n % h:
            continue
        w = n // h
        for hh, ww in ((h, w), (w, h)):
            score = abs(math.log(max(float(ww) / float(hh), 1e-9) / max(aspect, 1e-9)))
            if score < best_score:
                best_score, best = score, (int(hh), int(ww))
    return best


def payload_size(payload: Optional[Dict[str, Any]]):
    """Image size the script reads from a point's payload (lines 274-290): resized, else cropped, else original."""
    payload = payload or {}
    w = payload.get("resized_width") or payload.get("cropped_width") or payload.get("original_width")
    h = payload.get("resized_height") or payload.get("cropped_height") or payload.get("original_height")
    try:
        return (int(w) if w is not None else None), (int(h) if h is not None else None)
    except Exception:
        return None, None


def repool_from_initial(emb: np.ndarray, payload: Optional[Dict[str, Any]] = None, max_mean_pool_vectors: int = 32):
    """Per-point arithmetic of the script (lines 292-327), everything in fp32: adaptive row-mean pooling on the
    inferred grid (cap min(max_mean_pool_vectors, grid_h)), gaussian / triangular k=3 smoothing of the UNROUNDED
    fp32 rows, global = mean of those rows. Returns {name: fp32 array}; `experimental_pooling` is the gaussian."""
    emb = np.asarray(emb, dtype=np.float32)
    w, h = payload_size(payload)
    gh, gw = infer_grid(emb.shape[0], width=w, height=h)
    cap = int(max_mean_pool_vectors)
    mp = adaptive_row_mean_pooling_from_grid(emb, grid_h=gh, grid_w=gw, target_rows=(gh if cap <= 0 else min(cap, gh)),
                                             output_dtype=np.float32)
    g = weighted_row_smoothing_same_length(mp, window_size=3, kernel="gaussian", output_dtype=np.float32)
    t = weighted_row_smoothing_same_length(mp, window_size=3, kernel="triangular", output_dtype=np.float32)
    return {"mean_pooling": mp, "global_pooling": mp.mean(axis=0).astype(np.float32), "experimental_pooling": g,
            "experimental_pooling_gaussian": g, "experimental_pooling_triangular": t}
