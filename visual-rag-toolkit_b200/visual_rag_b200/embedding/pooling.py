"""Drop-in for visual_rag/embedding/pooling.py: same function names, positional/keyword arguments, defaults,
dtype rules and ValueErrors — every mean / weighted mean / MaxSim is computed by libvrag_b200 on the GPU.

Host-side logic kept here (as in the reference): input conversion, output-dtype inference, argument
validation and the tap weights of weighted_row_smoothing_same_length.  There is no numpy arithmetic path: the
functions raise if the CUDA library or a GPU is missing.

The model-aware dispatch of VisualEmbedder (visual_embedder.py:735-923) and the per-page orchestration of
ProcessingPipeline (pipeline.py:400-507) are mirrored by `mean_pool_visual_embedding`,
`experimental_pool_visual_embedding`, `global_pool_from_mean_pool` and `pool_page` (taking the model name
instead of an embedder object, since model inference is out of scope).
"""

from __future__ import annotations

import ctypes as C
import threading
from typing import Any, Dict, List, Literal, Optional, Sequence, Union

import numpy as np

from .. import _native as N

DEVICE = 0  # CUDA device used by the single-page functions


# ------------------------------------------------------------------------------------------------ helpers
def _infer_output_dtype(embedding, output_dtype=None):
    """pooling.py:19-32."""
    if output_dtype is not None:
        return output_dtype
    try:
        import torch

        if isinstance(embedding, torch.Tensor):
            return np.float16 if embedding.dtype == torch.float16 else np.float32
    except ImportError:  # pragma: no cover
        pass
    if isinstance(embedding, np.ndarray) and embedding.dtype == np.float16:
        return np.float16
    return np.float32


def _to_host_rows(x) -> np.ndarray:
    """The reference upcasts every input to fp32 first (e.g. pooling.py:68-74). fp16 inputs are handed to the
    kernel as fp16 (the upcast is exact and happens in registers); everything else becomes fp32."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            x = x.detach().cpu()
            x = x.numpy() if x.dtype in (torch.float16, torch.float32) else x.float().numpy()
    except ImportError:  # pragma: no cover
        pass
    a = np.asarray(x)
    if a.dtype not in (np.float16, np.float32):
        a = a.astype(np.float32)
    if a.ndim != 2:
        raise ValueError(f"expected a [rows, dim] array, got shape {a.shape}")
    return np.ascontiguousarray(a)


def _run(spec: N.PoolSpec, rows: np.ndarray, out_dtype) -> np.ndarray:
    """One vrag_pool_page call. dim != 128 is handled by zero-padding columns (means are column-wise)."""
    lib = N.load()
    n, dim = rows.shape
    if dim > 128:
        raise ValueError("embedding dim > 128 is not supported by the B200 kernels")
    if dim < 128:
        padded = np.zeros((n, 128), dtype=rows.dtype)
        padded[:, :dim] = rows
        rows = padded
    out_np = np.dtype(out_dtype)
    kernel_out = np.float16 if out_np == np.float16 else np.float32
    n_out = C.c_int64()
    N.check(lib.vrag_pool_out_rows(C.byref(spec), n, C.byref(n_out)))
    out = np.empty((n_out.value, 128), dtype=kernel_out)
    got = C.c_int64()
    N.check(
        lib.vrag_pool_page(
            DEVICE, C.byref(spec), rows.ctypes.data_as(C.c_void_p), N.VRAG_F16 if rows.dtype == np.float16 else N.VRAG_F32,
            n, out.ctypes.data_as(C.c_void_p), N.VRAG_F16 if kernel_out == np.float16 else N.VRAG_F32, out.shape[0],
            C.byref(got),
        )
    )
    out = out[:, :dim]
    return out if out.dtype == out_np else out.astype(out_np)


# ------------------------------------------------------------------------------------------------ spec builders
def spec_tile_mean(patches_per_tile: int = 64) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_TILE_MEAN, patches_per_tile=int(patches_per_tile))


def spec_adaptive_rows(grid_h: int = 0, grid_w: int = 0, target_rows: int = 0, clamp_to_h: bool = False) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_ADAPTIVE_ROWS, grid_h=int(grid_h), grid_w=int(grid_w), target_rows=int(target_rows),
                      clamp_to_h=int(bool(clamp_to_h)))


def spec_colsmol_experimental(num_tiles: int = 0, patches_per_tile: int = 64) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_COLSMOL_EXPERIMENTAL, num_tiles=int(num_tiles), patches_per_tile=int(patches_per_tile))


def spec_legacy_conv(window_size: int = 3) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_LEGACY_CONV, window=int(window_size))


def smoothing_weights(window_size: int, kernel: str, sigma: Optional[float] = None) -> np.ndarray:
    """Normalised fp32 taps — the weight construction of pooling.py:329-355 (host logic)."""
    k = int(window_size)
    center = (k - 1) / 2.0
    dist = np.abs(np.arange(k, dtype=np.float32) - center)
    if kernel == "uniform":
        w = np.ones((k,), dtype=np.float32)
    elif kernel == "triangular":
        w = np.clip((center + 1.0) - dist, 0.0, None).astype(np.float32)
    else:
        if sigma is None:
            sigma_eff = max(0.5, float(center) / 2.0)
        else:
            sigma_eff = float(sigma)
            if sigma_eff <= 0:
                raise ValueError("sigma must be > 0")
        w = np.exp(-0.5 * (dist / sigma_eff) ** 2).astype(np.float32)
    w_sum = float(w.sum())
    if w_sum <= 0:
        return np.zeros((0,), dtype=np.float32)
    return (w / w_sum).astype(np.float32)


def spec_smooth(window_size: int = 3, kernel: str = "gaussian", sigma: Optional[float] = None) -> N.PoolSpec:
    k = int(window_size)
    if k > 16:
        raise ValueError("window_size > 16 is not supported by the B200 kernels")
    s = N.PoolSpec(kind=N.POOL_SMOOTH, window=k)
    if k > 1:
        w = smoothing_weights(k, kernel, sigma)
        if w.size == 0:          # degenerate weights: the reference returns the rows unchanged (pooling.py:353-354)
            s.window = 1
        else:
            s.n_weights = k
            for i in range(k):
                s.weights[i] = float(w[i])
    return s


def spec_tile_4n(n_rows: int = 0, n_cols: int = 0, has_global: bool = True, include_self: bool = True) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_TILE_4N, n_rows=int(n_rows), n_cols=int(n_cols), has_global=int(bool(has_global)),
                      include_self=int(bool(include_self)))


def spec_global_mean(via_f16: bool = False) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_GLOBAL_MEAN, via_f16=int(bool(via_f16)))


def spec_seq_chunks(target_rows: int) -> N.PoolSpec:
    return N.PoolSpec(kind=N.POOL_SEQ_CHUNKS, target_rows=int(target_rows))


def with_token_window(spec: N.PoolSpec, skip: int, count: int) -> N.PoolSpec:
    """Bulk pooling only (GpuCorpus.pool_store): pool rows [skip, skip+count) of every source page — the page's visual
    tokens (pipeline.py:400-430 pools `visual_embedding = embedding[visual_token_indices]`) when the `initial` store
    also holds the instruction tokens (ColPali-v1.3: 1024 visual + 6 text tokens)."""
    spec.in_row_skip = int(skip)
    spec.in_row_count = int(count)
    return spec


def derived_from(spec: N.PoolSpec, index: int) -> N.PoolSpec:
    """Chain `spec` to the OUTPUT of spec number `index` of the same GpuCorpus.pool_store call (the pipeline's
    experimental / global pooling of the mean-pooled rows, pipeline.py:452-507): computed in that spec's pass."""
    spec.input_spec = int(index) + 1
    return spec


# ------------------------------------------------------------------------------------------------ p1 .. p8
def tile_level_mean_pooling(embedding, num_tiles: int, patches_per_tile: int = 64, output_dtype=None) -> np.ndarray:
    """pooling.py:35-98. `num_tiles` only matters when it matches T / patches_per_tile (otherwise it is
    overridden by ceil(T / patches_per_tile), lines 79-84) — i.e. the result always has ceil(T/ppt) rows."""
    out_dtype = _infer_output_dtype(embedding, output_dtype)
    rows = _to_host_rows(embedding)
    if rows.shape[0] == 0:
        return np.array([], dtype=out_dtype)
    return _run(spec_tile_mean(patches_per_tile), rows, out_dtype)


def colpali_row_mean_pooling(embedding, grid_size: int = 32, output_dtype=None) -> np.ndarray:
    """pooling.py:101-124."""
    out_dtype = _infer_output_dtype(embedding, output_dtype)
    rows = _to_host_rows(embedding)
    expected = int(grid_size) * int(grid_size)
    if rows.shape[0] != expected:
        raise ValueError(f"Expected {expected} visual tokens for grid_size={grid_size}, got {rows.shape[0]}")
    return _run(spec_adaptive_rows(grid_size, grid_size, grid_size), rows, out_dtype)


def adaptive_row_mean_pooling_from_grid(embedding, *, grid_h: int, grid_w: int, target_rows: int = 32,
                                        output_dtype=None) -> np.ndarray:
    """pooling.py:127-185."""
    out_dtype = _infer_output_dtype(embedding, output_dtype)
    rows = _to_host_rows(embedding)
    expected = int(grid_h) * int(grid_w)
    if rows.shape[0] != expected:
        raise ValueError(f"Expected {expected} visual tokens for grid_h×grid_w={grid_h}×{grid_w}, got {rows.shape[0]}")
    if int(target_rows) <= 0:
        raise ValueError("target_rows must be > 0")
    return _run(spec_adaptive_rows(grid_h, grid_w, target_rows), rows, out_dtype)


def colsmol_experimental_pooling(embedding, num_tiles: int, patches_per_tile: int = 64, output_dtype=None) -> np.ndarray:
    """pooling.py:188-232."""
    out_dtype = _infer_output_dtype(embedding, output_dtype)
    rows = _to_host_rows(embedding)
    if num_tiles <= 0:
        raise ValueError("num_tiles must be > 0")
    if patches_per_tile <= 0:
        raise ValueError("patches_per_tile must be > 0")
    if rows.shape[0] == 0:
        raise ValueError(
            f"Not enough tokens for num_tiles={num_tiles}, patches_per_tile={patches_per_tile}: got 0"
        )
    return _run(spec_colsmol_experimental(num_tiles, patches_per_tile), rows, out_dtype)


def colpali_experimental_pooling_from_rows(row_vectors, *, window_size: int = 3, output_dtype=None) -> np.ndarray:
    """pooling.py:235-286."""
    out_dtype = _infer_output_dtype(row_vectors, output_dtype)
    rows = _to_host_rows(row_vectors)
    if rows.shape[0] < 1:
        raise ValueError("row_vectors must be non-empty")
    window_size = int(window_size)
    if window_size < 1:
        raise ValueError("window_size must be >= 1")
    if window_size % 2 == 0:
        raise ValueError("window_size must be odd")
    return _run(spec_legacy_conv(window_size), rows, out_dtype)


def weighted_row_smoothing_same_length(row_vectors, *, window_size: int = 3,
                                       kernel: Literal["uniform", "triangular", "gaussian"] = "gaussian",
                                       sigma: Optional[float] = None, output_dtype=None) -> np.ndarray:
    """pooling.py:289-375."""
    out_dtype = _infer_output_dtype(row_vectors, output_dtype)
    rows = _to_host_rows(row_vectors)
    n = rows.shape[0]
    if n < 1:
        raise ValueError("row_vectors must be non-empty")
    k = int(window_size)
    if k < 1:
        raise ValueError("window_size must be >= 1")
    if k == 1 or n == 1:
        return _run(spec_smooth(1), rows, out_dtype)
    kernel = str(kernel).lower().strip()
    if kernel not in ("uniform", "triangular", "gaussian"):
        raise ValueError(f"Unknown kernel={kernel}. Choose uniform|triangular|gaussian.")
    return _run(spec_smooth(k, kernel, sigma), rows, out_dtype)


def colsmol_tile_4n_pooling_from_tiles(tile_vectors, *, n_rows: int, n_cols: int, has_global: bool = True,
                                       include_self: bool = True, output_dtype=None) -> np.ndarray:
    """pooling.py:378-436."""
    out_dtype = _infer_output_dtype(tile_vectors, output_dtype)
    rows = _to_host_rows(tile_vectors)
    n_rows, n_cols = int(n_rows), int(n_cols)
    if n_rows <= 0 or n_cols <= 0:
        raise ValueError("n_rows and n_cols must be > 0")
    grid_n = n_rows * n_cols
    if rows.shape[0] < grid_n:
        raise ValueError(
            f"Expected at least {grid_n} tile vectors for n_rows×n_cols={n_rows}×{n_cols}, got {rows.shape[0]}"
        )
    if not include_self and grid_n == 1:
        raise ValueError("need at least one array to stack")
    return _run(spec_tile_4n(n_rows, n_cols, has_global, include_self), rows, out_dtype)


def global_mean_pooling(embedding, output_dtype=None) -> np.ndarray:
    """pooling.py:439-465. fp16 numpy/torch input is NOT upcast by the reference, so numpy rounds the mean to
    fp16 before the final cast (line 463-465); `via_f16` reproduces that."""
    out_dtype = _infer_output_dtype(embedding, output_dtype)
    rows = _to_host_rows(embedding)
    via_f16 = rows.dtype == np.float16
    return _run(spec_global_mean(via_f16), rows, out_dtype)[0]


# ------------------------------------------------------------------------------------------------ a1 / a2
_scratch = {}             # device -> GpuCorpus kept for the per-call scoring functions below
_scratch_lock = threading.Lock()


def _scratch_corpus():
    """One long-lived corpus handle per device for compute_maxsim_score / compute_maxsim_batch: creating a handle
    (stream, events, pinned staging, device-property query) costs milliseconds, a per-pair score must not."""
    from ..corpus import GpuCorpus

    c = _scratch.get(DEVICE)
    if c is None:
        c = _scratch[DEVICE] = GpuCorpus(DEVICE)
    return c


def _page_scores(query_embedding, docs: Sequence[np.ndarray], normalize: bool) -> List[float]:
    mats = [_to_host_rows(d) for d in docs]
    q = np.asarray(query_embedding, dtype=np.float32)
    with _scratch_lock:
        return _scratch_corpus().score_pages(q, mats, normalize=normalize).tolist()


def compute_maxsim_score(query_embedding: np.ndarray, doc_embedding: np.ndarray, normalize: bool = True) -> float:
    """pooling.py:468-514 on the GPU. Note: documents are held in the fp16 store dtype; pass fp16-representable
    values (as every reference caller does after the Qdrant round trip) for bit-comparable results."""
    return _page_scores(query_embedding, [doc_embedding], normalize)[0]


def compute_maxsim_batch(query_embedding: np.ndarray, doc_embeddings: list, normalize: bool = True) -> list:
    """pooling.py:517-552: one upload + one scan for the whole list."""
    if len(doc_embeddings) == 0:
        return []
    return _page_scores(query_embedding, doc_embeddings, normalize)


# ------------------------------------------------------------------------------------------------ p9 / p10 / p11
def _model_flags(model_name: str):
    m = (model_name or "").lower()
    return "colsmol" in m, ("colqwen2.5" in m or "colqwen2_5" in m)


def mean_pool_visual_embedding(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
                               target_vectors: Optional[int] = 32, output_dtype=np.float32) -> np.ndarray:
    """VisualEmbedder.mean_pool_visual_embedding, visual_embedder.py:735-835 (output_dtype = embedder.output_dtype)."""
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
    rows = _to_host_rows(visual_embedding).astype(np.float32, copy=False)
    info = token_info or {}
    if is_colsmol:
        nr, nc = info.get("n_rows"), info.get("n_cols")
        num_tiles = int(nr) * int(nc) + 1 if nr and nc else 13
        return tile_level_mean_pooling(rows, num_tiles=num_tiles, patches_per_tile=64, output_dtype=output_dtype)
    t = int(rows.shape[0])
    if is_colqwen25:
        gh, gw = info.get("grid_h_eff"), info.get("grid_w_eff")
        if gh and gw and int(gh) * int(gw) == t:
            target = int(gh) if cap is None else min(int(cap), int(gh))
            return adaptive_row_mean_pooling_from_grid(rows, grid_h=int(gh), grid_w=int(gw), target_rows=target,
                                                       output_dtype=output_dtype)
    g = int(round(float(t) ** 0.5))
    if g * g == t:
        eff = int(g) if (is_colqwen25 and cap is None) else int(cap)
        if g == eff:
            return colpali_row_mean_pooling(rows, grid_size=eff, output_dtype=output_dtype)
        return adaptive_row_mean_pooling_from_grid(rows, grid_h=g, grid_w=g, target_rows=eff, output_dtype=output_dtype)
    return _run(spec_seq_chunks(int(cap or 32)), rows, output_dtype)


def global_pool_from_mean_pool(mean_pool: np.ndarray, output_dtype=np.float32) -> np.ndarray:
    """visual_embedder.py:837-840."""
    if mean_pool.size == 0:
        return np.zeros((128,), dtype=output_dtype)
    rows = _to_host_rows(mean_pool)
    return _run(spec_global_mean(rows.dtype == np.float16), rows, output_dtype)[0]


def experimental_pool_visual_embedding(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
                                       target_vectors: Optional[int] = 32, mean_pool: Optional[np.ndarray] = None,
                                       window_size: Optional[int] = None, kernel: Optional[str] = None,
                                       output_dtype=np.float32) -> np.ndarray:
    """VisualEmbedder.experimental_pool_visual_embedding, visual_embedder.py:842-923."""
    is_colsmol, is_colqwen25 = _model_flags(model_name)
    info = token_info or {}
    if is_colsmol:
        rows = _to_host_rows(visual_embedding).astype(np.float32, copy=False)
        if mean_pool is not None and getattr(mean_pool, "shape", None) is not None and int(mean_pool.shape[0]) > 0:
            num_tiles = int(mean_pool.shape[0])
        else:
            num_tiles = info.get("num_tiles")
            if num_tiles is None:
                nvt = info.get("num_visual_tokens")
                if nvt is None:
                    nvt = int(rows.shape[0])
                num_tiles = -(-int(nvt) // 64)
            num_tiles = int(num_tiles)
        return colsmol_experimental_pooling(rows, num_tiles=num_tiles, patches_per_tile=64, output_dtype=output_dtype)
    rows = mean_pool if mean_pool is not None else mean_pool_visual_embedding(
        model_name, visual_embedding, token_info, target_vectors=target_vectors, output_dtype=output_dtype)
    k = (kernel or ("gaussian" if is_colqwen25 else "legacy")).lower().strip()
    if k in ("legacy", "legacy_conv", "conv"):
        window = int(window_size) if window_size is not None else (5 if is_colqwen25 else 3)
        return colpali_experimental_pooling_from_rows(rows, window_size=window, output_dtype=output_dtype)
    window = int(window_size) if window_size is not None else 3
    kern = "gaussian" if k == "gaussian" else ("triangular" if k == "triangular" else "uniform")
    return weighted_row_smoothing_same_length(rows, window_size=window, kernel=kern, output_dtype=output_dtype)


def pool_page(model_name: str, visual_embedding, token_info: Optional[Dict[str, Any]] = None, *,
              max_mean_pool_vectors: Optional[int] = 32, pooling_windows: Optional[Sequence[int]] = None,
              experimental_pooling_kernel: str = "auto", colsmol_experimental_2d: bool = False,
              output_dtype=np.float32) -> Dict[str, np.ndarray]:
    """The pooling block of ProcessingPipeline._process_single_page (pipeline.py:400-507): named vectors of one
    page (mean_pooling, experimental_pooling[...], global_pooling), each rounded to output_dtype before the next
    step consumes it."""
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
    common = dict(target_vectors=tv, mean_pool=mean_pool, output_dtype=output_dtype)
    if is_colqwen25:
        g = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, window_size=3, kernel="gaussian", **common)
        t = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, window_size=3, kernel="triangular", **common)
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
            e = experimental_pool_visual_embedding(model_name, visual_embedding, token_info, window_size=k, kernel=kern, **common)
            out[f"experimental_pooling_{k}"] = e
            if k == ks[0]:
                out["experimental_pooling"] = e
    if is_colsmol and colsmol_experimental_2d:
        nr, nc = (token_info or {}).get("n_rows"), (token_info or {}).get("n_cols")
        if nr and nc:
            try:
                out["experimental_pooling_2d"] = colsmol_tile_4n_pooling_from_tiles(
                    mean_pool, n_rows=int(nr), n_cols=int(nc), has_global=True, include_self=True, output_dtype=output_dtype)
            except Exception:
                pass
    out["global_pooling"] = global_pool_from_mean_pool(mean_pool, output_dtype)
    return out
