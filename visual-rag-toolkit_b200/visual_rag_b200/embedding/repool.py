"""Bulk re-pooling of a whole collection from its stored `initial` vectors, on the device — the GPU form of
scripts/qdrant_recompute_colqwen_pooling_from_initial.py (reference): for every page infer the patch grid from the
token count and the payload's image size (`_infer_grid`, 64-105), adaptive row-mean pool with the cap, derive the
gaussian / triangular experimental stores and the global vector from the fp32 (unrounded) pooled rows (292-327),
and round everything to the fp16 store dtype on write. One pass over the tokens; nothing leaves HBM."""

from __future__ import annotations

import math
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from . import pooling as GP


def infer_grid(num_tokens: int, *, width: Optional[int] = None, height: Optional[int] = None) -> Tuple[int, int]:
    """scripts/qdrant_recompute_colqwen_pooling_from_initial.py:64-105 (host logic, same enumeration order)."""
    n = int(num_tokens)
    if n <= 0:
        raise ValueError("num_tokens must be > 0")
    if width and height and int(width) > 0 and int(height) > 0:
        aspect = float(width) / float(height)
    else:
        aspect = 1.0
    best = None
    best_score = float("inf")
    for h in range(1, int(math.isqrt(n)) + 1):
        if n % h != 0:
            continue
        w = n // h
        for hh, ww in ((h, w), (w, h)):
            score = abs(math.log(max(float(ww) / float(hh), 1e-9) / max(aspect, 1e-9)))
            if score < best_score:
                best_score = score
                best = (int(hh), int(ww))
    return best


def _payload_size(payload: Optional[Dict[str, Any]]):
    payload = payload or {}
    w = payload.get("resized_width") or payload.get("cropped_width") or payload.get("original_width")
    h = payload.get("resized_height") or payload.get("cropped_height") or payload.get("original_height")
    try:
        return (int(w) if w is not None else None), (int(h) if h is not None else None)
    except Exception:
        return None, None


def recompute_pooling_from_initial(corpus, payloads: Optional[Sequence[Optional[dict]]] = None, *,
                                   max_mean_pool_vectors: int = 32, src: str = "initial") -> Dict[str, float]:
    """Rebuild mean_pooling / experimental_pooling(_gaussian,_triangular) / global_pooling of every page of `corpus`
    from store `src`. payloads: per-page payload dicts (image sizes) or None. Returns {"ms": device time,
    "pages": n}. Raises ValueError for empty pages (the script skips points without vectors)."""
    n = corpus.n_pages(src)
    grids = np.empty((n, 2), dtype=np.int32)
    cache: Dict[Tuple[int, Optional[int], Optional[int]], Tuple[int, int]] = {}
    for p in range(n):
        _, t = corpus.page_range(src, p)
        w, h = _payload_size(payloads[p] if payloads is not None else None)
        key = (t, w, h)
        g = cache.get(key)
        if g is None:
            g = cache[key] = infer_grid(t, width=w, height=h)
        grids[p] = g
    cap = int(max_mean_pool_vectors)
    parent = GP.spec_adaptive_rows(0, 0, cap if cap > 0 else 0, clamp_to_h=True)
    parent.derive_from_f32 = 1
    specs = [parent, GP.derived_from(GP.spec_smooth(3, "gaussian"), 0), GP.derived_from(GP.spec_smooth(3, "gaussian"), 0),
             GP.derived_from(GP.spec_smooth(3, "triangular"), 0), GP.derived_from(GP.spec_global_mean(False), 0)]
    names = ["mean_pooling", "experimental_pooling", "experimental_pooling_gaussian", "experimental_pooling_triangular",
             "global_pooling"]
    ms = corpus.pool_store(src, specs, names, grid_hw=grids)
    return {"ms": ms, "pages": n}
