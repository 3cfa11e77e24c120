"""Saliency scores for returned pages — the numeric part of visual_rag/visualization/saliency.py
(generate_saliency_map, lines 53-107). The image overlay (PIL / matplotlib, create_saliency_overlay) is presentation
and stays out of this backend; what it consumes is produced here:

  patch_scores       [T]            max over query tokens of cos(q_i, d_t)   — computed by the CUDA saliency kernel
  patch_scores_norm  [T]            min-max normalised to [0, 1] (zeros if the range is < 1e-8)
  tile_scores        [n_rows,n_cols] mean of 64-patch tiles when token_info carries the ColSmol tile grid, else None
"""

from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np


def saliency_scores(corpus, query_embedding, page_id: int, token_info: Optional[Dict[str, Any]] = None,
                    vector_name: str = "initial") -> Dict[str, Any]:
    patch_scores = corpus.saliency(vector_name, query_embedding, page_id)
    score_min, score_max = patch_scores.min(), patch_scores.max()
    if score_max - score_min > 1e-8:                                   # saliency.py:81-86
        norm = (patch_scores - score_min) / (score_max - score_min)
    else:
        norm = np.zeros_like(patch_scores)
    tile_scores = None
    if token_info and token_info.get("n_rows") and token_info.get("n_cols"):   # saliency.py:88-110
        n_rows, n_cols = int(token_info["n_rows"]), int(token_info["n_cols"])
        ppt = 64
        g = n_rows * n_cols
        grid = norm[: g * ppt]
        if grid.shape[0] == g * ppt:
            tile_scores = grid.reshape(g, ppt).mean(axis=1).reshape(n_rows, n_cols)
    return {"patch_scores": patch_scores, "patch_scores_norm": norm, "tile_scores": tile_scores}
