"""Shared helpers of the retriever mirrors."""

from __future__ import annotations

import contextlib
import gc
import time
from typing import Tuple

import numpy as np

# Backwards-compatible stage-1 vocabulary (two_stage.py:131-139, run_qdrant_beir.py:1471-1476).
STAGE1_ALIASES = {
    "pooled_query_vs_tiles": "pooled_query_vs_standard_pooling",
    "tokens_vs_tiles": "tokens_vs_standard_pooling",
    "pooled_query_vs_experimental": "pooled_query_vs_experimental_pooling",
    "tokens_vs_experimental": "tokens_vs_experimental_pooling",
}


def to_numpy(embedding) -> np.ndarray:
    """torch (any float dtype, any device) or array-like -> fp32 numpy (two_stage.py:428-434)."""
    try:
        import torch

        if isinstance(embedding, torch.Tensor):
            return embedding.detach().cpu().float().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.array(embedding, dtype=np.float32)


@contextlib.contextmanager
def no_gc():
    """Pause the cyclic garbage collector while a batch of result dictionaries is built: tens of thousands of new
    (acyclic) containers otherwise trigger full collections that walk every object of the process — measured 40 ms of
    a 57 ms batched three-stage call (256 queries x 100 results) against 3.3 ms on the device."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def wire(client, array: np.ndarray):
    """A query as it travels to the client: the reference serialises to nested lists for Qdrant's JSON/gRPC wire
    (two_stage.py:142-159); an in-process client that declares `accepts_numpy` gets the fp32 array itself (a
    20 x 128 `.tolist()` + re-parse costs ~0.1 ms per call, a tenth of a two-stage search on the GPU store)."""
    return array if getattr(client, "accepts_numpy", False) else array.tolist()


def resolve_stage1(stage1_mode: str, pooled_name: str, experimental_name: str, global_name: str) -> Tuple[bool, str]:
    """stage1_mode -> (pool the query?, named vector to scan) — two_stage.py:141-157. Both the current and
    the legacy vocabulary are accepted everywhere (the reference's client-side search() only knew the
    legacy names)."""
    mode = STAGE1_ALIASES.get(stage1_mode, stage1_mode)
    if mode == "pooled_query_vs_standard_pooling":
        return True, pooled_name
    if mode == "tokens_vs_standard_pooling":
        return False, pooled_name
    if mode == "pooled_query_vs_experimental_pooling":
        return True, experimental_name
    if mode == "tokens_vs_experimental_pooling":
        return False, experimental_name
    if mode == "pooled_query_vs_global":
        return True, global_name
    raise ValueError(f"Unknown stage1_mode: {stage1_mode}")


def retry_call(fn, max_retries: int, retry_sleep: float):
    """Exponential-backoff retry of two_stage.py:89-100 / three_stage.py:35-48."""
    last_err = None
    for attempt in range(max_retries):
        try:
            return fn()
        except (ValueError, TypeError, NotImplementedError, KeyError):
            raise       # deterministic: a bad argument does not get better with a back-off (the reference's loop guards
                        # network hiccups of a remote Qdrant)
        except Exception as e:  # noqa: BLE001 - mirror of the reference
            if type(e).__name__ == "VragError":   # an error reported by the library is deterministic too
                raise
            last_err = e
            if attempt >= max_retries - 1:
                break
            time.sleep(retry_sleep * (2**attempt))
    if last_err is not None:
        raise last_err
