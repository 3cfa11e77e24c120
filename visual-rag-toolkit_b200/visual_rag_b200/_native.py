"""ctypes binding of libvrag_b200.so (C ABI: include/vrag_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 GPU is present when
a compute entry point is called, the call raises.  Build the library with `python __graft_entry__.py`
(or `make -C visual-rag-toolkit_b200/csrc`).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_LIB_NAME = "libvrag_b200.so"
# VRAG_LIB: A/B experiments with another build of the same library (developer knob; the default is the in-tree build)
_LIB_PATH = os.environ.get("VRAG_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)

VRAG_F16 = 0
VRAG_F32 = 1
VRAG_Q_NORMALIZE = 1
VRAG_Q_POOL = 2
VRAG_Q_FP16 = 4

# pooling kinds (mirrors include/vrag_b200.h)
POOL_TILE_MEAN = 0
POOL_ROW_MEAN = 1
POOL_ADAPTIVE_ROWS = 2
POOL_COLSMOL_EXPERIMENTAL = 3
POOL_LEGACY_CONV = 4
POOL_SMOOTH = 5
POOL_TILE_4N = 6
POOL_GLOBAL_MEAN = 7
POOL_SEQ_CHUNKS = 8

class PoolSpec(C.Structure):
    """vrag_pool_spec_t (include/vrag_b200.h)."""

    _fields_ = [
        ("kind", C.c_int),
        ("patches_per_tile", C.c_int),
        ("grid_h", C.c_int),
        ("grid_w", C.c_int),
        ("target_rows", C.c_int),
        ("clamp_to_h", C.c_int),
        ("num_tiles", C.c_int),
        ("window", C.c_int),
        ("n_weights", C.c_int),
        ("weights", C.c_float * 16),
        ("n_rows", C.c_int),
        ("n_cols", C.c_int),
        ("has_global", C.c_int),
        ("include_self", C.c_int),
        ("via_f16", C.c_int),
        ("derive_from_f32", C.c_int),
        ("input_spec", C.c_int),
        ("in_row_skip", C.c_int),
        ("in_row_count", C.c_int),
    ]


_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); every name here must be exported by the library (tests/test_abi.py)
SIGNATURES = {
    "vrag_last_error": (C.c_char_p, []),
    "vrag_abi_version": (C.c_int, []),
    "vrag_corpus_create": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "vrag_corpus_destroy": (C.c_int, [C.c_void_p]),
    "vrag_store_add": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, _i64p, C.c_int64, C.c_int64]),
    "vrag_store_append": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, _i64p, C.c_int64, C.c_int64]),
    "vrag_store_replace_pages": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, C.c_int64, C.c_void_p, C.c_int, C.c_int, _i64p, C.c_int64]),
    "vrag_store_delete_pages": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, C.c_int64]),
    "vrag_store_truncate": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "vrag_store_compact": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vrag_store_add_synthetic": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, C.c_int64, C.c_int64, C.c_uint64, C.c_int64]),
    "vrag_store_info": (C.c_int, [C.c_void_p, C.c_char_p, _i64p, _i64p, _i64p, _i64p]),
    "vrag_store_read_rows": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "vrag_store_page_range": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, _i64p, _i64p]),
    "vrag_store_page_rows": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, _i64p]),
    "vrag_store_drop": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vrag_search": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int, C.c_uint32, _i64p, C.c_int64, C.c_int, _f32p, _i64p, _i32p]),
    "vrag_score": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int, C.c_uint32, _i64p, C.c_int64, _f32p]),
    "vrag_score_pages": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_uint32, C.POINTER(C.c_void_p), _i64p, C.c_int64, C.c_int, _f32p]),
    "vrag_host_f32_to_f16": (C.c_int, [_f32p, C.POINTER(C.c_uint16), C.c_int64, C.c_int, C.c_int]),
    "vrag_search_multistage": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _u32p, _i32p, _f32p, C.c_int, _i32p, _i64p, C.c_int64, _f32p, _i64p, _i32p]),
    "vrag_filter_create": (C.c_int, [C.c_void_p, _u32p, C.c_int64, _i32p]),
    "vrag_filter_destroy": (C.c_int, [C.c_void_p, C.c_int]),
    "vrag_search_multistage_filtered": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), _u32p, _i32p, _f32p, C.c_int, _i32p, _f32p, _i64p, _i32p]),
    "vrag_search_multistage_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _u32p, _i32p, C.c_int, _f32p, _i32p, C.c_int, _f32p, _i64p, _i32p]),
    "vrag_search_multistage_batch_final": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _u32p, _i32p, C.c_int, _f32p, _i32p, C.c_int, _f32p, _i64p, _f32p, _i32p]),
    "vrag_batch_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, _i32p, C.c_int]),
    "vrag_batch_stage_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_uint32, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrag_batch_prefilter_failed": (C.c_int, [C.c_void_p, C.c_void_p, _i32p]),
    "vrag_topk_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrag_saliency": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int, C.c_int64, _f32p, C.c_int64, _i64p]),
    "vrag_score_dev": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vrag_topk_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrag_comm_unique_id": (C.c_int, [C.c_void_p]),
    "vrag_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vrag_comm_info": (C.c_int, [C.c_void_p, _i32p, _i32p]),
    "vrag_comm_transport": (C.c_int, [C.c_void_p, _i32p]),
    "vrag_comm_destroy": (C.c_int, [C.c_void_p]),
    "vrag_stage_hits_dev": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "vrag_allgather_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vrag_merge_hits_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrag_allreduce_max_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vrag_search_multistage_dev": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _u32p, _i32p, C.c_void_p, C.c_int, _i32p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrag_last_comm_timing": (C.c_int, [C.c_void_p, _f32p, C.c_int, _i32p]),
    "vrag_last_comm_offsets": (C.c_int, [C.c_void_p, _f32p, C.c_int, _i32p]),
    "vrag_pool_out_rows": (C.c_int, [C.POINTER(PoolSpec), C.c_int64, _i64p]),
    "vrag_pool_page": (C.c_int, [C.c_int, C.POINTER(PoolSpec), C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, _i64p]),
    "vrag_store_pool": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(PoolSpec), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]),
    "vrag_last_timing": (C.c_int, [C.c_void_p, _f32p, C.c_int]),
    "vrag_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


class VragError(RuntimeError):
    """An error reported by libvrag_b200 (non-zero status)."""


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found. Build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' at the repo root). "
            "visual_rag_b200 has no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().vrag_last_error()
        raise VragError(msg.decode("utf-8", "replace") if msg else f"libvrag_b200 error {status}")
