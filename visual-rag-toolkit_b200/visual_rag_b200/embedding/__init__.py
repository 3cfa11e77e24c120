"""Pooling + MaxSim entry points of the reference's visual_rag.embedding package (embedding/__init__.py:10),
backed by the sm_100a kernels."""

from .pooling import (  # noqa: F401
    adaptive_row_mean_pooling_from_grid,
    colpali_experimental_pooling_from_rows,
    colpali_row_mean_pooling,
    colsmol_experimental_pooling,
    colsmol_tile_4n_pooling_from_tiles,
    compute_maxsim_batch,
    compute_maxsim_score,
    global_mean_pooling,
    tile_level_mean_pooling,
    weighted_row_smoothing_same_length,
)
