// Plain-data interface between the host side of the library (vrag_lib.cu) and the scan kernels
// (maxsim_scan.cuh, instantiated in scan_kernels_*.cu): tile geometry constants and the kernel parameter block.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrag {

constexpr int kDim = 128;            // embedding dim (qdrant_indexer.py:133)
constexpr int kTileRows = 128;       // UMMA M
constexpr int kTileBytes = kTileRows * kDim * 2;  // 32 KB of fp16 per stage
constexpr int kHalfBytes = kTileBytes / 2;        // one K-half (64 fp16 = 128 B per row)
constexpr int kBoxRowsSmall = 32;    // partial tiles are fetched in boxes of at most 32 rows
constexpr int kBoxStep = 4;          // ... sized to the rows that are really needed, in steps of 4 rows: a gathered page of
constexpr int kNumSmallMaps = kBoxRowsSmall / kBoxStep;   // 24 rows is fetched as ONE 24-row box, not as a 32-row box
// Tensor maps {64 cols, 4*(i+1) rows} over the same rows, i = 0..7 (box size is a property of the map).
struct RowMaps {
  CUtensorMap m[kNumSmallMaps];
};
// threads: warp0 TMA, warp1 MMA, then 4 epilogue warps per epilogue group (ScanCfg::threads)
// inv_norm rows travel by 1-D TMA whose global start must be 16-byte aligned: fetch from (row & ~3) with a
// box 4 floats longer and let the epilogue index with the misalignment (row & 3).
// A tile is split into 128/slot_rows slots (1 slot unless candidates are packed); slot j keeps its scales at
// sc[j*(slot_rows+32) + (row & 3) + i] so that boxes of different candidates never overlap.
constexpr int kScaleStride = 256;    // floats per stage (4 slots * (32 + 32))
constexpr int kScaleBoxBig = kTileRows + 4;
constexpr int kScaleBoxSmall = kBoxRowsSmall + 4;
constexpr float kLoScale = 2048.0f;  // lo half of the query is stored scaled by 2^11 (keeps it fp16-normal)

struct ScanParams {
  const long long* offsets;   // [n_pages+1] row offsets of the store (used when fixed_rows == 0)
  const long long* page_end;  // != nullptr: the store has a page table: page p owns rows [offsets[p], page_end[p])
  const uint32_t* mask;       // != nullptr: payload-filter bitmask over the shard's pages (bit p set = page p passes the
                              //   filter): a page that does not pass is treated as an EMPTY page — no tile of it is
                              //   fetched or multiplied (LARGE pages) and its score is -inf (single-query kernels)
  long long fixed_rows;       // > 0: page p owns rows [p*fixed_rows, (p+1)*fixed_rows)
  long long n_pages;          // pages in this store (this shard)
  const long long* cand;      // nullptr: item i is page i. else: item i is page cand[i] - cand_base
  long long cand_base;        // first global page id of this shard
  long long n_items;
  const uint8_t* qimg;        // pre-swizzled B operand image, 2*QP rows x 128 fp16 (see query_prep.cu)
  float* scores;              // [n_items]; items whose page is not in this shard get -inf
  int q_valid;                // real query rows (<= QP)
  int use_scale;              // 1: multiply by inv_norm rows (normalize=True)
  int slot_rows;              // rows reserved per candidate inside a tile (32/64/128); 128 unless PACKED + cand
  int pages_per_tile;         // PACKED + dense + fixed_rows: floor(128 / fixed_rows)
  const int* tile_page0;      // PACKED + dense + variable rows: [n_tiles+1] first page of each tile
  const long long* tile_row0; // PACKED + dense + variable rows: [n_tiles+1] first row of each tile
  int slot_mode;              // PACKED: 1 -> every item is fetched separately into its own slot of slot_rows tile rows
  int hi_only;                // 1: contract only the fp16 hi half of the query (N = QP); QP >= 16
  int shfl_rows;              // PACKED: > 0 -> every page sits in its own power-of-two slot of shfl_rows (<= 32) tile
                              //   rows, so the per-page max is a segmented warp butterfly (no smem round trip)
  int pad_rows;               // PACKED + dense + fixed_rows not a power of two (<= 32): > 0 -> pages are fetched through a
                              //   3-D tensor map {128 cols, fixed_rows, n_pages} with a {64, shfl_rows, 128/shfl_rows} box: rows
                              //   fixed_rows..shfl_rows-1 of every page are out of bounds, so TMA zero-fills them and every
                              //   page lands in its own power-of-two slot of the tile (value = fixed_rows)
  long long n_tiles;          // PACKED: number of tiles (work units) per group
  // ---- query groups (batched candidate lists; BSW kernels). Group g = query g with its own operand image
  // qimg + g*qimg_stride, its own candidate list cand[g*n_items + i] and its own scores[g*n_items + i];
  // n_items / n_tiles above are PER GROUP. n_groups == 1: the single-query layout.
  int n_groups;
  long long qimg_stride;      // bytes between consecutive groups' operand images
  const int* q_valid_arr;     // [n_groups] real query rows of each group (nullptr: q_valid for all)
  // ---- sub-queries (dense batched scans; kernels with QS < QP). One operand image holds QP/QS queries of QS
  // columns each (query j = columns [j*QS, (j+1)*QS) of the hi and of the lo half). Query j of the launch writes
  // scores[j*score_stride + item]; q_valid_arr[j] = its real rows; only the first n_sub queries exist.
  long long score_stride;
  int n_sub;
  // ---- dense scans only: sampling and fused top-k prefilter
  int tile_stride;            // >= 1. > 1: only every tile_stride-th unit (LARGE: page, PACKED fixed rows: tile) is
                              //   scored and scores are written compactly (sample pass of the threshold estimate)
  const float* f_thr;         // != nullptr (QS < QP kernels): instead of writing the score matrix, append
  int* f_cnt;                 //   (score, page) keys of scores > f_thr[j] to f_keys[j*f_cap + atomicAdd(f_cnt[j])]
  unsigned long long* f_keys; //   key = order-preserving score bits << 32 | ~page index (same as the top-k kernels)
  int f_cap;
};

}  // namespace vrag
