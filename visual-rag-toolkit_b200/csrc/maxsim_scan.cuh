// MaxSim scan kernels for sm_100a (tcgen05 + TMEM + TMA), warp-specialised, one persistent CTA per SM.
//
// Computes, for every work item (a page of the corpus store or a candidate page id):
//     score = sum_{q < Q} max_{t in page} <qhat_q, d_t> * inv_norm_t
// which is the reference's compute_maxsim_score (visual_rag/embedding/pooling.py:468-514) with the
// document-side L2 normalisation folded into a per-row scale (inv_norm_t = 1/(||d_t||+1e-8)).
//
// Mapping onto the tensor core (one tcgen05.mma = 128 x N x 16):
//   A (M = 128 rows)  : 128 consecutive document rows (fp16, K-major, 128B swizzle), staged by TMA
//   B (N = 2*QP rows) : the query, resident in smem for the whole kernel. Rows [0,QP) hold fp16(qhat),
//                       rows [QP,2QP) hold fp16((qhat - fp16(qhat)) * 2^11): the fp32 query is carried
//                       as a hi/lo fp16 pair so the contraction is fp32-accurate on the query side.
//   D (TMEM)          : lane = document row, column = query token (hi | lo halves), fp32.
// Epilogue threads own one document row (TMEM lane) each: s_q = (hi_q + lo_q*2^-11) * inv_norm_row.
//
// Two work layouts share the producer / MMA roles:
//   LARGE pages (rows/page > 128): tiles are page-aligned, the epilogue keeps a running per-thread max
//     over the page's tiles and only reduces across lanes once per page.
//   PACKED pages (rows/page <= 128): several pages share one 128-row tile; the epilogue transposes the
//     scaled scores through smem and does a segmented max per (page, q).
#pragma once
#include "ptx.cuh"
#include "scan_params.h"

namespace vrag {

__device__ __forceinline__ unsigned long long score_key(float f, uint32_t idx) {
  uint32_t o;
  if (f != f) o = 0u;
  else {
    const uint32_t u = __float_as_uint(f);
    o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  return (static_cast<unsigned long long>(o) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

template <int QP>
struct ScanCfg {
  static constexpr int N = 2 * QP;                         // UMMA N (hi | lo)
  static constexpr int ACC = (N <= 128) ? 4 : 2;           // TMEM accumulator stages
  static constexpr int tmem_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }
  static constexpr int TMEM_COLS = tmem_pow2(ACC * N);     // allocations are powers of two >= 32 columns
  static constexpr int B_BYTES = N * kDim * 2;
  static constexpr int SC_PITCH = kTileRows + 4;           // PACKED: transposed score buffer pitch (16 B aligned rows)
  // PACKED: two epilogue groups of 4 warps alternate tiles (the per-tile epilogue is a dependent chain on one
  // warp per SM sub-partition; a second group per sub-partition hides its latency). One score buffer per group.
  static constexpr int EPI_GROUPS_PACKED = (QP <= 64) ? 2 : 1;
  static constexpr int MISC_BYTES = 8192;                  // barriers, reduce scratch, segment tables
  // multi (sub-query kernels, QP == 128): four epilogue groups, each owning 32 accumulator columns
  static constexpr int epi_groups(bool packed, bool multi = false) { return multi ? 4 : (packed ? EPI_GROUPS_PACKED : 1); }
  // PACKED operand-switching kernels (every query of a batch gathers its own candidate pages): a second producer warp
  // (the two alternate tiles: ~300 instructions per tile in ONE warp were the bound of the small-page gather, not HBM)
  // and a warp whose only job is to stream the query operand images (it needs no tile bookkeeping at all)
  static constexpr int extra_warps(bool packed, bool bsw) { return (packed && bsw) ? 2 : 0; }
  static constexpr int threads(bool packed, bool multi = false, bool bsw = false) {
    return 64 + 128 * epi_groups(packed, multi) + 32 * extra_warps(packed, bsw);
  }
  static constexpr int sc_bytes(bool packed, bool multi = false) {
    return packed ? (multi ? 4 * 32 : EPI_GROUPS_PACKED * QP) * SC_PITCH * 4 : 0;
  }
  // PACKED kernels keep the per-tile side data (inverse norms, slot metadata) in a ring of kMetaRing entries indexed by the
  // tile's sequence number instead of inside the row stage: the 32 KB row stage is handed back to the producer as soon as
  // the tile's MMAs retire, not when its epilogue gets round to it (gathers are bound by tiles in flight x latency). The
  // ring is deeper than stages + accumulator stages, the furthest the producer can run ahead of the epilogue.
  static constexpr int kMetaRing = 16;
  static constexpr int meta_entries(bool packed, int n_stages) { return packed ? kMetaRing : n_stages; }
  static constexpr int stage_bytes(bool packed) { return kTileBytes + (packed ? 0 : kScaleStride * 4); }
  // bsw: the query operand is double-buffered so that a CTA can switch between query groups mid-kernel
  static constexpr int stages(bool packed, bool bsw = false, bool multi = false) {
    const int budget = 227 * 1024 - 1024 /*align slack*/ - (bsw ? 2 : 1) * B_BYTES - MISC_BYTES - sc_bytes(packed, multi) -
                       (packed ? kMetaRing * kScaleStride * 4 : 0);
    const int s = budget / stage_bytes(packed);
    return s > 6 ? 6 : s;
  }
  static constexpr size_t smem_bytes(bool packed, bool bsw = false, bool multi = false) {
    return 1024 + size_t(stages(packed, bsw, multi)) * stage_bytes(packed) + (packed ? kMetaRing * kScaleStride * 4 : 0) +
           (bsw ? 2 : 1) * B_BYTES + MISC_BYTES + sc_bytes(packed, multi);
  }
};

// Resolve work item -> (first row, row count). Returns false when the item is not a page of this shard.
__device__ __forceinline__ bool page_allowed(const ScanParams& p, long long page) {
  return p.mask == nullptr || ((__ldg(p.mask + (page >> 5)) >> (page & 31)) & 1u) != 0u;
}
__device__ __forceinline__ bool resolve_page(const ScanParams& p, long long page, long long& row0, int& nrows) {
  if (page < 0 || page >= p.n_pages || !page_allowed(p, page)) {   // outside the shard, or filtered out: an empty page
    row0 = 0;
    nrows = 0;
    return false;
  }
  if (p.fixed_rows > 0) {
    row0 = page * p.fixed_rows;
    nrows = static_cast<int>(p.fixed_rows);
  } else {
    const long long a = __ldg(p.offsets + page), b = p.page_end ? __ldg(p.page_end + page) : __ldg(p.offsets + page + 1);
    row0 = a;
    nrows = static_cast<int>(b - a);
  }
  return true;
}
__device__ __forceinline__ long long item_page(const ScanParams& p, long long item, int g = 0) {
  return p.cand ? (__ldg(p.cand + g * p.n_items + item) - p.cand_base) : item * p.tile_stride;
}

// Work units of one CTA. One group: units blockIdx.x, +gridDim.x, ... (neighbouring CTAs stream neighbouring
// tiles). Several groups: a contiguous range of the (group-major) unit space, so that a CTA changes its query
// operand only a few times. Unit U -> (group U / units_per_group, local unit U % units_per_group).
struct UnitRange {
  long long first, step, count, per_group;
  __device__ __forceinline__ void decode(long long i, int& g, long long& u) const {
    const long long U = first + i * step;
    if (per_group == 0) { g = 0; u = U; }
    else { g = static_cast<int>(U / per_group); u = U - g * per_group; }
  }
};
__device__ __forceinline__ UnitRange unit_range(const ScanParams& p, long long units_per_group) {
  UnitRange r;
  if (p.n_groups <= 1) {
    r.first = blockIdx.x;
    r.step = gridDim.x;
    r.count = units_per_group > r.first ? (units_per_group - r.first + r.step - 1) / r.step : 0;
    r.per_group = 0;
  } else {
    const long long total = units_per_group * p.n_groups;
    const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    r.first = lo;
    r.step = 1;
    r.count = hi - lo;
    r.per_group = units_per_group;
  }
  return r;
}

// Walks the units of a UnitRange without a 64-bit division per unit (the per-tile instruction streams of the producer,
// the MMA thread and the epilogue warps are what bounds gathers of small pages): seek() divides once, advance() adds.
struct UnitIter {
  int g;
  long long u;
  __device__ __forceinline__ void seek(const UnitRange& r, long long i) { r.decode(i, g, u); }
  __device__ __forceinline__ void advance(const UnitRange& r, long long n) {
    if (r.per_group == 0) {
      u += n * r.step;
    } else {
      u += n;   // step == 1 in the multi-group layout
      while (u >= r.per_group) {
        u -= r.per_group;
        ++g;
      }
    }
  }
};

// Row ranges one PACKED tile fetches: dense layouts -> one contiguous range; slot mode -> up to 4 items.
struct PackedTileMeta {
  long long r0[4];
  int nr[4];
  int cnt;
};
__device__ __forceinline__ void packed_tile_meta(const ScanParams& p, int g, long long u, PackedTileMeta& m) {
  if (!p.slot_mode) {
    m.cnt = 1;
    if (p.fixed_rows > 0) {
      const long long pg0 = u * p.tile_stride * p.pages_per_tile;
      const long long pg1 = min(pg0 + p.pages_per_tile, p.n_pages);
      m.r0[0] = pg0 * p.fixed_rows;
      m.nr[0] = static_cast<int>((pg1 - pg0) * p.fixed_rows);
    } else {
      const long long a = __ldg(p.tile_row0 + u), b = __ldg(p.tile_row0 + u + 1);
      m.r0[0] = a;
      m.nr[0] = static_cast<int>(b - a);
    }
  } else {
    const int per_tile = kTileRows / p.slot_rows;
    const long long i0 = u * per_tile;
    m.cnt = static_cast<int>(min(static_cast<long long>(per_tile), p.n_items - i0));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m.r0[j] = 0;
      m.nr[j] = 0;
      if (j < m.cnt) resolve_page(p, item_page(p, i0 + j, g), m.r0[j], m.nr[j]);
    }
  }
}

// General PACKED path: entry `et` of tile u's segment table (item, first tile row, end tile row) and the table size.
__device__ __forceinline__ void packed_segment(const ScanParams& p, int g, long long u, int et, int& nseg, int& item,
                                               int& rb, int& re) {
  item = rb = re = 0;
  if (!p.slot_mode) {
    long long pg0, pg1;
    if (p.fixed_rows > 0) {
      pg0 = u * p.pages_per_tile;
      pg1 = min(pg0 + p.pages_per_tile, p.n_pages);
    } else {
      pg0 = __ldg(p.tile_page0 + u);
      pg1 = __ldg(p.tile_page0 + u + 1);
    }
    nseg = static_cast<int>(pg1 - pg0);
    if (et < nseg) {
      long long rbase, r0;
      int tmp, nr;
      if (p.fixed_rows > 0) rbase = pg0 * p.fixed_rows;
      else rbase = __ldg(p.tile_row0 + u);
      (void)tmp;
      const bool live = resolve_page(p, pg0 + et, r0, nr);
      item = static_cast<int>(pg0 + et);   // dense: item == page (n_items == n_pages < 2^31 per shard)
      rb = live ? static_cast<int>(r0 - rbase) : 0;   // filtered-out page: an empty segment (-inf)
      re = rb + nr;
    }
  } else {
    const int per_tile = kTileRows / p.slot_rows;
    const long long i0 = u * per_tile;
    nseg = static_cast<int>(min(static_cast<long long>(per_tile), p.n_items - i0));
    if (et < nseg) {
      long long r0;
      int nr;
      resolve_page(p, item_page(p, i0 + et, g), r0, nr);
      item = static_cast<int>(i0 + et);
      rb = et * p.slot_rows;
      re = rb + nr;   // nr == 0 -> empty segment -> -inf
    }
  }
}

// Fetch `nrows` (1..128) document rows starting at global row `row` into tile rows [dst_row, dst_row+nrows)
// of stage buffer `a` (+ their inv_norm scales at sc[(row & 3) ...]). dst_row is a multiple of 32.
// Returns the bytes the mbarrier must expect.
__device__ __forceinline__ uint32_t issue_rows(uint8_t* a, float* sc, uint64_t* bar, const CUtensorMap* tm128,
                                               const RowMaps* small, const CUtensorMap* ts128,
                                               const CUtensorMap* ts32, long long row, int nrows, int dst_row,
                                               bool use_scale) {
  const int32_t r32 = static_cast<int32_t>(row);
  if (dst_row == 0 && nrows > 96) {
    tma_load_2d(a, tm128, bar, 0, r32);
    tma_load_2d(a + kHalfBytes, tm128, bar, 64, r32);
    if (use_scale) tma_load_1d(sc, ts128, bar, r32 & ~3);
    return kTileBytes + (use_scale ? kScaleBoxBig * 4 : 0);
  }
  // boxes of at most 32 rows; the last one is sized to the rows that are left (rounded up to 4): the rows of the tile
  // beyond it keep stale shared memory, which the epilogue never reads (it masks by the page's row count)
  const int nb = (nrows + kBoxRowsSmall - 1) / kBoxRowsSmall;
  uint32_t bytes = 0;
  for (int j = 0; j < nb; ++j) {
    const int32_t r = r32 + j * kBoxRowsSmall;
    const int d = dst_row + j * kBoxRowsSmall;
    const int left = nrows - j * kBoxRowsSmall;
    const int mi = (min(left, kBoxRowsSmall) + kBoxStep - 1) / kBoxStep - 1;
    const CUtensorMap* tm = &small->m[mi];
    tma_load_2d(a + d * 128, tm, bar, 0, r);
    tma_load_2d(a + kHalfBytes + d * 128, tm, bar, 64, r);
    bytes += (mi + 1) * kBoxStep * kDim * 2;
    // consecutive scale boxes of one slot overlap by 4 floats in smem; both write identical values there
    if (use_scale) tma_load_1d(sc + j * kBoxRowsSmall, ts32, bar, r & ~3);
  }
  return bytes + nb * (use_scale ? kScaleBoxSmall * 4 : 0);
}

// Bytes issue_rows() makes the mbarrier expect for the same arguments (pure arithmetic: lets a converged warp know the
// count while only one elected lane issues the copies).
__device__ __forceinline__ uint32_t issue_rows_bytes(int nrows, int dst_row, bool use_scale) {
  if (dst_row == 0 && nrows > 96) return kTileBytes + (use_scale ? kScaleBoxBig * 4 : 0);
  const int nb = (nrows + kBoxRowsSmall - 1) / kBoxRowsSmall;
  const int last = nrows - (nb - 1) * kBoxRowsSmall;
  const int last_rows = ((last + kBoxStep - 1) / kBoxStep) * kBoxStep;
  return ((nb - 1) * kBoxRowsSmall + last_rows) * (kDim * 2) + nb * (use_scale ? kScaleBoxSmall * 4 : 0);
}

// In-place butterfly max over the 32 lanes of a warp for CNT (power of two <= 32) values per lane.
// On return v[0] of lane L holds the max over all lanes of value index (L >> (5 - log2 CNT)).
template <int CNT>
__device__ __forceinline__ void warp_transpose_max(float* v, int lane) {
  int cnt = CNT;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      const int half = cnt >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < half) {
          const float keep = upper ? v[j + half] : v[j];
          const float send = upper ? v[j] : v[j + half];
          const float recv = __shfl_xor_sync(0xffffffffu, send, off);
          v[j] = fmaxf(keep, recv);
        }
      }
      cnt = half;
    } else {
      v[0] = fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
    }
  }
}

// Packed fp16 pairs are carried as plain 32-bit registers (arrays of them stay in registers under full unrolling).
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_pack(float lo, float hi) {   // round-to-nearest-even, {lo -> bits 0-15, hi -> bits 16-31}
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
constexpr uint32_t kH2NegInf = 0xFC00FC00u;
// The same butterfly over packed fp16 pairs: 16 pairs (32 columns) per lane. On return lane L holds, in v[0], the maxima
// over all lanes of the column pair (L >> 1); half the shuffles and selects of the fp32 version.
__device__ __forceinline__ void warp_transpose_max_h2x16(uint32_t (&v)[16], int lane) {
  int cnt = 16;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      const int half = cnt >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < half) {
          const uint32_t keep = upper ? v[j + half] : v[j];
          const uint32_t send = upper ? v[j] : v[j + half];
          v[j] = h2_max(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
      }
      cnt = half;
    } else {
      v[0] = h2_max(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
    }
  }
}
__device__ __forceinline__ float h2_column_max(uint32_t v0, int lane) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>((lane & 1) ? (v0 >> 16) : (v0 & 0xFFFFu))));
}

// First (approximate) pass of a batched exhaustive scan. The operand image interleaves the two query blocks of an epilogue
// group in quads of token columns — A[0:4] B[0:4] A[4:8] B[4:8] ... (query_prep_group_kernel, two_block) — so ONE TMEM load
// of W columns starting at column 2*C of the group brings tokens [C, C + W/2) of both queries, and a pair of queries of 20
// tokens reads exactly 40 columns. Values are scaled, rounded to fp16 pairs and folded into the running maxima.
template <int C, int W>
__device__ __forceinline__ void first_pass_cols(uint32_t tg, bool live, float scale, uint32_t (&runA)[16], uint32_t (&runB)[16]) {
  uint32_t v[W];
  tmem_ld<W>(tg + 2 * C, v);
  tmem_ld_wait();
  if (live) {
#pragma unroll
    for (int j = 0; j < W; j += 8) {   // one quad of each block
      const int t = C + j / 2;
      runA[t / 2] = h2_max(runA[t / 2], h2_pack(__uint_as_float(v[j]) * scale, __uint_as_float(v[j + 1]) * scale));
      runA[t / 2 + 1] = h2_max(runA[t / 2 + 1], h2_pack(__uint_as_float(v[j + 2]) * scale, __uint_as_float(v[j + 3]) * scale));
      runB[t / 2] = h2_max(runB[t / 2], h2_pack(__uint_as_float(v[j + 4]) * scale, __uint_as_float(v[j + 5]) * scale));
      runB[t / 2 + 1] = h2_max(runB[t / 2 + 1], h2_pack(__uint_as_float(v[j + 6]) * scale, __uint_as_float(v[j + 7]) * scale));
    }
  }
}

// PACKED fast path: the SR (power of two <= 32) lanes of a slot hold QP values each (one tile row per lane).
// Segmented transpose-max over the slot's lanes, then sum over q. Returns the slot's score in all of its lanes.
template <int QP, int SR>
__device__ __forceinline__ void slot_colmax(float* v, int lane) {
  int cnt = QP;
#pragma unroll
  for (int off = SR / 2; off >= 1; off >>= 1) {
    if (cnt > 1) {
      const int half = cnt >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < half) {
          const float keep = upper ? v[j + half] : v[j];
          const float send = upper ? v[j] : v[j + half];
          v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
      }
      cnt = half;
    } else {
      v[0] = fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
    }
  }
}
template <int QP, int SR>
__device__ __forceinline__ float slot_maxsim(float* v, int lane, int q_valid) {
  slot_colmax<QP, SR>(v, lane);
  constexpr int CF = QP >= SR ? QP / SR : 1;    // q values left per lane
  constexpr int DUP = QP >= SR ? 1 : SR / QP;   // lanes holding the same q
  const int b = lane & (SR - 1);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < CF; ++i) {
    const int q = (b / DUP) * CF + i;
    if (q < q_valid && (b % DUP) == 0) sum += v[i];
  }
#pragma unroll
  for (int off = SR / 2; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  return sum;
}

// QS: columns per query inside the operand image. QS == QP: one query per image (single-query scans and BSW
// candidate scans). QS < QP (dense batched scans): QP/QS queries share every document tile.
// NBLK == 2 (LARGE pages, QS == 32): the operand image holds EIGHT queries as plain fp16 (rows [0,128) = queries 0-3,
// rows [128,256) = queries 4-7; no lo halves): the first pass of a batched exhaustive scan — twice the queries per
// document tile for the same tensor work; the host re-scores the top candidates exactly (hi|lo) afterwards.
template <int QP, int QS, bool PACKED, bool BSW, int NBLK = 1>
__global__ void __launch_bounds__(ScanCfg<QP>::threads(PACKED, QS < QP, BSW), 1)
maxsim_scan_kernel(const __grid_constant__ CUtensorMap tm_rows128, const __grid_constant__ RowMaps tm_small,
                   const __grid_constant__ CUtensorMap tm_scale128, const __grid_constant__ CUtensorMap tm_scale32,
                   const ScanParams p) {
  using Cfg = ScanCfg<QP>;
  constexpr int N = Cfg::N;
  constexpr int ACC = Cfg::ACC;
  constexpr bool MULTI = QS < QP;
  constexpr int STAGES = Cfg::stages(PACKED, BSW, MULTI);
  constexpr int NB = BSW ? 2 : 1;              // query operand buffers
  constexpr int NTHREADS = Cfg::threads(PACKED, MULTI, BSW);
  constexpr int PRODUCERS = Cfg::extra_warps(PACKED, BSW) ? 2 : 1;   // producer warps (slot-mode tiles alternate between them)
  constexpr int WARP_PROD_B = 2 + 4 * Cfg::epi_groups(PACKED, MULTI);  // second producer warp / operand loader warp
  constexpr int WARP_OPERANDS = WARP_PROD_B + 1;                       //   (exist only when PRODUCERS == 2)
  constexpr int EPI_GROUPS = Cfg::epi_groups(PACKED, MULTI);
  constexpr int QE = MULTI ? 32 : QP;          // accumulator columns one epilogue group handles
  constexpr int QR = QE <= 8 ? 8 : QE <= 16 ? 16 : QE <= 32 ? 32 : QE;   // QE padded to a power of two (reductions)
  constexpr int QG = (QR + 31) / 32;           // 32-wide column groups per epilogue group
  constexpr int QW = QR < 32 ? QR : 32;        // columns per 32-wide group
  constexpr int RS = MULTI ? QP : QR;          // LARGE: row stride of the cross-warp reduce scratch
  static_assert(QE == QR || !PACKED, "non-power-of-two query widths are only built for LARGE pages");
  constexpr int EPI_ARRIVALS = MULTI ? 16 : 4; // epilogue warps that consume every tile
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(QS == QP || ((QS == 1 || QS == 32) && QP == 128 && !BSW), "sub-query layouts: 128x1 or 4x32 columns");
  static_assert(NBLK == 1 || (NBLK == 2 && QS == 32 && QP == 128 && !PACKED && !BSW), "two query blocks: LARGE 8x32 only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kTileBytes;
  float* sScale = reinterpret_cast<float*>(sB + NB * Cfg::B_BYTES);
  constexpr int META = Cfg::meta_entries(PACKED, STAGES);             // side-data entries: ring (PACKED) / one per stage
  static_assert(!PACKED || META >= 6 + ACC + 1, "side-data ring shorter than the producer's lead over the epilogue");
  float* sSc = sScale + META * kScaleStride;                         // PACKED: [EPI_GROUPS][QP][SC_PITCH]
  uint8_t* misc = reinterpret_cast<uint8_t*>(sSc) + Cfg::sc_bytes(PACKED, MULTI);
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC;
  uint64_t* bfull = tempty + ACC;            // BSW: operand buffer filled (bulk copy) / drained (MMAs retired)
  uint64_t* bempty = bfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + 2);
  int* sMis = reinterpret_cast<int*>(misc + 256);                 // [META][4] scale misalignment (+ slot rows) per slot
  float* sRed = reinterpret_cast<float*>(misc + 512);             // LARGE: [2][4][QP]
  int* sSeg = reinterpret_cast<int*>(misc + 512);                 // PACKED: [EPI_GROUPS][3][128] ints (item, begin, end)
  float* sThr = reinterpret_cast<float*>(misc + 7168);            // QS < QP: [128] prefilter thresholds
  int* sStageCnt = reinterpret_cast<int*>(misc + 7680);           // QS < QP: [128] survivors staged by this CTA per query
  // Prefilter survivors of the slot/shuffle epilogue are staged per CTA in the (otherwise unused) transpose buffer and
  // flushed with one global atomicAdd per (CTA, query): thousands of same-address atomics from all SMs serialise in L2.
  constexpr int kStageKeys = 8192;                                // 64 KB of keys, split evenly between the launch's queries
  unsigned long long* sStage = reinterpret_cast<unsigned long long*>(sSc);
  const bool stage_on = PACKED && MULTI && p.f_thr != nullptr && p.shfl_rows > 0;
  const int stage_cap = kStageKeys / (QP / QS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool use_scale = p.use_scale != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows128);
    for (int i = 0; i < kNumSmallMaps; ++i) tma_prefetch_desc(&tm_small.m[i]);
    tma_prefetch_desc(&tm_scale128);
    tma_prefetch_desc(&tm_scale32);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      // LARGE: tcgen05.commit + one arrive per consuming epilogue warp (it reads the stage's scale rows).
      // PACKED: the commit alone — the epilogue reads its side data from the ring, never from the stage.
      mbar_init(&empty[s], PACKED ? 1 : 1 + EPI_ARRIVALS);
    }
    for (int a = 0; a < ACC; ++a) {
      // PACKED: + a plain arrive of the MMA thread after it observed full[stage]: the epilogue acquires the TMA-written
      // side data through tfull (it must not wait on full[stage] itself: the stage may already be in its next round)
      mbar_init(&tfull[a], PACKED ? 2 : 1);
      mbar_init(&tempty[a], EPI_ARRIVALS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bfull[b], 1);
      mbar_init(&bempty[b], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if constexpr (!BSW) {  // query operand image: global -> smem (already in the swizzled UMMA layout)
    const uint4* src = reinterpret_cast<const uint4*>(p.qimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = threadIdx.x; i < Cfg::B_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    fence_proxy_async_smem();
  }
  if constexpr (QS < QP) {
    if (p.f_thr && threadIdx.x < 128) sThr[threadIdx.x] = static_cast<int>(threadIdx.x) < p.n_sub ? __ldg(p.f_thr + threadIdx.x) : INFINITY;
    if (threadIdx.x < 128) sStageCnt[threadIdx.x] = 0;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // QS < QP kernels: deliver the score of (query j of the launch, page): into the score matrix at column `col`, or,
  // with the prefilter on, into query j's candidate list when it beats the query's threshold.
  auto emit = [&](int j, long long col, long long page, float val) {
    if (p.f_thr) {
      if (val > sThr[j]) {
        const unsigned long long key = score_key(val, static_cast<uint32_t>(page));
        if (stage_on) {
          const int slot = atomicAdd(&sStageCnt[j], 1);
          if (slot < stage_cap) {
            sStage[j * stage_cap + slot] = key;
            return;
          }
        }
        const int pos = atomicAdd(p.f_cnt + j, 1);
        if (pos < p.f_cap) p.f_keys[static_cast<long long>(j) * p.f_cap + pos] = key;
      }
    } else {
      p.scores[j * p.score_stride + col] = val;
    }
  };

  // Work units of this CTA (LARGE: items, PACKED: tiles)
  const UnitRange ur = unit_range(p, PACKED ? p.n_tiles : p.n_items);

  const bool is_prod_b = PRODUCERS == 2 && warp == WARP_PROD_B;
  if (PRODUCERS == 2 && warp == WARP_OPERANDS) {
    // ===================================================================== operand loader (PACKED + BSW)
    // The CTA's units are a contiguous range of the group-major unit space, so it visits query groups g_first..g_last in
    // order: stream their operand images into the two buffers, each as soon as the MMAs of the group that used the buffer
    // before have retired. No tile bookkeeping, and the tile producers never wait for an operand switch.
    if constexpr (BSW) {
      if (lane == 0 && ur.count > 0) {
        int g_first, g_last;
        long long u;
        ur.decode(0, g_first, u);
        ur.decode(ur.count - 1, g_last, u);
        int n_sw = -1;
        for (int g = g_first; g <= g_last; ++g) {
          ++n_sw;
          const int slot = n_sw & 1;
          if (n_sw >= 2) mbar_wait(&bempty[slot], ((n_sw >> 1) - 1) & 1);
          bulk_load(sB + slot * Cfg::B_BYTES, p.qimg + g * p.qimg_stride, Cfg::B_BYTES, &bfull[slot]);
          mbar_arrive_expect_tx(&bfull[slot], Cfg::B_BYTES);
        }
      }
    }
  } else if (warp == 0 || is_prod_b) {
    // ===================================================================== TMA producer(s)
    uint32_t stage = 0, phase = 0;
    int cur_g = -1, n_sw = -1;
    // BSW with a single producer (LARGE pages): entering query group g -> fill the other operand buffer once its previous
    // MMAs retired (lane 0 only). With two producers the operand loader warp above does this.
    // (called by the whole, converged warp: the wait is uniform, the copy is issued by one elected lane)
    auto switch_group = [&](int g) {
      if constexpr (BSW && PRODUCERS == 1) {
        if (g != cur_g) {
          ++n_sw;
          const int slot = n_sw & 1;
          if (n_sw >= 2) mbar_wait(&bempty[slot], ((n_sw >> 1) - 1) & 1);
          if (elect_one_sync()) {
            bulk_load(sB + slot * Cfg::B_BYTES, p.qimg + g * p.qimg_stride, Cfg::B_BYTES, &bfull[slot]);
            mbar_arrive_expect_tx(&bfull[slot], Cfg::B_BYTES);
          }
          __syncwarp();
          cur_g = g;
        }
      }
    };
    (void)cur_g; (void)n_sw;
    bool slot_path = false;
    if constexpr (PACKED) slot_path = p.slot_mode != 0;
    if (slot_path) {
      if constexpr (PACKED) {
        // Slot mode (gathered pages, one slot per item): item -> candidate id -> page offsets is a chain of dependent
        // global loads (~2 us). The whole warp resolves 32 items at a time (lane l: item l of the batch), one batch
        // ahead of the tiles being issued, so the chain is off the TMA issue path.
        // Producer warp pw takes the tiles pw, pw + PRODUCERS, ... of the CTA's range ("own" tiles, counted by k); tile i
        // uses stage i % STAGES whoever issues it, so the MMA thread still consumes the stages in order.
        // (a CTA's share of the tiles fits 32 bits: the per-tile index arithmetic below is on the single-warp critical path)
        const int pw = is_prod_b ? 1 : 0;
        const uint32_t n_tiles_cta = static_cast<uint32_t>(ur.count);
        const uint32_t n_own = n_tiles_cta > static_cast<uint32_t>(pw) ? (n_tiles_cta - pw + PRODUCERS - 1) / PRODUCERS : 0u;
        if (p.slot_rows == kBoxRowsSmall) {
          // ---- one page per 32-row slot, 4 slots per tile: the gather of small (pooled) pages. What bounds it is the
          // producers' instruction stream, so the per-item work is split in two: all 32 lanes resolve and pre-compute one
          // batch of 32 items (8 tiles) at a time — row, box map, byte count, side data — and publish the batch in shared
          // memory; the tile loop then runs converged and ONE elected lane issues the tile's copies from four 16-byte reads.
          // The dependent global loads (item -> candidate id -> page rows) are software-pipelined over the batches: ids two
          // batches ahead, row ranges one batch ahead, so the in-order warp never waits for either.
          int4* pbuf = reinterpret_cast<int4*>(misc + (MULTI ? 6656 : 4096)) + pw * 32;   // (free space between the segment tables and sThr)
          auto batch_pages = [&](uint32_t k0) -> long long {
            const long long i = pw + PRODUCERS * static_cast<long long>(k0 + (lane >> 2));
            long long page = -1;
            if (i < ur.count) {
              int g;
              long long u;
              ur.decode(i, g, u);
              const long long it = u * 4 + (lane & 3);
              if (it < p.n_items) page = item_page(p, it, g);
            }
            return page;
          };
          long long r0_c, pg_n;
          int nr_c;
          resolve_page(p, batch_pages(0), r0_c, nr_c);
          pg_n = batch_pages(8);
          stage = static_cast<uint32_t>(pw) % STAGES;
          phase = 0;
          uint32_t me = static_cast<uint32_t>(pw);
          for (uint32_t k0 = 0; k0 < n_own; k0 += 8) {
            {
              const int mi = (nr_c + kBoxStep - 1) / kBoxStep - 1;
              int4 e;
              e.x = static_cast<int32_t>(r0_c);                                   // first row (TMA coordinate)
              e.y = nr_c > 0 ? (static_cast<int>(r0_c & 3) | (nr_c << 2)) : 0;    // side data: scale misalignment | rows << 2
              e.z = mi * static_cast<int>(sizeof(CUtensorMap));                   // which box map
              e.w = nr_c > 0 ? (mi + 1) * (kBoxStep * kDim * 2) + (use_scale ? kScaleBoxSmall * 4 : 0) : 0;   // bytes
              pbuf[lane] = e;
            }
            __syncwarp();
            resolve_page(p, pg_n, r0_c, nr_c);   // batch k0 + 8 (its ids were loaded one batch ago)
            pg_n = batch_pages(k0 + 16);
            const uint32_t nt = min(8u, n_own - k0);
            for (uint32_t t = 0; t < nt; ++t) {
              mbar_wait(&empty[stage], phase ^ 1);   // the whole warp (uniform)
              if (elect_one_sync()) {
                uint8_t* a = sA + stage * kTileBytes;
                float* sc = sScale + me * kScaleStride;
                uint32_t bytes = 0;
                int misv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int4 e = pbuf[t * 4 + j];
                  if (e.w > 0) {
                    const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const char*>(&tm_small) + e.z);
                    uint8_t* dst = a + j * (kBoxRowsSmall * 128);
                    tma_load_2d(dst, tm, &full[stage], 0, e.x);
                    tma_load_2d(dst + kHalfBytes, tm, &full[stage], 64, e.x);
                    if (use_scale) tma_load_1d(sc + j * (kBoxRowsSmall + 32), &tm_scale32, &full[stage], e.x & ~3);
                  }
                  bytes += static_cast<uint32_t>(e.w);
                  misv[j] = e.y;
                }
                *reinterpret_cast<int4*>(sMis + me * 4) = make_int4(misv[0], misv[1], misv[2], misv[3]);
                mbar_arrive_expect_tx(&full[stage], bytes);   // same lane as the copies and the side-data store
              }
              __syncwarp();
              stage += PRODUCERS;
              if (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
              me = (me + PRODUCERS) % META;
            }
          }
        } else {
        // ---- 64- / 128-row slots (2 or 1 items per tile): several boxes per item (issue_rows)
        const int per_tile = kTileRows / p.slot_rows;
        const int tiles_per_batch = 32 / per_tile;
        auto resolve_batch = [&](uint32_t k0, long long& r0, int& nr) {
          const long long i = pw + PRODUCERS * (k0 + lane / per_tile);
          r0 = 0;
          nr = 0;
          if (i < ur.count) {
            int g;
            long long u;
            ur.decode(i, g, u);
            const long long it = u * per_tile + (lane % per_tile);
            if (it < p.n_items) resolve_page(p, item_page(p, it, g), r0, nr);
          }
        };
        long long cur_r0, nxt_r0;
        int cur_nr, nxt_nr;
        resolve_batch(0, cur_r0, cur_nr);
        for (uint32_t k0 = 0; k0 < n_own; k0 += tiles_per_batch) {
          resolve_batch(k0 + tiles_per_batch, nxt_r0, nxt_nr);
          for (int t = 0; t < tiles_per_batch; ++t) {
            if (k0 + t >= n_own) break;   // warp-uniform
            const uint32_t i = pw + PRODUCERS * (k0 + t);          // tile index in the CTA's range
            stage = i % STAGES;
            phase = (i / STAGES) & 1u;
            mbar_wait(&empty[stage], phase ^ 1);   // the whole warp (uniform)
            uint8_t* a = sA + stage * kTileBytes;
            const int me = static_cast<int>(i % META);          // side-data ring entry of this tile
            float* sc = sScale + me * kScaleStride;
            // The tile's items are issued one after the other from warp-uniform values (their row ranges are broadcast
            // from the lanes that resolved them): uniform operands keep every copy at a handful of instructions;
            // per-lane operands cost a register -> uniform-register waterfall per copy.
            uint32_t bytes = 0;
            for (int j = 0; j < per_tile; ++j) {
              const int src = (t * per_tile + j) & 31;
              const long long r0 = __shfl_sync(0xffffffffu, cur_r0, src);
              const int nr = __shfl_sync(0xffffffffu, cur_nr, src);
              if (nr > 0) bytes += issue_rows_bytes(nr, j * p.slot_rows, use_scale);
              if (elect_one_sync()) {
                if (nr > 0)
                  issue_rows(a, sc + j * (p.slot_rows + 32), &full[stage], &tm_rows128, &tm_small, &tm_scale128, &tm_scale32, r0, nr,
                             j * p.slot_rows, use_scale);
                // low 2 bits: scale misalignment; rest: rows of the slot (0 for unused / empty slots)
                sMis[me * 4 + j] = static_cast<int>(r0 & 3) | (nr << 2);
              }
              __syncwarp();
            }
            if (elect_one_sync()) {
              for (int j = per_tile; j < 4; ++j) sMis[me * 4 + j] = 0;
              mbar_arrive_expect_tx(&full[stage], bytes);   // same lane as the copies and the sMis stores
            }
            __syncwarp();
          }
          cur_r0 = nxt_r0;
          cur_nr = nxt_nr;
        }
        }
      }
    } else if (!is_prod_b) {
      // streaming layouts (LARGE pages, dense PACKED tiles): the warp runs converged, one elected lane issues the copies
      PackedTileMeta cur, nxt;
      cur.cnt = nxt.cnt = 0;
      long long row0 = 0, row0_n = 0;
      int nrows = 0, nrows_n = 0;
      for (long long i = 0; i < ur.count; ++i) {
        int g;
        long long u;
        ur.decode(i, g, u);
        switch_group(g);
        if constexpr (!PACKED) {
          // the next item's row range is resolved before this item's tiles are issued (dependent global loads)
          if (i == 0) resolve_page(p, item_page(p, u, g), row0, nrows);
          if (i + 1 < ur.count) {
            int gn;
            long long un;
            ur.decode(i + 1, gn, un);
            resolve_page(p, item_page(p, un, gn), row0_n, nrows_n);
          }
          for (int t0 = 0; t0 < nrows; t0 += kTileRows) {
            const int rows = min(kTileRows, nrows - t0);
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a = sA + stage * kTileBytes;
            float* sc = sScale + stage * kScaleStride;
            if (elect_one_sync()) {
              // copies first, then one arrive.expect_tx with the exact byte count: the phase cannot complete
              // before the arrive, and the tx-count may go transiently negative.
              const uint32_t bytes = issue_rows(a, sc, &full[stage], &tm_rows128, &tm_small, &tm_scale128,
                                                &tm_scale32, row0 + t0, rows, 0, use_scale);
              sMis[stage * 4] = static_cast<int>((row0 + t0) & 3);
              mbar_arrive_expect_tx(&full[stage], bytes);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          row0 = row0_n;
          nrows = nrows_n;
        } else {
          // dense: the tile's pages are contiguous rows -> one fetch; the row range is loaded one tile ahead
          if (i == 0) packed_tile_meta(p, g, u, cur);
          if (i + 1 < ur.count) {
            int gn;
            long long un;
            ur.decode(i + 1, gn, un);
            packed_tile_meta(p, gn, un, nxt);
          }
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a = sA + stage * kTileBytes;
          const int me = static_cast<int>(i % META);                     // side-data ring entry of this tile
          float* sc = sScale + me * kScaleStride;
          uint32_t bytes = 0;
          if (elect_one_sync()) {
          if (p.pad_rows > 0) {
            // padded slots: the first map is the store's 3-D {cols, rows-in-page, pages} view; one box per K-half
            const int32_t pg0 = static_cast<int32_t>(cur.r0[0] / p.pad_rows);
            tma_load_3d(a, &tm_rows128, &full[stage], 0, 0, pg0);
            tma_load_3d(a + kHalfBytes, &tm_rows128, &full[stage], 64, 0, pg0);
            bytes = kTileBytes;   // out-of-bounds rows / pages are zero-filled and counted
            if (use_scale) {
              tma_load_1d(sc, &tm_scale128, &full[stage], static_cast<int32_t>(cur.r0[0]) & ~3);
              bytes += kScaleBoxBig * 4;
            }
          } else if (cur.nr[0] > 0)
            bytes = issue_rows(a, sc, &full[stage], &tm_rows128, &tm_small, &tm_scale128, &tm_scale32, cur.r0[0],
                               cur.nr[0], 0, use_scale);
          sMis[me * 4] = static_cast<int>(cur.r0[0] & 3);
          mbar_arrive_expect_tx(&full[stage], bytes);
          }
          __syncwarp();
          cur = nxt;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp runs this loop converged (waits included) and ONE elected lane issues the tcgen05 instructions: with
    // uniform control flow the operand descriptors and the loop state live in the uniform datapath. (With `if (lane == 0)`
    // around the loop every UTCHMMA needed ~19 instructions of per-thread descriptor arithmetic and register -> uniform
    // register waterfalls; that single instruction stream, ~200 per tile, was what the scan ran at once the SM clock dropped
    // under the power cap.)
    {
      const uint32_t idesc = umma_idesc_f16(kTileRows, (p.hi_only && NBLK == 1) ? ((QP + 15) / 16) * 16 : N);
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t b_addr = smem_u32(sB);
      uint32_t stage = 0, phase = 0, acc = 0, accphase = 0;
      int cur_g = -1, n_sw = -1;
      UnitIter it;
      it.seek(ur, 0);
      for (long long i = 0; i < ur.count; ++i, it.advance(ur, 1)) {
        const int g = it.g;
        const long long u = it.u;
        if constexpr (BSW) {
          if (g != cur_g) {
            if (n_sw >= 0 && elect_one_sync()) umma_commit(&bempty[n_sw & 1]);   // previous group's operand buffer is free once its MMAs retire
            __syncwarp();
            ++n_sw;
            mbar_wait(&bfull[n_sw & 1], (n_sw >> 1) & 1);
            b_addr = smem_u32(sB + (n_sw & 1) * Cfg::B_BYTES);
            cur_g = g;
          }
        }
        int ntiles = 1;
        if constexpr (!PACKED) {
          long long row0;
          int nrows;
          resolve_page(p, item_page(p, u, g), row0, nrows);
          ntiles = (nrows + kTileRows - 1) / kTileRows;
        }
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&tempty[acc], accphase ^ 1);
          mbar_wait(&full[stage], phase);
          const uint32_t a_addr = smem_u32(sA + stage * kTileBytes);
          const uint32_t d_addr = tbase + acc * N;
          // one descriptor per operand and tile; the eight K-steps only move the start-address field (16-byte units)
          const uint64_t ad0 = umma_desc_k_sw128(a_addr), bd0 = umma_desc_k_sw128(b_addr);
          if (elect_one_sync()) {
            if constexpr (PACKED) mbar_arrive(&tfull[acc]);   // releases what this thread acquired: the tile's side data
            tc_fence_after_sync();
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = ad0 + static_cast<uint64_t>((kh * kHalfBytes + kk * 32) >> 4);
                const uint64_t bd = bd0 + static_cast<uint64_t>((kh * (N * 128) + kk * 32) >> 4);
                umma_f16_ss(d_addr, ad, bd, idesc, (kh | kk) != 0);
              }
            }
            umma_commit(&empty[stage]);   // smem stage may be refilled once these MMAs retire
            umma_commit(&tfull[acc]);     // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++acc == ACC) { acc = 0; accphase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue (4 warps = 128 TMEM lanes per group)
    // Single-query kernels: one group (LARGE) or two groups alternating tiles (PACKED). Sub-query kernels
    // (QS < QP): four groups, group k owns accumulator columns [32k, 32k+32) of EVERY tile — one 32-column query
    // (QS == 32) or 32 single-column queries (QS == 1) — so that each SM sub-partition has four warps to hide the
    // TMEM-load / shuffle latency chains behind each other.
    const int grp = (warp - 2) >> 2;         // epilogue group
    const int ew = (warp - 2) & 3;           // warp within the group
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int trow = lg * 32 + lane;         // tile row (= TMEM lane) owned by this thread
    const int et = ew * 32 + lane;           // 0..127 thread id within the group
    const int col0 = MULTI ? grp * 32 : 0;   // first accumulator column of this group
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t bar_id = 1 + grp;
    constexpr float kInvLo = 1.0f / kLoScale;
    const int seq0 = MULTI ? 0 : grp, seq_step = MULTI ? 1 : EPI_GROUPS;
    (void)seq0; (void)seq_step; (void)et;

    if constexpr (!PACKED) {
      uint32_t stage = 0, phase = 0, acc = 0, accphase = 0, par = 0;
      int q_valid = 0, q_valid2 = 0, qv_g = -1;
      (void)q_valid2;
      UnitIter it;
      it.seek(ur, 0);
      for (long long i = 0; i < ur.count; ++i, it.advance(ur, 1)) {
        const int g = it.g;
        const long long u = it.u;
        // per-group query width: a global load, so it is refreshed only when the group changes (never on the tile path)
        if (g != qv_g) {
          if constexpr (MULTI) q_valid = (QS == 32 && grp < p.n_sub) ? __ldg(p.q_valid_arr + grp) : 0;
          else q_valid = p.q_valid_arr ? __ldg(p.q_valid_arr + g) : p.q_valid;
          if constexpr (NBLK == 2) q_valid2 = (grp + 4 < p.n_sub) ? __ldg(p.q_valid_arr + grp + 4) : 0;
          qv_g = g;
        }
        long long row0;
        int nrows;
        const bool ok = resolve_page(p, item_page(p, u, g), row0, nrows);
        float run[NBLK == 2 ? 1 : QR];
        // NBLK == 2: 64 running maxima per thread (two query blocks) would not fit the register file of an 18-warp CTA, so
        // this first, approximate pass keeps them as packed fp16 pairs: max commutes with the (monotone) rounding, i.e. each
        // per-token maximum carries one fp16 rounding (<= 2^-12 for |cos| <= 1), which the host's exactness guard accounts for
        uint32_t runA[16], runB[16];   // (dead, and removed by the compiler, in the other kernels)
#pragma unroll
        for (int q = 0; q < (NBLK == 2 ? 1 : QR); ++q) run[q] = -INFINITY;
#pragma unroll
        for (int q = 0; q < 16; ++q) runA[q] = runB[q] = kH2NegInf;
        const int ncol = !MULTI ? QE : (QS == 32 ? ((q_valid + 7) & ~7) : ((min(32, max(0, p.n_sub - col0)) + 7) & ~7));
        const int ncol2 = (q_valid2 + 3) & ~3;
        const int nmax = NBLK == 2 ? max((q_valid + 3) & ~3, ncol2) : 0;   // first pass: columns read per block (steps of 4)
        (void)nmax;
        for (int t0 = 0; t0 < nrows; t0 += kTileRows) {
          const int valid = min(kTileRows, nrows - t0);
          mbar_wait(&tfull[acc], accphase);
          tc_fence_after_sync();
          const uint32_t ta = lane_addr + acc * N + col0;
          float scale = 1.0f;
          if (use_scale) {
            mbar_wait(&full[stage], phase);  // acquire the TMA-written scale rows
            scale = sScale[stage * kScaleStride + trow + (sMis[stage * 4] & 3)];
          }
          // 16 columns (hi and lo) per TMEM round trip; 8 when the group's width is not a multiple of 16. Sub-query
          // kernels are bound by the TMEM read port (64 B/clk: a 128x256 fp32 accumulator takes 2048 clk to drain), so
          // they read only the columns that hold real query rows (ncol, a multiple of 8).
          constexpr int LW = (QE % 16 == 0 && !MULTI) ? 16 : 8;
          // (a warp whose 32 rows all lie beyond the page's last row — three of four on the 6-row tail tile of a 1030-row
          // page — skips the tile: the batched kernels are bound by the TMEM read port and the epilogue's instruction issue)
          const bool warp_live = lg * 32 < valid;
          if constexpr (NBLK == 2) {
            // two independent query blocks, both plain fp16: block A at columns [col0, col0+32), block B at QP + the same.
            // Only the columns that hold query rows are read (in steps of 4); both blocks read the wider of the two widths —
            // the operand rows beyond a query's length are zero and its sum ignores them.
            if (warp_live && nmax > 0) {
              const bool live = trow < valid;
              const uint32_t tg = lane_addr + acc * N + grp * 64;   // the group's 64 interleaved columns
              // tokens [8k, 8k+8) of both queries per round trip; the first four when the queries end there
              if (nmax >= 8) first_pass_cols<0, 16>(tg, live, scale, runA, runB);
              else first_pass_cols<0, 8>(tg, live, scale, runA, runB);
              if (nmax >= 16) first_pass_cols<8, 16>(tg, live, scale, runA, runB);
              else if (nmax > 8) first_pass_cols<8, 8>(tg, live, scale, runA, runB);
              if (nmax >= 24) first_pass_cols<16, 16>(tg, live, scale, runA, runB);
              else if (nmax > 16) first_pass_cols<16, 8>(tg, live, scale, runA, runB);
              if (nmax >= 32) first_pass_cols<24, 16>(tg, live, scale, runA, runB);
              else if (nmax > 24) first_pass_cols<24, 8>(tg, live, scale, runA, runB);
            }
          } else if (warp_live) {
#pragma unroll
          for (int c = 0; c < QE; c += LW) {
            if (MULTI && c >= ncol) break;
            uint32_t hi[LW], lo[LW];
            if constexpr (LW == 16) tmem_ld_x16(ta + c, hi);
            else tmem_ld_x8(ta + c, hi);
            if (!p.hi_only) {
              if constexpr (LW == 16) tmem_ld_x16(ta + QP + c, lo);
              else tmem_ld_x8(ta + QP + c, lo);
            } else {
#pragma unroll
              for (int j = 0; j < LW; ++j) lo[j] = 0u;
            }
            tmem_ld_wait();
            if (trow < valid) {
#pragma unroll
              for (int j = 0; j < LW; ++j) {
                const float s = fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale;
                run[c + j] = fmaxf(run[c + j], s);
              }
            }
          }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&tempty[acc]);
            mbar_arrive(&empty[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++acc == ACC) { acc = 0; accphase ^= 1; }
        }
        // page done: max across the 128 rows owned by the group's threads, then sum over q
        float* red = sRed + par * 4 * RS + col0;
        if constexpr (NBLK == 2) {
          warp_transpose_max_h2x16(runA, lane);
          red[ew * RS + lane] = h2_column_max(runA[0], lane);
        } else {
#pragma unroll
        for (int gq = 0; gq < QG; ++gq) {
          warp_transpose_max<QW>(run + gq * 32, lane);
          constexpr int rep = 32 / QW;  // lanes holding the same q
          if ((lane & (rep - 1)) == 0) red[ew * RS + gq * 32 + (lane / rep)] = run[gq * 32];
        }
        }
        named_bar_sync(bar_id, 128);
        if (ew == 0) {
          if constexpr (MULTI) {
            const float m = fmaxf(fmaxf(red[lane], red[RS + lane]), fmaxf(red[2 * RS + lane], red[3 * RS + lane]));
            const float dead = (ok && nrows > 0) ? 0.0f : -INFINITY;
            if constexpr (QS == 32) {
              float sum = lane < q_valid ? m : 0.0f;
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
              if (lane == 0 && grp < p.n_sub) emit(grp, u, u * p.tile_stride, sum + dead);
            } else {
              if (col0 + lane < p.n_sub) emit(col0 + lane, u, u * p.tile_stride, m + dead);
            }
          } else {
            float sum = 0.0f;
#pragma unroll
            for (int gq = 0; gq < QG; ++gq) {
              const int q = gq * 32 + lane;
              if (q < QR) {
                const float m = fmaxf(fmaxf(red[q], red[RS + q]), fmaxf(red[2 * RS + q], red[3 * RS + q]));
                if (q < q_valid) sum += m;
              }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0) p.scores[g * p.n_items + u] = (ok && nrows > 0) ? sum : -INFINITY;
          }
        }
        par ^= 1;
        if constexpr (NBLK == 2) {
          // second query block: same reduction through the other scratch buffer (the barrier above separates its writes
          // from the reads of the previous page that used it)
          float* red2 = sRed + par * 4 * RS + col0;
          warp_transpose_max_h2x16(runB, lane);
          red2[ew * RS + lane] = h2_column_max(runB[0], lane);
          named_bar_sync(bar_id, 128);
          if (ew == 0) {
            const float m = fmaxf(fmaxf(red2[lane], red2[RS + lane]), fmaxf(red2[2 * RS + lane], red2[3 * RS + lane]));
            float sum = lane < q_valid2 ? m : 0.0f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0 && grp + 4 < p.n_sub) emit(grp + 4, u, u * p.tile_stride, sum + ((ok && nrows > 0) ? 0.0f : -INFINITY));
          }
          par ^= 1;
        }
      }
    } else {
      // ---------------- PACKED: several pages per tile
      int* seg = sSeg + grp * 3 * kTileRows;
      float* sc = sSc + grp * QE * Cfg::SC_PITCH;
      int seg_n = 0, seg_item = 0, seg_rb = 0, seg_re = 0;   // general path: this thread's entry of the tile's segment table
      bool seg_ready = false;
      int q_valid = 0, qv_g = -1;
      const bool row_fast = MULTI && QS == 1 && p.f_thr != nullptr && p.shfl_rows == 1 && !p.slot_mode && p.pad_rows == 0 &&
                            p.tile_stride == 1 && p.fixed_rows == 1;
      (void)row_fast;
      UnitIter it;
      it.seek(ur, seq0);
      for (long long seq = seq0; seq < ur.count; seq += seq_step, it.advance(ur, seq_step)) {
        const int g = it.g;
        const long long u = it.u;
        // per-group query width: a global load, so it is refreshed only when the group changes (never on the tile path)
        if (g != qv_g) {
          if constexpr (MULTI) q_valid = (QS == 32 && grp < p.n_sub) ? __ldg(p.q_valid_arr + grp) : 0;
          else q_valid = p.q_valid_arr ? __ldg(p.q_valid_arr + g) : p.q_valid;
          qv_g = g;
        }
        float* const scores_g = p.scores + g * p.n_items;
        // side data of this tile: ring entry `me` (written by the producer / TMA, acquired through tfull); the row stage
        // itself is never touched here
        const uint32_t seq32 = static_cast<uint32_t>(seq);
        const int me = static_cast<int>(seq32 % META);
        const uint32_t acc = seq32 % ACC, accphase = (seq32 / ACC) & 1u;
        const uint32_t ta = lane_addr + acc * N + col0;
        if constexpr (MULTI && QS == 1) {
          // ---- one page per tile row (global_pooling) under the top-k prefilter: the dense stage-1 scan of a batch of
          // pooled queries. Every thread owns one page and 32 queries; the kernel is bound by the epilogue's instruction
          // stream (4 warps per scheduler, ~1000 instructions per warp and tile on the generic path), so this path is
          // written for the minimum: two x16 TMEM round trips, FFMA+FMUL+FSETP+LOP per score, one warp-wide OR of the
          // hit masks, and the (rare: ~2 per warp and tile) survivors appended from a warp-uniform loop.
          if (row_fast) {
            const long long page = u * kTileRows + trow;
            mbar_wait(&tfull[acc], accphase);
            tc_fence_after_sync();
            float scale = 1.0f;
            if (use_scale) scale = sScale[me * kScaleStride + trow + (sMis[me * 4] & 3)];
            float v[32];
            uint32_t pass = 0u;
            const float4* thr4 = reinterpret_cast<const float4*>(sThr + col0);   // +inf for absent queries
#pragma unroll
            for (int c = 0; c < 32; c += 16) {
              uint32_t hi[16], lo[16];
              tmem_ld_x16(ta + c, hi);
              tmem_ld_x16(ta + QP + c, lo);
              tmem_ld_wait();
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 t = thr4[c / 4 + j4];
                const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const int j = j4 * 4 + jj;
                  v[c + j] = fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale;
                  if (v[c + j] > tt[jj]) pass |= 1u << (c + j);
                }
              }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (page >= p.n_pages) pass = 0u;
            unsigned any = __reduce_or_sync(0xffffffffu, pass);
            while (any) {   // warp-uniform
              const int i = __ffs(any) - 1;
              any &= any - 1;
              float val = v[0];
#pragma unroll
              for (int j = 1; j < 32; ++j) val = (i == j) ? v[j] : val;
              if ((pass >> i) & 1u) emit(col0 + i, page, page, val);
            }
            continue;
          }
        }
        bool fast = false;
        if constexpr (QE <= 32) fast = p.shfl_rows > 0;
        if (fast) {
          if constexpr (QE <= 32) {
            // fast path: page j of the tile owns tile rows [j*SR, j*SR + nr)
            const int SR = p.shfl_rows;
            const int slot = trow / SR, rin = trow - slot * SR;
            const long long item = u * (kTileRows / SR) + slot;                   // score column (compact when sampling)
            const long long page = u * p.tile_stride * (kTileRows / SR) + slot;   // dense: the page scored
            int nr = 0;
            bool item_ok;
            if (!p.slot_mode) {
              item_ok = page < p.n_pages;
              nr = (item_ok && page_allowed(p, page)) ? (p.pad_rows > 0 ? p.pad_rows : SR) : 0;
            } else {
              item_ok = item < p.n_items;
            }
            mbar_wait(&tfull[acc], accphase);
            tc_fence_after_sync();
            if (p.slot_mode) nr = item_ok ? (sMis[me * 4 + slot] >> 2) : 0;   // slot == 32-row slot here (SR == 32)
            float scale = 1.0f;
            if (use_scale) {
              if (p.pad_rows > 0) {
                // padded slots: the scale rows are the tile's real rows back to back (slot j starts at j * fixed_rows)
                scale = rin < nr ? sScale[me * kScaleStride + slot * p.pad_rows + rin + (sMis[me * 4] & 3)] : 0.0f;
              } else {
                const int sslot = trow / p.slot_rows;
                scale = sScale[me * kScaleStride + sslot * (p.slot_rows + 32) + (trow - sslot * p.slot_rows) +
                               (sMis[me * 4 + sslot] & 3)];
              }
            }
            float v[QE];
            const bool live = rin < nr;
            constexpr int LW = (QE >= 16 && !MULTI) ? 16 : 8;
            const int ncol = !MULTI ? QE : (QS == 32 ? ((q_valid + 7) & ~7) : ((min(32, max(0, p.n_sub - col0)) + 7) & ~7));
            // one page per tile row (SR == 1, e.g. global_pooling) in a single-column-query kernel: every use of v[] below
            // is guarded by item_ok or by a compare against +inf thresholds, so the per-element "row is live" select is
            // dropped (the dense batched global stage is bound by epilogue instruction issue, not by HBM)
            const bool row_pages = MULTI && QS == 1 && SR == 1;
#pragma unroll
            for (int c = 0; c < QE; c += LW) {
              if (MULTI && c >= ncol) {   // columns without query rows are not read from TMEM (see the LARGE path)
#pragma unroll
                for (int j = 0; j < LW; ++j) v[c + j] = -INFINITY;
                continue;
              }
              uint32_t hi[LW], lo[LW];
              if constexpr (LW == 16) {
                tmem_ld_x16(ta + c, hi);
                tmem_ld_x16(ta + QP + c, lo);
              } else {
                tmem_ld_x8(ta + c, hi);
                tmem_ld_x8(ta + QP + c, lo);
              }
              tmem_ld_wait();
              if (row_pages) {
#pragma unroll
                for (int j = 0; j < LW; ++j) v[c + j] = fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale;
              } else {
#pragma unroll
                for (int j = 0; j < LW; ++j)
                  v[c + j] = live ? fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale : -INFINITY;
              }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if constexpr (MULTI && QS == 1) {
              // 32 single-column queries: after the segmented butterfly lane rin of a slot holds the column maxima
              // of columns col0 + rin*cf .. +cf (cf = 32 / SR)
              switch (SR) {
                case 32: slot_colmax<32, 32>(v, lane); break;
                case 16: slot_colmax<32, 16>(v, lane); break;
                case 8: slot_colmax<32, 8>(v, lane); break;
                case 4: slot_colmax<32, 4>(v, lane); break;
                case 2: slot_colmax<32, 2>(v, lane); break;
                default: break;
              }
              const int cf = 32 / SR;
              const int qb = col0 + rin * cf;
              if (p.f_thr) {
                // prefilter: one compare + one warp vote per score column; survivors are rare (~2 per warp and tile), so
                // the append path runs under a warp-uniform branch (sThr is +inf for absent queries)
                if (SR == 1) {
                  const float4* thr4 = reinterpret_cast<const float4*>(sThr + col0);   // all lanes: queries col0..col0+31
#pragma unroll
                  for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 t = thr4[i4];
                    const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const bool hit = v[i4 * 4 + j] > tt[j] && item_ok;
                      if (__any_sync(0xffffffffu, hit)) {
                        if (hit) emit(col0 + i4 * 4 + j, item, page, v[i4 * 4 + j]);
                      }
                    }
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    if (i < cf) {   // warp-uniform
                      const bool hit = v[i] > sThr[qb + i] && item_ok;
                      if (__any_sync(0xffffffffu, hit)) {
                        if (hit) emit(qb + i, item, page, v[i]);
                      }
                    }
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < cf && item_ok && qb + i < p.n_sub) p.scores[(qb + i) * p.score_stride + item] = v[i];
              }
            } else {
              float sum;
              switch (SR) {
                case 32: sum = slot_maxsim<QE, 32>(v, lane, q_valid); break;
                case 16: sum = slot_maxsim<QE, 16>(v, lane, q_valid); break;
                case 8: sum = slot_maxsim<QE, 8>(v, lane, q_valid); break;
                case 4: sum = slot_maxsim<QE, 4>(v, lane, q_valid); break;
                case 2: sum = slot_maxsim<QE, 2>(v, lane, q_valid); break;
                default: sum = slot_maxsim<QE, 1>(v, lane, q_valid); break;
              }
              if constexpr (MULTI) {
                if (rin == 0 && item_ok && grp < p.n_sub) emit(grp, item, page, nr > 0 ? sum : -INFINITY);
              } else {
                if (rin == 0 && item_ok) scores_g[item] = nr > 0 ? sum : -INFINITY;
              }
            }
          }
          continue;
        }
        // general path
        // 1. segment table of this tile (item, first tile row, end tile row) -> smem. The table of the group's NEXT
        //    tile is loaded into registers now (tile -> page -> offsets is a chain of dependent global loads).
        if (!seg_ready) {
          packed_segment(p, g, u, et, seg_n, seg_item, seg_rb, seg_re);
          seg_ready = true;
        }
        const int nseg = seg_n;
        if (et < nseg) {
          seg[et] = seg_item;
          seg[kTileRows + et] = seg_rb;
          seg[2 * kTileRows + et] = seg_re;
        }
        if (seq + seq_step < ur.count) {
          int gn;
          long long un;
          ur.decode(seq + seq_step, gn, un);
          packed_segment(p, gn, un, et, seg_n, seg_item, seg_rb, seg_re);
        }
        // 2. scaled scores of my row -> transposed smem buffer sc[q][row]
        mbar_wait(&tfull[acc], accphase);
        tc_fence_after_sync();
        float scale = 1.0f;
        if (use_scale) {
          const int slot = trow / p.slot_rows;
          scale = sScale[me * kScaleStride + slot * (p.slot_rows + 32) + (trow - slot * p.slot_rows) +
                         (sMis[me * 4 + slot] & 3)];
        }
#pragma unroll
        for (int c = 0; c < QE; c += 8) {
          uint32_t hi[8], lo[8];
          tmem_ld_x8(ta + c, hi);
          tmem_ld_x8(ta + QP + c, lo);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sc[(c + j) * Cfg::SC_PITCH + trow] =
                fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale;
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        named_bar_sync(bar_id, 128);
        // 3. segmented max: a group of QW lanes handles one segment; lane -> q (+32 per extra query group);
        //    rows are read 4 at a time (16-byte aligned, conflict-free at pitch 132)
        constexpr int SEGS_PER_WARP = 32 / QW;
        const int sub = lane / QW, ql = lane % QW;
        for (int s0 = ew * SEGS_PER_WARP; s0 < nseg; s0 += 4 * SEGS_PER_WARP) {
          const int sg = s0 + sub;
          float sum = 0.0f;
          int item = -1;
          bool nonempty = false;
          if (sg < nseg) {
            item = seg[sg];
            const int rb = seg[kTileRows + sg], re = seg[2 * kTileRows + sg];
            nonempty = re > rb;
#pragma unroll
            for (int gq = 0; gq < QG; ++gq) {
              const int q = gq * 32 + ql;
              const float* row = sc + q * Cfg::SC_PITCH;
              float m0 = -INFINITY, m1 = -INFINITY;
              for (int r4 = rb & ~3; r4 < re; r4 += 4) {
                const float4 x = *reinterpret_cast<const float4*>(row + r4);
                const float a0 = (r4 >= rb) ? x.x : -INFINITY;
                const float a1 = (r4 + 1 >= rb && r4 + 1 < re) ? x.y : -INFINITY;
                const float a2 = (r4 + 2 >= rb && r4 + 2 < re) ? x.z : -INFINITY;
                const float a3 = (r4 + 3 < re) ? x.w : -INFINITY;
                m0 = fmaxf(m0, fmaxf(a0, a1));
                m1 = fmaxf(m1, fmaxf(a2, a3));
              }
              if constexpr (MULTI && QS == 1) {
                // QW == 32: sg is warp-uniform; every column is a query
                if (col0 + q < p.n_sub) emit(col0 + q, item, item, nonempty ? fmaxf(m0, m1) : -INFINITY);
              } else {
                if (q < q_valid) sum += fmaxf(m0, m1);
              }
            }
          }
          if constexpr (MULTI && QS == 1) continue;
#pragma unroll
          for (int off = QW / 2; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          if constexpr (MULTI) {
            if (sg < nseg && ql == 0 && grp < p.n_sub) emit(grp, item, item, nonempty ? sum : -INFINITY);
          } else {
            if (sg < nseg && ql == 0) scores_g[item] = nonempty ? sum : -INFINITY;
          }
        }
        named_bar_sync(bar_id, 128);   // the group's buffers are reused by its next tile
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if constexpr (QS < QP) {
    if (stage_on) {   // flush this CTA's staged survivors: one range reservation per query
      for (int j = warp; j < p.n_sub; j += NTHREADS / 32) {
        const int n = min(sStageCnt[j], stage_cap);
        int base = 0;
        if (lane == 0 && n > 0) base = atomicAdd(p.f_cnt + j, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < n; i += 32)
          if (base + i < p.f_cap) p.f_keys[static_cast<long long>(j) * p.f_cap + base + i] = sStage[j * stage_cap + i];
      }
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace vrag
