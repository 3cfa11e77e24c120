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

namespace vrag {

constexpr int kDim = 128;            // embedding dim (qdrant_indexer.py:133)
constexpr int kTileRows = 128;       // UMMA M
constexpr int kTileBytes = kTileRows * kDim * 2;  // 32 KB of fp16 per stage
constexpr int kHalfBytes = kTileBytes / 2;        // one K-half (64 fp16 = 128 B per row)
constexpr int kBoxRowsSmall = 32;    // partial tiles are fetched in 32-row boxes
constexpr int kScanThreads = 192;    // warp0 TMA, warp1 MMA, warps 2..5 epilogue
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
  long long n_tiles;          // PACKED: number of tiles (work units)
};

template <int QP>
struct ScanCfg {
  static constexpr int N = 2 * QP;                         // UMMA N (hi | lo)
  static constexpr int ACC = (N <= 128) ? 4 : 2;           // TMEM accumulator stages
  static constexpr int TMEM_COLS = (ACC * N < 32) ? 32 : ACC * N;
  static constexpr int B_BYTES = N * kDim * 2;
  static constexpr int SC_PITCH = kTileRows + 1;           // PACKED: transposed score buffer pitch
  static constexpr int SC_BUFS = (QP <= 32) ? 2 : 1;       // PACKED: score buffers (1 => extra barrier per tile)
  static constexpr int MISC_BYTES = 8192;                  // barriers, reduce scratch, segment tables
  static constexpr int stages(bool packed) {
    const int budget = 227 * 1024 - 1024 /*align slack*/ - B_BYTES - MISC_BYTES -
                       (packed ? SC_BUFS * QP * SC_PITCH * 4 : 0);
    const int s = budget / (kTileBytes + kScaleStride * 4);
    return s > 6 ? 6 : s;
  }
  static constexpr size_t smem_bytes(bool packed) {
    return 1024 + size_t(stages(packed)) * (kTileBytes + kScaleStride * 4) + B_BYTES + MISC_BYTES +
           (packed ? SC_BUFS * QP * SC_PITCH * 4 : 0);
  }
};

// Resolve work item -> (first row, row count). Returns false when the item is not a page of this shard.
__device__ __forceinline__ bool resolve_page(const ScanParams& p, long long page, long long& row0, int& nrows) {
  if (page < 0 || page >= p.n_pages) {
    row0 = 0;
    nrows = 0;
    return false;
  }
  if (p.fixed_rows > 0) {
    row0 = page * p.fixed_rows;
    nrows = static_cast<int>(p.fixed_rows);
  } else {
    const long long a = __ldg(p.offsets + page), b = __ldg(p.offsets + page + 1);
    row0 = a;
    nrows = static_cast<int>(b - a);
  }
  return true;
}
__device__ __forceinline__ long long item_page(const ScanParams& p, long long item) {
  return p.cand ? (__ldg(p.cand + item) - p.cand_base) : item;
}

// Fetch `nrows` (1..128) document rows starting at global row `row` into tile rows [dst_row, dst_row+nrows)
// of stage buffer `a` (+ their inv_norm scales at sc[(row & 3) ...]). dst_row is a multiple of 32.
// Returns the bytes the mbarrier must expect.
__device__ __forceinline__ uint32_t issue_rows(uint8_t* a, float* sc, uint64_t* bar, const CUtensorMap* tm128,
                                               const CUtensorMap* tm32, const CUtensorMap* ts128,
                                               const CUtensorMap* ts32, long long row, int nrows, int dst_row,
                                               bool use_scale) {
  const int32_t r32 = static_cast<int32_t>(row);
  if (dst_row == 0 && nrows > 96) {
    tma_load_2d(a, tm128, bar, 0, r32);
    tma_load_2d(a + kHalfBytes, tm128, bar, 64, r32);
    if (use_scale) tma_load_1d(sc, ts128, bar, r32 & ~3);
    return kTileBytes + (use_scale ? kScaleBoxBig * 4 : 0);
  }
  const int nb = (nrows + kBoxRowsSmall - 1) / kBoxRowsSmall;
  for (int j = 0; j < nb; ++j) {
    const int32_t r = r32 + j * kBoxRowsSmall;
    const int d = dst_row + j * kBoxRowsSmall;
    tma_load_2d(a + d * 128, tm32, bar, 0, r);
    tma_load_2d(a + kHalfBytes + d * 128, tm32, bar, 64, r);
    // consecutive scale boxes of one slot overlap by 4 floats in smem; both write identical values there
    if (use_scale) tma_load_1d(sc + j * kBoxRowsSmall, ts32, bar, r & ~3);
  }
  return nb * (kBoxRowsSmall * kDim * 2 + (use_scale ? kScaleBoxSmall * 4 : 0));
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

template <int QP, bool PACKED>
__global__ void __launch_bounds__(kScanThreads, 1)
maxsim_scan_kernel(const __grid_constant__ CUtensorMap tm_rows128, const __grid_constant__ CUtensorMap tm_rows32,
                   const __grid_constant__ CUtensorMap tm_scale128, const __grid_constant__ CUtensorMap tm_scale32,
                   const ScanParams p) {
  using Cfg = ScanCfg<QP>;
  constexpr int N = Cfg::N;
  constexpr int ACC = Cfg::ACC;
  constexpr int STAGES = Cfg::stages(PACKED);
  constexpr int QG = (QP + 31) / 32;           // 32-wide query groups
  constexpr int QW = QP < 32 ? QP : 32;        // queries per group
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kTileBytes;
  float* sScale = reinterpret_cast<float*>(sB + Cfg::B_BYTES);
  float* sSc = sScale + STAGES * kScaleStride;                       // PACKED: [SC_BUFS][QP][SC_PITCH]
  uint8_t* misc = reinterpret_cast<uint8_t*>(sSc + (PACKED ? Cfg::SC_BUFS * QP * Cfg::SC_PITCH : 0));
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC);
  int* sMis = reinterpret_cast<int*>(misc + 256);                 // [STAGES][4] scale misalignment per slot
  float* sRed = reinterpret_cast<float*>(misc + 512);             // LARGE: [2][4][QP]
  int* sSeg = reinterpret_cast<int*>(misc + 512);                 // PACKED: [2][3][128] ints (item, begin, end)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool use_scale = p.use_scale != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows128);
    tma_prefetch_desc(&tm_rows32);
    tma_prefetch_desc(&tm_scale128);
    tma_prefetch_desc(&tm_scale32);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1 + 4);   // tcgen05.commit + one arrive per epilogue warp (scale rows consumed)
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  {  // query operand image: global -> smem (already in the swizzled UMMA layout)
    const uint4* src = reinterpret_cast<const uint4*>(p.qimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = threadIdx.x; i < Cfg::B_BYTES / 16; i += kScanThreads) dst[i] = __ldg(src + i);
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Work units of this CTA: u = blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long n_units = PACKED ? p.n_tiles : p.n_items;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        if constexpr (!PACKED) {
          long long row0;
          int nrows;
          resolve_page(p, item_page(p, u), row0, nrows);
          for (int t0 = 0; t0 < nrows; t0 += kTileRows) {
            const int rows = min(kTileRows, nrows - t0);
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a = sA + stage * kTileBytes;
            float* sc = sScale + stage * kScaleStride;
            // copies first, then one arrive.expect_tx with the exact byte count: the phase cannot complete
            // before the arrive, and the tx-count may go transiently negative.
            const uint32_t bytes = issue_rows(a, sc, &full[stage], &tm_rows128, &tm_rows32, &tm_scale128,
                                              &tm_scale32, row0 + t0, rows, 0, use_scale);
            sMis[stage * 4] = static_cast<int>((row0 + t0) & 3);
            mbar_arrive_expect_tx(&full[stage], bytes);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a = sA + stage * kTileBytes;
          float* sc = sScale + stage * kScaleStride;
          if (p.cand == nullptr) {
            // dense: pages [pg0, pg1) are contiguous rows -> one fetch
            long long pg0, pg1;
            if (p.fixed_rows > 0) {
              pg0 = u * p.pages_per_tile;
              pg1 = min(pg0 + p.pages_per_tile, p.n_pages);
            } else {
              pg0 = __ldg(p.tile_page0 + u);
              pg1 = __ldg(p.tile_page0 + u + 1);
            }
            long long r0, r1;
            int tmp;
            resolve_page(p, pg0, r0, tmp);
            resolve_page(p, pg1 - 1, r1, tmp);
            const int rows = static_cast<int>(r1 + tmp - r0);
            uint32_t bytes = 0;
            if (rows > 0)
              bytes = issue_rows(a, sc, &full[stage], &tm_rows128, &tm_rows32, &tm_scale128, &tm_scale32, r0, rows,
                                 0, use_scale);
            sMis[stage * 4] = static_cast<int>(r0 & 3);
            mbar_arrive_expect_tx(&full[stage], bytes);
          } else {
            // candidates: each occupies its own slot of slot_rows rows
            const int per_tile = kTileRows / p.slot_rows;
            const long long i0 = u * per_tile;
            const int cnt = static_cast<int>(min(static_cast<long long>(per_tile), p.n_items - i0));
            uint32_t bytes = 0;
            for (int j = 0; j < cnt; ++j) {
              long long r0;
              int nr;
              resolve_page(p, item_page(p, i0 + j), r0, nr);
              if (nr > 0)
                bytes += issue_rows(a, sc + j * (p.slot_rows + 32), &full[stage], &tm_rows128, &tm_rows32,
                                    &tm_scale128, &tm_scale32, r0, nr, j * p.slot_rows, use_scale);
              sMis[stage * 4 + j] = static_cast<int>(r0 & 3);
            }
            for (int j = cnt; j < 4; ++j) sMis[stage * 4 + j] = 0;   // unused slots: rows are never read back
            mbar_arrive_expect_tx(&full[stage], bytes);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileRows, N);
      const uint32_t b_addr = smem_u32(sB);
      uint32_t stage = 0, phase = 0, acc = 0, accphase = 0;
      for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        int ntiles = 1;
        if constexpr (!PACKED) {
          long long row0;
          int nrows;
          resolve_page(p, item_page(p, u), row0, nrows);
          ntiles = (nrows + kTileRows - 1) / kTileRows;
        }
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(&tempty[acc], accphase ^ 1);
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + stage * kTileBytes);
          const uint32_t d_addr = tmem_base + acc * N;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = umma_desc_k_sw128(a_addr + kh * kHalfBytes + kk * 32);
              const uint64_t bd = umma_desc_k_sw128(b_addr + kh * (N * 128) + kk * 32);
              umma_f16_ss(d_addr, ad, bd, idesc, (kh | kk) != 0);
            }
          }
          umma_commit(&empty[stage]);   // smem stage may be refilled once these MMAs retire
          umma_commit(&tfull[acc]);     // accumulator ready for the epilogue
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++acc == ACC) { acc = 0; accphase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue (4 warps = 128 TMEM lanes)
    const int ew = warp - 2;                 // 0..3
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int trow = lg * 32 + lane;         // tile row (= TMEM lane) owned by this thread
    const int et = ew * 32 + lane;           // 0..127 epilogue thread id
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    uint32_t stage = 0, phase = 0, acc = 0, accphase = 0, par = 0;
    constexpr float kInvLo = 1.0f / kLoScale;

    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      if constexpr (!PACKED) {
        long long row0;
        int nrows;
        const bool ok = resolve_page(p, item_page(p, u), row0, nrows);
        float run[QP];
#pragma unroll
        for (int q = 0; q < QP; ++q) run[q] = -INFINITY;
        for (int t0 = 0; t0 < nrows; t0 += kTileRows) {
          const int valid = min(kTileRows, nrows - t0);
          mbar_wait(&tfull[acc], accphase);
          tc_fence_after_sync();
          const uint32_t ta = lane_addr + acc * N;
          float scale = 1.0f;
          if (use_scale) {
            mbar_wait(&full[stage], phase);  // acquire the TMA-written scale rows
            scale = sScale[stage * kScaleStride + trow + sMis[stage * 4]];
          }
#pragma unroll
          for (int c = 0; c < QP; c += 8) {
            uint32_t hi[8], lo[8];
            tmem_ld_x8(ta + c, hi);
            tmem_ld_x8(ta + QP + c, lo);
            tmem_ld_wait();
            if (trow < valid) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float s = fmaf(__uint_as_float(lo[j]), kInvLo, __uint_as_float(hi[j])) * scale;
                run[c + j] = fmaxf(run[c + j], s);
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
        // page done: max across the 128 rows owned by the epilogue threads, then sum over q
        float* red = sRed + par * 4 * QP;
#pragma unroll
        for (int g = 0; g < QG; ++g) {
          warp_transpose_max<QW>(run + g * 32, lane);
          constexpr int rep = 32 / QW;  // lanes holding the same q
          if ((lane & (rep - 1)) == 0) red[ew * QP + g * 32 + (lane / rep)] = run[g * 32];
        }
        named_bar_sync(1, 128);
        if (ew == 0) {
          float sum = 0.0f;
#pragma unroll
          for (int g = 0; g < QG; ++g) {
            const int q = g * 32 + lane;
            if (q < QP) {
              const float m = fmaxf(fmaxf(red[q], red[QP + q]), fmaxf(red[2 * QP + q], red[3 * QP + q]));
              if (q < p.q_valid) sum += m;
            }
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          if (lane == 0) p.scores[u] = (ok && nrows > 0) ? sum : -INFINITY;
        }
        par ^= 1;
      } else {
        // ---------------- PACKED: several pages per tile
        // 1. segment table of this tile (item, first tile row, end tile row) -> smem (double buffered)
        int* seg = sSeg + par * 3 * kTileRows;
        int nseg;
        if (p.cand == nullptr) {
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
            resolve_page(p, pg0, rbase, tmp);
            resolve_page(p, pg0 + et, r0, nr);
            seg[et] = static_cast<int>(pg0 + et);   // dense: item == page (n_items == n_pages < 2^31 per shard)
            seg[kTileRows + et] = static_cast<int>(r0 - rbase);
            seg[2 * kTileRows + et] = static_cast<int>(r0 - rbase) + nr;
          }
        } else {
          const int per_tile = kTileRows / p.slot_rows;
          const long long i0 = u * per_tile;
          nseg = static_cast<int>(min(static_cast<long long>(per_tile), p.n_items - i0));
          if (et < nseg) {
            long long r0;
            int nr;
            resolve_page(p, item_page(p, i0 + et), r0, nr);
            seg[et] = static_cast<int>(i0 + et);
            seg[kTileRows + et] = et * p.slot_rows;
            seg[2 * kTileRows + et] = et * p.slot_rows + nr;   // nr == 0 -> empty segment -> -inf
          }
        }
        // 2. scaled scores of my row -> transposed smem buffer sc[q][row]
        float* sc = sSc + (Cfg::SC_BUFS == 2 ? par : 0) * QP * Cfg::SC_PITCH;
        mbar_wait(&tfull[acc], accphase);
        tc_fence_after_sync();
        const uint32_t ta = lane_addr + acc * N;
        float scale = 1.0f;
        if (use_scale) {
          mbar_wait(&full[stage], phase);
          const int slot = trow / p.slot_rows;
          scale = sScale[stage * kScaleStride + slot * (p.slot_rows + 32) + (trow - slot * p.slot_rows) +
                         sMis[stage * 4 + slot]];
        }
#pragma unroll
        for (int c = 0; c < QP; c += 8) {
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
        if (lane == 0) {
          mbar_arrive(&tempty[acc]);
          mbar_arrive(&empty[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++acc == ACC) { acc = 0; accphase ^= 1; }
        named_bar_sync(1, 128);
        // 3. segmented max: a group of QW lanes handles one segment; lane -> q (+32 per extra group)
        constexpr int SEGS_PER_WARP = 32 / QW;
        const int sub = lane / QW, ql = lane % QW;
        for (int s0 = ew * SEGS_PER_WARP; s0 < nseg; s0 += 4 * SEGS_PER_WARP) {
          const int s = s0 + sub;
          float sum = 0.0f;
          int item = -1;
          bool nonempty = false;
          if (s < nseg) {
            item = seg[s];
            const int rb = seg[kTileRows + s], re = seg[2 * kTileRows + s];
            nonempty = re > rb;
#pragma unroll
            for (int g = 0; g < QG; ++g) {
              const int q = g * 32 + ql;
              float m = -INFINITY;
              for (int r = rb; r < re; ++r) m = fmaxf(m, sc[q * Cfg::SC_PITCH + r]);
              if (q < p.q_valid) sum += m;
            }
          }
#pragma unroll
          for (int off = QW / 2; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
          if (s < nseg && ql == 0) p.scores[item] = nonempty ? sum : -INFINITY;
        }
        if constexpr (Cfg::SC_BUFS == 1) named_bar_sync(1, 128);  // single score buffer: drain before reuse
        par ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace vrag
