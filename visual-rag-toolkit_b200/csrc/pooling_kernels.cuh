// Indexing-time pooling kernels (CUDA cores, HBM-bound): the arithmetic of visual_rag/embedding/pooling.py.
//
// Every pooled row is an fp32 mean (or weighted mean) of input rows accumulated IN THE SAME ORDER as numpy's
// axis-0 reductions (row after row, round-to-nearest adds, one IEEE division at the end), so results match the
// reference to the last bit in fp32 before the final store-dtype rounding.  One warp owns one output row; a
// lane owns 4 of the 128 dims (8-byte fp16 / 16-byte fp32 loads, a full row per warp-load, coalesced).
//
//   pool_tokens_kernel : token-level kinds, "mean over a contiguous range of input rows":
//        TILE_MEAN (p1, pooling.py:35-98), ADAPTIVE_ROWS (p2/p3, pooling.py:101-185; row means staged in smem,
//        then overlapping bins), SEQ_CHUNKS (visual_embedder.py:824-835), COLSMOL_EXPERIMENTAL (p4,
//        pooling.py:188-232), GLOBAL_MEAN (p8, pooling.py:439-465), LEGACY_CONV (p5, pooling.py:235-286).
//   pool_rows_kernel   : neighbourhood kinds over a page's (few) pooled rows staged in smem, several derived
//        outputs per pass: SMOOTH (p6, pooling.py:289-375), TILE_4N (p7, pooling.py:378-436), LEGACY_CONV,
//        GLOBAL_MEAN (visual_embedder.py:837-840).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrag {

enum PoolKind : int {
  kPoolTileMean = 0,
  kPoolRowMean = 1,            // alias of ADAPTIVE_ROWS with target == grid_h (host side)
  kPoolAdaptiveRows = 2,
  kPoolColsmolExperimental = 3,
  kPoolLegacyConv = 4,
  kPoolSmooth = 5,
  kPoolTile4n = 6,
  kPoolGlobalMean = 7,
  kPoolSeqChunks = 8,
};

constexpr int kPoolMaxWeights = 16;
constexpr int kPoolMaxSpecs = 8;

struct PoolSpecDev {
  int kind;
  int ppt;          // TILE_MEAN / COLSMOL_EXPERIMENTAL
  int grid_h, grid_w;
  int target_rows, clamp_to_h;
  int num_tiles;
  int window;
  int n_weights;
  float weights[kPoolMaxWeights];
  int n_rows, n_cols, has_global, include_self;
  int via_f16;
  int keep_f32;     // derived specs read this spec's fp32 rows before the store-dtype rounding
  // output
  void* out;
  int out_f32;
  const long long* out_off;   // [n_pages+1] device, or nullptr when out_fixed > 0
  long long out_fixed;
  // TILE_MEAN only: optional fused colsmol_experimental_pooling output (same tile means for the first tiles,
  // then the raw rows of the last tile), so the tokens are read once for both stores
  void* out2;
  const long long* out2_off;
  long long out2_fixed;
};

struct PoolInput {
  const void* in;
  int in_f32;
  const long long* in_off;    // [n_pages+1] device, or nullptr when in_fixed > 0
  long long in_fixed;
  long long n_pages;
  const int* grid_hw;         // [n_pages][2] device (ADAPTIVE_ROWS / TILE_4N per-page grids) or nullptr
  int row_skip, row_count;    // token-level pass: pool rows [row_skip, row_skip + row_count) of every page (count <= 0: rest)
};

__device__ __forceinline__ float4 pool_load4(const void* base, int f32, long long row, int lane) {
  if (f32) return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + row * 128) + lane);
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + row * 128) + lane);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float pool_round_f16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ void pool_store4(void* out, int f32, long long row, int lane, float4 v, int via_f16) {
  if (f32) {
    if (via_f16) v = make_float4(pool_round_f16(v.x), pool_round_f16(v.y), pool_round_f16(v.z), pool_round_f16(v.w));
    reinterpret_cast<float4*>(static_cast<float*>(out) + row * 128)[lane] = v;
  } else {
    uint2 o;
    *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(v.z, v.w);
    reinterpret_cast<uint2*>(static_cast<__half*>(out) + row * 128)[lane] = o;
  }
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {   // no FMA contraction, plain rn adds
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_div(float4 a, float d) {
  return make_float4(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d), __fdiv_rn(a.w, d));
}

// mean of global rows [lo, hi) of `base`, accumulated row after row. 16 row loads are kept in flight per warp
// (4 KB): with ~40 resident warps per SM that is what it takes to cover HBM latency at full bandwidth.
struct PoolRaw {
  uint4 w;   // fp32: one float4; fp16: .x/.y hold 4 halves
};
__device__ __forceinline__ PoolRaw pool_load_raw(const void* base, int f32, long long row, int lane) {
  PoolRaw r;
  if (f32) {
    r.w = __ldg(reinterpret_cast<const uint4*>(static_cast<const float*>(base) + row * 128) + lane);
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + row * 128) + lane);
    r.w = make_uint4(v.x, v.y, 0u, 0u);
  }
  return r;
}
__device__ __forceinline__ float4 pool_raw_to_f4(const PoolRaw& r, int f32) {
  if (f32) return make_float4(__uint_as_float(r.w.x), __uint_as_float(r.w.y), __uint_as_float(r.w.z), __uint_as_float(r.w.w));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.w.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.w.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
// acc += rows [r, r + N) of `base`, in row order; the N loads are in flight together
template <int N>
__device__ __forceinline__ void pool_add_rows(float4& acc, const void* base, int f32, long long r, int lane) {
  PoolRaw v[N];
#pragma unroll
  for (int j = 0; j < N; ++j) v[j] = pool_load_raw(base, f32, r + j, lane);
#pragma unroll
  for (int j = 0; j < N; ++j) acc = f4_add(acc, pool_raw_to_f4(v[j], f32));
}
__device__ __forceinline__ float4 range_mean_global(const void* base, int f32, long long lo, long long hi, int lane) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long long r = lo;
  for (; r + 16 <= hi; r += 16) pool_add_rows<16>(acc, base, f32, r, lane);
  // the remaining 1..15 rows in at most four batches (8, 4, 2, 1) instead of one HBM latency per row: ColQwen2.5 grids
  // have 17..31 tokens per grid row. No predicated work: the token pass is as much issue- as latency-bound.
  const int left = static_cast<int>(hi - r);
  if (left & 8) { pool_add_rows<8>(acc, base, f32, r, lane); r += 8; }
  if (left & 4) { pool_add_rows<4>(acc, base, f32, r, lane); r += 4; }
  if (left & 2) { pool_add_rows<2>(acc, base, f32, r, lane); r += 2; }
  if (left & 1) pool_add_rows<1>(acc, base, f32, r, lane);
  return f4_div(acc, static_cast<float>(hi - lo));
}
__device__ __forceinline__ float4 range_mean_smem(const float* rows, int lo, int hi, int lane) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = lo; r < hi; ++r) acc = f4_add(acc, reinterpret_cast<const float4*>(rows + r * 128)[lane]);
  return f4_div(acc, static_cast<float>(hi - lo));
}

// np.linspace(0, h, r+1) bin i -> [floor(e_i), ceil(e_{i+1})) clamped (pooling.py:176-182), fp64 like numpy
__device__ __forceinline__ void adaptive_bin(int h, int r, int i, int& lo, int& hi) {
  const double step = static_cast<double>(h) / static_cast<double>(r);
  const double e0 = static_cast<double>(i) * step;
  const double e1 = (i + 1 == r) ? static_cast<double>(h) : static_cast<double>(i + 1) * step;
  lo = static_cast<int>(floor(e0));
  hi = static_cast<int>(ceil(e1));
  lo = max(0, min(lo, h - 1));
  hi = max(lo + 1, min(hi, h));
}

__device__ __forceinline__ void pool_page_rows(const long long* off, long long fixed, long long page, long long& r0, int& n) {
  if (fixed > 0) {
    r0 = page * fixed;
    n = static_cast<int>(fixed);
  } else {
    r0 = off[page];
    n = static_cast<int>(off[page + 1] - r0);
  }
}

// ------------------------------------------------------------------------------------------------ row-level
struct PoolRowsArgs {
  PoolInput in;
  int n_specs;
  PoolSpecDev specs[kPoolMaxSpecs];
};

// Output row `o` of a row-level spec over the n rows of one page staged in smem as fp32 (nr x nc: TILE_4N grid).
__device__ __forceinline__ float4 pool_row_level(const PoolSpecDev& s, const float* rows, int n, int o, int lane, int nr,
                                                 int nc) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.kind == kPoolSmooth) {
    if (s.window == 1 || n == 1) {
      v = reinterpret_cast<const float4*>(rows + o * 128)[lane];
    } else {
      // pooling.py:358-374: taps i-left+t, products rounded to fp32, summed in tap order; weight mass in fp64
      const int left = s.window / 2;
      double mass = 0.0;
      for (int t = 0; t < s.window; ++t) {
        const int j = o - left + t;
        if (j < 0 || j >= n) continue;
        const float w = s.weights[t];
        const float4 x = reinterpret_cast<const float4*>(rows + j * 128)[lane];
        v = f4_add(v, make_float4(__fmul_rn(w, x.x), __fmul_rn(w, x.y), __fmul_rn(w, x.z), __fmul_rn(w, x.w)));
        mass += static_cast<double>(w);
      }
      // interior rows: the taps sum to exactly 1.0f and x / 1 == x, so the (slow, IEEE) division is skipped
      const float fm = static_cast<float>(mass);
      if (mass > 0.0) { if (fm != 1.0f) v = f4_div(v, fm); }
      else v = reinterpret_cast<const float4*>(rows + o * 128)[lane];
    }
  } else if (s.kind == kPoolTile4n) {
    const int g = nr * nc;
    if (o >= g) {
      v = reinterpret_cast<const float4*>(rows + g * 128)[lane];   // global tile copied through
    } else {
      const int r = o / nc, c = o - r * nc;
      int cnt = 0;
      if (s.include_self) { v = f4_add(v, reinterpret_cast<const float4*>(rows + o * 128)[lane]); ++cnt; }
      if (r > 0) { v = f4_add(v, reinterpret_cast<const float4*>(rows + (o - nc) * 128)[lane]); ++cnt; }
      if (r + 1 < nr) { v = f4_add(v, reinterpret_cast<const float4*>(rows + (o + nc) * 128)[lane]); ++cnt; }
      if (c > 0) { v = f4_add(v, reinterpret_cast<const float4*>(rows + (o - 1) * 128)[lane]); ++cnt; }
      if (c + 1 < nc) { v = f4_add(v, reinterpret_cast<const float4*>(rows + (o + 1) * 128)[lane]); ++cnt; }
      v = f4_div(v, static_cast<float>(cnt));
    }
  } else if (s.kind == kPoolLegacyConv) {
    const int r = s.window / 2;
    int lo, hi;
    if (s.window == 1 || n == 1) { lo = o; hi = o + 1; }
    else if (s.window == 3 && n == 2) { lo = (o == 2) ? 1 : 0; hi = (o == 0) ? 1 : 2; }
    else { lo = max(0, o - 2 * r); hi = min(n - 1, o) + 1; }
    v = (hi - lo == 1) ? reinterpret_cast<const float4*>(rows + lo * 128)[lane] : range_mean_smem(rows, lo, hi, lane);
  } else if (s.kind == kPoolGlobalMean) {
    if (n > 0) v = range_mean_smem(rows, 0, n, lane);   // empty page -> zeros (visual_embedder.py:838-839)
  }
  return v;
}

// All row-level specs of `a` for one page whose n rows are staged in `rows` (smem, fp32). kWarps warps cooperate.
template <int kWarps>
__device__ __forceinline__ void pool_derive_page(const PoolRowsArgs& a, const float* rows, int n, long long page, int warp,
                                                 int lane) {
  for (int si = 0; si < a.n_specs; ++si) {
    const PoolSpecDev& s = a.specs[si];
    long long o0;
    int n_out;
    pool_page_rows(s.out_off, s.out_fixed, page, o0, n_out);
    int nr = s.n_rows, nc = s.n_cols;
    if (s.kind == kPoolTile4n && a.in.grid_hw) {
      nr = a.in.grid_hw[2 * page];
      nc = a.in.grid_hw[2 * page + 1];
    }
    for (int o = warp; o < n_out; o += kWarps)
      pool_store4(s.out, s.out_f32, o0 + o, lane, pool_row_level(s, rows, n, o, lane, nr, nc), s.via_f16);
  }
}

// block = 128 threads (4 warps); one page per block iteration; dynamic smem = max_rows*128 floats.
__global__ void __launch_bounds__(128) pool_rows_kernel(const PoolRowsArgs a) {
  extern __shared__ float pool_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = 4;
  for (long long page = blockIdx.x; page < a.in.n_pages; page += gridDim.x) {
    long long r0;
    int n;
    pool_page_rows(a.in.in_off, a.in.in_fixed, page, r0, n);
    for (int r = warp; r < n; r += kWarps)
      reinterpret_cast<float4*>(pool_smem + r * 128)[lane] = pool_load4(a.in.in, a.in.in_f32, r0 + r, lane);
    __syncthreads();
    pool_derive_page<kWarps>(a, pool_smem, n, page, warp, lane);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ token-level
// block = 256 threads (8 warps); one page per block iteration; dynamic smem = grid_h_max*128 floats (ADAPTIVE).
// `d`: row-level specs DERIVED from this spec's output (d.n_specs may be 0): the page's pooled rows, rounded to the
// store dtype exactly as a reader of the stored rows would see them (the reference's dtype chain), are
// kept in smem (at float offset d_off) and the derived stores are written in the same pass — the pooled store is
// never re-read from HBM.
// IN_F32: the input dtype at compile time (a run-time flag leaves both conversion paths in the instruction stream as
// predicated-off instructions, and ncu shows the pass issue-bound as much as latency-bound: 10.7 G -> 8.5 G instructions
// for 400k ColQwen2.5 pages).
template <bool DERIVE, bool IN_F32>
__global__ void __launch_bounds__(256, 4) pool_tokens_kernel(const PoolInput in, const PoolSpecDev s, const PoolRowsArgs d,
                                                             const int d_off) {
  extern __shared__ float pool_smem[];
  constexpr int kInF32 = IN_F32 ? 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = 8;
  float* drows = pool_smem + d_off;
  auto keep = [&](int o, float4 v) {   // stage output row o for the derived specs, as stored
    if constexpr (DERIVE) {
      if ((!s.out_f32 || s.via_f16) && !s.keep_f32) v = make_float4(pool_round_f16(v.x), pool_round_f16(v.y), pool_round_f16(v.z), pool_round_f16(v.w));
      reinterpret_cast<float4*>(drows + o * 128)[lane] = v;
    }
  };
  for (long long page = blockIdx.x; page < in.n_pages; page += gridDim.x) {
    long long r0, o0;
    int t, n_out;
    pool_page_rows(in.in_off, in.in_fixed, page, r0, t);
    {   // token window (the page's visual tokens)
      const int skip = min(max(in.row_skip, 0), t);
      r0 += skip;
      t -= skip;
      if (in.row_count > 0) t = min(t, in.row_count);
    }
    pool_page_rows(s.out_off, s.out_fixed, page, o0, n_out);
    if (s.kind == kPoolAdaptiveRows) {
      int gh = s.grid_h, gw = s.grid_w;
      if (in.grid_hw) {
        gh = in.grid_hw[2 * page];
        gw = in.grid_hw[2 * page + 1];
      }
      // 1) row means of the gh x gw grid -> smem (pooling.py:162-163)
      for (int h = warp; h < gh; h += kWarps) {
        const float4 m = range_mean_global(in.in, kInF32, r0 + static_cast<long long>(h) * gw,
                                           r0 + static_cast<long long>(h + 1) * gw, lane);
        reinterpret_cast<float4*>(pool_smem + h * 128)[lane] = m;
      }
      __syncthreads();
      // 2) adaptive bins over the row means (identity when n_out == gh, repeat when gh == 1)
      for (int o = warp; o < n_out; o += kWarps) {
        float4 v;
        if (n_out == gh) {
          v = reinterpret_cast<const float4*>(pool_smem + o * 128)[lane];
        } else if (gh == 1) {
          v = reinterpret_cast<const float4*>(pool_smem)[lane];
        } else {
          int lo, hi;
          adaptive_bin(gh, n_out, o, lo, hi);
          v = range_mean_smem(pool_smem, lo, hi, lane);
        }
        pool_store4(s.out, s.out_f32, o0 + o, lane, v, 0);
        keep(o, v);
      }
      __syncthreads();
      if constexpr (DERIVE) {
        pool_derive_page<kWarps>(d, drows, n_out, page, warp, lane);
        __syncthreads();
      }
      continue;
    }
    for (int o = warp; o < n_out; o += kWarps) {
      long long lo = 0, hi = 0;
      switch (s.kind) {
        case kPoolTileMean:
          lo = static_cast<long long>(o) * s.ppt;
          hi = min(lo + static_cast<long long>(s.ppt), static_cast<long long>(t));
          break;
        case kPoolSeqChunks: {
          int a, b;
          adaptive_bin(t, n_out, o, a, b);
          lo = a;
          hi = b;
          break;
        }
        case kPoolColsmolExperimental: {
          // n_out = (nt-1) + rows of the last tile; outputs < nt-1 are tile means, the rest raw rows
          int nt = s.num_tiles > 0 ? s.num_tiles : (t + s.ppt - 1) / s.ppt;
          if (static_cast<long long>(nt - 1) * s.ppt >= t) nt = (t + s.ppt - 1) / s.ppt;   // pooling.py:209-219
          const int head = nt - 1;
          if (o < head) {
            lo = static_cast<long long>(o) * s.ppt;
            hi = lo + s.ppt;
          } else {
            lo = static_cast<long long>(head) * s.ppt + (o - head);
            hi = lo + 1;
          }
          break;
        }
        case kPoolGlobalMean:
          lo = 0;
          hi = t;
          break;
        case kPoolLegacyConv: {
          const int r = s.window / 2;
          if (s.window == 1 || t == 1) {
            lo = o;
            hi = o + 1;
          } else if (s.window == 3 && t == 2) {     // pooling.py:277-279
            lo = (o == 2) ? 1 : 0;
            hi = (o == 0) ? 1 : 2;
          } else {
            lo = max(0, o - 2 * r);
            hi = min(t - 1, o) + 1;
          }
          break;
        }
        default:
          break;
      }
      float4 v;
      if (hi - lo == 1) {
        v = pool_load4(in.in, kInF32, r0 + lo, lane);   // x/1 == x: raw rows pass through exactly
      } else {
        v = range_mean_global(in.in, kInF32, r0 + lo, r0 + hi, lane);
      }
      pool_store4(s.out, s.out_f32, o0 + o, lane, v, s.via_f16);
      keep(o, v);
      if (s.kind == kPoolTileMean && s.out2 && o < n_out - 1) {
        long long e0;
        int e_n;
        pool_page_rows(s.out2_off, s.out2_fixed, page, e0, e_n);
        pool_store4(s.out2, s.out_f32, e0 + o, lane, v, 0);
      }
    }
    if (s.kind == kPoolTileMean && s.out2 && n_out > 0) {
      // raw rows of the last tile (pooling.py:221-232); n_out == ceil(t / ppt) == num_tiles here
      long long e0;
      int e_n;
      pool_page_rows(s.out2_off, s.out2_fixed, page, e0, e_n);
      const int head = n_out - 1;
      for (int o = head + warp; o < e_n; o += kWarps)
        pool_store4(s.out2, s.out_f32, e0 + o, lane,
                    pool_load4(in.in, kInF32, r0 + static_cast<long long>(head) * s.ppt + (o - head), lane), 0);
    }
    if constexpr (DERIVE) {
      __syncthreads();
      pool_derive_page<kWarps>(d, drows, n_out, page, warp, lane);
      __syncthreads();
    }
  }
}

}  // namespace vrag
