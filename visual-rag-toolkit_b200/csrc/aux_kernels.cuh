// Small CUDA-core kernels around the scan: query operand preparation, per-row inverse norms,
// synthetic corpus generation and exact top-k selection.
#pragma once
#include "ptx.cuh"

namespace vrag {

// ------------------------------------------------------------------------------------------------
// Query preparation.  Mirrors the query side of compute_maxsim_score (pooling.py:495-503):
//   qhat = q / (||q||_2 + 1e-8) row-wise (skipped when normalize == 0),
// and the mean-pooled query of the pooled_query_* stage-1 modes (two_stage.py:142,148,154):
//   qbar = q.mean(axis=0)  (un-normalised mean; cosine normalisation afterwards).
// Output: the UMMA B-operand image (2*QP rows x 128 fp16, K-major, 128B swizzle): rows [0,QP) = fp16(qhat),
// rows [QP,2QP) = fp16((qhat - fp16(qhat)) * 2^11); rows >= Q_eff are zero.
// grid = QP blocks (one per operand row pair), 128 threads (one per dim).
__global__ void query_prep_kernel(const float* __restrict__ q, int Q, int pool, int normalize, int QP,
                                  uint8_t* __restrict__ qimg) {
  const int r = blockIdx.x;
  const int d = threadIdx.x;
  const int q_eff = pool ? 1 : Q;
  __shared__ float wsum[4];
  float x = 0.0f;
  if (r < q_eff) {
    if (pool) {
      float s = 0.0f;
      for (int i = 0; i < Q; ++i) s += q[i * 128 + d];  // numpy reduces axis 0 row by row
      x = s / static_cast<float>(Q);
    } else {
      x = q[r * 128 + d];
    }
  }
  float ss = x * x;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((d & 31) == 0) wsum[d >> 5] = ss;
  __syncthreads();
  if (normalize) {
    const float nrm = sqrtf((wsum[0] + wsum[1]) + (wsum[2] + wsum[3]));
    x = x / (nrm + 1e-8f);
  }
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
  const uint32_t rows = 2u * QP;
  *reinterpret_cast<__half*>(qimg + sw128_offset(rows, r, d)) = (r < q_eff) ? hi : __float2half_rn(0.0f);
  *reinterpret_cast<__half*>(qimg + sw128_offset(rows, QP + r, d)) = (r < q_eff) ? lo : __float2half_rn(0.0f);
}

// Batched variant: one operand image per query (blockIdx.y = query b), rows [q_begin[b], q_end[b]) of `q`.
// Same arithmetic as query_prep_kernel. q_valid_out[b] = effective query rows (1 when pooled).
__global__ void query_prep_batch_kernel(const float* __restrict__ q, const int* __restrict__ q_begin,
                                        const int* __restrict__ q_end, int pool, int normalize, int QP,
                                        uint8_t* __restrict__ qimg, long long qimg_stride,
                                        int* __restrict__ q_valid_out) {
  const int b = blockIdx.y;
  const int r = blockIdx.x;
  const int d = threadIdx.x;
  const int r0 = q_begin[b];
  const int Q = q_end[b] - r0;
  const int q_eff = pool ? 1 : Q;
  const float* qb = q + static_cast<long long>(r0) * 128;
  __shared__ float wsum[4];
  float x = 0.0f;
  if (r < q_eff) {
    if (pool) {
      float s = 0.0f;
      for (int i = 0; i < Q; ++i) s += qb[i * 128 + d];
      x = s / static_cast<float>(Q);
    } else {
      x = qb[r * 128 + d];
    }
  }
  float ss = x * x;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((d & 31) == 0) wsum[d >> 5] = ss;
  __syncthreads();
  if (normalize) {
    const float nrm = sqrtf((wsum[0] + wsum[1]) + (wsum[2] + wsum[3]));
    x = x / (nrm + 1e-8f);
  }
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
  const uint32_t rows = 2u * QP;
  uint8_t* img = qimg + b * qimg_stride;
  *reinterpret_cast<__half*>(img + sw128_offset(rows, r, d)) = (r < q_eff) ? hi : __float2half_rn(0.0f);
  *reinterpret_cast<__half*>(img + sw128_offset(rows, QP + r, d)) = (r < q_eff) ? lo : __float2half_rn(0.0f);
  if (r == 0 && d == 0) q_valid_out[b] = q_eff;
}

// Dense batched scans: G = QP/QS queries share one operand image (blockIdx.y = image); image row r holds row
// r % QS of query (image*G + r / QS), hi half at row r and lo half at row QP + r; absent rows/queries are zero.
// two_block (QS == 32): the image holds 2 * QP / QS queries as PLAIN fp16 — blockIdx.z = 0 writes queries 0..3, blockIdx.z = 1
// queries 4..7; queries g and g + 4 share the 64 rows [64g, 64g + 64), interleaved in quads of token rows — and eps_out[b] accumulates the bound the host's exactness guard
// needs for query b: sum over its rows of ||qhat - fp16(qhat)||_2 (the most a unit-norm document row can move that
// row's cosine) + 2^-12 (the fp16 rounding of the row's running maximum in the first-pass epilogue).
__global__ void query_prep_group_kernel(const float* __restrict__ q, const int* __restrict__ q_begin,
                                        const int* __restrict__ q_end, int nq, int pool, int normalize, int QP, int QS,
                                        uint8_t* __restrict__ qimg, long long qimg_stride,
                                        int* __restrict__ q_valid_out, int two_block = 0, float* __restrict__ eps_out = nullptr) {
  const int r = blockIdx.x;
  const int d = threadIdx.x;
  const int G = (two_block ? 2 : 1) * (QP / QS);
  const int b = blockIdx.y * G + blockIdx.z * (QP / QS) + r / QS;
  const int t = r % QS;
  __shared__ float wsum[4];
  float x = 0.0f;
  int q_eff = 0;
  if (b < nq) {
    const int r0 = q_begin[b];
    const int Q = q_end[b] - r0;
    q_eff = pool ? 1 : Q;
    const float* qb = q + static_cast<long long>(r0) * 128;
    if (t < q_eff) {
      if (pool) {
        float s = 0.0f;
        for (int i = 0; i < Q; ++i) s += qb[i * 128 + d];
        x = s / static_cast<float>(Q);
      } else {
        x = qb[t * 128 + d];
      }
    }
    if (t == 0 && d == 0) q_valid_out[b] = q_eff;
  }
  float ss = x * x;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((d & 31) == 0) wsum[d >> 5] = ss;
  __syncthreads();
  if (normalize) {
    const float nrm = sqrtf((wsum[0] + wsum[1]) + (wsum[2] + wsum[3]));
    x = x / (nrm + 1e-8f);
  }
  const bool live = b < nq && t < q_eff;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
  const uint32_t rows = 2u * QP;
  uint8_t* img = qimg + blockIdx.y * qimg_stride;
  if (two_block) {
    // the two query blocks of an epilogue group (queries g and g + 4: blockIdx.z = 0 / 1) are interleaved in quads of token
    // columns inside the group's 64 accumulator columns: A[0:4] B[0:4] A[4:8] B[4:8] ... (maxsim_scan.cuh, first_pass_cols)
    const uint32_t row2 = (r / QS) * 64u + (t / 4) * 8u + blockIdx.z * 4u + (t % 4);
    *reinterpret_cast<__half*>(img + sw128_offset(rows, row2, d)) = live ? hi : __float2half_rn(0.0f);
    // rounding error of this row, for the guard
    float e = live ? (x - __half2float(hi)) : 0.0f;
    e *= e;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    __syncthreads();   // wsum was read above by every thread
    if ((d & 31) == 0) wsum[d >> 5] = e;
    __syncthreads();
    if (d == 0 && live && eps_out) atomicAdd(eps_out + b, sqrtf((wsum[0] + wsum[1]) + (wsum[2] + wsum[3])) + 2.44140625e-4f);
    return;
  }
  *reinterpret_cast<__half*>(img + sw128_offset(rows, r, d)) = live ? hi : __float2half_rn(0.0f);
  *reinterpret_cast<__half*>(img + sw128_offset(rows, QP + r, d)) = live ? lo : __float2half_rn(0.0f);
}

// ------------------------------------------------------------------------------------------------
// inv_norm[row] = 1 / (||row||_2 + 1e-8) over fp16 rows, fp32 math (doc side of pooling.py:500).
// 16 lanes per row, 16-byte loads.
__global__ void inv_norm_kernel(const __half* __restrict__ rows, long long n_rows, float* __restrict__ inv) {
  const long long gt = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long row = gt >> 4;
  const int sub = static_cast<int>(gt & 15);
  float ss = 0.0f;
  if (row < n_rows) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rows + row * 128) + sub);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      ss = fmaf(f.x, f.x, ss);
      ss = fmaf(f.y, f.y, ss);
    }
  }
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (row < n_rows && sub == 0) inv[row] = 1.0f / (sqrtf(ss) + 1e-8f);
}

// fp32 -> fp16 conversion of an embedding matrix (the store-dtype cast of qdrant_indexer.py:423-441).
__global__ void f32_to_f16_kernel(const float* __restrict__ src, long long n, __half* __restrict__ dst) {
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<__half2*>(dst + i) = a;
    *reinterpret_cast<__half2*>(dst + i + 2) = b;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2half_rn(src[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// Synthetic corpus: counter-based gaussian rows, L2-normalised in fp32, rounded to fp16 (what a Col*
// model + fp16 store produces), plus inv_norm of the ROUNDED row (what the oracle will see).
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void synth_rows_kernel(__half* __restrict__ rows, float* __restrict__ inv, long long row_begin,
                                  long long n_rows, uint64_t seed, long long row_seed_base) {
  const long long gt = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long lrow = gt >> 4;
  const int sub = static_cast<int>(gt & 15);
  const bool ok = lrow < n_rows;
  const long long row = row_begin + lrow;
  float v[8];
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t h = mix64(seed ^ mix64(static_cast<uint64_t>(row_seed_base + row) * 64ull + sub * 4 + j));
    const float u1 = (static_cast<float>(static_cast<uint32_t>(h >> 40)) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (static_cast<float>(static_cast<uint32_t>(h) >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307f * u2, &s, &c);
    v[2 * j] = rad * c;
    v[2 * j + 1] = rad * s;
    ss += v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1];
  }
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float invn = 1.0f / (sqrtf(ss) + 1e-8f);
  uint4 out;
  __half2* h2 = reinterpret_cast<__half2*>(&out);
  float ss16 = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h2[j] = __floats2half2_rn(v[2 * j] * invn, v[2 * j + 1] * invn);
    const float2 f = __half22float2(h2[j]);
    ss16 = fmaf(f.x, f.x, ss16);
    ss16 = fmaf(f.y, f.y, ss16);
  }
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) ss16 += __shfl_xor_sync(0xffffffffu, ss16, off);
  if (ok) {
    *(reinterpret_cast<uint4*>(rows + row * 128) + sub) = out;
    if (sub == 0) inv[row] = 1.0f / (sqrtf(ss16) + 1e-8f);
  }
}

// ------------------------------------------------------------------------------------------------ peer-memory windows
// (the collectives that use them are at the end of this file; the top-k kernels below send / receive through them too)
constexpr int kP2PMaxRanks = 16;
struct P2PWindow {
  uint8_t* win[kP2PMaxRanks];   // window base of every rank as mapped into THIS process (win[me] = the local one)
  unsigned long long cap;       // bytes per (slot, source) region
  int me, R;
  unsigned epoch;
  unsigned* ctr;                // local: blocks that finished their stores (last one publishes the flags)
  unsigned long long timeout_ns;   // watchdog of the flag wait (VRAG_P2P_TIMEOUT_S, default 120 s)
  unsigned long long ll_off;    // byte offset of the LL area inside a window: ll[2 slots][R sources][ll_cap bytes]
  unsigned long long ll_cap;    // bytes per (slot, source) LL region = 2 x the largest LL payload
};
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned* p2p_flag(const P2PWindow& w, int rank, int slot, int src) {
  return reinterpret_cast<unsigned*>(w.win[rank] + 2ull * w.R * w.cap) + slot * kP2PMaxRanks + src;
}
__device__ __forceinline__ unsigned p2p_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- LL lines (flag-in-data, the idea of NCCL's low-latency protocol): small messages travel as 16-byte lines
// (w0, epoch, w1, epoch) written with ONE vector store each — an 8-byte half is delivered atomically, so a reader that sees
// `epoch` in both halves has the payload; no fence, no separate flag, no counter: the latency of an exchange is one NVLink
// store plus the poll. The lines of collective `epoch` live in slot epoch & 1 of a window area of their own (a payload word
// of a bulk message must never be mistaken for a flag); stale lines carry older epochs.
__device__ __forceinline__ uint8_t* ll_region(const P2PWindow& w, int rank, int src) {
  return w.win[rank] + w.ll_off + (static_cast<unsigned long long>(w.epoch & 1u) * w.R + src) * w.ll_cap;
}
__device__ __forceinline__ void ll_store(uint8_t* line, uint32_t w0, uint32_t w1, uint32_t e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(w0), "r"(e), "r"(w1), "r"(e) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const uint8_t* line, uint32_t e, uint32_t& w0, uint32_t& w1) {
  uint32_t f0, f1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(f0), "=r"(w1), "=r"(f1) : "l"(line) : "memory");
  return f0 == e && f1 == e;
}
// pair p (8 payload bytes) of this rank's message -> region [me] of EVERY rank's window (the local one included)
__device__ __forceinline__ void ll_send_pair(const P2PWindow& w, long long p, uint32_t w0, uint32_t w1) {
  for (int r = 0; r < w.R; ++r) ll_store(ll_region(w, r, w.me) + p * 16, w0, w1, w.epoch);
}
// pair p of rank src's message, from the local window; spins until it has arrived
__device__ __forceinline__ void ll_recv_pair(const P2PWindow& w, int src, long long p, uint32_t& w0, uint32_t& w1) {
  const uint8_t* line = ll_region(w, w.me, src) + p * 16;
  if (ll_try_load(line, w.epoch, w0, w1)) return;
  const unsigned long long t0 = p2p_now_ns();
  unsigned spins = 0;
  while (!ll_try_load(line, w.epoch, w0, w1)) {
    // a peer that is merely late (host-side skew between the ranks) is waited for; only a peer that died or ranks whose
    // collective calls diverged end in the watchdog, which fails the call instead of hanging the GPU
    if ((++spins & 1023u) == 0 && p2p_now_ns() - t0 > w.timeout_ns) {
      printf("vrag: peer exchange watchdog (LL) rank=%d waits for rank=%d epoch=%u\n", w.me, src, w.epoch);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Exact top-k.  Keys are 64-bit: (order-preserving bits of the fp32 score) << 32 | (0xFFFFFFFF - item
// index), so "largest key first" = "highest score first, ties -> lower item index", which is what
// Python's stable list.sort(reverse=True) over pages in index order yields (quick_test.py:165,
// two_stage.py:424).
//
// One level = one kernel: every block takes a chunk of CHUNK keys into shared memory and leaves the chunk's
// K2 = pow2ceil(k) largest keys sorted descending at the front:
//   1. bitonic-sort runs of K2 keys, even runs descending / odd runs ascending;
//   2. repeat: element-wise max of each (descending, ascending) run pair = the pair's K2 largest keys as a
//      bitonic sequence; drop the other half; bitonic-merge the surviving runs (alternating directions again).
// Shared-memory traffic, which bounds this kernel, shrinks geometrically after step 1, and step 1 is
// O(log^2 K2) instead of O(log^2 CHUNK) stages.  Levels repeat until one chunk is left; that last level
// writes (score, id) pairs.
constexpr int kTopkMaxK = 4096;

// Packed top-k entry exchanged between shards (== vrag_hit_t, include/vrag_b200.h): one 16-byte record per result so
// that a shard's whole local list travels as ONE message. aux bit 0: the list came from a top-k estimate that missed
// (sampled threshold / prefilter) and the search must be repeated exactly; id < 0 marks padding.
struct __align__(16) Hit {
  float score;
  uint32_t aux;
  long long id;
};
constexpr uint32_t kHitMiss = 1u;
// entry idx of a packed list as two LL lines
__device__ __forceinline__ void ll_send_hit(const P2PWindow& w, long long idx, const Hit& h) {
  ll_send_pair(w, 2 * idx, __float_as_uint(h.score), h.aux);
  ll_send_pair(w, 2 * idx + 1, static_cast<uint32_t>(static_cast<unsigned long long>(h.id) & 0xFFFFFFFFull),
               static_cast<uint32_t>(static_cast<unsigned long long>(h.id) >> 32));
}
__device__ __forceinline__ Hit ll_recv_hit(const P2PWindow& w, int src, long long idx) {
  uint32_t a, b, lo, hi;
  ll_recv_pair(w, src, 2 * idx, a, b);
  ll_recv_pair(w, src, 2 * idx + 1, lo, hi);
  Hit h;
  h.score = __uint_as_float(a);
  h.aux = b;
  h.id = static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo);
  return h;
}

__device__ __forceinline__ uint32_t score_to_ord(float f) {
  if (f != f) return 0u;  // NaN sorts last
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_score(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct TopkArgs {
  const float* scores;        // level 0 input (nullptr on merge levels)
  const unsigned long long* keys_in;  // merge-level input
  long long n;                // number of input elements
  int k;
  int k2;                     // pow2ceil(k), <= CHUNK
  unsigned long long* keys_out;  // [n_chunks][k] (nullptr on the final level)
  // final level outputs
  float* out_scores;          // [k]
  long long* out_ids;         // [k]
  int* out_pos;               // [k] item index of each result (optional)
  int* out_count;             // number of valid results (optional)
  const long long* ids;       // item index -> id (nullptr: id = id_base + index)
  long long id_base;
  long long n_total;          // number of real items (for out_count / padding)
  // batch (blockIdx.y = query): element strides between consecutive queries (0 when unbatched)
  long long in_stride;        // scores / keys_in
  long long keys_out_stride;
  long long out_stride;       // out_scores / out_ids / out_pos
  long long ids_stride;
  const int* n_dyn;           // optional per-query element count (device); overrides n / n_total
  int* fail_flag;             // optional (with n_dyn): set when a query's count is < need or > n (= the list capacity):
  int need;                   //   the sampled / prefiltered candidate list is unusable and the caller must redo exactly
  // ---- sharded search (exchange of per-shard lists)
  const Hit* hits_in;         // merge input: gathered per-shard lists instead of scores / keys. Element i of query b is
  int hits_k_src;             //   hits_in[(i / hits_k_src) * hits_rank_stride + b * hits_k_src + i % hits_k_src]
  long long hits_rank_stride; //   ([rank][query][k_src] as an all-gather leaves them); ties -> lower i = lower rank = lower
                              //   global id; padding entries (id < 0) are ignored; a set kHitMiss bit raises fail_flag
  Hit* out_hits;              // final level: also write the results as packed entries [k] (the all-gather send buffer)
  const int* aux_src;         // optional device flag OR-ed into bit 0 (kHitMiss) of every out_hits entry
  // ---- exchange fused into the two top-k kernels of a sharded scanning stage (peer-memory transport, lists that fit the LL
  // area): the final level of the LOCAL top-k sends its packed entries as LL lines straight into every rank's window
  // (ll_send), the merge level polls the R lists from the local window while it loads its keys (ll_recv, instead of
  // hits_in): no exchange kernel, no fence, no staging copy between them.
  int ll_send, ll_recv;
  P2PWindow ll;
  // gathered lists that are each sorted descending (what every rank's local top-k emits): run r of the sort buffer IS list r
  // (padded to k2 with the smallest key, odd runs loaded back to front = ascending), so the run-sorting stages are skipped
  // and the kernel goes straight to the merge-and-prune rounds. Number of lists, 0 = off; needs hits_k_src <= k2.
  int presorted_src;
};
// element i of query b of the gathered lists ([rank][query][k_src]), from the all-gather buffer or the LL window
__device__ __forceinline__ Hit topk_gathered_hit(const TopkArgs& a, long long b, long long i) {
  const long long r = i / a.hits_k_src, j = i - r * a.hits_k_src;
  if (a.ll_recv) return ll_recv_hit(a.ll, static_cast<int>(r), b * a.hits_k_src + j);
  return a.hits_in[r * a.hits_rank_stride + b * a.hits_k_src + j];
}

// Result j of query b from its sorted key (0: padding of a gathered list / beyond the valid results).
__device__ __forceinline__ void topk_emit(const TopkArgs& a, long long b, long long j, unsigned long long key) {
  const long long ob = b * a.out_stride;
  uint32_t aux = 0u;
  if (a.out_hits || a.ll_send) {
    if (a.n_dyn && a.fail_flag && ((a.n_dyn[b] < a.need) || (a.n_dyn[b] > a.n))) aux |= kHitMiss;
    if (a.aux_src && *a.aux_src) aux |= kHitMiss;
  }
  float sc = -INFINITY;
  long long id = -1;
  int pos = -1;
  if (key != 0ull) {
    const uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
    sc = ord_to_score(static_cast<uint32_t>(key >> 32));
    if (a.hits_in || a.ll_recv) {
      id = topk_gathered_hit(a, b, idx).id;   // (LL: the line has arrived, this is a plain re-read)
    } else {
      id = a.ids ? a.ids[b * a.ids_stride + idx] : (a.id_base + idx);
    }
    pos = static_cast<int>(idx);
    if (id < 0) sc = -INFINITY;   // padding that reached the output (fewer than k real entries)
  }
  if (a.out_scores) a.out_scores[ob + j] = sc;
  if (a.out_ids) a.out_ids[ob + j] = id;
  if (a.out_pos) a.out_pos[ob + j] = pos;
  if (a.out_hits || a.ll_send) {
    Hit h;
    h.score = sc;
    h.aux = aux;
    h.id = id;
    if (a.out_hits) a.out_hits[ob + j] = h;
    if (a.ll_send) ll_send_hit(a.ll, ob + j, h);
  }
}

template <int CHUNK, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) topk_kernel(const TopkArgs a) {
  extern __shared__ unsigned long long skeys[];
  constexpr int PER = CHUNK / 2 / THREADS;   // compare-exchanges per thread per full-width stage
  const long long b = blockIdx.y;
  long long n_in = a.n_dyn ? a.n_dyn[b] : a.n;
  if (a.n_dyn && a.fail_flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && (n_in < a.need || n_in > a.n)) atomicOr(a.fail_flag, 1);
    if (n_in > a.n) n_in = a.n;
  }
  const long long n_real = a.n_dyn ? n_in : a.n_total;
  const long long base = static_cast<long long>(blockIdx.x) * CHUNK;
  for (int j = threadIdx.x; j < CHUNK; j += THREADS) {
    long long i = base + j;
    unsigned long long key = 0ull;
    bool have = i < n_in;
    if (a.presorted_src > 0) {
      const int run = j / a.k2, pos = j - run * a.k2;
      const int e = (run & 1) ? a.k2 - 1 - pos : pos;
      have = run < a.presorted_src && e < a.hits_k_src;
      i = static_cast<long long>(run) * a.hits_k_src + e;
    }
    if (have) {
      if (a.hits_in || a.ll_recv) {
        const Hit h = topk_gathered_hit(a, b, i);
        if (h.id >= 0)
          key = (static_cast<unsigned long long>(score_to_ord(h.score)) << 32) |
                static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
        if ((h.aux & kHitMiss) && a.fail_flag) atomicOr(a.fail_flag, 1);
      } else if (a.scores) {
        key = (static_cast<unsigned long long>(score_to_ord(a.scores[b * a.in_stride + i])) << 32) |
              static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
      } else {
        key = a.keys_in[b * a.in_stride + i];
      }
    }
    skeys[j] = key;
  }
  const int K2 = a.k2;
  // Only the first `eff` slots can hold real keys (the rest is zero padding, the smallest key): the sort network is
  // sized for eff = the smallest power-of-two multiple of K2 that covers this block's keys, not for CHUNK — a
  // prefiltered candidate list of ~2k keys in an 8192-key buffer sorts 4x less.
  int eff = CHUNK;
  {
    const long long here = a.presorted_src > 0 ? static_cast<long long>(a.presorted_src) * K2 : n_in - base;   // keys of this block
    while ((eff >> 1) >= K2 && (eff >> 1) >= here) eff >>= 1;
  }
  // 1. sorted runs of K2, run r descending iff r is even (already so when the input lists were sorted)
  for (int size = 2; size <= (a.presorted_src > 0 ? 1 : K2); size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < PER; ++t) {
        const int e = threadIdx.x + t * THREADS;
        if (2 * e >= eff) break;
        const int pos = 2 * e - (e & (stride - 1));
        const unsigned long long x = skeys[pos], y = skeys[pos + stride];
        const bool desc = (pos & size) == 0;
        if ((x < y) == desc) {
          skeys[pos] = y;
          skeys[pos + stride] = x;
        }
      }
    }
  }
  // 2. merge-and-prune rounds
  for (int live = eff; live > K2; live >>= 1) {
    __syncthreads();
    unsigned long long keep[PER];
    const int half = live >> 1;
#pragma unroll
    for (int t = 0; t < PER; ++t) {
      const int e = threadIdx.x + t * THREADS;
      keep[t] = 0ull;
      if (e < half) {
        const int j = e / K2, i = e - j * K2;
        const unsigned long long x = skeys[(2 * j) * K2 + i], y = skeys[(2 * j + 1) * K2 + i];
        keep[t] = x > y ? x : y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < PER; ++t) {
      const int e = threadIdx.x + t * THREADS;
      if (e < half) skeys[e] = keep[t];
    }
    for (int stride = K2 >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < PER; ++t) {
        const int e = threadIdx.x + t * THREADS;
        if (e < (half >> 1)) {
          const int pos = 2 * e - (e & (stride - 1));
          const unsigned long long x = skeys[pos], y = skeys[pos + stride];
          const bool desc = ((pos / K2) & 1) == 0;
          if ((x < y) == desc) {
            skeys[pos] = y;
            skeys[pos + stride] = x;
          }
        }
      }
    }
  }
  __syncthreads();
  if (a.keys_out) {
    for (int j = threadIdx.x; j < a.k; j += THREADS)
      a.keys_out[b * a.keys_out_stride + static_cast<long long>(blockIdx.x) * a.k + j] = skeys[j];
  } else {
    const long long nvalid = n_real < a.k ? n_real : a.k;
    for (int j = threadIdx.x; j < a.k; j += THREADS) topk_emit(a, b, j, j < nvalid ? skeys[j] : 0ull);   // k may exceed the sort buffer: never read past nvalid
    if (a.out_count && threadIdx.x == 0) a.out_count[b] = static_cast<int>(nvalid);
  }
}


// ------------------------------------------------------------------------------------------------
// Radix select for large inputs (n > 4096): find the k-th largest 32-bit score key T with three histogram passes
// (11 + 11 + 10 bits), gather the keys > T and the needed keys == T (lowest indices first), then sort only those k
// keys with topk_kernel. Cost is ~3 reads of the score array regardless of k. blockIdx.y = query of a batch.
constexpr int kSelBins = 2048;
constexpr int kSelItemsPerBlock = 4096;
struct SelState {
  unsigned int prefix;     // selected high bits of T so far
  int k_rem;               // how many of the k still have to come from the current prefix bucket
  int count_gt;            // items known to be > T
  int eq_count;            // items == T (after the last pass)
  int counter;             // compaction cursor
  int pad[3];
};
struct SelArgs {
  const float* scores;     // [batch][n]
  long long n;
  int k;
  SelState* state;         // [batch]
  unsigned int* hist;      // [batch][3][kSelBins]
  unsigned long long* keys_out;   // [batch][keys_stride], k keys written per query
  long long keys_stride;   // elements between the key lists of consecutive queries (>= k)
};

__device__ __forceinline__ int sel_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ unsigned int sel_mask_above(int pass) {   // bits already fixed before this pass
  return pass == 0 ? 0u : (pass == 1 ? 0xFFE00000u : 0xFFFFFC00u);
}

__global__ void __launch_bounds__(256) sel_hist_kernel(const SelArgs a, int pass) {
  __shared__ unsigned int h[kSelBins];
  const long long b = blockIdx.y;
  for (int i = threadIdx.x; i < kSelBins; i += 256) h[i] = 0;
  __syncthreads();
  const unsigned int prefix = a.state[b].prefix, above = sel_mask_above(pass);
  const int shift = sel_shift(pass);
  const unsigned int bin_mask = pass == 2 ? 1023u : 2047u;
  const float* sc = a.scores + b * a.n;
  const long long base = static_cast<long long>(blockIdx.x) * kSelItemsPerBlock;
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int j = 0; j < kSelItemsPerBlock / 256; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    int bin = -1;
    if (i < a.n) {
      const unsigned int key = score_to_ord(sc[i]);
      if ((key & above) == (prefix & above)) bin = static_cast<int>((key >> shift) & bin_mask);
    }
    // warp-aggregated shared atomics: scores cluster in a few bins
    const unsigned int peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&h[bin], __popc(peers));
  }
  __syncthreads();
  unsigned int* gh = a.hist + (b * 3 + pass) * kSelBins;
  for (int i = threadIdx.x; i < kSelBins; i += 256)
    if (h[i]) atomicAdd(&gh[i], h[i]);
}

// one block per query: pick the bucket that contains the k_rem-th largest key of the current prefix bucket
__global__ void __launch_bounds__(256) sel_scan_kernel(const SelArgs a, int pass) {
  __shared__ unsigned int part[256];
  const long long b = blockIdx.x;
  const unsigned int* gh = a.hist + (b * 3 + pass) * kSelBins;
  SelState* st = a.state + b;
  const int k_rem = st->k_rem;
  // thread t owns bins [8t, 8t+8); suffix sums from the top
  unsigned int mine[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[j] = gh[threadIdx.x * 8 + j]; s += mine[j]; }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {   // inclusive suffix scan
    const unsigned int add = (threadIdx.x + off < 256) ? part[threadIdx.x + off] : 0u;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned int incl = part[threadIdx.x];            // items in bins >= 8t
  const unsigned int excl = incl - s;                     // items in bins >= 8t+8
  if (excl < static_cast<unsigned int>(k_rem) && incl >= static_cast<unsigned int>(k_rem)) {
    unsigned int above = excl;
    for (int j = 7; j >= 0; --j) {
      if (above + mine[j] >= static_cast<unsigned int>(k_rem)) {
        st->prefix |= static_cast<unsigned int>(threadIdx.x * 8 + j) << sel_shift(pass);
        st->k_rem = k_rem - static_cast<int>(above);
        st->count_gt += static_cast<int>(above);
        st->eq_count = static_cast<int>(mine[j]);
        break;
      }
      above += mine[j];
    }
  }
}

// gather keys > T (any order) and, when every key == T is needed, those too
__global__ void __launch_bounds__(256) sel_compact_kernel(const SelArgs a) {
  const long long b = blockIdx.y;
  SelState* st = a.state + b;
  const unsigned int T = st->prefix;
  const bool take_eq = st->eq_count == st->k_rem;
  const float* sc = a.scores + b * a.n;
  unsigned long long* out = a.keys_out + b * a.keys_stride;
  const long long base = static_cast<long long>(blockIdx.x) * kSelItemsPerBlock;
  for (int j = 0; j < kSelItemsPerBlock / 256; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i < a.n) {
      const unsigned int key = score_to_ord(sc[i]);
      if (key > T || (take_eq && key == T)) {
        const int pos = atomicAdd(&st->counter, 1);
        out[pos] = (static_cast<unsigned long long>(key) << 32) |
                   static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
      }
    }
  }
}

// ties at T beyond what is needed: take the k_rem lowest indices (index order scan, one block per query)
__global__ void __launch_bounds__(1024) sel_ties_kernel(const SelArgs a) {
  const long long b = blockIdx.x;
  SelState* st = a.state + b;
  if (st->eq_count == st->k_rem) return;
  __shared__ int wsum[32];
  __shared__ int running;
  const unsigned int T = st->prefix;
  const int need = st->k_rem, start = st->count_gt;
  const float* sc = a.scores + b * a.n;
  unsigned long long* out = a.keys_out + b * a.keys_stride;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < a.n; base += 1024) {
    const long long i = base + threadIdx.x;
    const bool hit = i < a.n && score_to_ord(sc[i]) == T;
    const unsigned int bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (hit && pos < need)
      out[start + pos] = (static_cast<unsigned long long>(T) << 32) |
                         static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wsum[w];
      running += tot;
    }
    __syncthreads();
    if (running >= need) break;
  }
}

// ------------------------------------------------------------------------------------------------
// k > kTopkMaxK (the reference accepts any prefetch_k / limit): the k selected keys of a query (radix select above, or
// every score when k >= n) are sorted by a bitonic network over GLOBAL memory — P = pow2ceil(k) keys per query, zero-padded,
// descending; compare-exchange distances below kBigChunk run inside shared memory (one launch per merge size), the few
// longer ones one launch each — and topk_emit_kernel writes the results. Rare path: a few launches, a few MB.
constexpr int kBigChunk = 4096;
__global__ void __launch_bounds__(256) keys_from_scores_kernel(const float* __restrict__ scores, long long n, long long in_stride,
                                                               unsigned long long* __restrict__ keys, long long P) {
  const long long b = blockIdx.y;
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= P) return;
  keys[b * P + i] = i < n ? ((static_cast<unsigned long long>(score_to_ord(scores[b * in_stride + i])) << 32) |
                             static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i)))
                          : 0ull;
}
__global__ void __launch_bounds__(256) keys_zero_tail_kernel(unsigned long long* __restrict__ keys, long long from, long long P) {
  const long long i = from + blockIdx.x * 256ll + threadIdx.x;
  if (i < P) keys[blockIdx.y * P + i] = 0ull;
}
// all compare-exchange steps with distance < kBigChunk of the merge sizes [size_lo, size_hi] (size_lo == size_hi > kBigChunk:
// the in-chunk tail of one global merge; size_lo = 2: the initial sort of every chunk)
__global__ void __launch_bounds__(512) bitonic_chunk_kernel(unsigned long long* __restrict__ keys, long long P, long long size_lo,
                                                            long long size_hi) {
  __shared__ unsigned long long sk[kBigChunk];
  unsigned long long* g = keys + blockIdx.y * P + static_cast<long long>(blockIdx.x) * kBigChunk;
  const long long gbase = static_cast<long long>(blockIdx.x) * kBigChunk;
  for (int j = threadIdx.x; j < kBigChunk; j += 512) sk[j] = g[j];
  for (long long size = size_lo; size <= size_hi; size <<= 1) {
    const int s0 = static_cast<int>(size >> 1 < kBigChunk / 2 ? size >> 1 : kBigChunk / 2);
    for (int stride = s0; stride > 0; stride >>= 1) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < kBigChunk / 2 / 512; ++t) {
        const int e = threadIdx.x + t * 512;
        const int pos = 2 * e - (e & (stride - 1));
        const unsigned long long x = sk[pos], y = sk[pos + stride];
        const bool desc = ((gbase + pos) & size) == 0;   // size == P: always descending
        if ((x < y) == desc) {
          sk[pos] = y;
          sk[pos + stride] = x;
        }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < kBigChunk; j += 512) g[j] = sk[j];
}
__global__ void __launch_bounds__(256) bitonic_global_kernel(unsigned long long* __restrict__ keys, long long P, long long size,
                                                             long long stride) {
  const long long e = blockIdx.x * 256ll + threadIdx.x;
  if (e >= (P >> 1)) return;
  unsigned long long* g = keys + blockIdx.y * P;
  const long long pos = 2 * e - (e & (stride - 1));
  const unsigned long long x = g[pos], y = g[pos + stride];
  const bool desc = (pos & size) == 0;
  if ((x < y) == desc) {
    g[pos] = y;
    g[pos + stride] = x;
  }
}
__global__ void __launch_bounds__(256) topk_emit_kernel(const TopkArgs a, const unsigned long long* __restrict__ keys, long long P) {
  const long long b = blockIdx.y;
  const long long j = blockIdx.x * 256ll + threadIdx.x;
  const long long nvalid = a.n_total < a.k ? a.n_total : a.k;
  if (j < a.k) topk_emit(a, b, j, j < nvalid ? keys[b * P + j] : 0ull);
  if (a.out_count && j == 0) a.out_count[b] = static_cast<int>(nvalid);
}

// ------------------------------------------------------------------------------------------------
// Per-token saliency of one page (visual_rag/visualization/saliency.py:69-79):
//   patch_scores[t] = max_q <q_q / (||q_q|| + 1e-8), d_t / (||d_t|| + 1e-8)>
// the column-max twin of MaxSim's row-max. One warp per document token (a lane owns 4 dims), the normalised query
// lives in shared memory as fp32; 8 warps per block. Only the top-k returned pages are ever scored, so this is a
// CUDA-core kernel (~2.6 MFLOP per page).
__global__ void __launch_bounds__(256) saliency_kernel(const __half* __restrict__ rows, const float* __restrict__ inv,
                                                       long long row0, int n_rows, const float* __restrict__ q, int Q,
                                                       float* __restrict__ out, int combine) {
  extern __shared__ float sq[];   // [Q][128] normalised query
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < Q; r += 8) {
    const float4 x = reinterpret_cast<const float4*>(q + r * 128)[lane];
    float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float s = 1.0f / (sqrtf(ss) + 1e-8f);
    reinterpret_cast<float4*>(sq + r * 128)[lane] = make_float4(x.x * s, x.y * s, x.z * s, x.w * s);
  }
  __syncthreads();
  for (int t = blockIdx.x * 8 + warp; t < n_rows; t += gridDim.x * 8) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(rows + (row0 + t) * 128) + lane);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    const float dn = __ldg(inv + row0 + t);
    float best = -INFINITY;
    for (int r = 0; r < Q; ++r) {
      const float4 x = reinterpret_cast<const float4*>(sq + r * 128)[lane];
      float d = a.x * x.x + a.y * x.y + b.x * x.z + b.y * x.w;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      best = fmaxf(best, d * dn);
    }
    if (lane == 0) out[t] = combine ? fmaxf(out[t], best) : best;   // combine: a later chunk of a long query
  }
}

// acc[i] += part[i]: MaxSim is a sum over query tokens, so a query longer than one operand image (128 rows) is scored
// in row chunks whose partial page scores add up (-inf = "page not in this shard" stays -inf).
__global__ void add_scores_kernel(float* __restrict__ acc, const float* __restrict__ part, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) acc[i] = __fadd_rn(acc[i], part[i]);
}

// ------------------------------------------------------------------------------------------------
// Fused top-k prefilter helpers (dense batched scans). thr[b] = the m-th best score of query b's sample.
__global__ void prefilter_thr_kernel(const float* __restrict__ sample_top, int m, int batch, float* __restrict__ thr,
                                     int* __restrict__ cnt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) {
    thr[b] = sample_top[static_cast<long long>(b) * m + (m - 1)];
    cnt[b] = 0;
  }
}
// Threshold straight from the sample, one block per query: linear 2048-bin histogram of the query's sample scores
// between their min and max, then the lower edge of the bin in which the count from the top reaches m. The
// threshold only has to let roughly m sample scores through (the survivor count of the full scan is verified
// afterwards), so no exact selection is needed: two passes over an L2-resident 256 KB row.
// CACHE (single-query sampled top-k, n_sample <= 32768): the strided sample is gathered once into dynamic shared memory
// and the histogram pass reads it from there. One block does the whole estimate; every 4-byte sample costs a 32-byte
// sector from L2, which is what its ~17 us are made of (a compact sample written by the scan epilogue would cut it).
constexpr int kThrCacheMax = 32768;
template <bool CACHE>
__global__ void __launch_bounds__(1024) prefilter_sample_thr_kernel(const float* __restrict__ sample, long long n_sample,
                                                                    int m, float* __restrict__ thr, int* __restrict__ cnt,
                                                                    long long elem_stride = 1) {
  extern __shared__ float s_cache[];   // CACHE: [n_sample]
  __shared__ unsigned int hist[2048];
  __shared__ float red_lo[32], red_hi[32];
  __shared__ float s_lo, s_w;
  __shared__ int s_bin;
  const long long b = blockIdx.x;
  const float* sc = sample + b * n_sample;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lo = INFINITY, hi = -INFINITY;
  if constexpr (CACHE) {
    for (long long i0 = threadIdx.x; i0 < n_sample; i0 += 4096) {
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long i = i0 + j * 1024;
        x[j] = i < n_sample ? __ldg(sc + i * elem_stride) : __int_as_float(0x7fc00000);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long i = i0 + j * 1024;
        if (i < n_sample) s_cache[i] = x[j];
        if (x[j] > -INFINITY && x[j] < INFINITY) { lo = fminf(lo, x[j]); hi = fmaxf(hi, x[j]); }
      }
    }
  } else {
    for (long long i = threadIdx.x; i < n_sample; i += 1024) {
      const float x = sc[i * elem_stride];
      if (x > -INFINITY && x < INFINITY) { lo = fminf(lo, x); hi = fmaxf(hi, x); }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
  for (int i = threadIdx.x; i < 2048; i += 1024) hist[i] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 32; ++w) { lo = fminf(lo, red_lo[w]); hi = fmaxf(hi, red_hi[w]); }
    s_lo = lo;
    s_w = (hi > lo) ? (hi - lo) / 2048.0f : 0.0f;
    s_bin = -1;
  }
  __syncthreads();
  const float base = s_lo, w = s_w;
  if (w > 0.0f) {
    const float inv_w = 1.0f / w;
    for (long long i = threadIdx.x; i < n_sample; i += 1024) {
      const float x = CACHE ? s_cache[i] : sc[i * elem_stride];
      if (x > -INFINITY && x < INFINITY) atomicAdd(&hist[min(2047, max(0, static_cast<int>((x - base) * inv_w)))], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // warp 0: count from the top, 64 bins per lane
    unsigned int mine = 0;
    for (int j = 0; j < 64; ++j) mine += hist[2047 - (lane * 64 + j)];
    unsigned int incl = mine;   // inclusive prefix over lanes (lane 0 = top bins)
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const unsigned int excl = incl - mine;
    if (excl < static_cast<unsigned int>(m) && incl >= static_cast<unsigned int>(m)) {
      unsigned int acc = excl;
      for (int j = 0; j < 64; ++j) {
        acc += hist[2047 - (lane * 64 + j)];
        if (acc >= static_cast<unsigned int>(m)) { s_bin = 2047 - (lane * 64 + j); break; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // fewer than m finite sample scores (or a degenerate range): let everything through; the survivor check decides
    thr[b] = (s_bin >= 0 && w > 0.0f) ? (base + static_cast<float>(s_bin) * w) : -INFINITY;
    cnt[b] = 0;
  }
}

// The candidate list of query b is usable iff it holds at least `need` and at most `cap` entries; otherwise the
// estimate failed and the host reruns the batch with the exact (unfiltered) path. Counts are clamped to cap.
__global__ void prefilter_check_kernel(int* __restrict__ cnt, int batch, int need, int cap, int* __restrict__ flag) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) {
    const int c = cnt[b];
    if (c < need || c > cap) atomicOr(flag, 1);
    if (c > cap) cnt[b] = cap;
  }
}

// Sampled top-k, pass 2: append the keys of all scores above the threshold to keys[0, cap) (any order; the sort
// kernel orders them). One warp-aggregated atomic per warp and 128 scores. thr == -inf lets everything above -inf through.
__global__ void __launch_bounds__(256) topk_compact_thr_kernel(const float* __restrict__ scores, long long n,
                                                               const float* __restrict__ thr, int* __restrict__ cnt,
                                                               unsigned long long* __restrict__ keys, int cap) {
  const float t = __ldg(thr);
  const int lane = threadIdx.x & 31;
  const long long n4 = n >> 2;
  // warp-uniform trip count (the shuffles below need every lane): vb = first float4 index of the warp's 32
  for (long long vb = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) - lane; vb < n4 + 1;
       vb += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = vb + lane;
    float x[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (v < n4) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(scores) + v);
      x[0] = f.x; x[1] = f.y; x[2] = f.z; x[3] = f.w;
    } else if (v == n4) {
      for (int j = 0; j < static_cast<int>(n & 3); ++j) x[j] = scores[v * 4 + j];
    }
    unsigned hits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x[j] > t) hits |= 1u << j;   // NaN and -inf never pass
    const int mine = __popc(hits);
    // exclusive prefix of `mine` over the warp + one atomic for the warp's total
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;   // warp-uniform
    int base = 0;
    if (lane == 31) base = atomicAdd(cnt, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((hits >> j) & 1u) {
        if (base < cap)
          keys[base] = (static_cast<unsigned long long>(score_to_ord(x[j])) << 32) |
                       static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(v * 4 + j));
        ++base;
      }
  }
}

// For every final result (query q, rank j) look its page id up in an earlier stage's list and fetch the score it had
// there (NaN if absent): the score_stage1 / score_stage2 fields of ThreeStageRetriever's result dicts
// (three_stage.py:160-173) without shipping the whole stage lists to the host. One warp per (q, j).
__global__ void __launch_bounds__(256) gather_stage_scores_kernel(const long long* __restrict__ final_ids, int k_final, int nq,
                                                                  const long long* __restrict__ stage_ids,
                                                                  const float* __restrict__ stage_scores, int k_stage,
                                                                  float* __restrict__ out, int out_stride, int out_col) {
  const long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<long long>(nq) * k_final) return;
  const int q = static_cast<int>(w / k_final);
  const long long id = final_ids[w];
  float found = __int_as_float(0x7fc00000);
  if (id >= 0) {
    const long long* ids = stage_ids + static_cast<long long>(q) * k_stage;
    for (int i0 = 0; i0 < k_stage; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < k_stage && ids[i] == id;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        found = stage_scores[static_cast<long long>(q) * k_stage + i0 + (__ffs(m) - 1)];
        break;
      }
    }
  }
  if (lane == 0) out[w * out_stride + out_col] = found;
}

// Second half of the approximate-first-pass scheme (batched exhaustive scans): per query, the candidates of the fp16 first
// pass with their EXACT scores -> sort keys (score, lower page first), and the guard: at least `need` candidates must beat
// (strictly) the best score any page outside the candidate list can have, h_min + eps (h_min = the first-pass score of the
// last candidate; a missing candidate, id < 0, means every page of the store is in the list). One block per query.
__global__ void __launch_bounds__(256) approx_finalize_kernel(const float* __restrict__ exact, const float* __restrict__ approx,
                                                              const long long* __restrict__ ids, int kc, int cap, long long id_base,
                                                              const float* __restrict__ eps, int need,
                                                              unsigned long long* __restrict__ keys, int* __restrict__ n_keys,
                                                              int* __restrict__ fail_flag) {
  const long long b = blockIdx.x;
  __shared__ int s_better, s_valid;
  if (threadIdx.x == 0) s_better = s_valid = 0;
  __syncthreads();
  const long long last_id = ids[b * kc + kc - 1];
  const float bound = last_id >= 0 ? approx[b * kc + kc - 1] + eps[b] : -INFINITY;
  int better = 0, valid = 0;
  for (int j = threadIdx.x; j < kc; j += 256) {
    const long long id = ids[b * kc + j];
    const float e = exact[b * kc + j];
    unsigned long long key = 0ull;
    if (id >= 0) {
      key = (static_cast<unsigned long long>(score_to_ord(e)) << 32) |
            static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(id - id_base));
      ++valid;
      if (e > bound) ++better;
    }
    keys[b * cap + j] = key;
  }
  for (int j = kc + threadIdx.x; j < cap; j += 256) keys[b * cap + j] = 0ull;
  atomicAdd(&s_better, better);
  atomicAdd(&s_valid, valid);
  __syncthreads();
  if (threadIdx.x == 0) {
    n_keys[b] = kc;
    if (s_better < min(need, s_valid)) atomicOr(fail_flag, 1);
  }
}

// Gathered per-shard lists that are too long for one sort block (n_src * k_src > 8192): unpack them into plain score / id
// arrays [n_lists][n_src * k_src] for the radix-select path. Padding entries get a NaN score (the lowest key), so that
// they sort behind every real entry including real -inf ones; a kHitMiss bit raises the flag.
__global__ void hits_unpack_kernel(const Hit* __restrict__ hits, int n_src, int n_lists, int k_src, float* __restrict__ scores,
                                   long long* __restrict__ ids, int* __restrict__ fail_flag) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long per = static_cast<long long>(n_src) * k_src;
  if (t >= per * n_lists) return;
  const long long b = t / per, i = t - b * per;
  const long long r = i / k_src;
  const Hit h = hits[(r * n_lists + b) * k_src + (i - r * k_src)];
  scores[t] = h.id >= 0 ? h.score : __int_as_float(0x7fc00000);
  ids[t] = h.id;
  if ((h.aux & kHitMiss) && fail_flag) atomicOr(fail_flag, 1);
}

// Store compaction: page p's rows [begin[p], begin[p] + n) -> rows [new_off[p], new_off[p+1]) of the new buffers (and the
// matching inverse norms). One block walks pages; a warp moves one 256-byte row per step (16-byte lanes x 16).
__global__ void __launch_bounds__(256) compact_rows_kernel(const __half* __restrict__ rows, const float* __restrict__ inv,
                                                           const long long* __restrict__ begin, const long long* __restrict__ new_off,
                                                           long long n_pages, __half* __restrict__ out_rows, float* __restrict__ out_inv) {
  const int sub = threadIdx.x & 15, rl = threadIdx.x >> 4;   // 16 rows in flight per block step
  for (long long p = blockIdx.x; p < n_pages; p += gridDim.x) {
    const long long src = begin[p], dst = new_off[p], n = new_off[p + 1] - dst;
    for (long long r = rl; r < n; r += 16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(rows + (src + r) * 128) + sub);
      *(reinterpret_cast<uint4*>(out_rows + (dst + r) * 128) + sub) = v;
      if (sub == 0) out_inv[dst + r] = inv[src + r];
    }
  }
}

__global__ void fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void sel_init_kernel(SelState* st, int k, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) {
    st[b].prefix = 0u;
    st[b].k_rem = k;
    st[b].count_gt = 0;
    st[b].eq_count = 0;
    st[b].counter = 0;
  }
}

// ------------------------------------------------------------------------------------------------ peer-memory exchange
// One-shot collectives of the sharded search over NVLink / NVSwitch peer memory (the messages are 4 KB .. a few MB and
// latency-bound: a NCCL all-gather of 256 packed hits costs 25-58 us on 8 GPUs, mostly protocol). Every rank owns a
// WINDOW that all peers have mapped (CUDA IPC): data[2 slots][R sources][cap bytes] + flags[2][R] (+ the LL area, above). A collective with
// sequence number `epoch` uses slot epoch & 1: each rank stores its message into region [slot][me] of EVERY rank's window
// (plain stores through the peer mapping), fences system-wide, and the last block to finish publishes `epoch` in
// flags[slot][me] of every window; then it waits until its own flags[slot][*] all show `epoch` and consumes the R regions
// from local memory. Two slots suffice: a rank can be at most one collective ahead of the slowest peer (completing
// collective e needs every peer's flag for e, which a peer only sends after it finished e - 1 in stream order).
// message of this rank (n16 16-byte words) -> every rank's window; publish; wait for all sources. All blocks of the grid
// must be co-resident (they spin): launched with at most 128 blocks of 256 threads (see p2p_blocks).
__device__ __forceinline__ void p2p_exchange(const P2PWindow& w, const uint4* __restrict__ msg, long long n16, long long n_tail = 0) {
  const int slot = w.epoch & 1;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int r = 0; r < w.R; ++r) {
    uint4* dst = reinterpret_cast<uint4*>(w.win[r] + (static_cast<unsigned long long>(slot) * w.R + w.me) * w.cap);
    for (long long i = gtid; i < n16; i += gsz) dst[i] = msg[i];
    for (long long i = gtid; i < n_tail; i += gsz)   // trailing 4-byte words (a message that is not a multiple of 16 bytes;
      reinterpret_cast<unsigned*>(dst + n16)[i] = reinterpret_cast<const unsigned*>(msg + n16)[i];   // all of it when unaligned)
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(w.ctr, 1u);
    if (prev == gridDim.x - 1) {
      *w.ctr = 0u;               // next collective (same stream) starts from zero
      __threadfence_system();    // cumulative: every block's stores (observed through the counter) before the flags
      for (int r = 0; r < w.R; ++r) p2p_st_release_sys(p2p_flag(w, r, slot, w.me), w.epoch);
    }
  }
  if (threadIdx.x < w.R) {
    const unsigned* f = p2p_flag(w, w.me, slot, threadIdx.x);
    const unsigned long long t0 = p2p_now_ns();
    unsigned spins = 0;
    while (static_cast<int>(p2p_ld_acquire_sys(f) - w.epoch) < 0) {
      // a peer that is merely late (host-side skew between the ranks) is waited for; only a peer that died or ranks whose
      // collective calls diverged end in the watchdog, which fails the call instead of hanging the GPU
      if ((++spins & 1023u) == 0 && p2p_now_ns() - t0 > w.timeout_ns) {
        printf("vrag: peer exchange watchdog rank=%d waits for rank=%d epoch=%u\n", w.me, static_cast<int>(threadIdx.x), w.epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}
// all-gather: gathered[src][0..n16) for src = 0..R-1 (the layout ncclAllGather produces)
__global__ void __launch_bounds__(256) p2p_allgather_kernel(const P2PWindow w, const uint4* __restrict__ msg, long long n16,
                                                            uint4* __restrict__ gathered) {
  p2p_exchange(w, msg, n16);
  const int slot = w.epoch & 1;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int s = 0; s < w.R; ++s) {
    const uint4* src = reinterpret_cast<const uint4*>(w.win[w.me] + (static_cast<unsigned long long>(slot) * w.R + s) * w.cap);
    for (long long i = gtid; i < n16; i += gsz) gathered[s * n16 + i] = __ldcg(src + i);   // written by a peer: not through L1
  }
}
// max-all-reduce of n floats in place. A buffer that is not 16-byte aligned goes through the 4-byte path entirely (the
// choice must not change which collective path a rank takes: the ranks' sequence numbers have to stay in step).
__global__ void __launch_bounds__(256) p2p_allreduce_max_kernel(const P2PWindow w, float* __restrict__ buf, long long n) {
  const long long n4 = (reinterpret_cast<unsigned long long>(buf) & 15ull) == 0 ? (n >> 2) : 0;
  const long long n_tail = n - 4 * n4;
  p2p_exchange(w, reinterpret_cast<const uint4*>(buf), n4, n_tail);
  const int slot = w.epoch & 1;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
  const uint8_t* base = w.win[w.me] + static_cast<unsigned long long>(slot) * w.R * w.cap;
  for (long long i = gtid; i < n4; i += gsz) {
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int s = 0; s < w.R; ++s) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(base + s * w.cap) + i);   // written by a peer: not through L1
      m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
    }
    reinterpret_cast<float4*>(buf)[i] = m;
  }
  for (long long i = gtid; i < n_tail; i += gsz) {
    float m = -INFINITY;
    for (int s = 0; s < w.R; ++s) m = fmaxf(m, __ldcg(reinterpret_cast<const float*>(base + s * w.cap) + n4 * 4 + i));
    buf[n4 * 4 + i] = m;
  }
}

// ------------------------------------------------------------------------------------------------ owned candidates
// A candidate stage of a BATCH over the sharded corpus receives replicated lists [n_queries][n_cand] of global page ids, of
// which a rank owns ~1/R. A foreign candidate fetches nothing, but it still costs its slot in the gather's tile pipeline and a
// chain of dependent id -> page -> row-range loads in the rerank, so the stage stopped getting cheaper with more ranks. These
// kernels restrict the scan to the rank's own candidates: count per query (+ the maximum over the batch, which the host
// reads to size the lists), order-preserving compaction to [n_queries][stride] (ids padded with -1, original positions
// kept), and the scatter of the compact scores back into the -inf-filled [n_queries][n_cand] matrix the exchange expects.
__global__ void __launch_bounds__(256) own_count_kernel(const long long* __restrict__ ids, int n_cand, long long base, long long n_pages,
                                                        int* __restrict__ cnt, int* __restrict__ max_cnt) {
  const long long* row = ids + static_cast<long long>(blockIdx.x) * n_cand;
  int mine = 0;
  for (int i = threadIdx.x; i < n_cand; i += 256) {
    const long long id = row[i];
    mine += (id >= base && id < base + n_pages) ? 1 : 0;
  }
  __shared__ int acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&acc, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt[blockIdx.x] = acc;
    atomicMax(max_cnt, acc);
  }
}
__global__ void __launch_bounds__(256) own_compact_kernel(const long long* __restrict__ ids, int n_cand, long long base, long long n_pages,
                                                          int stride, long long* __restrict__ out_ids, int* __restrict__ out_pos) {
  __shared__ int wsum[8];
  __shared__ int running;
  const long long* row = ids + static_cast<long long>(blockIdx.x) * n_cand;
  long long* oid = out_ids + static_cast<long long>(blockIdx.x) * stride;
  int* opos = out_pos + static_cast<long long>(blockIdx.x) * stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n_cand; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const long long id = i < n_cand ? row[i] : -1;
    const bool own = id >= base && id < base + n_pages;
    const unsigned bal = __ballot_sync(0xffffffffu, own);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (own && pos < stride) {
      oid[pos] = id;
      opos[pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += wsum[w];
      running += tot;
    }
    __syncthreads();
  }
  for (int j = running + threadIdx.x; j < stride; j += 256) {   // padding: a foreign id scores -inf and is never scattered
    oid[j] = -1;
    opos[j] = -1;
  }
}
__global__ void __launch_bounds__(256) own_scatter_kernel(const float* __restrict__ sc_c, const int* __restrict__ pos, int stride, int n_cand,
                                                          int n_queries, float* __restrict__ raw) {
  const long long t = blockIdx.x * 256ll + threadIdx.x;
  if (t >= static_cast<long long>(n_queries) * stride) return;
  const long long b = t / stride;
  const int p = pos[t];
  if (p >= 0) raw[b * n_cand + p] = sc_c[t];
}

// The same two collectives for small messages (<= ll_cap / 2 bytes per rank) as LL lines: send, then poll — no fence, flag or
// counter. At most 128 blocks: every block both sends and waits, so all of them have to be resident.
__global__ void __launch_bounds__(256) p2p_ll_allgather_kernel(const P2PWindow w, const uint4* __restrict__ msg, long long n16,
                                                               uint4* __restrict__ gathered) {
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = gtid; i < n16; i += gsz) {
    const uint4 v = msg[i];
    ll_send_pair(w, 2 * i, v.x, v.y);
    ll_send_pair(w, 2 * i + 1, v.z, v.w);
  }
  for (int s = 0; s < w.R; ++s)
    for (long long i = gtid; i < n16; i += gsz) {
      uint4 v;
      ll_recv_pair(w, s, 2 * i, v.x, v.y);
      ll_recv_pair(w, s, 2 * i + 1, v.z, v.w);
      gathered[s * n16 + i] = v;
    }
}
__global__ void __launch_bounds__(256) p2p_ll_allreduce_max_kernel(const P2PWindow w, float* __restrict__ buf, long long n) {
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long np = (n + 1) >> 1;
  for (long long p = gtid; p < np; p += gsz) {
    const float x = buf[2 * p], y = 2 * p + 1 < n ? buf[2 * p + 1] : -INFINITY;
    ll_send_pair(w, p, __float_as_uint(x), __float_as_uint(y));
  }
  for (long long p = gtid; p < np; p += gsz) {   // the same thread that sent pair p overwrites it
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int s = 0; s < w.R; ++s) {
      uint32_t a, b;
      ll_recv_pair(w, s, p, a, b);
      m0 = fmaxf(m0, __uint_as_float(a));
      m1 = fmaxf(m1, __uint_as_float(b));
    }
    buf[2 * p] = m0;
    if (2 * p + 1 < n) buf[2 * p + 1] = m1;
  }
}

}  // namespace vrag
