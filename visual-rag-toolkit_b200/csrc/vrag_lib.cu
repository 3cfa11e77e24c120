// libvrag_b200.so — host side of the C ABI declared in include/vrag_b200.h.
// Owns the GPU-resident corpus stores and launches the sm_100a kernels. No CPU compute path exists:
// every scoring call ends in a kernel launch or an error.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vrag_b200.h"
#include "aux_kernels.cuh"
#include "host_convert.h"
#include "scan_launch.h"
#include "pooling_kernels.cuh"

using namespace vrag;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                    \
  } while (0)
#define VRAG_LOCK(c)                             \
  if (!(c)) return fail("corpus is NULL");       \
  std::lock_guard<std::recursive_mutex> _vrag_lock((c)->mu)
#define TRY(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

extern "C" const char* vrag_last_error(void) { return g_err.c_str(); }
extern "C" int vrag_abi_version(void) { return 2; }

// Function attributes (opt-in dynamic shared memory) are per device: every launch site keeps one flag per device of
// the process. Returns true the first time a site is reached on the current device.
struct PerDeviceOnce {
  bool done[64] = {false};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
// Experiment / fallback knobs read from the environment once per process.
static bool env_flag_is(const char* name, char value) {
  const char* e = getenv(name);
  return e && e[0] == value;
}
static bool knob_no_pad() { static const bool v = getenv("VRAG_NO_PAD") != nullptr; return v; }
static bool knob_hi_only() { static const bool v = env_flag_is("VRAG_QUERY_SPLIT", '0'); return v; }
static bool knob_no_prefilter() { return env_flag_is("VRAG_PREFILTER", '0'); }   // per batch call: tests toggle it at run time

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// rows: [total_rows][128] fp16, box = {64 cols, box_rows}, 128B swizzle (UMMA K-major SW128 operand layout)
static int make_rows_map(CUtensorMap* m, void* base, int64_t total_rows, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {128, static_cast<cuuint64_t>(total_rows)};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(rows, box=%d) failed: %d", box_rows, (int)r);
  return 0;
}
// 3-D view {128 cols, rows_per_page, n_pages} of a fixed-rows store; box {64, slot_rows, 128/slot_rows}: the rows of a
// page beyond rows_per_page are out of bounds and arrive as zeros, so every page fills a power-of-two slot of the tile.
static int make_padded_map(CUtensorMap* m, void* base, int64_t n_pages, int rows_per_page, int slot_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[3] = {128, static_cast<cuuint64_t>(rows_per_page), static_cast<cuuint64_t>(n_pages)};
  cuuint64_t strides[2] = {256, static_cast<cuuint64_t>(rows_per_page) * 256};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(slot_rows), static_cast<cuuint32_t>(kTileRows / slot_rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(padded, rows=%d slot=%d) failed: %d", rows_per_page, slot_rows, (int)r);
  return 0;
}
static int make_scale_map(CUtensorMap* m, void* base, int64_t total_rows, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[1] = {static_cast<cuuint64_t>(total_rows)};
  cuuint64_t strides[1] = {0};
  cuuint32_t box[1] = {static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[1] = {1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(scale, box=%d) failed: %d", box_rows, (int)r);
  return 0;
}

// ------------------------------------------------------------------------------------------------ corpus
struct Store {
  __half* rows = nullptr;        // [total_rows][128]
  float* inv = nullptr;          // [total_rows] 1/(||row||+1e-8)
  long long* offsets = nullptr;  // device [n_pages+1] (nullptr when fixed_rows > 0)
  std::vector<int64_t> h_offsets;
  int* tile_page0 = nullptr;     // packed dense variable-length: [n_tiles+1]
  long long* tile_row0 = nullptr;  // first row of each tile [n_tiles+1]
  int64_t n_tiles = 0;
  int64_t n_pages = 0, total_rows = 0, fixed_rows = 0, max_rows = 0;
  int64_t cap_rows = 0;          // allocated rows (>= total_rows; vrag_store_append grows geometrically)
  bool dirty = false;            // appended to since the page tables / tensor maps were last built
  bool packed = false;
  CUtensorMap tm128, ts128, ts32;
  RowMaps tm_small;              // row boxes of 4, 8, ..., 32 rows
  CUtensorMap tm3d;              // fixed_rows <= 32, not a power of two: {128, fixed_rows, n_pages} view with a padded box
  int pad_slot = 0;              // its slot height (next power of two), 0 when the view does not exist
  // Page-table indirection (upserts that change a page's row count, deletes): page p owns rows [h_begin[p], h_end[p])
  // anywhere in the row buffer. A replaced page that grew lives behind the last row, its old rows are garbage until
  // vrag_store_compact; a deleted page has end == begin and keeps its index. Dense stores (indirect == false) stay on
  // the contiguous fast paths.
  bool indirect = false;
  std::vector<int64_t> h_begin, h_end;
  long long* d_begin = nullptr;
  long long* d_end = nullptr;
  int64_t garbage_rows = 0;
};

// rows [r0, r0 + n) of page p (host-side page table)
static void page_rows_h(const Store& s, int64_t p, int64_t* r0, int64_t* n) {
  if (s.indirect) {
    *r0 = s.h_begin[p];
    *n = s.h_end[p] - s.h_begin[p];
  } else if (s.fixed_rows > 0) {
    *r0 = p * s.fixed_rows;
    *n = s.fixed_rows;
  } else {
    *r0 = s.h_offsets[p];
    *n = s.h_offsets[p + 1] - s.h_offsets[p];
  }
}
static void make_indirect(Store& s) {
  if (s.indirect) return;
  s.h_begin.resize(s.n_pages);
  s.h_end.resize(s.n_pages);
  for (int64_t p = 0; p < s.n_pages; ++p) {
    int64_t r0, n;
    page_rows_h(s, p, &r0, &n);
    s.h_begin[p] = r0;
    s.h_end[p] = r0 + n;
  }
  s.indirect = true;
  s.fixed_rows = 0;
  s.h_offsets.clear();
  s.dirty = true;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max<size_t>(n, 256);
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct BatchCtx {   // the batch of queries last uploaded to the device
  bool valid = false;
  int nq = 0, n_stages = 0, qs = 1;
  bool per_stage = false;
  std::vector<int> max_rows;   // largest query row count per stage
};

// NCCL, resolved at run time (dlopen): the few entry points the exchange needs, declared here so that the build has no
// NCCL header / link dependency. ncclUniqueId is a 128-byte struct passed by value; ncclComm_t is an opaque pointer.
struct NcclId { char internal[VRAG_UNIQUE_ID_BYTES]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static const int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclMax = 2, kNcclMin = 3;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

static const int kMaxCommEvents = 24;
struct CommState {
  int rank = 0, nranks = 1;
  void* nccl = nullptr;              // ncclComm_t
  DevBuf<Hit> send, recv;            // packed local lists / gathered lists
  DevBuf<float> raw;                 // batched candidate stages: [n_queries][n_cand] scores (max-all-reduced across shards)
  DevBuf<float> raw_c;               // the same for the rank's OWN candidates only: [n_queries][max own count]
  DevBuf<long long> own_ids;         // [n_queries][max own count] compacted candidate ids (-1 padded)
  DevBuf<int> own_pos;               // their positions in the replicated lists
  DevBuf<int> own_cnt;               // [n_queries] own counts, [n_queries] = the maximum
  DevBuf<float> m_scores;            // unpacked gathered lists (very long merges)
  DevBuf<long long> m_ids;
  cudaEvent_t ev[2 * kMaxCommEvents] = {nullptr};
  int n_ev = 0;                      // collectives timed in the current host-facing search
  bool timing = false;               // record events around collectives (host-facing searches on the library stream)
  float last_us[kMaxCommEvents] = {0};
  float last_begin_us[kMaxCommEvents] = {0};   // start of each collective relative to the start of the search (ev0)
  int last_n = 0;
  // peer-memory exchange (aux_kernels.cuh): this rank's window, the peers' windows as mapped here, the collective counter
  bool p2p = false;
  uint8_t* win = nullptr;
  uint8_t* peer_win[kP2PMaxRanks] = {nullptr};
  size_t p2p_cap = 0;                // bytes per (slot, source) region
  unsigned p2p_epoch = 0;
  unsigned* p2p_ctr = nullptr;
  unsigned long long p2p_timeout_ns = 120ull * 1000000000ull;
};

struct vrag_corpus {
  // One handle = one stream + one set of scratch buffers: entry points serialise on this lock, so a handle may be shared
  // by threads (the reference's ingest runs uploader threads, run_qdrant_beir.py:720-768; ctypes releases the GIL).
  // Recursive: vrag_search / vrag_store_append call other entry points.
  std::recursive_mutex mu;
  int device = 0;
  BatchCtx batch;
  int64_t page_base = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::map<std::string, Store> stores;
  DevBuf<float> d_query, d_scores, d_out_scores;
  DevBuf<unsigned long long> d_skeys;   // sampled top-k: survivor keys
  DevBuf<float> d_sthr;           // sampled top-k: threshold [1]
  DevBuf<int> d_sstate;           // sampled top-k: [0] survivor count, [1] "estimate failed" flag
  int sampled_runs = 0, sampled_fallbacks = 0;
  DevBuf<float> d_scores_part;    // partial page scores of the later row chunks of a > 128-token query
  DevBuf<uint8_t> d_qimg;
  DevBuf<unsigned long long> d_keys_a, d_keys_b;
  DevBuf<uint8_t> d_sel_state;
  DevBuf<unsigned int> d_sel_hist;
  DevBuf<long long> d_cand, d_out_ids;
  DevBuf<int> d_counts;
  DevBuf<uint8_t> d_qimg_batch;   // batched search: one operand image per query
  DevBuf<int> d_qmeta;            // batched search: [n_stages][2][nq] query row ranges + [nq] effective rows
  DevBuf<float> d_fthr, d_ftop;   // prefilter: thresholds [nq], sample top-m scores [nq][m]
  DevBuf<long long> d_ftop_ids;
  DevBuf<float> d_stage_sc;       // final-only batch results: earlier-stage scores of the final ids
  DevBuf<int> d_fcnt;             // prefilter: candidate counts [nq] + flag [1]
  DevBuf<unsigned long long> d_fkeys;   // prefilter: candidate keys [nq][cap]
  DevBuf<float> d_ap_sc, d_ap_exact, d_ap_eps;   // approximate first pass: candidate scores (fp16 pass / exact) [nq][kc], bounds [nq]
  DevBuf<long long> d_ap_id;                     //   candidate ids [nq][kc]
  DevBuf<unsigned long long> d_ap_keys;          //   sort keys of the exact scores [nq][key_cap]
  DevBuf<int> d_ap_cnt;                          //   key counts [nq]
  int64_t approx_runs = 0;
  int* h_flag = nullptr;          // pinned
  int64_t prefilter_runs = 0, prefilter_fallbacks = 0;
  int* h_qmeta = nullptr;         // pinned staging of d_qmeta
  size_t h_qmeta_cap = 0;
  size_t h_query_cap = 0;         // rows
  float* h_query = nullptr;       // pinned staging
  float* h_out_scores = nullptr;
  long long* h_out_ids = nullptr;
  int* h_counts = nullptr;
  size_t h_out_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  uint16_t* h_up = nullptr;       // pinned staging of host uploads: two chunks of kUploadChunk fp16 values, filled by the
  cudaEvent_t ev_up[2] = {nullptr, nullptr};   //   host worker pool while the other one is on its way (upload_host_pages)
  float last_ms[2] = {0, 0};
  int64_t launches = 0;
  bool attrs_set = false;
  CommState comm;
  std::map<int, DevBuf<uint32_t>> filters;   // payload-filter page bitmasks (vrag_filter_create)
  std::map<int, int64_t> filter_pages;       // pages each mask covers
  int next_filter = 1;
};

static const int kOperandRows = 128;    // query rows of one MMA operand image; longer token queries are scored in chunks
static const int kMaxQueryRows = 1024;  // staging capacity for raw query tokens (pooled queries may be long)
static const int kMaxStages = 8;

static int set_device(vrag_corpus* c) {
  CUDA_OK(cudaSetDevice(c->device));
  return 0;
}

static void free_store(Store& s) {
  if (s.rows) cudaFree(s.rows);
  if (s.inv) cudaFree(s.inv);
  if (s.offsets) cudaFree(s.offsets);
  if (s.tile_page0) cudaFree(s.tile_page0);
  if (s.tile_row0) cudaFree(s.tile_row0);
  if (s.d_begin) cudaFree(s.d_begin);
  if (s.d_end) cudaFree(s.d_end);
  s = Store();
}

extern "C" int vrag_corpus_create(int device, int64_t page_base, vrag_corpus_t** out) {
  if (!out) return fail("out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (%s): libvrag_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("device %d out of range (have %d)", device, ndev);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  vrag_corpus* c = new vrag_corpus();
  c->device = device;
  c->page_base = page_base;
  c->num_sms = prop.multiProcessorCount;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&c->ev0));
  CUDA_OK(cudaEventCreate(&c->ev1));
  CUDA_OK(cudaEventCreate(&c->evk0));
  CUDA_OK(cudaEventCreate(&c->evk1));
  CUDA_OK(cudaMallocHost(&c->h_query, kMaxQueryRows * 128 * sizeof(float)));
  c->h_query_cap = kMaxQueryRows;
  CUDA_OK(cudaMallocHost(&c->h_counts, (kMaxStages + 1) * sizeof(int)));   // + the sampled top-k flag
  CUDA_OK(cudaMallocHost(&c->h_flag, sizeof(int)));
  *c->h_flag = 0;
  TRY(c->d_query.ensure(kMaxQueryRows * 128));
  TRY(c->d_qimg.ensure(256 * 256));
  TRY(c->d_counts.ensure(kMaxStages + 1));
  *out = c;
  return 0;
}

static void comm_release(vrag_corpus* c);

extern "C" int vrag_corpus_destroy(vrag_corpus_t* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->stores) free_store(kv.second);
  c->d_query.release();
  c->d_scores.release();
  c->d_out_scores.release();
  c->d_scores_part.release();
  c->d_skeys.release();
  c->d_sthr.release();
  c->d_sstate.release();
  c->d_qimg.release();
  c->d_keys_a.release();
  c->d_keys_b.release();
  c->d_sel_state.release();
  c->d_sel_hist.release();
  c->d_cand.release();
  c->d_out_ids.release();
  c->d_counts.release();
  c->d_qimg_batch.release();
  c->d_qmeta.release();
  c->d_fthr.release();
  c->d_ftop.release();
  c->d_ftop_ids.release();
  c->d_fcnt.release();
  c->d_stage_sc.release();
  c->d_fkeys.release();
  c->d_ap_sc.release();
  c->d_ap_exact.release();
  c->d_ap_eps.release();
  c->d_ap_id.release();
  c->d_ap_keys.release();
  c->d_ap_cnt.release();
  comm_release(c);
  for (auto& kv : c->filters) kv.second.release();
  if (c->h_flag) cudaFreeHost(c->h_flag);
  if (c->h_qmeta) cudaFreeHost(c->h_qmeta);
  if (c->h_query) cudaFreeHost(c->h_query);
  if (c->h_out_scores) cudaFreeHost(c->h_out_scores);
  if (c->h_out_ids) cudaFreeHost(c->h_out_ids);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->h_up) cudaFreeHost(c->h_up);
  for (int i = 0; i < 2; ++i)
    if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->evk0);
  cudaEventDestroy(c->evk1);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

static int ensure_host_out(vrag_corpus* c, size_t n) {
  if (n <= c->h_out_cap) return 0;
  if (c->h_out_scores) cudaFreeHost(c->h_out_scores);
  if (c->h_out_ids) cudaFreeHost(c->h_out_ids);
  c->h_out_cap = 0;
  size_t want = std::max<size_t>(n, 4096);
  CUDA_OK(cudaMallocHost(&c->h_out_scores, want * sizeof(float)));
  CUDA_OK(cudaMallocHost(&c->h_out_ids, want * sizeof(long long)));
  c->h_out_cap = want;
  return 0;
}

// page layout bookkeeping + tensor maps once rows/inv are in place
static int finish_store(vrag_corpus* c, Store& s, const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows) {
  if (s.offsets) cudaFree(s.offsets);
  if (s.tile_page0) cudaFree(s.tile_page0);
  if (s.tile_row0) cudaFree(s.tile_row0);
  s.offsets = nullptr;
  s.tile_page0 = nullptr;
  s.tile_row0 = nullptr;
  s.n_tiles = 0;
  s.dirty = false;
  s.n_pages = n_pages;
  s.fixed_rows = fixed_rows;
  if (s.d_begin) cudaFree(s.d_begin);
  if (s.d_end) cudaFree(s.d_end);
  s.d_begin = s.d_end = nullptr;
  if (s.indirect) {
    int64_t mx = 0;
    for (int64_t i = 0; i < n_pages; ++i) mx = std::max(mx, s.h_end[i] - s.h_begin[i]);
    s.max_rows = mx;
    if (n_pages > 0) {
      CUDA_OK(cudaMalloc(&s.d_begin, n_pages * sizeof(long long)));
      CUDA_OK(cudaMalloc(&s.d_end, n_pages * sizeof(long long)));
      CUDA_OK(cudaMemcpyAsync(s.d_begin, s.h_begin.data(), n_pages * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
      CUDA_OK(cudaMemcpyAsync(s.d_end, s.h_end.data(), n_pages * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
    }
  } else if (fixed_rows > 0) {
    s.max_rows = fixed_rows;
  } else {
    if (page_offsets != s.h_offsets.data()) s.h_offsets.assign(page_offsets, page_offsets + n_pages + 1);
    int64_t mx = 0;
    for (int64_t i = 0; i < n_pages; ++i) mx = std::max(mx, page_offsets[i + 1] - page_offsets[i]);
    s.max_rows = mx;
    CUDA_OK(cudaMalloc(&s.offsets, (n_pages + 1) * sizeof(long long)));
    CUDA_OK(cudaMemcpyAsync(s.offsets, page_offsets, (n_pages + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  s.packed = s.max_rows <= kTileRows;
  if (s.packed && fixed_rows == 0 && n_pages > 0 && !s.indirect) {
    // greedy packing of consecutive pages into 128-row tiles
    std::vector<int> t0;
    int64_t pg = 0;
    while (pg < n_pages) {
      t0.push_back(static_cast<int>(pg));
      int64_t rows = 0;
      int64_t e = pg;
      while (e < n_pages && rows + (page_offsets[e + 1] - page_offsets[e]) <= kTileRows && e - pg < kTileRows) {
        rows += page_offsets[e + 1] - page_offsets[e];
        ++e;
      }
      pg = e;
    }
    t0.push_back(static_cast<int>(n_pages));
    s.n_tiles = static_cast<int64_t>(t0.size()) - 1;
    CUDA_OK(cudaMalloc(&s.tile_page0, t0.size() * sizeof(int)));
    CUDA_OK(cudaMemcpyAsync(s.tile_page0, t0.data(), t0.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    std::vector<long long> tr(t0.size());
    for (size_t i = 0; i < t0.size(); ++i) tr[i] = page_offsets[t0[i]];
    CUDA_OK(cudaMalloc(&s.tile_row0, tr.size() * sizeof(long long)));
    CUDA_OK(cudaMemcpyAsync(s.tile_row0, tr.data(), tr.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  s.pad_slot = 0;
  if (s.total_rows > 0 && fixed_rows > 0 && fixed_rows <= 32 && (fixed_rows & (fixed_rows - 1)) != 0) {
    int slot = 4;
    while (slot < fixed_rows) slot <<= 1;
    TRY(make_padded_map(&s.tm3d, s.rows, n_pages, static_cast<int>(fixed_rows), slot));
    s.pad_slot = slot;
  }
  if (s.total_rows > 0) {
    TRY(make_rows_map(&s.tm128, s.rows, s.total_rows, kTileRows));
    for (int i = 0; i < kNumSmallMaps; ++i) TRY(make_rows_map(&s.tm_small.m[i], s.rows, s.total_rows, kBoxStep * (i + 1)));
    TRY(make_scale_map(&s.ts128, s.inv, s.total_rows, kScaleBoxBig));
    TRY(make_scale_map(&s.ts32, s.inv, s.total_rows, kScaleBoxSmall));
  }
  return 0;
}

static int check_layout(const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows, int64_t* total_rows) {
  if (n_pages < 0) return fail("n_pages < 0");
  if (n_pages >= (1ll << 31)) return fail("a shard holds at most 2^31-1 pages");
  if (fixed_rows > 0) {
    *total_rows = n_pages * fixed_rows;
  } else {
    if (!page_offsets) return fail("page_offsets is NULL and fixed_rows == 0");
    if (page_offsets[0] != 0) return fail("page_offsets[0] must be 0");
    for (int64_t i = 0; i < n_pages; ++i)
      if (page_offsets[i + 1] < page_offsets[i]) return fail("page_offsets must be non-decreasing (page %lld)", (long long)i);
    *total_rows = page_offsets[n_pages];
  }
  if (*total_rows >= (1ll << 31)) return fail("a store holds at most 2^31-1 rows per shard (TMA coordinates are int32)");
  return 0;
}

static int alloc_store(vrag_corpus* c, const char* name, int64_t total_rows, Store** out) {
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it != c->stores.end()) {
    if (total_rows > 0 && it->second.cap_rows >= total_rows &&
        (it->second.cap_rows <= (1ll << 18) || it->second.cap_rows <= 4 * total_rows)) {   // never pin >4x (or >64 MB) of slack
      // replacing a store by one that fits its buffers (the scratch store of compute_maxsim_score / _batch is rewritten on
      // every call): keep the row / scale allocations; finish_store rebuilds page tables and tensor maps
      it->second.total_rows = total_rows;
      it->second.h_offsets.clear();
      it->second.indirect = false;
      it->second.h_begin.clear();
      it->second.h_end.clear();
      it->second.garbage_rows = 0;
      *out = &it->second;
      return 0;
    }
    free_store(it->second);
    c->stores.erase(it);
  }
  Store s;
  s.total_rows = total_rows;
  s.cap_rows = total_rows;
  if (total_rows > 0) {
    CUDA_OK(cudaMalloc(&s.rows, static_cast<size_t>(total_rows) * 128 * sizeof(__half)));
    cudaError_t e = cudaMalloc(&s.inv, static_cast<size_t>(total_rows) * sizeof(float));
    if (e != cudaSuccess) {
      cudaFree(s.rows);
      return fail("cudaMalloc(inv) failed: %s", cudaGetErrorString(e));
    }
  }
  c->stores[name] = s;
  *out = &c->stores[name];
  return 0;
}

// Host rows -> fp16 rows on the device, pipelined: the pages (each a contiguous [rows][128] block of fp32 or fp16 in
// ordinary host memory) are treated as one stream of values, cut into chunks of kUploadChunk; the host worker pool casts
// (numpy's astype(float16): round to nearest even) or copies a chunk into one of two pinned staging buffers while the DMA
// engine still moves the previous one. Half the PCIe bytes of an fp32 upload, no device staging buffer, no cast kernel,
// and the caller need not concatenate its documents. Enqueues on c->stream; does not synchronise at the end.
static const size_t kUploadChunk = size_t(2) << 20;   // values per staging buffer (4 MB of fp16)
static int upload_host_pages(vrag_corpus* c, __half* dst, const void* const* pages, const int64_t* page_rows, int64_t n_pages,
                             int dtype) {
  std::vector<size_t> pre(static_cast<size_t>(n_pages) + 1, 0);   // value offset of every page in the stream
  for (int64_t i = 0; i < n_pages; ++i) pre[i + 1] = pre[i] + static_cast<size_t>(page_rows[i]) * 128;
  const size_t n_el = pre[n_pages];
  if (n_el == 0) return 0;
  if (!c->h_up) {
    CUDA_OK(cudaMallocHost(&c->h_up, 2 * kUploadChunk * sizeof(uint16_t)));
    for (int i = 0; i < 2; ++i) CUDA_OK(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
  }
  const size_t esz = dtype == VRAG_F32 ? 4 : 2;
  // values [e0, e1) of the stream -> out
  auto fill = [&](size_t e0, size_t e1, uint16_t* out) {
    size_t pg = static_cast<size_t>(std::upper_bound(pre.begin(), pre.end(), e0) - pre.begin()) - 1;
    while (e0 < e1) {
      const size_t n = std::min(e1, pre[pg + 1]) - e0;
      if (n > 0) {
        const char* src = static_cast<const char*>(pages[pg]) + (e0 - pre[pg]) * esz;
        if (dtype == VRAG_F32) vrag::host_f32_to_f16(reinterpret_cast<const float*>(src), out, n);
        else memcpy(out, src, n * 2);
        out += n;
        e0 += n;
      }
      ++pg;
    }
  };
  const int threads = vrag::host_pool_threads();
  int k = 0;
  for (size_t e0 = 0; e0 < n_el; e0 += kUploadChunk, ++k) {
    const size_t e1 = std::min(n_el, e0 + kUploadChunk);
    const int b = k & 1;
    uint16_t* buf = c->h_up + b * kUploadChunk;
    CUDA_OK(cudaEventSynchronize(c->ev_up[b]));   // the copy that last read this buffer (this call or an earlier one)
    const size_t n = e1 - e0;
    const int tasks = static_cast<int>(std::min<size_t>(threads, (n + 65535) / 65536));   // >= 64k values per task
    if (tasks <= 1) {
      fill(e0, e1, buf);
    } else {
      const size_t per = ((n + tasks - 1) / tasks + 15) & ~size_t(15);
      vrag::host_parallel_for(tasks, [&](int t) {
        const size_t a = e0 + std::min(n, per * t), z = e0 + std::min(n, per * (t + 1));
        if (a < z) fill(a, z, buf + (a - e0));
      });
    }
    CUDA_OK(cudaMemcpyAsync(dst + e0, buf, n * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaEventRecord(c->ev_up[b], c->stream));
  }
  return 0;
}

// rows (host or device, fp16 or fp32) -> fp16 rows at dst + their inverse norms. fp32 is cast exactly like
// QdrantIndexer._build_qdrant_points (qdrant_indexer.py:423-441).
static int upload_rows(vrag_corpus* c, __half* dst, float* dst_inv, const void* rows, int dtype, int rows_on_device,
                       int64_t n_rows) {
  const size_t n_el = static_cast<size_t>(n_rows) * 128;
  if (!rows_on_device) {
    const void* pages[1] = {rows};
    const int64_t prow[1] = {n_rows};
    TRY(upload_host_pages(c, dst, pages, prow, 1, dtype));
  } else if (dtype == VRAG_F16) {
    // all copies go through the library stream: the kernels below run on it (a non-blocking stream does not
    // order against the legacy default stream a plain cudaMemcpy uses)
    CUDA_OK(cudaMemcpyAsync(dst, rows, n_el * sizeof(__half), cudaMemcpyDeviceToDevice, c->stream));
  } else {
    f32_to_f16_kernel<<<static_cast<unsigned>((n_el / 4 + 256) / 256), 256, 0, c->stream>>>(static_cast<const float*>(rows), n_el, dst);
    c->launches++;
  }
  const long long threads = n_rows * 16;
  inv_norm_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, c->stream>>>(dst, n_rows, dst_inv);
  c->launches++;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int vrag_store_add(vrag_corpus_t* c, const char* name, const void* rows, int dtype, int rows_on_device,
                              const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  TRY(set_device(c));
  if (dtype != VRAG_F16 && dtype != VRAG_F32) return fail("unknown dtype %d", dtype);
  int64_t total_rows = 0;
  TRY(check_layout(page_offsets, n_pages, fixed_rows, &total_rows));
  if (total_rows > 0 && !rows) return fail("rows is NULL");
  Store* s = nullptr;
  TRY(alloc_store(c, name, total_rows, &s));
  if (total_rows > 0) TRY(upload_rows(c, s->rows, s->inv, rows, dtype, rows_on_device, total_rows));
  return finish_store(c, *s, page_offsets, n_pages, fixed_rows);
}

// Grow the row / scale buffers of a store to at least `need` rows (contents preserved).
static int grow_store(vrag_corpus* c, Store& s, const char* name, int64_t need) {
  if (need <= s.cap_rows) return 0;
  if (need >= (1ll << 31)) return fail("a store holds at most 2^31-1 rows per shard (TMA coordinates are int32)");
  const int64_t cap = std::min<int64_t>((1ll << 31) - 1, std::max<int64_t>(need, s.cap_rows + s.cap_rows / 2 + 1024));
  __half* nr = nullptr;
  float* ni = nullptr;
  CUDA_OK(cudaMalloc(&nr, static_cast<size_t>(cap) * 128 * sizeof(__half)));
  if (cudaMalloc(&ni, static_cast<size_t>(cap) * sizeof(float)) != cudaSuccess) {
    cudaFree(nr);
    return fail("cudaMalloc(inv) failed while growing store '%s'", name);
  }
  if (s.total_rows > 0) {
    cudaError_t e = cudaMemcpyAsync(nr, s.rows, static_cast<size_t>(s.total_rows) * 256, cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ni, s.inv, static_cast<size_t>(s.total_rows) * sizeof(float), cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      cudaFree(nr);
      cudaFree(ni);
      return fail("copy into the grown store '%s' failed: %s", name, cudaGetErrorString(e));
    }
  }
  if (s.rows) cudaFree(s.rows);
  if (s.inv) cudaFree(s.inv);
  s.rows = nr;
  s.inv = ni;
  s.cap_rows = cap;
  s.dirty = true;   // tensor maps point at the old buffers
  return 0;
}

// Append pages to a named store (created on first use): the ingest path of QdrantIndexer.upload_batch
// (qdrant_indexer.py:341-507) — every batch of points adds its pages behind the existing ones; page index = upload
// order. Device buffers grow geometrically; page tables and tensor maps are rebuilt lazily on the next use.
extern "C" int vrag_store_append(vrag_corpus_t* c, const char* name, const void* rows, int dtype, int rows_on_device,
                                 const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return vrag_store_add(c, name, rows, dtype, rows_on_device, page_offsets, n_pages, fixed_rows);
  TRY(set_device(c));
  if (dtype != VRAG_F16 && dtype != VRAG_F32) return fail("unknown dtype %d", dtype);
  int64_t new_rows = 0;
  TRY(check_layout(page_offsets, n_pages, fixed_rows, &new_rows));
  if (n_pages == 0) return 0;
  if (new_rows > 0 && !rows) return fail("rows is NULL");
  Store& s = it->second;
  const int64_t total = s.total_rows + new_rows;
  if (total >= (1ll << 31)) return fail("a store holds at most 2^31-1 rows per shard (TMA coordinates are int32)");
  if (s.n_pages + n_pages >= (1ll << 31)) return fail("a shard holds at most 2^31-1 pages");
  // ---- page layout: stay fixed-rows only if every new page has the same row count. The new layout is built on the side
  // and committed only after the allocation and the upload have succeeded: a failed append leaves the store untouched.
  bool stay_fixed = s.fixed_rows > 0;
  if (stay_fixed) {
    if (fixed_rows > 0) stay_fixed = fixed_rows == s.fixed_rows;
    else
      for (int64_t i = 0; i < n_pages && stay_fixed; ++i) stay_fixed = (page_offsets[i + 1] - page_offsets[i]) == s.fixed_rows;
  }
  std::vector<int64_t> new_tail;     // row offsets of the appended pages (absolute), when the store is / becomes variable
  bool materialise = false;          // the existing pages' offsets have to be written out first (fixed -> variable)
  if (s.indirect) stay_fixed = false;
  if (!stay_fixed) {
    materialise = !s.indirect && (s.fixed_rows > 0 || s.h_offsets.empty());
    int64_t at = s.total_rows;
    new_tail.reserve(n_pages);
    for (int64_t i = 0; i < n_pages; ++i) {
      at += fixed_rows > 0 ? fixed_rows : (page_offsets[i + 1] - page_offsets[i]);
      new_tail.push_back(at);
    }
  }
  // ---- grow (same rows, larger buffers: the store is still consistent if the upload below fails)
  TRY(grow_store(c, s, name, total));
  if (new_rows > 0) TRY(upload_rows(c, s.rows + static_cast<size_t>(s.total_rows) * 128, s.inv + s.total_rows, rows, dtype, rows_on_device, new_rows));
  // ---- commit
  if (s.indirect) {
    int64_t at = s.total_rows;
    for (int64_t i = 0; i < n_pages; ++i) {
      s.h_begin.push_back(at);
      s.h_end.push_back(new_tail[i]);
      at = new_tail[i];
    }
  } else if (!stay_fixed) {
    if (materialise) {
      s.h_offsets.resize(s.n_pages + 1);
      for (int64_t i = 0; i <= s.n_pages; ++i) s.h_offsets[i] = i * s.fixed_rows;
    }
    s.h_offsets.insert(s.h_offsets.end(), new_tail.begin(), new_tail.end());
    s.fixed_rows = 0;
  }
  s.total_rows = total;
  s.n_pages += n_pages;
  s.dirty = true;
  return 0;
}

extern "C" int vrag_store_replace_pages(vrag_corpus_t* c, const char* name, const int64_t* local_pages, int64_t n_pages,
                                        const void* rows, int dtype, int rows_on_device, const int64_t* page_offsets,
                                        int64_t fixed_rows) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return fail("unknown vector store '%s'", name);
  TRY(set_device(c));
  if (dtype != VRAG_F16 && dtype != VRAG_F32) return fail("unknown dtype %d", dtype);
  int64_t new_rows = 0;
  TRY(check_layout(page_offsets, n_pages, fixed_rows, &new_rows));
  if (n_pages == 0) return 0;
  if (!local_pages) return fail("local_pages is NULL");
  if (new_rows > 0 && !rows) return fail("rows is NULL");
  Store& s = it->second;
  // validate everything first: either all pages are replaced or none
  bool same_shapes = true;
  int64_t extra_rows = 0;   // rows that have to go behind the last row (pages that grew)
  for (int64_t i = 0; i < n_pages; ++i) {
    const int64_t pg = local_pages[i];
    if (pg < 0 || pg >= s.n_pages) return fail("page %lld out of range (store '%s' has %lld pages)", (long long)pg, name, (long long)s.n_pages);
    int64_t r0, have;
    page_rows_h(s, pg, &r0, &have);
    const int64_t got = fixed_rows > 0 ? fixed_rows : (page_offsets[i + 1] - page_offsets[i]);
    if (have != got) same_shapes = false;
    if (got > have) extra_rows += got;
  }
  const size_t el = dtype == VRAG_F16 ? sizeof(__half) : sizeof(float);
  if (!same_shapes) {
    // a page changed its row count: the store gets a page table. A page that shrank is overwritten in place, a page
    // that grew moves behind the last row (its old rows become garbage until vrag_store_compact).
    if (s.total_rows + extra_rows >= (1ll << 31)) return fail("a store holds at most 2^31-1 rows per shard (TMA coordinates are int32)");
    TRY(grow_store(c, s, name, s.total_rows + extra_rows));
    make_indirect(s);
  }
  for (int64_t i = 0; i < n_pages; ++i) {
    const int64_t pg = local_pages[i];
    int64_t r0, have;
    page_rows_h(s, pg, &r0, &have);
    const int64_t src0 = fixed_rows > 0 ? i * fixed_rows : page_offsets[i];
    const int64_t nr = fixed_rows > 0 ? fixed_rows : (page_offsets[i + 1] - page_offsets[i]);
    if (nr > have) {          // moves to the end of the row buffer
      s.garbage_rows += have;
      r0 = s.total_rows;
      s.total_rows += nr;
    } else if (nr < have) {
      s.garbage_rows += have - nr;
    }
    if (nr != have) {
      s.h_begin[pg] = r0;
      s.h_end[pg] = r0 + nr;
      s.dirty = true;
    }
    if (nr == 0) continue;
    TRY(upload_rows(c, s.rows + static_cast<size_t>(r0) * 128, s.inv + r0,
                    static_cast<const char*>(rows) + static_cast<size_t>(src0) * 128 * el, dtype, rows_on_device, nr));
  }
  return 0;   // equal shapes: page tables, tile packing and tensor maps stay valid; otherwise rebuilt on next use (dirty)
}

// Delete pages: a deleted page keeps its index (the id tables of the host stay valid) and owns no rows, so it scores
// -inf in every scan and is never returned. Its rows are reclaimed by vrag_store_compact.
extern "C" int vrag_store_delete_pages(vrag_corpus_t* c, const char* name, const int64_t* local_pages, int64_t n_pages) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return fail("unknown vector store '%s'", name);
  if (n_pages == 0) return 0;
  if (!local_pages) return fail("local_pages is NULL");
  Store& s = it->second;
  for (int64_t i = 0; i < n_pages; ++i)
    if (local_pages[i] < 0 || local_pages[i] >= s.n_pages)
      return fail("page %lld out of range (store '%s' has %lld pages)", (long long)local_pages[i], name, (long long)s.n_pages);
  make_indirect(s);
  for (int64_t i = 0; i < n_pages; ++i) {
    const int64_t pg = local_pages[i];
    s.garbage_rows += s.h_end[pg] - s.h_begin[pg];
    s.h_end[pg] = s.h_begin[pg];
  }
  s.dirty = true;
  return 0;
}

// Drop the last pages of a store (the rollback of a multi-store batch upload whose later store failed).
extern "C" int vrag_store_truncate(vrag_corpus_t* c, const char* name, int64_t n_pages) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return fail("unknown vector store '%s'", name);
  Store& s = it->second;
  if (n_pages < 0 || n_pages > s.n_pages) return fail("cannot truncate store '%s' (%lld pages) to %lld", name, (long long)s.n_pages, (long long)n_pages);
  if (n_pages == s.n_pages) return 0;
  if (s.indirect) {
    s.h_begin.resize(n_pages);
    s.h_end.resize(n_pages);
    int64_t top = 0;
    for (int64_t p = 0; p < n_pages; ++p) top = std::max(top, s.h_end[p]);
    s.total_rows = top;
  } else if (s.fixed_rows > 0) {
    s.total_rows = n_pages * s.fixed_rows;
  } else {
    s.h_offsets.resize(n_pages + 1);
    s.total_rows = s.h_offsets[n_pages];
  }
  s.n_pages = n_pages;
  s.dirty = true;
  return 0;
}

// Rewrite the rows of a store contiguously in page order (one device pass) and return it to the dense layout: garbage
// left by replaced / deleted pages is reclaimed and the contiguous fast paths apply again. Deleted pages stay as
// zero-row pages, so page indices do not change.
static int compact_store(vrag_corpus* c, Store& s, const char* name) {
  if (!s.indirect) return 0;
  std::vector<int64_t> off(s.n_pages + 1, 0);
  for (int64_t p = 0; p < s.n_pages; ++p) off[p + 1] = off[p] + (s.h_end[p] - s.h_begin[p]);
  const int64_t total = off[s.n_pages];
  __half* nr = nullptr;
  float* ni = nullptr;
  long long *d_b = nullptr, *d_o = nullptr;
  if (total > 0) {
    CUDA_OK(cudaMalloc(&nr, static_cast<size_t>(total) * 256));
    if (cudaMalloc(&ni, static_cast<size_t>(total) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&d_b, s.n_pages * sizeof(long long)) != cudaSuccess ||
        cudaMalloc(&d_o, (s.n_pages + 1) * sizeof(long long)) != cudaSuccess) {
      cudaFree(nr);
      if (ni) cudaFree(ni);
      if (d_b) cudaFree(d_b);
      return fail("cudaMalloc failed while compacting store '%s'", name);
    }
    cudaMemcpyAsync(d_b, s.h_begin.data(), s.n_pages * sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_o, off.data(), (s.n_pages + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(s.n_pages, static_cast<int64_t>(c->num_sms) * 16));
    compact_rows_kernel<<<grid, 256, 0, c->stream>>>(s.rows, s.inv, d_b, d_o, s.n_pages, nr, ni);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_b);
    cudaFree(d_o);
    if (e != cudaSuccess) {
      cudaFree(nr);
      cudaFree(ni);
      return fail("compaction of store '%s' failed: %s", name, cudaGetErrorString(e));
    }
  }
  if (s.rows) cudaFree(s.rows);
  if (s.inv) cudaFree(s.inv);
  s.rows = nr;
  s.inv = ni;
  s.cap_rows = total;
  s.total_rows = total;
  s.indirect = false;
  s.h_begin.clear();
  s.h_end.clear();
  s.garbage_rows = 0;
  bool same = s.n_pages > 0;
  for (int64_t p = 1; p < s.n_pages && same; ++p) same = (off[p + 1] - off[p]) == off[1];
  const int64_t fixed = (same && off[1] > 0) ? off[1] : 0;
  s.h_offsets = off;
  return finish_store(c, s, fixed > 0 ? nullptr : s.h_offsets.data(), s.n_pages, fixed);
}

extern "C" int vrag_store_compact(vrag_corpus_t* c, const char* name) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (!name || !*name) return fail("store name is empty");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return fail("unknown vector store '%s'", name);
  TRY(set_device(c));
  return compact_store(c, it->second, name);
}

extern "C" int vrag_store_add_synthetic(vrag_corpus_t* c, const char* name, const int64_t* page_offsets,
                                        int64_t n_pages, int64_t fixed_rows, uint64_t seed, int64_t row_seed_base) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  TRY(set_device(c));
  int64_t total_rows = 0;
  TRY(check_layout(page_offsets, n_pages, fixed_rows, &total_rows));
  Store* s = nullptr;
  TRY(alloc_store(c, name, total_rows, &s));
  const int64_t chunk = 1ll << 24;
  for (int64_t r = 0; r < total_rows; r += chunk) {
    const int64_t n = std::min(chunk, total_rows - r);
    synth_rows_kernel<<<static_cast<unsigned>((n * 16 + 255) / 256), 256, 0, c->stream>>>(s->rows, s->inv, r, n, seed,
                                                                                          row_seed_base);
    c->launches++;
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaGetLastError());
  return finish_store(c, *s, page_offsets, n_pages, fixed_rows);
}

static int find_store(vrag_corpus* c, const char* name, Store** out) {
  if (!c) return fail("corpus is NULL");
  if (!name) return fail("store name is NULL");
  auto it = c->stores.find(name);
  if (it == c->stores.end()) return fail("unknown vector store '%s'", name);
  *out = &it->second;
  if (it->second.dirty) {   // appended to: rebuild page tables, tile packing and tensor maps once, on first use
    Store& s = it->second;
    TRY(set_device(c));
    TRY(finish_store(c, s, (s.fixed_rows > 0 || s.indirect) ? nullptr : s.h_offsets.data(), s.n_pages, s.fixed_rows));
  }
  return 0;
}

extern "C" int vrag_store_info(vrag_corpus_t* c, const char* name, int64_t* n_pages, int64_t* total_rows,
                               int64_t* fixed_rows, int64_t* max_rows) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  if (n_pages) *n_pages = s->n_pages;
  if (total_rows) *total_rows = s->total_rows;
  if (fixed_rows) *fixed_rows = s->fixed_rows;
  if (max_rows) *max_rows = s->max_rows;
  return 0;
}

extern "C" int vrag_store_page_range(vrag_corpus_t* c, const char* name, int64_t local_page, int64_t* row0,
                                     int64_t* n_rows) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  if (local_page < 0 || local_page >= s->n_pages) return fail("page %lld out of range", (long long)local_page);
  page_rows_h(*s, local_page, row0, n_rows);
  return 0;
}

extern "C" int vrag_store_page_rows(vrag_corpus_t* c, const char* name, int64_t first_page, int64_t n, int64_t* out_rows) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  if (first_page < 0 || n < 0 || first_page + n > s->n_pages) return fail("pages [%lld, %lld) out of range", (long long)first_page, (long long)(first_page + n));
  if (n > 0 && !out_rows) return fail("out_rows is NULL");
  for (int64_t i = 0; i < n; ++i) {
    int64_t r0;
    page_rows_h(*s, first_page + i, &r0, out_rows + i);
  }
  return 0;
}

extern "C" int vrag_store_read_rows(vrag_corpus_t* c, const char* name, int64_t row0, int64_t n_rows,
                                    void* out_f16_host) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > s->total_rows) return fail("row range out of bounds");
  if (n_rows == 0) return 0;
  CUDA_OK(cudaMemcpyAsync(out_f16_host, s->rows + static_cast<size_t>(row0) * 128, static_cast<size_t>(n_rows) * 256,
                          cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int vrag_store_drop(vrag_corpus_t* c, const char* name) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  free_store(*s);
  c->stores.erase(name);
  return 0;
}

// ------------------------------------------------------------------------------------------------ launches
// kind: 0 single query (QS == QP), 1 operand switching (BSW), 2 sub-queries (QP == 128, qs in {1, 32})
static int launch_scan_variant(vrag_corpus* c, const Store& s, const ScanParams& p, long long n_units, cudaStream_t st,
                               int kind, int qp_or_qs) {
  ScanLaunch L;
  L.tm_rows = p.pad_rows > 0 ? &s.tm3d : &s.tm128;
  L.tm_small = &s.tm_small;
  L.tm_scale128 = &s.ts128;
  L.tm_scale32 = &s.ts32;
  L.p = p;
  L.n_units = n_units;
  L.num_sms = c->num_sms;
  L.stream = st;
  cudaError_t e = kind == 0 ? scan_launch_single(qp_or_qs, s.packed, L)
                : kind == 1 ? scan_launch_bsw(qp_or_qs, s.packed, L)
                            : scan_launch_multi(qp_or_qs, s.packed, L);
  c->launches++;
  if (e != cudaSuccess) return fail("scan kernel launch (kind %d, %d, %s) failed: %s", kind, qp_or_qs,
                                    s.packed ? "packed" : "large", cudaGetErrorString(e));
  return 0;
}

// Work layout of one scan over store `s`: every page (d_cand == nullptr) or n_items candidate ids (per query group).
static void fill_scan_params(vrag_corpus* c, const Store& s, const long long* d_cand, int64_t n_items, int QP,
                             bool normalize, float* d_scores, ScanParams* out, long long* n_units_out,
                             bool multi = false) {
  ScanParams& p = *out;
  memset(&p, 0, sizeof(p));
  p.offsets = s.indirect ? s.d_begin : s.offsets;
  p.page_end = s.indirect ? s.d_end : nullptr;
  p.fixed_rows = s.fixed_rows;
  p.n_pages = s.n_pages;
  p.cand = d_cand;
  p.cand_base = c->page_base;
  p.n_items = n_items;
  p.scores = d_scores;
  p.use_scale = normalize ? 1 : 0;
  p.slot_rows = kTileRows;
  p.n_groups = 1;
  p.tile_stride = 1;
  long long n_units = n_items;
  if (s.packed) {
    const bool small_rows = s.max_rows <= 32;
    if (d_cand || s.indirect) {
      // slot mode (candidate lists; stores with a page table): one slot per page, fetched on its own -> segmented-butterfly
      // epilogue when slots are 32 rows
      p.slot_mode = 1;
      p.slot_rows = small_rows ? 32 : (s.max_rows <= 64 ? 64 : 128);
      const int per_tile = kTileRows / p.slot_rows;
      p.n_tiles = (n_items + per_tile - 1) / per_tile;
      if (p.slot_rows == 32 && (QP <= 32 || (multi && !d_cand))) p.shfl_rows = 32;
    } else if (s.fixed_rows > 0 && s.pad_slot > 0 && (QP <= 32 || multi) && !knob_no_pad()) {
      // odd page sizes (ColSmol's 12/13 tiles, 3, 5, ...): TMA pads every page to a power-of-two slot for free
      p.pad_rows = static_cast<int>(s.fixed_rows);
      p.shfl_rows = s.pad_slot;
      p.pages_per_tile = kTileRows / s.pad_slot;
      p.n_tiles = (s.n_pages + p.pages_per_tile - 1) / p.pages_per_tile;
    } else if (s.fixed_rows > 0) {
      p.pages_per_tile = static_cast<int>(kTileRows / s.fixed_rows);
      p.n_tiles = (s.n_pages + p.pages_per_tile - 1) / p.pages_per_tile;
      const bool pow2 = (s.fixed_rows & (s.fixed_rows - 1)) == 0;
      if (pow2 && s.fixed_rows <= 32 && (QP <= 32 || multi)) p.shfl_rows = static_cast<int>(s.fixed_rows);
    } else if (small_rows && (QP <= 32 || multi) && s.total_rows * 2 >= s.n_pages * 32) {
      // variable pages of <= 32 rows, reasonably full: fetch every page into its own 32-row slot (the over-read of
      // the next page's first rows hits L2) so that the segmented-butterfly epilogue applies
      p.slot_mode = 1;
      p.slot_rows = 32;
      p.n_tiles = (n_items + 3) / 4;
      p.shfl_rows = 32;
    } else {
      p.tile_page0 = s.tile_page0;
      p.tile_row0 = s.tile_row0;
      p.n_tiles = s.n_tiles;
    }
    n_units = p.n_tiles;
  }
  *n_units_out = n_units;
}

static int fill_neg_inf(vrag_corpus* c, float* d, int64_t n, cudaStream_t st);

// Score a store (or a candidate list) into d_scores[n_items]. All pointers are device pointers.
static int launch_scan(vrag_corpus* c, const Store& s, const float* d_query, int n_query_rows, uint32_t flags,
                       const long long* d_cand, int64_t n_cand, float* d_scores, cudaStream_t st, bool time_kernel,
                       const uint32_t* d_mask = nullptr) {
  const bool pool = (flags & VRAG_Q_POOL) != 0;
  const bool normalize = (flags & VRAG_Q_NORMALIZE) != 0;
  if (n_query_rows < 1) return fail("query has no rows");
  if (n_query_rows > kMaxQueryRows) return fail("query has %d rows; at most %d supported", n_query_rows, kMaxQueryRows);
  const int q_eff = pool ? 1 : n_query_rows;
  if (q_eff > kOperandRows) {
    // longer than one operand image: score balanced row chunks and add the partial page scores (the sum over query
    // tokens of compute_maxsim_score, pooling.py:509-512, split over chunks)
    const int64_t n = d_cand ? n_cand : s.n_pages;
    if (n == 0) return 0;
    TRY(c->d_scores_part.ensure(n));
    const int n_chunks = (q_eff + kOperandRows - 1) / kOperandRows;
    int r0 = 0;
    for (int i = 0; i < n_chunks; ++i) {
      const int rows = (q_eff - r0 + (n_chunks - i) - 1) / (n_chunks - i);
      float* dst = i == 0 ? d_scores : c->d_scores_part.p;
      TRY(launch_scan(c, s, d_query + static_cast<size_t>(r0) * 128, rows, flags, d_cand, n_cand, dst, st, time_kernel && i == 0,
                      d_mask));
      if (i > 0) {
        add_scores_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d_scores, c->d_scores_part.p, n);
        c->launches++;
      }
      r0 += rows;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  // operand width: the next power of two, except 17..24 token queries over LARGE pages, which get a 24-row operand
  // (MMA N = 48 instead of 64: a quarter less tensor work and TMEM traffic where the scan is power-limited)
  const int QP = q_eff <= 8 ? 8 : q_eff <= 16 ? 16 : (q_eff <= 24 && !s.packed) ? 24 : q_eff <= 32 ? 32 : q_eff <= 64 ? 64 : 128;
  const int64_t n_items = d_cand ? n_cand : s.n_pages;
  if (n_items == 0) return 0;
  if (s.total_rows == 0) return fill_neg_inf(c, d_scores, n_items, st);   // pages without rows (a named vector no point has yet)

  query_prep_kernel<<<QP, 128, 0, st>>>(d_query, n_query_rows, pool ? 1 : 0, normalize ? 1 : 0, QP, c->d_qimg.p);
  c->launches++;

  ScanParams p;
  long long n_units = 0;
  fill_scan_params(c, s, d_cand, n_items, QP, normalize, d_scores, &p, &n_units);
  p.qimg = c->d_qimg.p;
  p.q_valid = q_eff;
  p.mask = d_mask;
  {  // VRAG_Q_FP16 (or the experiment knob VRAG_QUERY_SPLIT=0): contract only the fp16 hi half of the query
    // (half the tensor work; LARGE pages only — that is where the scan is power/bandwidth bound)
    p.hi_only = (((flags & VRAG_Q_FP16) != 0 || knob_hi_only()) && QP >= 16 && !s.packed) ? 1 : 0;
  }
  if (time_kernel) CUDA_OK(cudaEventRecord(c->evk0, st));
  int r = launch_scan_variant(c, s, p, n_units, st, 0, QP);
  if (r) return r;
  if (time_kernel) CUDA_OK(cudaEventRecord(c->evk1, st));
  return 0;
}


// Batched candidate scan: query b (operand image b of d_qimg_batch) scores its own candidate list
// d_cand[b*n_items + i] into d_scores[b*n_items + i], all queries in ONE launch (the CTAs switch operands).
// Returns 2 (no error set) when the shape is not covered by the batched kernels; the caller then runs the
// queries one launch each.
static int launch_scan_batch(vrag_corpus* c, const Store& s, const float* d_queries, const int* d_qbegin,
                             const int* d_qend, int* d_qvalid, int nq, int max_q_eff, uint32_t flags,
                             const long long* d_cand, int64_t n_items, float* d_scores, cudaStream_t st) {
  const bool pool = (flags & VRAG_Q_POOL) != 0;
  const bool normalize = (flags & VRAG_Q_NORMALIZE) != 0;
  const int q_eff = pool ? 1 : max_q_eff;
  if (q_eff > 64 || !d_cand) return 2;
  if (n_items == 0 || nq == 0) return 0;
  if (s.total_rows == 0) return fill_neg_inf(c, d_scores, n_items * nq, st);
  const int QP = q_eff <= 32 ? 32 : 64;
  const size_t img = static_cast<size_t>(2 * QP) * 256;
  TRY(c->d_qimg_batch.ensure(img * nq));
  query_prep_batch_kernel<<<dim3(QP, nq), 128, 0, st>>>(d_queries, d_qbegin, d_qend, pool ? 1 : 0, normalize ? 1 : 0, QP,
                                                        c->d_qimg_batch.p, static_cast<long long>(img), d_qvalid);
  c->launches++;
  ScanParams p;
  long long upg = 0;
  fill_scan_params(c, s, d_cand, n_items, QP, normalize, d_scores, &p, &upg);
  p.qimg = c->d_qimg_batch.p;
  p.qimg_stride = static_cast<long long>(img);
  p.q_valid_arr = d_qvalid;
  p.n_groups = nq;
  const long long n_units = upg * nq;
  return launch_scan_variant(c, s, p, n_units, st, 1, QP);
}

// Dense batched scan: every query scores every page of the store; G = 128/QS queries share each document tile
// (QS = 1: pooled / single-row queries, 128 per launch; QS = 32: up to 32 token rows, 4 per launch).
// d_scores is [nq][n_pages]. Returns 2 when the shape is not covered (caller falls back to one launch per query).
struct DenseOpts {
  int tile_stride = 1;            // > 1: sample pass, scores written compactly as [nq][n_sample]
  int64_t n_sample = 0;           // pages scored by the sample pass
  const float* thr = nullptr;     // prefilter pass: per-query thresholds / counters / key lists
  int* cnt = nullptr;
  unsigned long long* keys = nullptr;
  int cap = 0;
  bool skip_prep = false;         // operand images are already in d_qimg_batch (second pass of the same queries)
  bool two_block = false;         // approximate first pass: 8 plain-fp16 token queries per image (LARGE stores)
  float* eps = nullptr;           //   its per-query error bounds [nq] (zeroed and filled by the operand preparation)
};
static bool dense_batch_covers(int nq, int max_q_eff, uint32_t flags) {
  const int q_eff = (flags & VRAG_Q_POOL) ? 1 : max_q_eff;
  return q_eff <= 32 && nq >= 2;
}
static int launch_scan_dense_batch(vrag_corpus* c, const Store& s, const float* d_queries, const int* d_qbegin,
                                   const int* d_qend, int* d_qvalid, int nq, int max_q_eff, uint32_t flags,
                                   float* d_scores, cudaStream_t st, bool time_kernel, const DenseOpts& o = DenseOpts()) {
  const bool pool = (flags & VRAG_Q_POOL) != 0;
  const bool normalize = (flags & VRAG_Q_NORMALIZE) != 0;
  const int q_eff = pool ? 1 : max_q_eff;
  if (q_eff > 32 || nq < 1) return 2;
  if (s.n_pages == 0) return 0;
  if (s.total_rows == 0) return 2;   // pages without rows: the per-query path fills -inf
  const int QP = 128;
  const int QS = q_eff == 1 ? 1 : 32;
  const bool two = o.two_block && QS == 32 && !s.packed;
  const int G = (two ? 2 : 1) * (QP / QS);
  const int n_img = (nq + G - 1) / G;
  const size_t img = static_cast<size_t>(2 * QP) * 256;
  if (!o.skip_prep) {
    TRY(c->d_qimg_batch.ensure(img * n_img));
    if (two && o.eps) CUDA_OK(cudaMemsetAsync(o.eps, 0, static_cast<size_t>(nq) * sizeof(float), st));
    query_prep_group_kernel<<<dim3(QP, n_img, two ? 2 : 1), 128, 0, st>>>(d_queries, d_qbegin, d_qend, nq, pool ? 1 : 0,
                                                                          normalize ? 1 : 0, QP, QS, c->d_qimg_batch.p,
                                                                          static_cast<long long>(img), d_qvalid, two ? 1 : 0,
                                                                          two ? o.eps : nullptr);
    c->launches++;
  }
  if (time_kernel) CUDA_OK(cudaEventRecord(c->evk0, st));
  const int64_t stride_cols = o.tile_stride > 1 ? o.n_sample : s.n_pages;
  for (int g = 0; g < n_img; ++g) {
    ScanParams p;
    long long n_units = 0;
    fill_scan_params(c, s, nullptr, s.n_pages, QP, normalize, d_scores ? d_scores + static_cast<size_t>(g) * G * stride_cols : nullptr,
                     &p, &n_units, true);
    p.qimg = c->d_qimg_batch.p + img * g;
    p.q_valid = QS;
    p.q_valid_arr = d_qvalid + g * G;
    p.score_stride = stride_cols;
    p.n_sub = std::min(G, nq - g * G);
    {
      p.hi_only = (two || (((flags & VRAG_Q_FP16) != 0 || knob_hi_only()) && !s.packed)) ? 1 : 0;
    }
    if (o.tile_stride > 1) {   // sample pass: every tile_stride-th page (LARGE) / full tile (PACKED fixed rows)
      p.tile_stride = o.tile_stride;
      if (s.packed) {
        const int64_t full_tiles = s.n_pages / p.pages_per_tile;
        p.n_tiles = (full_tiles + o.tile_stride - 1) / o.tile_stride;
        n_units = p.n_tiles;
      } else {
        p.n_items = (s.n_pages + o.tile_stride - 1) / o.tile_stride;
        n_units = p.n_items;
      }
    }
    if (o.thr) {
      p.f_thr = o.thr + g * G;
      p.f_cnt = o.cnt + g * G;
      p.f_keys = o.keys + static_cast<size_t>(g) * G * o.cap;
      p.f_cap = o.cap;
    }
    TRY(launch_scan_variant(c, s, p, n_units, st, 2, two ? kMultiQs8x32 : QS));
  }
  if (time_kernel) CUDA_OK(cudaEventRecord(c->evk1, st));
  return 0;
}

// Exact top-k of d_scores[batch][n] -> (out_scores[batch][k], out_ids[batch][k]) sorted descending, ties -> lower
// item index. n <= 4096: one shared-memory sort; larger: radix select of the k best keys, then sort those.
template <int CHUNK, int THREADS>
static int launch_topk_sort(vrag_corpus* c, const TopkArgs& a, int batch, cudaStream_t st) {
  auto kern = topk_kernel<CHUNK, THREADS>;
  const size_t smem = CHUNK * sizeof(unsigned long long);
  static PerDeviceOnce once;
  if (once.first()) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<dim3(1, static_cast<unsigned>(batch)), THREADS, smem, st>>>(a);
  c->launches++;
  return 0;
}

// k > kTopkMaxK (a prefetch_k / limit beyond what the shared-memory sort holds; the reference accepts any value): radix
// select of the k best keys (or every key when k >= n) into a zero-padded power-of-two list, global bitonic sort, emit.
static const int kTopkHardMaxK = 1 << 20;
static int launch_topk_big(vrag_corpus* c, const float* d_scores, const long long* d_ids, int64_t id_base, int64_t n, int k,
                           float* out_scores, long long* out_ids, int* out_pos, int* out_count, cudaStream_t st, int batch,
                           long long ids_stride, Hit* out_hits, const int* aux_src, const P2PWindow* ll_send) {
  const long long k_sel = std::min<long long>(k, n);
  long long P = kBigChunk;
  while (P < k_sel) P <<= 1;
  TRY(c->d_keys_a.ensure(static_cast<size_t>(batch) * P));
  unsigned long long* keys = c->d_keys_a.p;
  const unsigned ub = static_cast<unsigned>(batch);
  if (k_sel >= n) {
    keys_from_scores_kernel<<<dim3(static_cast<unsigned>((P + 255) / 256), ub), 256, 0, st>>>(d_scores, n, n, keys, P);
    c->launches++;
  } else {
    TRY(c->d_sel_state.ensure(static_cast<size_t>(batch) * sizeof(SelState)));
    TRY(c->d_sel_hist.ensure(static_cast<size_t>(batch) * 3 * kSelBins));
    SelArgs sa;
    sa.scores = d_scores;
    sa.n = n;
    sa.k = k;
    sa.state = reinterpret_cast<SelState*>(c->d_sel_state.p);
    sa.hist = c->d_sel_hist.p;
    sa.keys_out = keys;
    sa.keys_stride = P;
    CUDA_OK(cudaMemsetAsync(sa.hist, 0, static_cast<size_t>(batch) * 3 * kSelBins * sizeof(unsigned int), st));
    sel_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(sa.state, k, batch);
    const dim3 grid(static_cast<unsigned>((n + kSelItemsPerBlock - 1) / kSelItemsPerBlock), ub);
    for (int pass = 0; pass < 3; ++pass) {
      sel_hist_kernel<<<grid, 256, 0, st>>>(sa, pass);
      sel_scan_kernel<<<batch, 256, 0, st>>>(sa, pass);
    }
    sel_compact_kernel<<<grid, 256, 0, st>>>(sa);
    sel_ties_kernel<<<batch, 1024, 0, st>>>(sa);
    c->launches += 9;
    if (P > k_sel) {
      keys_zero_tail_kernel<<<dim3(static_cast<unsigned>((P - k_sel + 255) / 256), ub), 256, 0, st>>>(keys, k_sel, P);
      c->launches++;
    }
  }
  const dim3 cgrid(static_cast<unsigned>(P / kBigChunk), ub);
  bitonic_chunk_kernel<<<cgrid, 512, 0, st>>>(keys, P, 2, kBigChunk);
  c->launches++;
  for (long long size = 2ll * kBigChunk; size <= P; size <<= 1) {
    for (long long stride = size >> 1; stride >= kBigChunk; stride >>= 1) {
      bitonic_global_kernel<<<dim3(static_cast<unsigned>((P / 2 + 255) / 256), ub), 256, 0, st>>>(keys, P, size, stride);
      c->launches++;
    }
    bitonic_chunk_kernel<<<cgrid, 512, 0, st>>>(keys, P, size, size);
    c->launches++;
  }
  TopkArgs a;
  memset(&a, 0, sizeof(a));
  a.k = k;
  a.out_scores = out_scores;
  a.out_ids = out_ids;
  a.out_pos = out_pos;
  a.out_count = out_count;
  a.ids = d_ids;
  a.id_base = id_base;
  a.n = n;
  a.n_total = n;
  a.out_stride = k;
  a.ids_stride = ids_stride;
  a.out_hits = out_hits;
  a.aux_src = aux_src;
  if (ll_send) {
    a.ll_send = 1;
    a.ll = *ll_send;
  }
  topk_emit_kernel<<<dim3(static_cast<unsigned>((k + 255) / 256), ub), 256, 0, st>>>(a, keys, P);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_topk(vrag_corpus* c, const float* d_scores, const long long* d_ids, int64_t id_base, int64_t n,
                       int k, float* out_scores, long long* out_ids, int* out_pos, int* out_count, cudaStream_t st,
                       int batch = 1, long long ids_stride = 0, Hit* out_hits = nullptr, const int* aux_src = nullptr,
                       const P2PWindow* ll_send = nullptr) {
  if (k < 1) return fail("k must be >= 1");
  if (k > kTopkHardMaxK) return fail("k=%d exceeds the supported maximum %d", k, kTopkHardMaxK);
  if (n >= (1ll << 32) - 1) return fail("too many items for top-k");
  if (batch < 1) return fail("batch must be >= 1");
  if (k > kTopkMaxK) return launch_topk_big(c, d_scores, d_ids, id_base, n, k, out_scores, out_ids, out_pos, out_count, st, batch,
                                            ids_stride, out_hits, aux_src, ll_send);
  int k2 = 1;
  while (k2 < k) k2 <<= 1;
  TopkArgs a;
  memset(&a, 0, sizeof(a));
  a.k = k;
  a.out_scores = out_scores;
  a.out_ids = out_ids;
  a.out_pos = out_pos;
  a.out_count = out_count;
  a.ids = d_ids;
  a.id_base = id_base;
  a.n_total = n;
  a.out_stride = k;
  a.ids_stride = ids_stride;
  a.out_hits = out_hits;
  a.aux_src = aux_src;
  if (ll_send) {
    a.ll_send = 1;
    a.ll = *ll_send;
  }
  long long m = n;   // elements the final sort sees
  if (n > 4096) {
    // ---- radix select: leaves exactly k keys per query in d_keys_a
    TRY(c->d_sel_state.ensure(static_cast<size_t>(batch) * sizeof(SelState)));
    TRY(c->d_sel_hist.ensure(static_cast<size_t>(batch) * 3 * kSelBins));
    TRY(c->d_keys_a.ensure(static_cast<size_t>(batch) * k));
    SelArgs sa;
    sa.scores = d_scores;
    sa.n = n;
    sa.k = k;
    sa.state = reinterpret_cast<SelState*>(c->d_sel_state.p);
    sa.hist = c->d_sel_hist.p;
    sa.keys_out = c->d_keys_a.p;
    sa.keys_stride = k;
    CUDA_OK(cudaMemsetAsync(sa.hist, 0, static_cast<size_t>(batch) * 3 * kSelBins * sizeof(unsigned int), st));
    sel_init_kernel<<<(batch + 127) / 128, 128, 0, st>>>(sa.state, k, batch);
    const dim3 grid(static_cast<unsigned>((n + kSelItemsPerBlock - 1) / kSelItemsPerBlock), static_cast<unsigned>(batch));
    for (int pass = 0; pass < 3; ++pass) {
      sel_hist_kernel<<<grid, 256, 0, st>>>(sa, pass);
      sel_scan_kernel<<<batch, 256, 0, st>>>(sa, pass);
    }
    sel_compact_kernel<<<grid, 256, 0, st>>>(sa);
    sel_ties_kernel<<<batch, 1024, 0, st>>>(sa);
    c->launches += 9;
    a.scores = nullptr;
    a.keys_in = c->d_keys_a.p;
    a.in_stride = k;
    m = k;
  } else {
    a.scores = d_scores;
    a.in_stride = n;
  }
  a.n = m;
  const int chunk = m <= 1024 ? 1024 : (m <= 2048 ? 2048 : 8192);
  a.k2 = std::min(k2, chunk);
  a.keys_out = nullptr;
  if (chunk == 8192) TRY((launch_topk_sort<8192, 1024>(c, a, batch, st)));
  else if (chunk == 2048) TRY((launch_topk_sort<2048, 1024>(c, a, batch, st)));
  else TRY((launch_topk_sort<1024, 512>(c, a, batch, st)));
  CUDA_OK(cudaGetLastError());
  return 0;
}

// Sorted top-k of per-query key lists keys[batch][cap] holding n_dyn[b] valid keys each (cap <= 8192).
static int launch_topk_keys(vrag_corpus* c, const unsigned long long* keys, const int* n_dyn, int cap, int k,
                            int64_t id_base, float* out_scores, long long* out_ids, cudaStream_t st, int batch,
                            const long long* ids = nullptr, int* out_count = nullptr, int* fail_flag = nullptr, int need = 0,
                            Hit* out_hits = nullptr, const int* aux_src = nullptr, const P2PWindow* ll_send = nullptr) {
  int k2 = 1;
  while (k2 < k) k2 <<= 1;
  TopkArgs a;
  memset(&a, 0, sizeof(a));
  if (ll_send) {
    a.ll_send = 1;
    a.ll = *ll_send;
  }
  a.k = k;
  a.keys_in = keys;
  a.in_stride = cap;
  a.n = cap;
  a.n_total = cap;
  a.n_dyn = n_dyn;
  a.out_scores = out_scores;
  a.out_ids = out_ids;
  a.id_base = id_base;
  a.ids = ids;
  a.out_count = out_count;
  a.fail_flag = fail_flag;
  a.need = need;
  a.out_stride = k;
  a.out_hits = out_hits;
  a.aux_src = aux_src;
  const int chunk = cap <= 1024 ? 1024 : (cap <= 2048 ? 2048 : 8192);
  a.k2 = std::min(k2, chunk);
  if (chunk == 8192) TRY((launch_topk_sort<8192, 1024>(c, a, batch, st)));
  else if (chunk == 2048) TRY((launch_topk_sort<2048, 1024>(c, a, batch, st)));
  else TRY((launch_topk_sort<1024, 512>(c, a, batch, st)));
  CUDA_OK(cudaGetLastError());
  return 0;
}

// Sampled top-k of ONE large score array (single-query searches: the radix select costs ~70 us of launches and
// match_any histograms per call, 9 % of a two-stage query): a threshold from a strided sample of the scores
// (prefilter_sample_thr_kernel), one compaction pass that keeps the keys above it, and a sort of the ~1k survivors.
// Exact whenever between k and cap keys survive (checked on the device; otherwise the caller repeats the search with
// the radix select) — ties at the k-th score are all above the threshold, so the key order still breaks them by index.
struct SampledPlan {
  bool on = false;
  long long stride = 1, n_sample = 0;
  int m = 0, cap = 0;
};
static SampledPlan plan_sampled_topk(int64_t n, int k) {
  SampledPlan pl;
  if (env_flag_is("VRAG_SAMPLED_TOPK", '0')) return pl;
  if (n < 8192 || k > 1024 || k >= n) return pl;
  const long long want = std::max<long long>(8192, (24ll * n + k - 1) / k);   // >= 24 expected sample hits above the k-th score
  double bound;
  if (n <= 65536 || want * 2 > n) {   // small arrays: "sample" everything, the threshold is the k-th score's histogram bin
    pl.stride = 1;
    pl.n_sample = n;
    pl.m = k;
    bound = k + 64.0 + n / 512.0;
  } else {
    const long long take = std::min<long long>(want, kThrCacheMax);   // the estimate kernel caches its sample in shared memory
    pl.stride = (n + take - 1) / take;
    pl.n_sample = (n + pl.stride - 1) / pl.stride;
    const double ratio = static_cast<double>(n) / pl.n_sample, target = k / ratio;
    // 4-sigma margins on both sides (a miss, ~3e-5 per query, costs one exact redo)
    int m = static_cast<int>(target) + 1;
    while (m - 4.0 * sqrt(static_cast<double>(m)) < target) ++m;
    pl.m = m;
    bound = m * ratio * (1.0 + 3.0 / sqrt(static_cast<double>(m)));   // list capacity: +3 sigma
  }
  pl.cap = bound <= 1024 ? 1024 : bound <= 2048 ? 2048 : bound <= 8192 ? 8192 : 0;
  pl.on = pl.cap > 0;
  return pl;
}
static int launch_topk_sampled(vrag_corpus* c, const float* d_scores, const long long* d_ids, int64_t id_base, int64_t n, int k,
                               float* out_scores, long long* out_ids, int* out_count, int* d_fail_flag, cudaStream_t st,
                               const SampledPlan& pl, Hit* out_hits = nullptr, const P2PWindow* ll_send = nullptr) {
  TRY(c->d_skeys.ensure(pl.cap));
  TRY(c->d_sthr.ensure(1));
  TRY(c->d_sstate.ensure(2));
  if (pl.n_sample <= kThrCacheMax) {
    static PerDeviceOnce once;
    if (once.first())
      CUDA_OK(cudaFuncSetAttribute(prefilter_sample_thr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThrCacheMax * 4));
    prefilter_sample_thr_kernel<true><<<1, 1024, static_cast<size_t>(pl.n_sample) * 4, st>>>(d_scores, pl.n_sample, pl.m, c->d_sthr.p,
                                                                                             c->d_sstate.p, pl.stride);
  } else {
    prefilter_sample_thr_kernel<false><<<1, 1024, 0, st>>>(d_scores, pl.n_sample, pl.m, c->d_sthr.p, c->d_sstate.p, pl.stride);
  }
  const long long vecs = n / 4 + 1;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(c->num_sms * 8ll, (vecs + 255) / 256));
  topk_compact_thr_kernel<<<grid, 256, 0, st>>>(d_scores, n, c->d_sthr.p, c->d_sstate.p, c->d_skeys.p, pl.cap);
  c->launches += 2;
  c->sampled_runs++;
  // the sort kernel also checks the survivor count (k <= count <= cap) and raises the flag otherwise
  return launch_topk_keys(c, c->d_skeys.p, c->d_sstate.p, pl.cap, k, id_base, out_scores, out_ids, st, 1, d_ids, out_count,
                          d_fail_flag, static_cast<int>(std::min<int64_t>(k, n)), out_hits, nullptr, ll_send);
}

// ------------------------------------------------------------------------------------------------ multi-GPU exchange
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int nccl_load() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.lib) return 0;
  void* h = nullptr;
  // a bare soname: if the process already loaded a NCCL (torch's bundled one), the dynamic loader hands that one back
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail("NCCL not found (dlopen libnccl.so.2: %s)", dlerror());
  NcclApi api;
  api.lib = h;
#define VRAG_SYM(field, sym)                                                   \
  *reinterpret_cast<void**>(&api.field) = dlsym(h, sym);                       \
  if (!api.field) return fail("NCCL symbol %s not found", sym);
  VRAG_SYM(GetUniqueId, "ncclGetUniqueId")
  VRAG_SYM(CommInitRank, "ncclCommInitRank")
  VRAG_SYM(CommDestroy, "ncclCommDestroy")
  VRAG_SYM(AllGather, "ncclAllGather")
  VRAG_SYM(AllReduce, "ncclAllReduce")
  VRAG_SYM(GetErrorString, "ncclGetErrorString")
#undef VRAG_SYM
  g_nccl = api;
  return 0;
}
#define NCCL_OK(expr)                                                                                       \
  do {                                                                                                      \
    int _r = (expr);                                                                                        \
    if (_r != 0) return fail("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

static void comm_release(vrag_corpus* c) {
  CommState& cm = c->comm;
  if (cm.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(cm.nccl);
  cm.nccl = nullptr;
  cm.send.release();
  cm.recv.release();
  cm.raw.release();
  cm.m_scores.release();
  cm.m_ids.release();
  cm.raw_c.release();
  cm.own_ids.release();
  cm.own_pos.release();
  cm.own_cnt.release();
  for (auto& e : cm.ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  for (int r = 0; r < kP2PMaxRanks; ++r) {
    if (cm.peer_win[r] && cm.peer_win[r] != cm.win) cudaIpcCloseMemHandle(cm.peer_win[r]);
    cm.peer_win[r] = nullptr;
  }
  if (cm.win) cudaFree(cm.win);
  if (cm.p2p_ctr) cudaFree(cm.p2p_ctr);
  cm.win = nullptr;
  cm.p2p_ctr = nullptr;
  cm.p2p = false;
  cm.p2p_epoch = 0;
  cm.rank = 0;
  cm.nranks = 1;
}

// Map every rank's exchange window into this process (CUDA IPC; handles travel through one NCCL all-gather). Any failure
// on any rank (no peer access, IPC not permitted in this container, VRAG_P2P=0) leaves ALL ranks on the NCCL collectives:
// the decision is agreed with a min-all-reduce.
static const size_t kP2PRegionBytes = size_t(4) << 20;
static const size_t kLLMaxPayload = size_t(256) << 10;   // messages up to this size travel as LL lines (2 x the bytes on the wire)
static size_t p2p_ll_offset(int R) { return (2 * static_cast<size_t>(R) * kP2PRegionBytes + 2 * kP2PMaxRanks * sizeof(unsigned) + 255) & ~size_t(255); }
static int comm_setup_p2p(vrag_corpus* c) {
  CommState& cm = c->comm;
  const int R = cm.nranks;
  int ok = (R <= kP2PMaxRanks && !env_flag_is("VRAG_P2P", '0')) ? 1 : 0;
  const size_t win_bytes = p2p_ll_offset(R) + 2 * static_cast<size_t>(R) * (2 * kLLMaxPayload);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && (cudaMalloc(&cm.win, win_bytes) != cudaSuccess || cudaMalloc(&cm.p2p_ctr, sizeof(unsigned)) != cudaSuccess)) ok = 0;
  if (ok && (cudaMemset(cm.win, 0, win_bytes) != cudaSuccess || cudaMemset(cm.p2p_ctr, 0, sizeof(unsigned)) != cudaSuccess)) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, cm.win) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // exchange handles (+ this rank's verdict so far) through NCCL
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Msg) % 16 == 0, "message size");
  Msg* d_msgs = nullptr;
  CUDA_OK(cudaMalloc(&d_msgs, sizeof(Msg) * (R + 1)));
  Msg m;
  memset(&m, 0, sizeof(m));
  m.h = mine;
  m.ok = ok;
  CUDA_OK(cudaMemcpyAsync(d_msgs + R, &m, sizeof(Msg), cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(g_nccl.AllGather(d_msgs + R, d_msgs, sizeof(Msg), kNcclInt8, cm.nccl, c->stream));
  std::vector<Msg> all(R);
  CUDA_OK(cudaMemcpyAsync(all.data(), d_msgs, sizeof(Msg) * R, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < R; ++r) ok &= all[r].ok;
  if (ok) {
    for (int r = 0; r < R && ok; ++r) {
      if (r == cm.rank) { cm.peer_win[r] = cm.win; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
      cm.peer_win[r] = static_cast<uint8_t*>(p);
    }
    cudaGetLastError();
  }
  // second agreement: every rank could map every window
  float* d_flag = reinterpret_cast<float*>(d_msgs);
  const float mine_ok = ok ? 1.0f : 0.0f;
  CUDA_OK(cudaMemcpyAsync(d_flag, &mine_ok, sizeof(float), cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(g_nccl.AllReduce(d_flag, d_flag, 1, kNcclFloat32, kNcclMin, cm.nccl, c->stream));
  float all_ok = 0.0f;
  CUDA_OK(cudaMemcpyAsync(&all_ok, d_flag, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  cudaFree(d_msgs);
  cm.p2p = all_ok == 1.0f;
  cm.p2p_cap = kP2PRegionBytes;
  cm.p2p_epoch = 0;
  if (const char* e = getenv("VRAG_P2P_TIMEOUT_S")) {
    const double sec = atof(e);
    if (sec > 0.0) cm.p2p_timeout_ns = static_cast<unsigned long long>(sec * 1e9);
  }
  if (!cm.p2p) {
    for (int r = 0; r < kP2PMaxRanks; ++r) {
      if (cm.peer_win[r] && cm.peer_win[r] != cm.win) cudaIpcCloseMemHandle(cm.peer_win[r]);
      cm.peer_win[r] = nullptr;
    }
    if (cm.win) cudaFree(cm.win);
    if (cm.p2p_ctr) cudaFree(cm.p2p_ctr);
    cm.win = nullptr;
    cm.p2p_ctr = nullptr;
    cudaGetLastError();
  }
  return 0;
}
static P2PWindow p2p_window(vrag_corpus* c) {
  CommState& cm = c->comm;
  P2PWindow w;
  for (int r = 0; r < kP2PMaxRanks; ++r) w.win[r] = cm.peer_win[r];
  w.cap = cm.p2p_cap;
  w.me = cm.rank;
  w.R = cm.nranks;
  w.epoch = ++cm.p2p_epoch;
  w.ctr = cm.p2p_ctr;
  w.timeout_ns = cm.p2p_timeout_ns;
  w.ll_off = p2p_ll_offset(cm.nranks);
  w.ll_cap = 2 * kLLMaxPayload;
  return w;
}
// Blocks of an exchange kernel: every block sends AND waits, so all of them must be resident at once — the kernel runs alone
// on its stream's turn (256 threads, no shared memory: one block per SM always fits), so up to 128 of the 148 SMs; small
// messages take one block per 256 16-byte words.
static unsigned p2p_blocks(long long n16) { return static_cast<unsigned>(std::max<long long>(1, std::min<long long>(128, (n16 + 255) / 256))); }

extern "C" int vrag_comm_unique_id(void* out_id128) {
  if (!out_id128) return fail("out_id128 is NULL");
  TRY(nccl_load());
  NcclId id;
  NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof(id));
  return 0;
}

extern "C" int vrag_comm_init(vrag_corpus_t* c, int rank, int nranks, const void* unique_id128) {
  VRAG_LOCK(c);
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("rank %d / nranks %d out of range", rank, nranks);
  if (c->comm.nccl) return fail("the handle already has a communicator");
  TRY(set_device(c));
  if (nranks > 1) {
    if (!unique_id128) return fail("unique_id128 is NULL");
    TRY(nccl_load());
    NcclId id;
    memcpy(&id, unique_id128, sizeof(id));
    NCCL_OK(g_nccl.CommInitRank(&c->comm.nccl, nranks, id, rank));
    for (auto& e : c->comm.ev) CUDA_OK(cudaEventCreate(&e));
  }
  c->comm.rank = rank;
  c->comm.nranks = nranks;
  if (nranks > 1) TRY(comm_setup_p2p(c));
  return 0;
}

extern "C" int vrag_comm_info(vrag_corpus_t* c, int* rank, int* nranks) {
  VRAG_LOCK(c);
  if (rank) *rank = c->comm.rank;
  if (nranks) *nranks = c->comm.nranks;
  return 0;
}

extern "C" int vrag_comm_transport(vrag_corpus_t* c, int* peer_memory) {
  VRAG_LOCK(c);
  if (peer_memory) *peer_memory = c->comm.p2p ? 1 : 0;
  return 0;
}

extern "C" int vrag_comm_destroy(vrag_corpus_t* c) {
  VRAG_LOCK(c);
  TRY(set_device(c));
  cudaStreamSynchronize(c->stream);
  comm_release(c);
  return 0;
}

extern "C" int vrag_last_comm_timing(vrag_corpus_t* c, float* out_us, int capacity, int* n) {
  VRAG_LOCK(c);
  if (!n) return fail("n is NULL");
  *n = c->comm.last_n;
  for (int i = 0; i < c->comm.last_n && i < capacity && out_us; ++i) out_us[i] = c->comm.last_us[i];
  return 0;
}

extern "C" int vrag_last_comm_offsets(vrag_corpus_t* c, float* out_begin_us, int capacity, int* n) {
  VRAG_LOCK(c);
  if (!n) return fail("n is NULL");
  *n = c->comm.last_n;
  for (int i = 0; i < c->comm.last_n && i < capacity && out_begin_us; ++i) out_begin_us[i] = c->comm.last_begin_us[i];
  return 0;
}

static_assert(sizeof(Hit) == sizeof(vrag_hit_t) && sizeof(Hit) == 16, "packed entries are 16 bytes on both sides of the ABI");
static bool sharded(const vrag_corpus* c) { return c->comm.nranks > 1; }
// event pair around a collective of a host-facing search (device time of the exchange, reported by vrag_last_comm_timing)
static void comm_mark(vrag_corpus* c, cudaStream_t st, bool begin) {
  CommState& cm = c->comm;
  if (!cm.timing || cm.n_ev >= kMaxCommEvents) return;
  cudaEventRecord(cm.ev[2 * cm.n_ev + (begin ? 0 : 1)], st);
  if (!begin) cm.n_ev++;
}
static void comm_collect_timing(vrag_corpus* c) {   // after the stream was synchronised
  CommState& cm = c->comm;
  cm.last_n = cm.n_ev;
  for (int i = 0; i < cm.n_ev; ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, cm.ev[2 * i], cm.ev[2 * i + 1]);
    cm.last_us[i] = ms * 1e3f;
    float off = 0.0f;
    if (cudaEventElapsedTime(&off, c->ev0, cm.ev[2 * i]) != cudaSuccess) {
      off = 0.0f;
      cudaGetLastError();
    }
    cm.last_begin_us[i] = off * 1e3f;
  }
  cm.n_ev = 0;
}

// gathered[r][list][k] <- rank r's local[list][k]: ONE collective for all lists of a stage
static int comm_allgather_hits(vrag_corpus* c, const Hit* local, int n_lists, int k, Hit* gathered, cudaStream_t st) {
  const size_t bytes = static_cast<size_t>(n_lists) * k * sizeof(Hit);
  if (!sharded(c)) {
    if (gathered != local) CUDA_OK(cudaMemcpyAsync(gathered, local, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  comm_mark(c, st, true);
  if (c->comm.p2p && bytes <= c->comm.p2p_cap && bytes > 0) {
    // one kernel: this rank's lists -> every rank's window over NVLink, wait for the others, copy out. Small messages as
    // LL lines (flag in the data: no fence), larger ones as plain stores + system fence + flag
    const long long n16 = static_cast<long long>(bytes / sizeof(uint4));
    if (bytes <= kLLMaxPayload)
      p2p_ll_allgather_kernel<<<p2p_blocks(n16), 256, 0, st>>>(p2p_window(c), reinterpret_cast<const uint4*>(local), n16,
                                                              reinterpret_cast<uint4*>(gathered));
    else
      p2p_allgather_kernel<<<p2p_blocks(n16), 256, 0, st>>>(p2p_window(c), reinterpret_cast<const uint4*>(local), n16,
                                                           reinterpret_cast<uint4*>(gathered));
    c->launches++;
    CUDA_OK(cudaGetLastError());
  } else {
    NCCL_OK(g_nccl.AllGather(local, gathered, bytes, kNcclInt8, c->comm.nccl, st));
  }
  comm_mark(c, st, false);
  return 0;
}
static int comm_allreduce_max(vrag_corpus* c, float* buf, int64_t n, cudaStream_t st) {
  if (!sharded(c) || n == 0) return 0;
  comm_mark(c, st, true);
  if (c->comm.p2p && static_cast<size_t>(n) * sizeof(float) <= c->comm.p2p_cap) {
    if (static_cast<size_t>(n) * sizeof(float) <= kLLMaxPayload)
      p2p_ll_allreduce_max_kernel<<<p2p_blocks((n + 1) >> 1), 256, 0, st>>>(p2p_window(c), buf, n);
    else
      p2p_allreduce_max_kernel<<<p2p_blocks(n >> 2), 256, 0, st>>>(p2p_window(c), buf, n);
    c->launches++;
    CUDA_OK(cudaGetLastError());
  } else {
    NCCL_OK(g_nccl.AllReduce(buf, buf, static_cast<size_t>(n), kNcclFloat32, kNcclMax, c->comm.nccl, st));
  }
  comm_mark(c, st, false);
  return 0;
}

// Merge gathered lists hits[src][list][k_src] -> global top-k per list (ties -> lower source rank = lower global id).
static int merge_hits(vrag_corpus* c, const Hit* hits, int n_src, int n_lists, int k_src, int k, float* out_scores,
                      long long* out_ids, int* fail_flag, cudaStream_t st, const P2PWindow* ll_recv = nullptr,
                      bool lists_sorted = false) {
  if (k < 1 || k > kTopkHardMaxK) return fail("k=%d out of range [1,%d]", k, kTopkHardMaxK);
  long long n = static_cast<long long>(n_src) * k_src;
  if (ll_recv && n > 8192) return fail("internal: fused exchange needs lists that fit one merge block");
  if (n <= 8192) {
    int k2 = 1;
    while (k2 < k) k2 <<= 1;
    // lists that are sorted already (a rank's local top-k): one run per list, no run-sorting stages
    const bool presorted = lists_sorted && k_src <= k2 && static_cast<long long>(n_src) * k2 <= 8192 &&
                           !env_flag_is("VRAG_MERGE_PRESORTED", '0');
    TopkArgs a;
    memset(&a, 0, sizeof(a));
    a.k = k;
    a.hits_in = hits;
    if (ll_recv) {
      a.hits_in = nullptr;
      a.ll_recv = 1;
      a.ll = *ll_recv;
    }
    a.hits_k_src = k_src;
    a.hits_rank_stride = static_cast<long long>(n_lists) * k_src;
    a.n = n;
    a.n_total = n;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.out_stride = k;
    a.fail_flag = fail_flag;
    if (presorted) {
      a.presorted_src = n_src;
      n = static_cast<long long>(n_src) * k2;   // slots of the sort buffer in use
    }
    const int chunk = n <= 1024 ? 1024 : (n <= 2048 ? 2048 : 8192);
    a.k2 = std::min(k2, chunk);
    if (chunk == 8192) TRY((launch_topk_sort<8192, 1024>(c, a, n_lists, st)));
    else if (chunk == 2048) TRY((launch_topk_sort<2048, 1024>(c, a, n_lists, st)));
    else TRY((launch_topk_sort<1024, 512>(c, a, n_lists, st)));
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  // very long lists (k > 1024 on 8 ranks): unpack and take the radix-select path
  const size_t tot = static_cast<size_t>(n) * n_lists;
  TRY(c->comm.m_scores.ensure(tot));
  TRY(c->comm.m_ids.ensure(tot));
  hits_unpack_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(hits, n_src, n_lists, k_src, c->comm.m_scores.p,
                                                                             c->comm.m_ids.p, fail_flag);
  c->launches++;
  return launch_topk(c, c->comm.m_scores.p, c->comm.m_ids.p, 0, n, k, out_scores, out_ids, nullptr, nullptr, st, n_lists, n);
}

// ------------------------------------------------------------------------------------------------ host-facing search
static int stage_query(vrag_corpus* c, const float* query, int n_query_rows) {
  if (!query) return fail("query is NULL");
  if (n_query_rows < 1 || n_query_rows > kMaxQueryRows) return fail("query rows %d out of range [1,%d]", n_query_rows, kMaxQueryRows);
  memcpy(c->h_query, query, static_cast<size_t>(n_query_rows) * 128 * sizeof(float));
  CUDA_OK(cudaMemcpyAsync(c->d_query.p, c->h_query, static_cast<size_t>(n_query_rows) * 128 * sizeof(float),
                          cudaMemcpyHostToDevice, c->stream));
  return 0;
}

extern "C" int vrag_score(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, uint32_t flags,
                          const int64_t* cand_ids, int64_t n_cand, float* out_scores) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  const int64_t n_items = cand_ids ? n_cand : s->n_pages;
  if (n_items == 0) return 0;
  if (!out_scores) return fail("out_scores is NULL");
  TRY(stage_query(c, query, n_query_rows));
  if (cand_ids) {
    TRY(c->d_cand.ensure(n_cand));
    CUDA_OK(cudaMemcpyAsync(c->d_cand.p, cand_ids, n_cand * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  }
  TRY(c->d_scores.ensure(n_items));
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  TRY(launch_scan(c, *s, c->d_query.p, n_query_rows, flags, cand_ids ? c->d_cand.p : nullptr, n_cand, c->d_scores.p,
                  c->stream, true));
  CUDA_OK(cudaEventRecord(c->ev1, c->stream));
  CUDA_OK(cudaMemcpyAsync(out_scores, c->d_scores.p, n_items * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1);
  cudaEventElapsedTime(&c->last_ms[1], c->evk0, c->evk1);
  return 0;
}

extern "C" int vrag_host_f32_to_f16(const float* src, uint16_t* dst, int64_t n, int threads, int force_scalar) {
  if (n < 0 || (n > 0 && (!src || !dst))) return fail("bad arguments");
  if (threads <= 1 || n < (1 << 17)) {
    vrag::host_f32_to_f16(src, dst, static_cast<size_t>(n), force_scalar != 0);
    return 0;
  }
  const int tasks = std::min<int64_t>(threads, (n + 65535) / 65536);
  const int64_t per = ((n + tasks - 1) / tasks + 15) & ~int64_t(15);
  vrag::host_parallel_for(tasks, [&](int t) {
    const int64_t a = std::min<int64_t>(n, per * t), z = std::min<int64_t>(n, per * (t + 1));
    if (a < z) vrag::host_f32_to_f16(src + a, dst + a, static_cast<size_t>(z - a), force_scalar != 0);
  });
  return 0;
}

// MaxSim of one query against documents that live in ordinary host memory, one pointer per document: the per-call twins
// of the reference (compute_maxsim_score / compute_maxsim_batch, pooling.py:468-552). The documents go through the
// pipelined host upload into a scratch store of the handle (buffers kept between calls) and are scored by one scan.
extern "C" int vrag_score_pages(vrag_corpus_t* c, const float* query, int n_query_rows, uint32_t flags,
                                const void* const* pages, const int64_t* page_rows, int64_t n_pages, int dtype,
                                float* out_scores) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  TRY(set_device(c));
  if (dtype != VRAG_F16 && dtype != VRAG_F32) return fail("unknown dtype %d", dtype);
  if (n_pages < 0) return fail("n_pages < 0");
  if (n_pages == 0) return 0;
  if (!pages || !page_rows || !out_scores) return fail("NULL argument");
  std::vector<int64_t> offs(static_cast<size_t>(n_pages) + 1, 0);
  for (int64_t i = 0; i < n_pages; ++i) {
    if (page_rows[i] < 0) return fail("page %lld has a negative row count", (long long)i);
    if (page_rows[i] > 0 && !pages[i]) return fail("page %lld is NULL", (long long)i);
    offs[i + 1] = offs[i] + page_rows[i];
  }
  int64_t total_rows = 0;
  TRY(check_layout(offs.data(), n_pages, 0, &total_rows));
  static const char* kScratch = "__score_pages__";
  Store* s = nullptr;
  TRY(alloc_store(c, kScratch, total_rows, &s));
  if (total_rows > 0) {
    TRY(upload_host_pages(c, s->rows, pages, page_rows, n_pages, dtype));
    const long long threads = total_rows * 16;
    inv_norm_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, c->stream>>>(s->rows, total_rows, s->inv);
    c->launches++;
  }
  // equal row counts (always so for a single document): the fixed-rows layout needs no offset table on the device
  bool uniform = page_rows[0] > 0;
  for (int64_t i = 1; i < n_pages && uniform; ++i) uniform = page_rows[i] == page_rows[0];
  if (uniform) TRY(finish_store(c, *s, nullptr, n_pages, page_rows[0]));
  else TRY(finish_store(c, *s, offs.data(), n_pages, 0));
  return vrag_score(c, kScratch, query, n_query_rows, flags, nullptr, 0, out_scores);
}

// Scores of candidates when this shard's store has no rows at all: every candidate is foreign -> -inf.
static int fill_neg_inf(vrag_corpus* c, float* d, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d, n, -INFINITY);
  c->launches++;
  return 0;
}

// All stages of one query on the device, enqueued on `stm`, no synchronisation. d_query: device query rows; d_cand:
// optional stage-0 candidate ids (rank-local when the handle is sharded). Stage s writes ks[s] entries at
// [sum(ks[:s]), ...) of d_out_scores / d_out_ids (unused slots (-inf, -1)) and, single shard only, its valid count to
// d_counts[s]. d_fail (may be null unless allow_sampled): OR-ed with 1 when a sampled top-k estimate missed on any rank.
// Collective when sharded (see include/vrag_b200.h): one all-gather of packed local lists per scanning stage, one
// max-all-reduce per stage over replicated candidates.
static int run_stages(vrag_corpus* c, int n_stages, Store* const* st, const uint32_t* flags, const int* ks,
                      const float* d_query, int n_query_rows, const int* q_offsets, const long long* d_cand, int64_t n_cand,
                      float* d_out_scores, long long* d_out_ids, int* d_counts, int* d_fail, bool allow_sampled,
                      cudaStream_t stm, bool* timed_out, bool* used_sampled_out, const uint32_t* d_mask = nullptr) {
  const bool sh = sharded(c);
  const int64_t n_first = d_cand ? n_cand : st[0]->n_pages;
  int64_t max_k = 0;
  for (int s = 0; s < n_stages; ++s) max_k = std::max<int64_t>(max_k, ks[s]);
  TRY(c->d_scores.ensure(std::max<int64_t>(std::max(n_first, max_k), 1)));
  if (sh) {
    TRY(c->comm.send.ensure(max_k));
    TRY(c->comm.recv.ensure(max_k * c->comm.nranks));
  }
  size_t off = 0;
  int64_t n_prev = n_first;   // items entering the stage
  const long long* d_prev_ids = d_cand;
  bool timed = false, used_sampled = false;
  for (int s = 0; s < n_stages; ++s) {
    const int64_t n_items = n_prev;
    const float* dq = d_query + (q_offsets ? static_cast<size_t>(q_offsets[s]) * 128 : 0);
    const int qrows = q_offsets ? (q_offsets[s + 1] - q_offsets[s]) : n_query_rows;
    const bool replicated = sh && s > 0;   // candidates = the previous stage's merged list, identical on every rank
    if (n_items > 0) {
      if (st[s]->total_rows == 0 && d_prev_ids) {
        TRY(fill_neg_inf(c, c->d_scores.p, n_items, stm));
      } else if (st[s]->n_pages > 0 || d_prev_ids) {
        // the dominant (timed) kernel is the first scan
        // a payload-filter bitmask applies to every stage: when fewer pages pass than a stage keeps, its list is padded with
        // filtered-out pages (score -inf), which must stay -inf in the stages behind it
        TRY(launch_scan(c, *st[s], dq, qrows, flags[s], d_prev_ids, n_items, c->d_scores.p, stm, !timed && timed_out, d_mask));
        timed = true;
      }
    }
    float* o_sc = d_out_scores + off;
    long long* o_id = d_out_ids + off;
    if (replicated) {
      TRY(comm_allreduce_max(c, c->d_scores.p, n_items, stm));
      TRY(launch_topk(c, c->d_scores.p, d_prev_ids, c->page_base, n_items, ks[s], o_sc, o_id, nullptr, nullptr, stm));
    } else {
      const SampledPlan sp = (allow_sampled && n_items > 0) ? plan_sampled_topk(n_items, ks[s]) : SampledPlan();
      // peer-memory transport and lists that fit one merge block: the exchange is fused into the two top-k kernels — the
      // local top-k's last kernel stores its packed entries as LL lines into every rank's window, the merge kernel polls
      // them while it loads its keys (the exchange costs no launch of its own)
      const bool fused = sh && c->comm.p2p && !env_flag_is("VRAG_P2P_FUSED", '0') &&
                         static_cast<long long>(c->comm.nranks) * ks[s] <= 8192 &&
                         static_cast<size_t>(ks[s]) * sizeof(Hit) <= kLLMaxPayload;
      P2PWindow win;
      if (fused) win = p2p_window(c);
      const P2PWindow* llw = fused ? &win : nullptr;
      Hit* hits = (sh && !fused) ? c->comm.send.p : nullptr;
      if (sp.on) {
        if (!used_sampled) CUDA_OK(cudaMemsetAsync(d_fail, 0, sizeof(int), stm));
        used_sampled = true;
        TRY(launch_topk_sampled(c, c->d_scores.p, d_prev_ids, c->page_base, n_items, ks[s], sh ? nullptr : o_sc,
                                sh ? nullptr : o_id, sh ? nullptr : d_counts + s, d_fail, stm, sp, hits, llw));
      } else {
        TRY(launch_topk(c, c->d_scores.p, d_prev_ids, c->page_base, n_items, ks[s], sh ? nullptr : o_sc, sh ? nullptr : o_id,
                        nullptr, sh ? nullptr : d_counts + s, stm, 1, 0, hits, nullptr, llw));
      }
      if (fused) {
        comm_mark(c, stm, true);    // reported as this stage's collective: the merge kernel = wait for the peers' lines + sort
        TRY(merge_hits(c, nullptr, c->comm.nranks, 1, ks[s], ks[s], o_sc, o_id, d_fail, stm, llw, true));
        comm_mark(c, stm, false);
      } else if (sh) {
        TRY(comm_allgather_hits(c, c->comm.send.p, 1, ks[s], c->comm.recv.p, stm));
        TRY(merge_hits(c, c->comm.recv.p, c->comm.nranks, 1, ks[s], ks[s], o_sc, o_id, d_fail, stm, nullptr, true));
      }
    }
    d_prev_ids = o_id;
    // sharded: the global survivor count is not known on the host; unused slots carry id -1 and score -inf downstream
    n_prev = sh ? ks[s] : std::min<int64_t>(ks[s], n_items);
    off += ks[s];
  }
  if (timed_out) *timed_out = timed;
  if (used_sampled_out) *used_sampled_out = used_sampled;
  return 0;
}

static int search_multistage_impl(vrag_corpus_t* c, int n_stages, const char* const* names,
                                  const uint32_t* flags, const int* ks, const float* query, int n_query_rows,
                                  const int* q_offsets, const int64_t* cand_ids, int64_t n_cand,
                                  float* out_scores, int64_t* out_ids, int* out_counts, bool allow_sampled, int filter_id = 0) {
  if (!c) return fail("corpus is NULL");
  if (n_stages < 1 || n_stages > kMaxStages) return fail("n_stages %d out of range [1,%d]", n_stages, kMaxStages);
  if (!names || !flags || !ks || !out_scores || !out_ids || !out_counts) return fail("NULL argument");
  TRY(set_device(c));
  const uint32_t* d_mask = nullptr;
  if (filter_id != 0) {
    auto f = c->filters.find(filter_id);
    if (f == c->filters.end()) return fail("unknown filter %d", filter_id);
    d_mask = f->second.p;
  }
  Store* st[kMaxStages];
  size_t total_k = 0;
  for (int s = 0; s < n_stages; ++s) {
    TRY(find_store(c, names[s], &st[s]));
    if (ks[s] < 1) return fail("stage %d: k must be >= 1", s);
    if (ks[s] > kTopkHardMaxK) return fail("stage %d: k=%d exceeds the supported maximum %d", s, ks[s], kTopkHardMaxK);
    if (st[s]->n_pages != st[0]->n_pages) return fail("stage %d: store '%s' has a different page count", s, names[s]);
    total_k += ks[s];
    if (q_offsets) {
      if (q_offsets[s] < 0 || q_offsets[s + 1] <= q_offsets[s] || q_offsets[s + 1] > n_query_rows)
        return fail("stage %d: bad query row range [%d,%d)", s, q_offsets[s], q_offsets[s + 1]);
    }
  }
  if (cand_ids && n_cand < 0) return fail("n_cand < 0");
  if (d_mask && c->filter_pages[filter_id] != st[0]->n_pages)
    return fail("filter %d covers %lld pages, store '%s' has %lld", filter_id, (long long)c->filter_pages[filter_id], names[0],
                (long long)st[0]->n_pages);
  TRY(stage_query(c, query, n_query_rows));
  TRY(c->d_out_scores.ensure(total_k));
  TRY(c->d_out_ids.ensure(total_k));
  TRY(ensure_host_out(c, total_k));
  if (cand_ids) {
    TRY(c->d_cand.ensure(std::max<int64_t>(n_cand, 1)));
    if (n_cand > 0)
      CUDA_OK(cudaMemcpyAsync(c->d_cand.p, cand_ids, n_cand * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  }
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  bool timed = false, used_sampled = false;
  int* const d_fail = c->d_counts.p + kMaxStages;   // travels to the host with the per-stage counts
  const bool sh = sharded(c);
  if (sh) {
    CUDA_OK(cudaMemsetAsync(d_fail, 0, sizeof(int), c->stream));
    c->comm.timing = true;
    c->comm.n_ev = 0;
  }
  int rc = run_stages(c, n_stages, st, flags, ks, c->d_query.p, n_query_rows, q_offsets, cand_ids ? c->d_cand.p : nullptr,
                      n_cand, c->d_out_scores.p, c->d_out_ids.p, c->d_counts.p, d_fail, allow_sampled, c->stream, &timed,
                      &used_sampled, d_mask);
  c->comm.timing = false;
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(c->ev1, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_out_scores, c->d_out_scores.p, total_k * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_out_ids, c->d_out_ids.p, total_k * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_counts, c->d_counts.p, (kMaxStages + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (sh) comm_collect_timing(c);
  if (used_sampled && c->h_counts[kMaxStages]) {
    // a sampled threshold kept too few / too many keys (heavy ties, mostly -inf scores, ...) on this or — sharded: the
    // flag travels in the exchanged entries, so every rank takes this branch together — any rank: exact radix-select path
    c->sampled_fallbacks++;
    return search_multistage_impl(c, n_stages, names, flags, ks, query, n_query_rows, q_offsets, cand_ids, n_cand, out_scores,
                                  out_ids, out_counts, false, filter_id);
  }
  memcpy(out_scores, c->h_out_scores, total_k * sizeof(float));
  memcpy(out_ids, c->h_out_ids, total_k * sizeof(long long));
  if (sh) {   // valid results of a merged list = the leading entries with a real page id
    size_t off = 0;
    for (int s = 0; s < n_stages; ++s) {
      int n = 0;
      while (n < ks[s] && c->h_out_ids[off + n] >= 0) ++n;
      out_counts[s] = n;
      off += ks[s];
    }
  } else {
    memcpy(out_counts, c->h_counts, n_stages * sizeof(int));
  }
  cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1);
  if (timed) cudaEventElapsedTime(&c->last_ms[1], c->evk0, c->evk1);
  return 0;
}

extern "C" int vrag_stage_hits_dev(vrag_corpus_t* c, const char* name, const float* query_dev, int n_query_rows, uint32_t flags,
                                   const int64_t* cand_ids_dev, int64_t n_cand, int k, vrag_hit_t* out_hits_dev, void* stream) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  if (!query_dev || !out_hits_dev) return fail("NULL device pointer");
  if (k < 1 || k > kTopkHardMaxK) return fail("k=%d out of range [1,%d]", k, kTopkHardMaxK);
  cudaStream_t stm = static_cast<cudaStream_t>(stream);
  const long long* cand = reinterpret_cast<const long long*>(cand_ids_dev);
  const int64_t n_items = cand ? n_cand : s->n_pages;
  TRY(c->d_scores.ensure(std::max<int64_t>(n_items, 1)));
  if (n_items > 0) {
    if (s->total_rows == 0 && cand) TRY(fill_neg_inf(c, c->d_scores.p, n_items, stm));
    else TRY(launch_scan(c, *s, query_dev, n_query_rows, flags, cand, n_items, c->d_scores.p, stm, false));
  }
  return launch_topk(c, c->d_scores.p, cand, c->page_base, n_items, k, nullptr, nullptr, nullptr, nullptr, stm, 1, 0,
                     reinterpret_cast<Hit*>(out_hits_dev));
}

extern "C" int vrag_allgather_topk(vrag_corpus_t* c, const vrag_hit_t* local_dev, int n_lists, int k, vrag_hit_t* gathered_dev,
                                   void* stream) {
  VRAG_LOCK(c);
  TRY(set_device(c));
  if (!local_dev || !gathered_dev) return fail("NULL device pointer");
  if (n_lists < 1 || k < 1) return fail("n_lists and k must be >= 1");
  return comm_allgather_hits(c, reinterpret_cast<const Hit*>(local_dev), n_lists, k, reinterpret_cast<Hit*>(gathered_dev),
                             static_cast<cudaStream_t>(stream));
}

extern "C" int vrag_merge_hits_dev(vrag_corpus_t* c, const vrag_hit_t* gathered_dev, int n_src, int n_lists, int k_src, int k,
                                   float* out_scores_dev, int64_t* out_ids_dev, int* flag_dev, void* stream) {
  VRAG_LOCK(c);
  TRY(set_device(c));
  if (!gathered_dev || !out_scores_dev || !out_ids_dev) return fail("NULL device pointer");
  if (n_src < 1 || n_lists < 1 || k_src < 1) return fail("n_src, n_lists and k_src must be >= 1");
  return merge_hits(c, reinterpret_cast<const Hit*>(gathered_dev), n_src, n_lists, k_src, k, out_scores_dev,
                    reinterpret_cast<long long*>(out_ids_dev), flag_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int vrag_allreduce_max_dev(vrag_corpus_t* c, float* scores_dev, int64_t n, void* stream) {
  VRAG_LOCK(c);
  TRY(set_device(c));
  if (!scores_dev && n > 0) return fail("NULL device pointer");
  return comm_allreduce_max(c, scores_dev, n, static_cast<cudaStream_t>(stream));
}

extern "C" int vrag_search_multistage_dev(vrag_corpus_t* c, int n_stages, const char* const* names, const uint32_t* flags,
                                          const int* ks, const float* query_dev, int n_query_rows, const int* q_offsets,
                                          float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
  VRAG_LOCK(c);
  if (n_stages < 1 || n_stages > kMaxStages) return fail("n_stages %d out of range [1,%d]", n_stages, kMaxStages);
  if (!names || !flags || !ks || !query_dev || !out_scores_dev || !out_ids_dev) return fail("NULL argument");
  TRY(set_device(c));
  Store* st[kMaxStages];
  for (int s = 0; s < n_stages; ++s) {
    TRY(find_store(c, names[s], &st[s]));
    if (ks[s] < 1 || ks[s] > kTopkHardMaxK) return fail("stage %d: k=%d out of range [1,%d]", s, ks[s], kTopkHardMaxK);
    if (st[s]->n_pages != st[0]->n_pages) return fail("stage %d: store '%s' has a different page count", s, names[s]);
    if (q_offsets && (q_offsets[s] < 0 || q_offsets[s + 1] <= q_offsets[s] || q_offsets[s + 1] > n_query_rows))
      return fail("stage %d: bad query row range", s);
  }
  return run_stages(c, n_stages, st, flags, ks, query_dev, n_query_rows, q_offsets, nullptr, 0, out_scores_dev,
                    reinterpret_cast<long long*>(out_ids_dev), c->d_counts.p, c->d_counts.p + kMaxStages, false,
                    static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" int vrag_search_multistage(vrag_corpus_t* c, int n_stages, const char* const* names,
                                      const uint32_t* flags, const int* ks, const float* query, int n_query_rows,
                                      const int* q_offsets, const int64_t* cand_ids, int64_t n_cand,
                                      float* out_scores, int64_t* out_ids, int* out_counts) {
  VRAG_LOCK(c);
  return search_multistage_impl(c, n_stages, names, flags, ks, query, n_query_rows, q_offsets, cand_ids, n_cand, out_scores,
                                out_ids, out_counts, true);
}

// ------------------------------------------------------------------------------------------------ payload filters
extern "C" int vrag_filter_create(vrag_corpus_t* c, const uint32_t* bits, int64_t n_pages, int* out_filter) {
  VRAG_LOCK(c);
  if (!out_filter) return fail("out_filter is NULL");
  if (n_pages < 0 || (n_pages > 0 && !bits)) return fail("bad filter arguments");
  TRY(set_device(c));
  const size_t words = static_cast<size_t>((n_pages + 31) / 32);
  const int id = c->next_filter++;
  DevBuf<uint32_t>& buf = c->filters[id];
  int rc = buf.ensure(std::max<size_t>(words, 1));
  if (rc == 0 && words > 0) {
    cudaError_t e = cudaMemcpyAsync(buf.p, bits, words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail("filter upload failed: %s", cudaGetErrorString(e));
  }
  if (rc) {
    buf.release();
    c->filters.erase(id);
    return rc;
  }
  c->filter_pages[id] = n_pages;
  *out_filter = id;
  return 0;
}

extern "C" int vrag_filter_destroy(vrag_corpus_t* c, int filter) {
  VRAG_LOCK(c);
  auto f = c->filters.find(filter);
  if (f == c->filters.end()) return fail("unknown filter %d", filter);
  TRY(set_device(c));
  cudaStreamSynchronize(c->stream);
  f->second.release();
  c->filters.erase(f);
  c->filter_pages.erase(filter);
  return 0;
}

extern "C" int vrag_search_multistage_filtered(vrag_corpus_t* c, int filter, int n_stages, const char* const* names,
                                               const uint32_t* flags, const int* ks, const float* query, int n_query_rows,
                                               const int* q_offsets, float* out_scores, int64_t* out_ids, int* out_counts) {
  VRAG_LOCK(c);
  return search_multistage_impl(c, n_stages, names, flags, ks, query, n_query_rows, q_offsets, nullptr, 0, out_scores, out_ids,
                                out_counts, true, filter);
}

extern "C" int vrag_search(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, uint32_t flags,
                           const int64_t* cand_ids, int64_t n_cand, int k, float* out_scores, int64_t* out_ids,
                           int* out_count) {
  VRAG_LOCK(c);
  const char* names[1] = {name};
  int cnt = 0;
  TRY(vrag_search_multistage(c, 1, names, &flags, &k, query, n_query_rows, nullptr, cand_ids, n_cand, out_scores,
                             out_ids, &cnt));
  if (out_count) *out_count = cnt;
  return 0;
}


// Fused top-k prefilter plan for a dense batched stage (top-k of n_pages scores per query, k << n_pages):
//   1. score a strided sample of S pages, take each query's m-th best sample score as its threshold;
//   2. scan the whole store, keeping only (score, page) pairs above the threshold (expected ~m*n/S >= k of them);
//   3. sort the survivors. Exact whenever every query keeps between k and cap candidates (checked on the device;
//      otherwise the caller reruns the batch without the filter).
// The n x nq score matrix is never written. m solves m - 5 sqrt(m) >= k*S/n (5-sigma margin on the count).
struct PrefilterPlan {
  bool on = false;
  int tile_stride = 1;
  int64_t n_sample = 0;
  int m = 0;
  int cap = 8192;
  // Approximate first pass + exact re-score (dense batched token scans of LARGE stores, >= 5 queries): the scan runs with 8
  // plain-fp16 queries per document tile (twice the queries for the same tensor work), keeps kc > k candidates per query,
  // re-scores them exactly (fp32 query as hi|lo pair, the operand-switching gather) and takes the top-k of those. A
  // device-side guard proves that no page outside the candidates can reach the top-k (else the batch is redone exactly).
  bool approx = false;
  int kc = 0;         // candidates kept by the first pass
  int key_cap = 0;    // key-list capacity of the final sort (power of two >= kc)
};
static bool knob_no_approx() { return env_flag_is("VRAG_APPROX_PASS", '0'); }
static PrefilterPlan plan_prefilter(const Store& s, int k, int nq, int max_q_eff, uint32_t flags) {
  PrefilterPlan pl;
  if (knob_no_prefilter()) return pl;
  const int64_t n = s.n_pages;
  if (n < (1 << 16) || !dense_batch_covers(nq, max_q_eff, flags)) return pl;
  int64_t unit_pages = 1, n_units = n;   // sampling granularity
  if (s.packed) {
    const bool pow2 = s.fixed_rows > 0 && (s.fixed_rows & (s.fixed_rows - 1)) == 0 && s.fixed_rows <= 32;
    if (!pow2) return pl;
    unit_pages = kTileRows / s.fixed_rows;
    n_units = n / unit_pages;
  }
  // sample size: >= 32 expected hits above the k-th score, and at least 65536 pages — or an eighth of a small store (the
  // shards of a strongly scaled corpus: 1M pages over 8 GPUs are 125k per rank; without the filter their stage costs a
  // [queries][pages] score matrix and a batched radix select)
  const int64_t floor_s = std::min<int64_t>(65536, std::max<int64_t>(8192, n / 8));
  const int64_t want = std::max<int64_t>(floor_s, (32 * n + k - 1) / k);
  if (want * 4 > n) return pl;
  if (!s.packed && want * 20 > n) return pl;   // full-token stores: the sample pass re-reads corpus bytes, the score matrix it
                                               // saves is tiny next to them — only worth it when the sample is < 5 % of the scan
  const int64_t stride = n_units / ((want + unit_pages - 1) / unit_pages);
  if (stride < 2) return pl;
  const int64_t sampled_units = (n_units + stride - 1) / stride;
  pl.n_sample = sampled_units * unit_pages;
  const double ratio = static_cast<double>(n) / pl.n_sample;
  const double target = k / ratio;
  int m = static_cast<int>(target) + 1;
  while (m - 5.0 * sqrt(static_cast<double>(m)) < target) ++m;
  if (m > kTopkMaxK || m * ratio * (1.0 + 5.0 / sqrt(static_cast<double>(m))) > pl.cap) return pl;
  pl.m = m;
  pl.tile_stride = static_cast<int>(stride);
  pl.on = true;
  return pl;
}

// ------------------------------------------------------------------------------------------------ batched queries
static int ensure_host_query(vrag_corpus* c, size_t rows) {
  if (rows <= c->h_query_cap) return 0;
  if (c->h_query) cudaFreeHost(c->h_query);
  c->h_query = nullptr;
  c->h_query_cap = 0;
  CUDA_OK(cudaMallocHost(&c->h_query, rows * 128 * sizeof(float)));
  c->h_query_cap = rows;
  return 0;
}

// ---- an uploaded batch of queries (vrag_search_multistage_batch and the device-level stage API share it)
static int batch_upload(vrag_corpus* c, int n_stages, const uint32_t* flags, int n_queries, const float* query_rows,
                        const int* q_offsets, int per_stage_queries) {
  BatchCtx& bc = c->batch;
  bc.valid = false;
  const int nq = n_queries;
  const int qs = per_stage_queries ? n_stages : 1;
  const int total_rows = q_offsets[static_cast<size_t>(nq) * qs];
  // per (stage, query) row range + the largest effective row count per stage
  bc.max_rows.assign(n_stages, 0);
  for (int b = 0; b < nq; ++b)
    for (int s = 0; s < n_stages; ++s) {
      const int i = b * qs + (per_stage_queries ? s : 0);
      const int r0 = q_offsets[i], r1 = q_offsets[i + 1];
      if (r0 < 0 || r1 <= r0 || r1 > total_rows) return fail("query %d stage %d: bad row range [%d,%d)", b, s, r0, r1);
      if (r1 - r0 > kMaxQueryRows) return fail("query %d has %d rows; at most %d supported", b, r1 - r0, kMaxQueryRows);
      bc.max_rows[s] = std::max(bc.max_rows[s], r1 - r0);
    }
  // ---- stage queries and metadata to the device
  TRY(ensure_host_query(c, total_rows));
  TRY(c->d_query.ensure(static_cast<size_t>(total_rows) * 128));
  memcpy(c->h_query, query_rows, static_cast<size_t>(total_rows) * 128 * sizeof(float));
  CUDA_OK(cudaMemcpyAsync(c->d_query.p, c->h_query, static_cast<size_t>(total_rows) * 128 * sizeof(float),
                          cudaMemcpyHostToDevice, c->stream));
  const size_t meta_n = static_cast<size_t>(n_stages) * 2 * nq + nq;   // [s][begin|end][b], then q_valid scratch [b]
  if (meta_n > c->h_qmeta_cap) {
    if (c->h_qmeta) cudaFreeHost(c->h_qmeta);
    c->h_qmeta = nullptr;
    c->h_qmeta_cap = 0;
    CUDA_OK(cudaMallocHost(&c->h_qmeta, meta_n * sizeof(int)));
    c->h_qmeta_cap = meta_n;
  }
  TRY(c->d_qmeta.ensure(meta_n));
  for (int s = 0; s < n_stages; ++s)
    for (int b = 0; b < nq; ++b) {
      const int i = b * qs + (per_stage_queries ? s : 0);
      c->h_qmeta[(static_cast<size_t>(s) * 2 + 0) * nq + b] = q_offsets[i];
      c->h_qmeta[(static_cast<size_t>(s) * 2 + 1) * nq + b] = q_offsets[i + 1];
    }
  CUDA_OK(cudaMemcpyAsync(c->d_qmeta.p, c->h_qmeta, (meta_n - nq) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  TRY(c->d_fcnt.ensure(nq + 1));
  CUDA_OK(cudaMemsetAsync(c->d_fcnt.p + nq, 0, sizeof(int), c->stream));
  bc.nq = nq;
  bc.n_stages = n_stages;
  bc.qs = qs;
  bc.per_stage = per_stage_queries != 0;
  bc.valid = true;
  return 0;
}

// One stage of the uploaded batch for queries [b0, b0+qc): scan (dense / candidate lists) + local top-k into
// o_sc / o_id ([qc][k]). d_prev_ids: nullptr (every page) or [qc][n_items] global candidate ids.
static int batch_stage_chunk(vrag_corpus* c, int s, Store& store, uint32_t flags, int k, const long long* d_prev_ids,
                             int64_t n_items, bool have_items, int b0, int qc, float* o_sc, long long* o_id, cudaStream_t stm,
                             const PrefilterPlan& plan, bool* timed, float* raw_out = nullptr, Hit* o_hits = nullptr) {
  // raw_out != nullptr (candidate stages only): write the [qc][n_items] score matrix there and skip the top-k
  // o_hits != nullptr (sharded scanning stage): the local lists go out as packed entries [qc][k] (o_sc / o_id may be null)
  BatchCtx& bc = c->batch;
  float* const d_sc = raw_out ? raw_out : c->d_scores.p;
  const int nq = bc.nq;
  const int* d_qb = c->d_qmeta.p + (static_cast<size_t>(s) * 2 + 0) * nq + b0;
  const int* d_qe = c->d_qmeta.p + (static_cast<size_t>(s) * 2 + 1) * nq + b0;
  int* d_qvalid = c->d_qmeta.p + static_cast<size_t>(bc.n_stages) * 2 * nq;
  const int max_rows = bc.max_rows[s];
  if (have_items) {
    int r = 2;
    // first-pass outputs: the stage's own lists, or — approximate first pass — the candidate lists that get re-scored
    const bool approx = plan.approx && !d_prev_ids;
    const int k1 = approx ? plan.kc : k;
    float* const l_sc = approx ? c->d_ap_sc.p : o_sc;
    long long* const l_id = approx ? c->d_ap_id.p : o_id;
    Hit* const l_hits = approx ? nullptr : o_hits;
    auto finish_approx = [&]() -> int {
      // exact scores of every query's candidates (operand-switching gather, fp32-exact query), then the final order
      int rr = launch_scan_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, c->d_ap_id.p, plan.kc,
                                 c->d_ap_exact.p, stm);
      if (rr) return rr == 2 ? fail("internal: candidate re-score not covered") : rr;
      approx_finalize_kernel<<<qc, 256, 0, stm>>>(c->d_ap_exact.p, c->d_ap_sc.p, c->d_ap_id.p, plan.kc, plan.key_cap, c->page_base,
                                                  c->d_ap_eps.p, k, c->d_ap_keys.p, c->d_ap_cnt.p, c->d_fcnt.p + nq);
      c->launches++;
      return launch_topk_keys(c, c->d_ap_keys.p, c->d_ap_cnt.p, plan.key_cap, k, c->page_base, o_sc, o_id, stm, qc, nullptr, nullptr,
                              nullptr, 0, o_hits, o_hits ? c->d_fcnt.p + nq : nullptr);
    };
    if (!d_prev_ids && plan.on) {
      // fused top-k prefilter: sample -> thresholds -> filtered scan -> sort the survivors
      DenseOpts o;
      o.tile_stride = plan.tile_stride;
      o.n_sample = plan.n_sample;
      o.two_block = approx;
      o.eps = c->d_ap_eps.p;
      TRY(launch_scan_dense_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, c->d_scores.p, stm,
                                  false, o));
      prefilter_sample_thr_kernel<false><<<qc, 1024, 0, stm>>>(c->d_scores.p, plan.n_sample, plan.m, c->d_fthr.p, c->d_fcnt.p);
      DenseOpts f;
      f.thr = c->d_fthr.p;
      f.cnt = c->d_fcnt.p;
      f.keys = c->d_fkeys.p;
      f.cap = plan.cap;
      f.skip_prep = true;
      f.two_block = approx;
      TRY(launch_scan_dense_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, nullptr, stm,
                                  timed && !*timed, f));
      if (timed) *timed = true;
      prefilter_check_kernel<<<(qc + 127) / 128, 128, 0, stm>>>(c->d_fcnt.p, qc, static_cast<int>(std::min<int64_t>(k1, store.n_pages)),
                                                                plan.cap, c->d_fcnt.p + nq);
      c->launches += 2;
      TRY(launch_topk_keys(c, c->d_fkeys.p, c->d_fcnt.p, plan.cap, k1, c->page_base, l_sc, l_id, stm, qc, nullptr, nullptr, nullptr, 0,
                           l_hits, l_hits ? c->d_fcnt.p + nq : nullptr));
      return approx ? finish_approx() : 0;
    }
    if (approx) {
      DenseOpts o;
      o.two_block = true;
      o.eps = c->d_ap_eps.p;
      TRY(launch_scan_dense_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, c->d_scores.p, stm,
                                  timed && !*timed, o));
      if (timed) *timed = true;
      TRY(launch_topk(c, c->d_scores.p, nullptr, c->page_base, n_items, k1, l_sc, l_id, nullptr, nullptr, stm, qc, 0));
      return finish_approx();
    }
    if (!d_prev_ids && dense_batch_covers(qc, max_rows, flags)) {
      r = launch_scan_dense_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, c->d_scores.p, stm,
                                  timed && !*timed);
      if (r == 1) return r;
      if (r == 0 && timed) *timed = true;
    } else if (d_prev_ids && store.total_rows == 0) {
      TRY(fill_neg_inf(c, d_sc, static_cast<int64_t>(qc) * n_items, stm));   // an empty shard owns no candidate
      r = 0;
    } else if (d_prev_ids) {
      const bool t = timed && !*timed;
      if (t) CUDA_OK(cudaEventRecord(c->evk0, stm));
      r = launch_scan_batch(c, store, c->d_query.p, d_qb, d_qe, d_qvalid + b0, qc, max_rows, flags, d_prev_ids, n_items,
                            d_sc, stm);
      if (r == 1) return r;
      if (t && r == 0) { CUDA_OK(cudaEventRecord(c->evk1, stm)); *timed = true; }
    }
    if (r == 2) {   // one launch per query (shapes the batched kernels do not cover)
      for (int b = 0; b < qc; ++b) {
        const int q0 = c->h_qmeta[(static_cast<size_t>(s) * 2 + 0) * nq + b0 + b];
        const int q1 = c->h_qmeta[(static_cast<size_t>(s) * 2 + 1) * nq + b0 + b];
        TRY(launch_scan(c, store, c->d_query.p + static_cast<size_t>(q0) * 128, q1 - q0, flags,
                        d_prev_ids ? d_prev_ids + static_cast<size_t>(b) * n_items : nullptr, n_items,
                        d_sc + static_cast<size_t>(b) * n_items, stm, false));
      }
    }
  }
  if (raw_out) return 0;
  return launch_topk(c, c->d_scores.p, d_prev_ids, c->page_base, have_items ? n_items : 0, k, o_sc, o_id, nullptr, nullptr, stm,
                     qc, d_prev_ids ? n_items : 0, o_hits);
}

// chunk size (queries) that bounds the stage-0 score matrix [chunk][n_pages] to ~1 GiB, and the scratch it needs
static int batch_prepare_stage(vrag_corpus* c, Store& store, int k, int64_t n_items, bool dense, bool allow_prefilter,
                               uint32_t flags, int s, PrefilterPlan* plan, int* qchunk_out) {
  BatchCtx& bc = c->batch;
  const int64_t max_scores = int64_t(1) << 28;
  const int64_t per_query = std::max<int64_t>(dense ? store.n_pages : n_items, 1);
  const int qchunk = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(bc.nq, max_scores / per_query)));
  TRY(c->d_scores.ensure(static_cast<int64_t>(qchunk) * per_query));
  *plan = PrefilterPlan();
  // approximate first pass + exact re-score: token queries (<= 32 rows), cosine scores, a full-token store, enough queries
  // to fill more than one 4-query image, and a candidate list that stays a small part of the store
  bool approx = false;
  int kc = 0;
  if (dense && allow_prefilter && !knob_no_approx() && !store.packed && (flags & VRAG_Q_NORMALIZE) &&
      !(flags & (VRAG_Q_POOL | VRAG_Q_FP16)) && bc.max_rows[s] > 1 && bc.max_rows[s] <= 32 && qchunk >= 5 && k >= 1) {
    kc = std::max(k + 118, 2 * k);
    approx = kc <= 2048 && static_cast<int64_t>(kc) * 8 <= store.n_pages;
  }
  if (dense && allow_prefilter && !knob_no_prefilter()) *plan = plan_prefilter(store, approx ? kc : k, qchunk, bc.max_rows[s], flags);
  if (plan->on) {
    TRY(c->d_fthr.ensure(qchunk));
    TRY(c->d_fkeys.ensure(static_cast<size_t>(qchunk) * plan->cap));
    c->prefilter_runs++;
  }
  if (approx) {
    plan->approx = true;
    plan->kc = kc;
    plan->key_cap = kc <= 1024 ? 1024 : 2048;
    TRY(c->d_ap_sc.ensure(static_cast<size_t>(qchunk) * kc));
    TRY(c->d_ap_exact.ensure(static_cast<size_t>(qchunk) * kc));
    TRY(c->d_ap_id.ensure(static_cast<size_t>(qchunk) * kc));
    TRY(c->d_ap_eps.ensure(qchunk));
    TRY(c->d_ap_keys.ensure(static_cast<size_t>(qchunk) * plan->key_cap));
    TRY(c->d_ap_cnt.ensure(qchunk));
    c->approx_runs++;
  }
  *qchunk_out = qchunk;
  return 0;
}

// The uploaded batch over a sharded corpus, stage-major: every stage's local work for ALL queries first (in query
// chunks that bound the dense score matrix — chunking is a local matter and never changes the number of collectives),
// then ONE collective for the whole batch — the all-gather of [n_queries][k] packed local lists for a scanning stage,
// the max-all-reduce of the [n_queries][n_cand] candidate scores for a restricted stage — and the batched merge.
// Results are stage-major in d_out_scores / d_out_ids like the single-shard path. The prefilter's "estimate missed"
// flag of any rank reaches every rank inside the exchanged entries (d_fcnt[nq] after the merge).
static int batch_run_sharded(vrag_corpus* c, int n_stages, Store* const* st, const uint32_t* flags, const int* ks,
                             bool allow_prefilter, cudaStream_t stm, bool* timed) {
  BatchCtx& bc = c->batch;
  const int nq = bc.nq, R = c->comm.nranks;
  size_t off = 0;
  const long long* d_prev_ids = nullptr;
  for (int s = 0; s < n_stages; ++s) {
    const int k = ks[s];
    float* o_sc = c->d_out_scores.p + off * nq;
    long long* o_id = c->d_out_ids.p + off * nq;
    PrefilterPlan plan;
    int qchunk = 1;
    if (s == 0) {
      const int64_t n_pages = st[0]->n_pages;
      TRY(batch_prepare_stage(c, *st[0], k, 0, true, allow_prefilter, flags[0], 0, &plan, &qchunk));
      TRY(c->comm.send.ensure(static_cast<size_t>(nq) * k));
      TRY(c->comm.recv.ensure(static_cast<size_t>(nq) * k * R));
      for (int b0 = 0; b0 < nq; b0 += qchunk) {
        const int qc = std::min(qchunk, nq - b0);
        TRY(batch_stage_chunk(c, 0, *st[0], flags[0], k, nullptr, n_pages, n_pages > 0, b0, qc, nullptr, nullptr, stm, plan, timed,
                              nullptr, c->comm.send.p + static_cast<size_t>(b0) * k));
      }
      TRY(comm_allgather_hits(c, c->comm.send.p, nq, k, c->comm.recv.p, stm));
      TRY(merge_hits(c, c->comm.recv.p, R, nq, k, k, o_sc, o_id, c->d_fcnt.p + nq, stm, nullptr, true));
    } else {
      const int64_t n_cand = ks[s - 1];
      TRY(c->comm.raw.ensure(static_cast<size_t>(nq) * n_cand));
      // the rank scores only the candidates it OWNS (~1/R of the replicated lists): count them per query, read the largest
      // count back (one small synchronisation per candidate stage), compact to [nq][m], scan, scatter into the -inf matrix
      bool compacted = false;
      if (!env_flag_is("VRAG_OWN_COMPACT", '0') && n_cand >= 64 && n_cand <= (1 << 24)) {
        TRY(c->comm.own_cnt.ensure(static_cast<size_t>(nq) + 1));
        CUDA_OK(cudaMemsetAsync(c->comm.own_cnt.p + nq, 0, sizeof(int), stm));
        own_count_kernel<<<nq, 256, 0, stm>>>(d_prev_ids, static_cast<int>(n_cand), c->page_base, st[s]->n_pages, c->comm.own_cnt.p,
                                             c->comm.own_cnt.p + nq);
        c->launches++;
        CUDA_OK(cudaMemcpyAsync(c->h_flag, c->comm.own_cnt.p + nq, sizeof(int), cudaMemcpyDeviceToHost, stm));
        CUDA_OK(cudaStreamSynchronize(stm));
        const int m_max = *c->h_flag;
        *c->h_flag = 0;
        const int64_t m = std::min<int64_t>(n_cand, (static_cast<int64_t>(std::max(m_max, 1)) + 31) & ~int64_t(31));
        if (m * 4 <= n_cand * 3) {   // worth it: at most 3/4 of the list left
          compacted = true;
          const long long tot = static_cast<long long>(nq) * n_cand;
          fill_f32_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stm>>>(c->comm.raw.p, tot, -INFINITY);
          c->launches++;
          if (m_max > 0 && st[s]->total_rows > 0) {
            TRY(c->comm.own_ids.ensure(static_cast<size_t>(nq) * m));
            TRY(c->comm.own_pos.ensure(static_cast<size_t>(nq) * m));
            TRY(c->comm.raw_c.ensure(static_cast<size_t>(nq) * m));
            own_compact_kernel<<<nq, 256, 0, stm>>>(d_prev_ids, static_cast<int>(n_cand), c->page_base, st[s]->n_pages,
                                                   static_cast<int>(m), c->comm.own_ids.p, c->comm.own_pos.p);
            c->launches++;
            TRY(batch_prepare_stage(c, *st[s], k, m, false, false, flags[s], s, &plan, &qchunk));
            for (int b0 = 0; b0 < nq; b0 += qchunk) {
              const int qc = std::min(qchunk, nq - b0);
              TRY(batch_stage_chunk(c, s, *st[s], flags[s], k, c->comm.own_ids.p + static_cast<size_t>(b0) * m, m, true, b0, qc,
                                    nullptr, nullptr, stm, PrefilterPlan(), timed, c->comm.raw_c.p + static_cast<size_t>(b0) * m));
            }
            const long long tc = static_cast<long long>(nq) * m;
            own_scatter_kernel<<<static_cast<unsigned>((tc + 255) / 256), 256, 0, stm>>>(c->comm.raw_c.p, c->comm.own_pos.p,
                                                                                        static_cast<int>(m), static_cast<int>(n_cand), nq,
                                                                                        c->comm.raw.p);
            c->launches++;
          }
          CUDA_OK(cudaGetLastError());
        }
      }
      if (!compacted) {
        TRY(batch_prepare_stage(c, *st[s], k, n_cand, false, false, flags[s], s, &plan, &qchunk));
        for (int b0 = 0; b0 < nq; b0 += qchunk) {
          const int qc = std::min(qchunk, nq - b0);
          TRY(batch_stage_chunk(c, s, *st[s], flags[s], k, d_prev_ids + static_cast<size_t>(b0) * n_cand, n_cand, true, b0, qc,
                                nullptr, nullptr, stm, PrefilterPlan(), timed, c->comm.raw.p + static_cast<size_t>(b0) * n_cand));
        }
      }
      TRY(comm_allreduce_max(c, c->comm.raw.p, static_cast<int64_t>(nq) * n_cand, stm));
      TRY(launch_topk(c, c->comm.raw.p, d_prev_ids, 0, n_cand, k, o_sc, o_id, nullptr, nullptr, stm, nq, n_cand));
    }
    d_prev_ids = o_id;
    off += k;
  }
  return 0;
}

static int search_multistage_batch_impl(vrag_corpus_t* c, int n_stages, const char* const* names,
                                        const uint32_t* flags, const int* ks, int n_queries, const float* query_rows,
                                        const int* q_offsets, int per_stage_queries, float* out_scores,
                                        int64_t* out_ids, int* out_counts, bool no_prefilter,
                                        float* out_stage_scores = nullptr, bool final_only = false) {
  // final_only: only the last stage's lists ([nq][k_last]) travel to the host, plus for every final result its score in
  // each earlier stage (out_stage_scores [nq][k_last][n_stages-1], NaN if absent); out_counts is [nq].
  if (!c) return fail("corpus is NULL");
  if (n_stages < 1 || n_stages > kMaxStages) return fail("n_stages %d out of range [1,%d]", n_stages, kMaxStages);
  if (!names || !flags || !ks || !out_scores || !out_ids || !out_counts || !query_rows || !q_offsets) return fail("NULL argument");
  if (final_only && n_stages > 1 && !out_stage_scores) return fail("out_stage_scores is NULL");
  if (n_queries < 0) return fail("n_queries < 0");
  if (n_queries == 0) return 0;
  TRY(set_device(c));
  Store* st[kMaxStages];
  size_t total_k = 0;
  for (int s = 0; s < n_stages; ++s) {
    TRY(find_store(c, names[s], &st[s]));
    if (ks[s] < 1) return fail("stage %d: k must be >= 1", s);
    if (ks[s] > kTopkMaxK) return fail("stage %d: k=%d exceeds the maximum of a BATCHED stage (%d); use the single-query call", s, ks[s], kTopkMaxK);
    if (st[s]->n_pages != st[0]->n_pages) return fail("stage %d: store '%s' has a different page count", s, names[s]);
    total_k += ks[s];
  }
  const int nq = n_queries;
  TRY(batch_upload(c, n_stages, flags, nq, query_rows, q_offsets, per_stage_queries));
  // ---- outputs, stage-major: stage s occupies [nq*sum(ks[:s]), +nq*ks[s]) as [nq][ks[s]]
  const size_t out_n = total_k * nq;
  TRY(c->d_out_scores.ensure(out_n));
  TRY(c->d_out_ids.ensure(out_n));
  const int k_last = ks[n_stages - 1];
  const size_t stage_sc_n = static_cast<size_t>(nq) * k_last * std::max(n_stages - 1, 1);
  TRY(ensure_host_out(c, std::max(out_n, stage_sc_n)));
  if (final_only) TRY(c->d_stage_sc.ensure(stage_sc_n));
  const int64_t n_pages = st[0]->n_pages;
  const bool sh = sharded(c);
  PrefilterPlan plan;
  int qchunk = 1;
  bool timed = false;
  if (sh) {
    CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    c->comm.timing = true;
    c->comm.n_ev = 0;
    const int rc = batch_run_sharded(c, n_stages, st, flags, ks, !no_prefilter, c->stream, &timed);
    c->comm.timing = false;
    if (rc) return rc;
  } else {
  TRY(batch_prepare_stage(c, *st[0], ks[0], 0, true, !no_prefilter, flags[0], 0, &plan, &qchunk));
  {
    int64_t need = static_cast<int64_t>(qchunk) * std::max<int64_t>(n_pages, 1);
    for (int s = 0; s + 1 < n_stages; ++s) need = std::max<int64_t>(need, static_cast<int64_t>(qchunk) * ks[s]);
    TRY(c->d_scores.ensure(need));
  }
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  for (int b0 = 0; b0 < nq; b0 += qchunk) {
    const int qc = std::min(qchunk, nq - b0);
    size_t off = 0;
    int64_t n_prev = n_pages;
    const long long* d_prev_ids = nullptr;
    for (int s = 0; s < n_stages; ++s) {
      const int64_t n_items = (s == 0) ? n_pages : ks[s - 1];   // candidate lists keep the full stride; missing ids are -1
      float* o_sc = c->d_out_scores.p + off * nq + static_cast<size_t>(b0) * ks[s];
      long long* o_id = c->d_out_ids.p + off * nq + static_cast<size_t>(b0) * ks[s];
      TRY(batch_stage_chunk(c, s, *st[s], flags[s], ks[s], d_prev_ids, n_items, n_prev > 0, b0, qc, o_sc, o_id, c->stream,
                            s == 0 ? plan : PrefilterPlan(), &timed));
      d_prev_ids = o_id;
      n_prev = std::min<int64_t>(ks[s], n_prev);
      if (!final_only) for (int b = 0; b < qc; ++b) out_counts[static_cast<size_t>(s) * nq + b0 + b] = static_cast<int>(n_prev);
      else if (s == n_stages - 1) for (int b = 0; b < qc; ++b) out_counts[b0 + b] = static_cast<int>(n_prev);
      off += ks[s];
    }
  }
  }
  const size_t off_last = (total_k - k_last) * nq;   // the last stage's [nq][k_last] block
  if (final_only) {
    size_t off = 0;
    for (int s = 0; s + 1 < n_stages; ++s) {
      const long long warps = static_cast<long long>(nq) * k_last;
      gather_stage_scores_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, c->stream>>>(
          c->d_out_ids.p + off_last, k_last, nq, c->d_out_ids.p + off * nq, c->d_out_scores.p + off * nq, ks[s], c->d_stage_sc.p,
          n_stages - 1, s);
      c->launches++;
      off += ks[s];
    }
  }
  CUDA_OK(cudaEventRecord(c->ev1, c->stream));
  if (final_only) {
    const size_t fn = static_cast<size_t>(nq) * k_last;
    CUDA_OK(cudaMemcpyAsync(c->h_out_ids, c->d_out_ids.p + off_last, fn * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(out_scores, c->d_out_scores.p + off_last, fn * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (n_stages > 1)
      CUDA_OK(cudaMemcpyAsync(c->h_out_scores, c->d_stage_sc.p, stage_sc_n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  } else {
    CUDA_OK(cudaMemcpyAsync(c->h_out_scores, c->d_out_scores.p, out_n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->h_out_ids, c->d_out_ids.p, out_n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  }
  const bool check_flag = plan.on || plan.approx || (sh && !no_prefilter);
  if (check_flag) CUDA_OK(cudaMemcpyAsync(c->h_flag, c->d_fcnt.p + nq, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (sh) comm_collect_timing(c);
  if (check_flag && *c->h_flag) {
    // a threshold estimate kept too few / too many candidates for some query: redo the batch with the exact path
    *c->h_flag = 0;
    c->prefilter_fallbacks++;
    return search_multistage_batch_impl(c, n_stages, names, flags, ks, n_queries, query_rows, q_offsets, per_stage_queries,
                                        out_scores, out_ids, out_counts, true, out_stage_scores, final_only);
  }
  if (final_only) {
    memcpy(out_ids, c->h_out_ids, static_cast<size_t>(nq) * k_last * sizeof(long long));
    if (n_stages > 1) memcpy(out_stage_scores, c->h_out_scores, stage_sc_n * sizeof(float));
  } else {
    memcpy(out_scores, c->h_out_scores, out_n * sizeof(float));
    memcpy(out_ids, c->h_out_ids, out_n * sizeof(long long));
  }
  if (sh) {   // merged lists: the valid results are the leading entries with a real page id
    auto count_valid = [](const long long* ids, int k) {
      int n = 0;
      while (n < k && ids[n] >= 0) ++n;
      return n;
    };
    if (final_only) {
      for (int b = 0; b < nq; ++b) out_counts[b] = count_valid(c->h_out_ids + static_cast<size_t>(b) * k_last, k_last);
    } else {
      size_t off = 0;
      for (int s = 0; s < n_stages; ++s) {
        for (int b = 0; b < nq; ++b)
          out_counts[static_cast<size_t>(s) * nq + b] = count_valid(c->h_out_ids + off * nq + static_cast<size_t>(b) * ks[s], ks[s]);
        off += ks[s];
      }
    }
  }
  cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1);
  if (timed) cudaEventElapsedTime(&c->last_ms[1], c->evk0, c->evk1);
  return 0;
}

extern "C" int vrag_search_multistage_batch_final(vrag_corpus_t* c, int n_stages, const char* const* names,
                                                  const uint32_t* flags, const int* ks, int n_queries,
                                                  const float* query_rows, const int* q_offsets, int per_stage_queries,
                                                  float* out_scores, int64_t* out_ids, float* out_stage_scores,
                                                  int* out_counts) {
  VRAG_LOCK(c);
  return search_multistage_batch_impl(c, n_stages, names, flags, ks, n_queries, query_rows, q_offsets, per_stage_queries,
                                      out_scores, out_ids, out_counts, false, out_stage_scores, true);
}

// ---- device-level batched stage API (sharded multi-GPU search: NCCL all-gathers run between the stages)
extern "C" int vrag_batch_upload(vrag_corpus_t* c, int n_stages, int n_queries, const float* query_rows, const int* q_offsets,
                                 int per_stage_queries) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  if (n_stages < 1 || n_stages > kMaxStages) return fail("n_stages %d out of range [1,%d]", n_stages, kMaxStages);
  if (n_queries < 1 || !query_rows || !q_offsets) return fail("bad batch arguments");
  TRY(set_device(c));
  TRY(batch_upload(c, n_stages, nullptr, n_queries, query_rows, q_offsets, per_stage_queries));
  CUDA_OK(cudaStreamSynchronize(c->stream));   // later stages may run on the caller's stream
  return 0;
}

extern "C" int vrag_batch_stage_dev(vrag_corpus_t* c, int stage, const char* name, uint32_t flags, int k,
                                    const int64_t* cand_ids_dev, int64_t n_cand, int allow_prefilter, float* out_scores_dev,
                                    int64_t* out_ids_dev, void* stream) {
  VRAG_LOCK(c);
  Store* st;
  TRY(find_store(c, name, &st));
  TRY(set_device(c));
  BatchCtx& bc = c->batch;
  if (!bc.valid) return fail("no uploaded batch: call vrag_batch_upload first");
  if (stage < 0 || stage >= bc.n_stages) return fail("stage %d out of range", stage);
  const bool raw = k == 0;   // candidate stage, scores only: [n_queries][n_cand] into out_scores_dev
  if (raw && !cand_ids_dev) return fail("k == 0 (raw scores) needs candidate lists");
  if (!raw && (k < 1 || k > kTopkMaxK)) return fail("k=%d out of range [1,%d]", k, kTopkMaxK);
  if (!out_scores_dev || (!raw && !out_ids_dev)) return fail("NULL device pointer");
  cudaStream_t stm = static_cast<cudaStream_t>(stream);
  const bool dense = cand_ids_dev == nullptr;
  PrefilterPlan plan;
  int qchunk = 1;
  TRY(batch_prepare_stage(c, *st, k, n_cand, dense, allow_prefilter != 0, flags, stage, &plan, &qchunk));
  const int64_t n_items = dense ? st->n_pages : n_cand;
  for (int b0 = 0; b0 < bc.nq; b0 += qchunk) {
    const int qc = std::min(qchunk, bc.nq - b0);
    TRY(batch_stage_chunk(c, stage, *st, flags, k, dense ? nullptr : reinterpret_cast<const long long*>(cand_ids_dev) + static_cast<size_t>(b0) * n_cand,
                          n_items, n_items > 0, b0, qc, out_scores_dev + static_cast<size_t>(b0) * k,
                          reinterpret_cast<long long*>(out_ids_dev) + static_cast<size_t>(b0) * k, stm, plan, nullptr,
                          raw ? out_scores_dev + static_cast<size_t>(b0) * n_cand : nullptr));
  }
  return 0;
}

// 1 if a prefiltered stage since the last upload kept too few / too many candidates for some query (the caller then
// repeats the stages with allow_prefilter = 0). Synchronises `stream`.
extern "C" int vrag_batch_prefilter_failed(vrag_corpus_t* c, void* stream, int* failed) {
  VRAG_LOCK(c);
  if (!c || !failed) return fail("NULL argument");
  TRY(set_device(c));
  if (!c->batch.valid) return fail("no uploaded batch");
  cudaStream_t stm = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemcpyAsync(c->h_flag, c->d_fcnt.p + c->batch.nq, sizeof(int), cudaMemcpyDeviceToHost, stm));
  CUDA_OK(cudaStreamSynchronize(stm));
  *failed = *c->h_flag ? 1 : 0;
  if (*c->h_flag) {
    *c->h_flag = 0;
    c->prefilter_fallbacks++;
    CUDA_OK(cudaMemsetAsync(c->d_fcnt.p + c->batch.nq, 0, sizeof(int), stm));
  }
  return 0;
}

// Batched exact top-k merge: scores/ids [nq][n] (device) -> [nq][k], ties -> lower position.
extern "C" int vrag_topk_batch_dev(vrag_corpus_t* c, const float* scores_dev, const int64_t* ids_dev, int64_t n, int k, int nq,
                                   float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  TRY(set_device(c));
  if (!scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev) return fail("NULL device pointer");
  return launch_topk(c, scores_dev, reinterpret_cast<const long long*>(ids_dev), 0, n, k, out_scores_dev,
                     reinterpret_cast<long long*>(out_ids_dev), nullptr, nullptr, static_cast<cudaStream_t>(stream), nq, n);
}

extern "C" int vrag_search_multistage_batch(vrag_corpus_t* c, int n_stages, const char* const* names,
                                            const uint32_t* flags, const int* ks, int n_queries, const float* query_rows,
                                            const int* q_offsets, int per_stage_queries, float* out_scores,
                                            int64_t* out_ids, int* out_counts) {
  VRAG_LOCK(c);
  return search_multistage_batch_impl(c, n_stages, names, flags, ks, n_queries, query_rows, q_offsets, per_stage_queries,
                                      out_scores, out_ids, out_counts, false);
}

// ------------------------------------------------------------------------------------------------ saliency
extern "C" int vrag_saliency(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, int64_t page_id,
                             float* out_scores, int64_t capacity, int64_t* out_rows) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  if (!out_rows) return fail("out_rows is NULL");
  const int64_t local = page_id - c->page_base;
  if (local < 0 || local >= s->n_pages) return fail("page id %lld is not in this shard", (long long)page_id);
  if (n_query_rows < 1 || n_query_rows > kMaxQueryRows) return fail("query rows %d out of range [1,%d]", n_query_rows, kMaxQueryRows);
  int64_t r0, n;
  page_rows_h(*s, local, &r0, &n);
  *out_rows = n;
  if (n > capacity) return fail("output buffer too small: need %lld scores", (long long)n);
  if (n == 0) return 0;
  if (!out_scores) return fail("out_scores is NULL");
  TRY(stage_query(c, query, n_query_rows));
  TRY(c->d_scores.ensure(n));
  const int kSalRows = 96;   // query rows held in shared memory per launch; longer queries max-combine over chunks
  static PerDeviceOnce once;
  if (once.first()) CUDA_OK(cudaFuncSetAttribute(saliency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSalRows * 512));
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n + 7) / 8, c->num_sms * 4));
  for (int q0 = 0; q0 < n_query_rows; q0 += kSalRows) {
    const int rows = std::min(kSalRows, n_query_rows - q0);
    saliency_kernel<<<grid, 256, static_cast<size_t>(rows) * 512, c->stream>>>(
        s->rows, s->inv, r0, static_cast<int>(n), c->d_query.p + static_cast<size_t>(q0) * 128, rows, c->d_scores.p, q0 > 0 ? 1 : 0);
    c->launches++;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(out_scores, c->d_scores.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ device-pointer variants
extern "C" int vrag_score_dev(vrag_corpus_t* c, const char* name, const float* query_dev, int n_query_rows,
                              uint32_t flags, const int64_t* cand_ids_dev, int64_t n_cand, float* out_scores_dev,
                              void* stream) {
  VRAG_LOCK(c);
  Store* s;
  TRY(find_store(c, name, &s));
  TRY(set_device(c));
  if (!query_dev || !out_scores_dev) return fail("NULL device pointer");
  return launch_scan(c, *s, query_dev, n_query_rows, flags, reinterpret_cast<const long long*>(cand_ids_dev), n_cand,
                     out_scores_dev, static_cast<cudaStream_t>(stream), false);
}

extern "C" int vrag_topk_dev(vrag_corpus_t* c, const float* scores_dev, const int64_t* ids_dev, int64_t id_base,
                             int64_t n, int k, float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
  VRAG_LOCK(c);
  if (!c) return fail("corpus is NULL");
  TRY(set_device(c));
  if (!out_scores_dev || !out_ids_dev) return fail("NULL device pointer");
  return launch_topk(c, scores_dev, reinterpret_cast<const long long*>(ids_dev), id_base, n, k, out_scores_dev,
                     reinterpret_cast<long long*>(out_ids_dev), nullptr, nullptr, static_cast<cudaStream_t>(stream));
}


// ------------------------------------------------------------------------------------------------ pooling
static int ceil_div_i64(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// Output rows + argument validation for one spec on a page of t rows (grid overrides for per-page grids).
static int pool_out_rows(const vrag_pool_spec_t& s, int64_t t, int gh, int gw, int64_t* out) {
  switch (s.kind) {
    case VRAG_POOL_TILE_MEAN:
      if (s.patches_per_tile <= 0) return fail("patches_per_tile must be > 0");
      *out = ceil_div_i64(t, s.patches_per_tile);
      return 0;
    case VRAG_POOL_ADAPTIVE_ROWS: {
      if (gh <= 0 || gw <= 0) return fail("grid_h and grid_w must be > 0");
      if (static_cast<int64_t>(gh) * gw != t)
        return fail("Expected %lld visual tokens for grid_h x grid_w=%dx%d, got %lld", (long long)gh * gw, gh, gw, (long long)t);
      if (gh > 256) return fail("grid_h=%d exceeds the supported maximum 256", gh);
      int r = s.target_rows > 0 ? s.target_rows : gh;
      if (s.clamp_to_h && r > gh) r = gh;
      *out = r;
      return 0;
    }
    case VRAG_POOL_SEQ_CHUNKS:
      if (s.target_rows <= 0) return fail("target_rows must be > 0");
      if (t < 1) return fail("embedding must be non-empty");
      *out = s.target_rows;
      return 0;
    case VRAG_POOL_COLSMOL_EXPERIMENTAL: {
      if (s.patches_per_tile <= 0) return fail("patches_per_tile must be > 0");
      int64_t nt = s.num_tiles > 0 ? s.num_tiles : ceil_div_i64(t, s.patches_per_tile);
      if ((nt - 1) * s.patches_per_tile >= t) nt = ceil_div_i64(t, s.patches_per_tile);
      if (nt <= 0) return fail("Not enough tokens for num_tiles=%d, patches_per_tile=%d: got %lld", s.num_tiles, s.patches_per_tile, (long long)t);
      const int64_t last = (nt - 1) * s.patches_per_tile;
      *out = (nt - 1) + std::min<int64_t>(s.patches_per_tile, t - last);
      return 0;
    }
    case VRAG_POOL_LEGACY_CONV:
      if (t < 1) return fail("row_vectors must be non-empty");
      if (s.window < 1) return fail("window_size must be >= 1");
      if (s.window % 2 == 0) return fail("window_size must be odd");
      if (s.window == 1 || t == 1) *out = t;
      else if (s.window == 3 && t == 2) *out = 3;
      else *out = t + 2 * (s.window / 2);
      return 0;
    case VRAG_POOL_SMOOTH:
      if (t < 1) return fail("row_vectors must be non-empty");
      if (s.window < 1) return fail("window_size must be >= 1");
      if (s.window > kPoolMaxWeights) return fail("window_size %d exceeds the supported maximum %d", s.window, kPoolMaxWeights);
      if (s.window > 1 && s.n_weights != s.window) return fail("SMOOTH needs window weights");
      *out = t;
      return 0;
    case VRAG_POOL_TILE_4N: {
      if (gh <= 0 || gw <= 0) return fail("n_rows and n_cols must be > 0");
      const int64_t g = static_cast<int64_t>(gh) * gw;
      if (t < g) return fail("Expected at least %lld tile vectors for n_rows x n_cols=%dx%d, got %lld", (long long)g, gh, gw, (long long)t);
      if (!s.include_self && g == 1) return fail("need at least one array to stack");
      *out = g + ((s.has_global && t > g) ? 1 : 0);
      return 0;
    }
    case VRAG_POOL_GLOBAL_MEAN:
      *out = 1;
      return 0;
    default:
      return fail("unknown pooling kind %d", s.kind);
  }
}

static bool pool_is_row_level(int kind) {
  return kind == VRAG_POOL_SMOOTH || kind == VRAG_POOL_TILE_4N;
}
static void pool_grid_of(const vrag_pool_spec_t& s, const int32_t* grid_hw, int64_t page, int* gh, int* gw) {
  if (s.kind == VRAG_POOL_TILE_4N) {
    *gh = s.n_rows;
    *gw = s.n_cols;
  } else {
    *gh = s.grid_h;
    *gw = s.grid_w;
  }
  if (grid_hw) {
    *gh = grid_hw[2 * page];
    *gw = grid_hw[2 * page + 1];
  }
}

extern "C" int vrag_pool_out_rows(const vrag_pool_spec_t* spec, int64_t in_rows, int64_t* out_rows) {
  if (!spec || !out_rows) return fail("NULL argument");
  int gh, gw;
  pool_grid_of(*spec, nullptr, 0, &gh, &gw);
  return pool_out_rows(*spec, in_rows, gh, gw, out_rows);
}

static PoolSpecDev to_dev_spec(const vrag_pool_spec_t& s) {
  PoolSpecDev d;
  memset(&d, 0, sizeof(d));
  d.kind = s.kind;
  d.ppt = s.patches_per_tile;
  d.grid_h = s.grid_h;
  d.grid_w = s.grid_w;
  d.target_rows = s.target_rows;
  d.clamp_to_h = s.clamp_to_h;
  d.num_tiles = s.num_tiles;
  d.window = s.window;
  d.n_weights = s.n_weights;
  for (int i = 0; i < kPoolMaxWeights; ++i) d.weights[i] = s.weights[i];
  d.n_rows = s.n_rows;
  d.n_cols = s.n_cols;
  d.has_global = s.has_global;
  d.include_self = s.include_self;
  d.via_f16 = s.via_f16;
  d.keep_f32 = s.derive_from_f32;
  return d;
}

// Launch the right kernel(s) for `specs` over `in`; every dev spec already has its output pointers set.
// Specs with input_spec == k > 0 are DERIVED from the output of spec k-1 (a token-level spec of the same call) and
// are computed inside that spec's pass from the pooled rows held in shared memory. max_out[i]: largest row count
// spec i produces for any page (shared-memory sizing).
static int launch_pool(const PoolInput& in, int n, const vrag_pool_spec_t* specs, PoolSpecDev* dev, int max_in_rows,
                       int max_grid_h, const int* max_out, int num_sms, cudaStream_t st, int64_t* launches) {
  static PerDeviceOnce once;
  if (once.first()) {
    CUDA_OK(cudaFuncSetAttribute(pool_tokens_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * 512));
    CUDA_OK(cudaFuncSetAttribute(pool_tokens_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * 512));
    CUDA_OK(cudaFuncSetAttribute(pool_tokens_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * 512));
    CUDA_OK(cudaFuncSetAttribute(pool_tokens_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * 512));
    CUDA_OK(cudaFuncSetAttribute(pool_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 512));
  }
  if (in.n_pages == 0) return 0;
  PoolRowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.in = in;
  for (int i = 0; i < n; ++i) {
    if (specs[i].input_spec == 0 && pool_is_row_level(specs[i].kind)) {
      if (ra.n_specs >= kPoolMaxSpecs) return fail("too many row-level pooling specs");
      ra.specs[ra.n_specs++] = dev[i];
    }
  }
  // LEGACY_CONV / GLOBAL_MEAN ride along with the row-level pass when the input pages are small enough
  const bool small_pages = max_in_rows <= 256;
  // TILE_MEAN + COLSMOL_EXPERIMENTAL (default num_tiles, same patches_per_tile) share one pass over the tokens
  int fused_exp = -1;
  for (int i = 0; i < n && fused_exp < 0; ++i) {
    if (specs[i].kind != VRAG_POOL_TILE_MEAN || specs[i].input_spec != 0) continue;
    for (int j = 0; j < n; ++j) {
      if (specs[j].kind == VRAG_POOL_COLSMOL_EXPERIMENTAL && specs[j].input_spec == 0 && specs[j].num_tiles <= 0 &&
          specs[j].patches_per_tile == specs[i].patches_per_tile && dev[j].out_f32 == dev[i].out_f32 &&
          specs[j].in_row_skip == specs[i].in_row_skip && specs[j].in_row_count == specs[i].in_row_count) {
        dev[i].out2 = dev[j].out;
        dev[i].out2_off = dev[j].out_off;
        dev[i].out2_fixed = dev[j].out_fixed;
        fused_exp = j;
        break;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    const int k = specs[i].kind;
    if (specs[i].input_spec != 0 || pool_is_row_level(k) || i == fused_exp) continue;
    if (ra.n_specs > 0 && small_pages && (k == VRAG_POOL_LEGACY_CONV || k == VRAG_POOL_GLOBAL_MEAN) &&
        ra.n_specs < kPoolMaxSpecs && specs[i].in_row_skip == 0 && specs[i].in_row_count <= 0) {
      ra.specs[ra.n_specs++] = dev[i];
      continue;
    }
    // specs derived from this one
    PoolRowsArgs d;
    memset(&d, 0, sizeof(d));
    d.in = in;
    for (int j = 0; j < n; ++j)
      if (specs[j].input_spec == i + 1) d.specs[d.n_specs++] = dev[j];
    const int grid_rows = (k == VRAG_POOL_ADAPTIVE_ROWS) ? std::max(max_grid_h, 1) : 0;
    const int keep_rows = d.n_specs > 0 ? std::max(max_out[i], 1) : 0;
    if (grid_rows + keep_rows > 384) return fail("pooling: %d staged rows per page exceed shared memory", grid_rows + keep_rows);
    const size_t smem = static_cast<size_t>(grid_rows + keep_rows) * 512;
    const unsigned grid = static_cast<unsigned>(std::min<long long>(in.n_pages, static_cast<long long>(num_sms) * 8));
    PoolInput win = in;
    win.row_skip = specs[i].in_row_skip;
    win.row_count = specs[i].in_row_count;
#define VRAG_TOKENS(D, F) pool_tokens_kernel<D, F><<<grid, 256, smem, st>>>(win, dev[i], d, grid_rows * 128)
    if (d.n_specs > 0) { if (in.in_f32) VRAG_TOKENS(true, true); else VRAG_TOKENS(true, false); }
    else { if (in.in_f32) VRAG_TOKENS(false, true); else VRAG_TOKENS(false, false); }
#undef VRAG_TOKENS
    if (launches) ++*launches;
  }
  if (ra.n_specs > 0) {
    if (max_in_rows > 256) return fail("SMOOTH / TILE_4N need pages of at most 256 rows (got %d)", max_in_rows);
    const size_t smem = static_cast<size_t>(std::max(max_in_rows, 1)) * 512;
    // small pages: latency-bound per block (load -> sync -> compute), so keep every SM full of blocks
    const int per_sm = std::max(1, std::min(16, static_cast<int>((200 * 1024) / std::max<size_t>(smem, 1))));
    const unsigned grid = static_cast<unsigned>(std::min<long long>(in.n_pages, static_cast<long long>(num_sms) * per_sm));
    pool_rows_kernel<<<grid, 128, smem, st>>>(ra);
    if (launches) ++*launches;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

// Per-device context of the single-page pooling calls (the drop-in for calling a pooling.py function on one numpy
// array, ~10 calls per indexed page): one stream, grow-only device buffers and pinned staging, created on first use and
// kept for the life of the process — a call is two copies, one launch and one synchronisation (no allocation, no stream
// creation, no device-property query on the call path). Calls on one device serialise on the context's lock.
struct PoolPageCtx {
  std::mutex mu;
  bool ready = false;
  int num_sms = 0;
  cudaStream_t st = nullptr;
  void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
  long long* d_off = nullptr;    // {0, 0}: the offsets of an empty page
  size_t in_cap = 0, out_cap = 0;
};
static PoolPageCtx g_pool_ctx[64];

static int pool_ctx_ensure(PoolPageCtx& cx, int device, size_t in_b, size_t out_b) {
  if (!cx.ready) {
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    cx.num_sms = prop.multiProcessorCount;
    CUDA_OK(cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&cx.d_off, 2 * sizeof(long long)));
    CUDA_OK(cudaMemset(cx.d_off, 0, 2 * sizeof(long long)));
    cx.ready = true;
  }
  if (in_b > cx.in_cap) {
    if (cx.d_in) cudaFree(cx.d_in);
    if (cx.h_in) cudaFreeHost(cx.h_in);
    cx.d_in = cx.h_in = nullptr;
    cx.in_cap = 0;
    const size_t want = std::max<size_t>(in_b + in_b / 2, size_t(1) << 20);
    CUDA_OK(cudaMalloc(&cx.d_in, want));
    CUDA_OK(cudaMallocHost(&cx.h_in, want));
    cx.in_cap = want;
  }
  if (out_b > cx.out_cap) {
    if (cx.d_out) cudaFree(cx.d_out);
    if (cx.h_out) cudaFreeHost(cx.h_out);
    cx.d_out = cx.h_out = nullptr;
    cx.out_cap = 0;
    const size_t want = std::max<size_t>(out_b + out_b / 2, size_t(1) << 18);
    CUDA_OK(cudaMalloc(&cx.d_out, want));
    CUDA_OK(cudaMallocHost(&cx.h_out, want));
    cx.out_cap = want;
  }
  return 0;
}

extern "C" int vrag_pool_page(int device, const vrag_pool_spec_t* spec, const void* in, int in_dtype, int64_t in_rows,
                              void* out, int out_dtype, int64_t out_capacity_rows, int64_t* out_rows) {
  if (!spec || !out_rows) return fail("NULL argument");
  if (spec->input_spec != 0) return fail("input_spec must be 0 for a single pooling call");
  if (spec->in_row_skip != 0 || spec->in_row_count != 0) return fail("token windows apply to vrag_store_pool only");
  if ((in_dtype != VRAG_F16 && in_dtype != VRAG_F32) || (out_dtype != VRAG_F16 && out_dtype != VRAG_F32))
    return fail("unknown dtype");
  if (in_rows < 0) return fail("in_rows < 0");
  int gh, gw;
  pool_grid_of(*spec, nullptr, 0, &gh, &gw);
  int64_t n_out = 0;
  TRY(pool_out_rows(*spec, in_rows, gh, gw, &n_out));
  *out_rows = n_out;
  if (n_out > out_capacity_rows) return fail("output buffer too small: need %lld rows", (long long)n_out);
  if (n_out == 0) return 0;
  if (!in && in_rows > 0) return fail("in is NULL");
  if (!out) return fail("out is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: libvrag_b200 has no CPU fallback");
  if (device < 0 || device >= ndev || device >= 64) return fail("device %d out of range (have %d)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  const size_t in_b = static_cast<size_t>(in_rows) * 128 * (in_dtype == VRAG_F32 ? 4 : 2);
  const size_t out_b = static_cast<size_t>(n_out) * 128 * (out_dtype == VRAG_F32 ? 4 : 2);
  PoolPageCtx& cx = g_pool_ctx[device];
  std::lock_guard<std::mutex> lock(cx.mu);
  TRY(pool_ctx_ensure(cx, device, in_b, out_b));
  if (in_b) {
    memcpy(cx.h_in, in, in_b);   // pinned staging: the copy below is a real async DMA, not a pageable-memory bounce
    CUDA_OK(cudaMemcpyAsync(cx.d_in, cx.h_in, in_b, cudaMemcpyHostToDevice, cx.st));
  }
  PoolInput pin;
  memset(&pin, 0, sizeof(pin));
  pin.in = cx.d_in;
  pin.in_f32 = in_dtype == VRAG_F32;
  pin.in_fixed = std::max<int64_t>(in_rows, 1);
  pin.n_pages = 1;
  PoolSpecDev dev = to_dev_spec(*spec);
  dev.out = cx.d_out;
  dev.out_f32 = out_dtype == VRAG_F32;
  dev.out_fixed = n_out;
  if (in_rows == 0) {   // empty page (GLOBAL_MEAN of an empty mean-pool): describe it with offsets {0,0}
    pin.in_off = cx.d_off;
    pin.in_fixed = 0;
  }
  const int mo = static_cast<int>(n_out);
  TRY(launch_pool(pin, 1, spec, &dev, static_cast<int>(in_rows), gh, &mo, cx.num_sms, cx.st, nullptr));
  CUDA_OK(cudaMemcpyAsync(cx.h_out, cx.d_out, out_b, cudaMemcpyDeviceToHost, cx.st));
  cudaError_t e = cudaStreamSynchronize(cx.st);
  if (e != cudaSuccess) return fail("pooling kernel failed: %s", cudaGetErrorString(e));
  memcpy(out, cx.h_out, out_b);
  return 0;
}

extern "C" int vrag_store_pool(vrag_corpus_t* c, const char* src, int n_specs, const vrag_pool_spec_t* specs,
                               const char* const* dst_names, const int32_t* grid_hw) {
  VRAG_LOCK(c);
  Store* sp;
  TRY(find_store(c, src, &sp));
  TRY(set_device(c));
  if (sp->indirect) TRY(compact_store(c, *sp, src));   // the pooling kernels read pages as contiguous row ranges
  if (n_specs < 1 || n_specs > kPoolMaxSpecs) return fail("n_specs %d out of range [1,%d]", n_specs, kPoolMaxSpecs);
  if (!specs || !dst_names) return fail("NULL argument");
  const int64_t n_pages = sp->n_pages;
  // 1. shapes
  std::vector<std::vector<int64_t>> offs(n_specs);
  std::vector<int64_t> fixed(n_specs, 0);
  std::vector<int> max_out(n_specs, 0);
  int max_gh = 0;
  for (int i = 0; i < n_specs; ++i) {
    if (!dst_names[i] || !*dst_names[i]) return fail("dst name %d is empty", i);
    if (std::string(dst_names[i]) == src) return fail("dst store must differ from src");
    for (int j = 0; j < i; ++j)
      if (std::string(dst_names[i]) == dst_names[j]) return fail("dst store '%s' is named twice", dst_names[i]);
    const int par = specs[i].input_spec - 1;   // -1: the source store
    if (par >= i) return fail("spec %d: input_spec must name an earlier spec", i);
    if (par >= 0) {
      const int pk = specs[par].kind;
      if (specs[par].input_spec != 0 || pool_is_row_level(pk))
        return fail("spec %d: derived specs must hang off a token-level spec of the source store", i);
      const int k = specs[i].kind;
      if (!(pool_is_row_level(k) || k == VRAG_POOL_LEGACY_CONV || k == VRAG_POOL_GLOBAL_MEAN))
        return fail("spec %d: kind %d cannot be derived from pooled rows", i, k);
    }
    offs[i].resize(n_pages + 1);
    offs[i][0] = 0;
    bool all_same = true;
    for (int64_t p = 0; p < n_pages; ++p) {
      int64_t t = par >= 0 ? (offs[par][p + 1] - offs[par][p])
                           : (sp->fixed_rows > 0 ? sp->fixed_rows : (sp->h_offsets[p + 1] - sp->h_offsets[p]));
      if (par < 0 && !pool_is_row_level(specs[i].kind)) {   // token window of a token-level spec
        const int64_t skip = std::min<int64_t>(std::max(specs[i].in_row_skip, 0), t);
        t -= skip;
        if (specs[i].in_row_count > 0) t = std::min<int64_t>(t, specs[i].in_row_count);
      }
      int gh, gw;
      pool_grid_of(specs[i], grid_hw, p, &gh, &gw);
      int64_t r = 0;
      TRY(pool_out_rows(specs[i], t, gh, gw, &r));
      if (specs[i].kind == VRAG_POOL_ADAPTIVE_ROWS) max_gh = std::max(max_gh, gh);
      offs[i][p + 1] = offs[i][p] + r;
      max_out[i] = std::max<int>(max_out[i], static_cast<int>(r));
      if (p > 0 && r != offs[i][1]) all_same = false;
    }
    if (all_same && n_pages > 0 && offs[i][1] > 0) fixed[i] = offs[i][1];
  }
  // 2. allocate destination stores, device offsets, per-page grids (the temporaries are freed on every return path)
  struct Temps {
    int* grid = nullptr;
    std::vector<long long*> offs;
    ~Temps() {
      if (grid) cudaFree(grid);
      for (auto p : offs)
        if (p) cudaFree(p);
    }
  } tmp;
  tmp.offs.assign(n_specs, nullptr);
  int*& d_grid = tmp.grid;
  std::vector<long long*>& d_offs = tmp.offs;
  if (grid_hw && n_pages > 0) {
    CUDA_OK(cudaMalloc(&d_grid, n_pages * 2 * sizeof(int)));
    CUDA_OK(cudaMemcpyAsync(d_grid, grid_hw, n_pages * 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  std::vector<PoolSpecDev> dev(n_specs);
  std::vector<Store*> dst(n_specs);
  // NOTE: alloc_store may rehash the map; std::map keeps element addresses stable, and `sp` stays valid.
  for (int i = 0; i < n_specs; ++i) {
    TRY(alloc_store(c, dst_names[i], offs[i][n_pages], &dst[i]));
    dev[i] = to_dev_spec(specs[i]);
    dev[i].out = dst[i]->rows;
    dev[i].out_f32 = 0;
    dev[i].out_fixed = fixed[i];
    if (fixed[i] == 0 && n_pages > 0) {
      CUDA_OK(cudaMalloc(&d_offs[i], (n_pages + 1) * sizeof(long long)));
      CUDA_OK(cudaMemcpyAsync(d_offs[i], offs[i].data(), (n_pages + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
      dev[i].out_off = d_offs[i];
    }
  }
  TRY(find_store(c, src, &sp));
  PoolInput pin;
  memset(&pin, 0, sizeof(pin));
  pin.in = sp->rows;
  pin.in_f32 = 0;
  pin.in_off = sp->offsets;
  pin.in_fixed = sp->fixed_rows;
  pin.n_pages = n_pages;
  pin.grid_hw = d_grid;
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  int rc = launch_pool(pin, n_specs, specs, dev.data(), static_cast<int>(sp->max_rows), max_gh, max_out.data(), c->num_sms,
                       c->stream, &c->launches);
  CUDA_OK(cudaEventRecord(c->ev1, c->stream));
  if (rc == 0) {
    for (int i = 0; i < n_specs; ++i) {
      const int64_t tr = offs[i][n_pages];
      if (tr > 0) {
        inv_norm_kernel<<<static_cast<unsigned>((tr * 16 + 255) / 256), 256, 0, c->stream>>>(dst[i]->rows, tr, dst[i]->inv);
        c->launches++;
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return fail("pooling kernels failed: %s", cudaGetErrorString(e));
  cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1);
  c->last_ms[1] = c->last_ms[0];
  for (int i = 0; i < n_specs; ++i)
    TRY(finish_store(c, *dst[i], fixed[i] > 0 ? nullptr : offs[i].data(), n_pages, fixed[i]));
  return 0;
}

extern "C" int vrag_last_timing(vrag_corpus_t* c, float* out_ms, int n) {
  VRAG_LOCK(c);
  if (!c || !out_ms) return fail("NULL argument");
  for (int i = 0; i < n && i < 2; ++i) out_ms[i] = c->last_ms[i];
  return 0;
}
extern "C" int64_t vrag_launch_count(vrag_corpus_t* c) { return c ? c->launches : 0; }
