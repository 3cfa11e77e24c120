/*
 * vrag_host.c — a plain C99 host of libvrag_b200.so: no Python, no torch, no C++.
 *
 * Shows (and, on a GPU box, checks) that the drop-in boundary of this repo is the C ABI of include/vrag_b200.h and nothing
 * above it: a host in any language that can call C does what this file does. The reference's own seam for the path is
 * `client.query_points(...)` on a Qdrant collection (visual_rag/retrieval/two_stage.py:102-191); here the collection is a
 * vrag_corpus_t and the two-stage query is ONE vrag_search_multistage call.
 *
 *   vrag_host single [pages] [tokens]   one GPU: synthetic `initial` store -> device pooling -> two-stage search; the
 *                                       lists are checked against vrag_score + a host-side stable sort.
 *   vrag_host sharded N [pages] [tokens] N processes (fork), rank r on GPU r, each owning pages [r*pages/N, (r+1)*pages/N):
 *                                       the communicator id travels through pipes, every search is collective, and every
 *                                       rank checks the merged lists against a second, unsharded handle holding the
 *                                       whole corpus (rank-local scan + one exchange per stage == one big shard).
 * Exit code 0 = every check passed.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "vrag_b200.h"

#define DIM 128
#define CHECK(call)                                                                      \
  do {                                                                                   \
    if ((call) != 0) {                                                                   \
      fprintf(stderr, "%s:%d: %s failed: %s\n", __FILE__, __LINE__, #call, vrag_last_error()); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static float frand(void) { /* uniform in (-1, 1), deterministic */
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return (float)((double)(lcg_state >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

/* order[] = indices of score[] by score descending, ties -> lower index (Python's stable sort(reverse=True)) */
static const float* g_sort_scores;
static int cmp_desc_stable(const void* a, const void* b) {
  const int64_t i = *(const int64_t*)a, j = *(const int64_t*)b;
  const float x = g_sort_scores[i], y = g_sort_scores[j];
  if (x > y) return -1;
  if (x < y) return 1;
  return i < j ? -1 : (i > j ? 1 : 0);
}

/* initial (tokens rows per page, synthetic, seeded by GLOBAL row index) + mean_pooling (32 rows per page, pooled on the device) */
static int build_shard(vrag_corpus_t* c, int64_t first_page, int64_t n_pages, int tokens) {
  vrag_pool_spec_t spec;
  const char* dst[1] = {"mean_pooling"};
  CHECK(vrag_store_add_synthetic(c, "initial", NULL, n_pages, tokens, 1234u, first_page * tokens));
  memset(&spec, 0, sizeof(spec));
  spec.kind = VRAG_POOL_SEQ_CHUNKS; /* works for any token count */
  spec.target_rows = 32;
  CHECK(vrag_store_pool(c, "initial", 1, &spec, dst, NULL));
  return 0;
}

/* two-stage search (pooled prefetch -> exact rerank) + the exhaustive top-k */
static int run_queries(vrag_corpus_t* c, const float* q, int q_rows, int prefetch_k, int top_k, float* sc2, int64_t* id2, int* cnt2,
                       float* sc1, int64_t* id1, int* cnt1) {
  const char* names[2] = {"mean_pooling", "initial"};
  const uint32_t flags[2] = {VRAG_Q_NORMALIZE, VRAG_Q_NORMALIZE};
  int ks[2];
  ks[0] = prefetch_k;
  ks[1] = top_k;
  CHECK(vrag_search_multistage(c, 2, names, flags, ks, q, q_rows, NULL, NULL, 0, sc2, id2, cnt2));
  CHECK(vrag_search(c, "initial", q, q_rows, VRAG_Q_NORMALIZE, NULL, 0, top_k, sc1, id1, cnt1));
  return 0;
}

static int check_against_scores(vrag_corpus_t* c, int64_t n_pages, const float* q, int q_rows, int prefetch_k, int top_k,
                                const float* sc2, const int64_t* id2, const int* cnt2, const float* sc1, const int64_t* id1, int cnt1) {
  float* all = (float*)malloc(sizeof(float) * (size_t)n_pages);
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_pages);
  float* re = (float*)malloc(sizeof(float) * (size_t)prefetch_k);
  int64_t i;
  int bad = 0;
  /* stage 1 == stable top-prefetch_k of the pooled scores */
  CHECK(vrag_score(c, "mean_pooling", q, q_rows, VRAG_Q_NORMALIZE, NULL, 0, all));
  for (i = 0; i < n_pages; ++i) order[i] = i;
  g_sort_scores = all;
  qsort(order, (size_t)n_pages, sizeof(int64_t), cmp_desc_stable);
  if (cnt2[0] != (prefetch_k < n_pages ? prefetch_k : (int)n_pages)) bad |= 1;
  for (i = 0; i < cnt2[0] && !bad; ++i)
    if (id2[i] != order[i] || sc2[i] != all[order[i]]) bad |= 2;
  /* stage 2 == stable top-k of the exact scores of the stage-1 survivors (candidate order = stage-1 order) */
  CHECK(vrag_score(c, "initial", q, q_rows, VRAG_Q_NORMALIZE, id2, cnt2[0], re));
  for (i = 0; i < cnt2[0]; ++i) order[i] = i;
  g_sort_scores = re;
  qsort(order, (size_t)cnt2[0], sizeof(int64_t), cmp_desc_stable);
  for (i = 0; i < cnt2[1] && !bad; ++i)
    if (id2[prefetch_k + i] != id2[order[i]] || sc2[prefetch_k + i] != re[order[i]]) bad |= 4;
  if (cnt2[1] != (top_k < cnt2[0] ? top_k : cnt2[0])) bad |= 8;
  /* exhaustive == stable top-k of all exact scores */
  CHECK(vrag_score(c, "initial", q, q_rows, VRAG_Q_NORMALIZE, NULL, 0, all));
  for (i = 0; i < n_pages; ++i) order[i] = i;
  g_sort_scores = all;
  qsort(order, (size_t)n_pages, sizeof(int64_t), cmp_desc_stable);
  for (i = 0; i < cnt1 && !bad; ++i)
    if (id1[i] != order[i] || sc1[i] != all[order[i]]) bad |= 16;
  free(all);
  free(order);
  free(re);
  if (bad) fprintf(stderr, "list check failed (mask %d)\n", bad);
  return bad;
}

static void make_query(float* q, int q_rows) {
  int i;
  for (i = 0; i < q_rows * DIM; ++i) q[i] = frand();
}

static int run_single(int64_t n_pages, int tokens) {
  enum { Q = 20, PRE = 256, K = 10 };
  vrag_corpus_t* c = NULL;
  float q[Q * DIM], sc2[PRE + K], sc1[K], ms[2];
  int64_t id2[PRE + K], id1[K];
  int cnt2[2], cnt1 = 0, t;
  CHECK(vrag_corpus_create(0, 0, &c));
  if (build_shard(c, 0, n_pages, tokens)) return 1;
  for (t = 0; t < 3; ++t) {
    make_query(q, Q);
    if (run_queries(c, q, Q, PRE, K, sc2, id2, cnt2, sc1, id1, &cnt1)) return 1;
    if (check_against_scores(c, n_pages, q, Q, PRE, K, sc2, id2, cnt2, sc1, id1, cnt1)) return 1;
  }
  CHECK(vrag_last_timing(c, ms, 2));
  printf("single: %lld pages x %d tokens, two-stage %d -> %d and exhaustive top-%d equal score + stable sort; last search %.3f ms; "
         "%lld kernel launches\n", (long long)n_pages, tokens, PRE, K, K, ms[0], (long long)vrag_launch_count(c));
  CHECK(vrag_corpus_destroy(c));
  return 0;
}

static int run_rank(int rank, int n_ranks, int64_t n_pages, int tokens, const unsigned char* id) {
  enum { Q = 20, PRE = 256, K = 10 };
  vrag_corpus_t *shard = NULL, *whole = NULL;
  const int64_t first = n_pages * rank / n_ranks, last = n_pages * (rank + 1) / n_ranks;
  float q[Q * DIM], sc2[PRE + K], sc1[K], wsc2[PRE + K], wsc1[K], us[8];
  int64_t id2[PRE + K], id1[K], wid2[PRE + K], wid1[K];
  int cnt2[2], cnt1 = 0, wcnt2[2], wcnt1 = 0, t, peer = 0, n_us = 0;
  CHECK(vrag_corpus_create(rank, first, &shard));
  if (build_shard(shard, first, last - first, tokens)) return 1;
  CHECK(vrag_comm_init(shard, rank, n_ranks, id));
  CHECK(vrag_comm_transport(shard, &peer));
  CHECK(vrag_corpus_create(rank, 0, &whole)); /* the same corpus as ONE shard, for the comparison */
  if (build_shard(whole, 0, n_pages, tokens)) return 1;
  for (t = 0; t < 4; ++t) {
    make_query(q, Q); /* same seed on every rank: the collective calls see the same query */
    if (run_queries(shard, q, Q, PRE, K, sc2, id2, cnt2, sc1, id1, &cnt1)) return 1;
    if (run_queries(whole, q, Q, PRE, K, wsc2, wid2, wcnt2, wsc1, wid1, &wcnt1)) return 1;
    if (cnt1 != wcnt1 || cnt2[0] != wcnt2[0] || cnt2[1] != wcnt2[1] || memcmp(id1, wid1, sizeof(int64_t) * (size_t)cnt1) ||
        memcmp(sc1, wsc1, sizeof(float) * (size_t)cnt1) || memcmp(id2, wid2, sizeof(int64_t) * (size_t)cnt2[0]) ||
        memcmp(sc2, wsc2, sizeof(float) * (size_t)cnt2[0]) || memcmp(id2 + PRE, wid2 + PRE, sizeof(int64_t) * (size_t)cnt2[1]) ||
        memcmp(sc2 + PRE, wsc2 + PRE, sizeof(float) * (size_t)cnt2[1])) {
      fprintf(stderr, "rank %d: sharded lists differ from the single-shard lists (query %d)\n", rank, t);
      return 1;
    }
  }
  { /* device time of the two collectives of a two-stage search (all-gather of 256 packed hits, max-all-reduce of 256
       candidate scores): median over REP searches */
    enum { REP = 31 };
    const char* names[2] = {"mean_pooling", "initial"};
    const uint32_t flags[2] = {VRAG_Q_NORMALIZE, VRAG_Q_NORMALIZE};
    const int ks[2] = {PRE, K};
    float a[REP], b[REP], ms[2], tot[REP];
    int i, j;
    for (t = 0; t < REP; ++t) {
      CHECK(vrag_search_multistage(shard, 2, names, flags, ks, q, Q, NULL, NULL, 0, sc2, id2, cnt2));
      CHECK(vrag_last_comm_timing(shard, us, 8, &n_us));
      CHECK(vrag_last_timing(shard, ms, 2));
      a[t] = n_us > 0 ? us[0] : 0.0f;
      b[t] = n_us > 1 ? us[1] : 0.0f;
      tot[t] = ms[0] * 1000.0f;
    }
    for (i = 1; i < REP; ++i) /* insertion sorts */
      for (j = i; j > 0; --j) {
        float x;
        if (a[j] < a[j - 1]) { x = a[j]; a[j] = a[j - 1]; a[j - 1] = x; }
        if (b[j] < b[j - 1]) { x = b[j]; b[j] = b[j - 1]; b[j - 1] = x; }
        if (tot[j] < tot[j - 1]) { x = tot[j]; tot[j] = tot[j - 1]; tot[j - 1] = x; }
      }
    if (rank == 0)
      printf("sharded: %d ranks x %lld pages, transport %s: exhaustive top-%d and two-stage %d -> %d lists bit-identical to the "
             "single-shard lists on every query; two-stage search device time median %.1f us, of which all-gather(256 hits) "
             "%.1f us (min %.1f), max-all-reduce(256 scores) %.1f us (min %.1f)\n", n_ranks, (long long)(n_pages / n_ranks),
             peer ? "NVLink peer memory" : "NCCL", K, PRE, K, tot[REP / 2], a[REP / 2], a[0], b[REP / 2], b[0]);
  }
  CHECK(vrag_comm_destroy(shard));
  CHECK(vrag_corpus_destroy(shard));
  CHECK(vrag_corpus_destroy(whole));
  return 0;
}

static int run_sharded(int n_ranks, int64_t n_pages, int tokens) {
  /* fork BEFORE the first CUDA call; rank 0 creates the communicator id and sends it down one pipe per rank */
  int pipes[16][2], r, status, failed = 0;
  pid_t pids[16];
  unsigned char id[VRAG_UNIQUE_ID_BYTES];
  if (n_ranks < 2 || n_ranks > 16) {
    fprintf(stderr, "n_ranks must be 2..16\n");
    return 2;
  }
  for (r = 1; r < n_ranks; ++r)
    if (pipe(pipes[r]) != 0) return 2;
  for (r = 1; r < n_ranks; ++r) {
    pids[r] = fork();
    if (pids[r] == 0) {
      if (read(pipes[r][0], id, sizeof(id)) != (ssize_t)sizeof(id)) _exit(3);
      _exit(run_rank(r, n_ranks, n_pages, tokens, id));
    }
  }
  if (vrag_comm_unique_id(id) != 0) {
    fprintf(stderr, "vrag_comm_unique_id: %s\n", vrag_last_error());
    memset(id, 0, sizeof(id));
    failed = 1;
  }
  for (r = 1; r < n_ranks; ++r)
    if (write(pipes[r][1], id, sizeof(id)) != (ssize_t)sizeof(id)) failed = 1;
  if (!failed) failed = run_rank(0, n_ranks, n_pages, tokens, id);
  for (r = 1; r < n_ranks; ++r) {
    waitpid(pids[r], &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) failed = 1;
  }
  return failed;
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "single";
  if (strcmp(mode, "single") == 0) {
    const int64_t pages = argc > 2 ? atoll(argv[2]) : 20000;
    const int tokens = argc > 3 ? atoi(argv[3]) : 256;
    return run_single(pages, tokens);
  }
  if (strcmp(mode, "sharded") == 0 && argc > 2) {
    const int64_t pages = argc > 3 ? atoll(argv[3]) : 40000;
    const int tokens = argc > 4 ? atoi(argv[4]) : 256;
    return run_sharded(atoi(argv[2]), pages, tokens);
  }
  if (strcmp(mode, "abi") == 0) { /* needs no GPU */
    printf("abi %d\n", vrag_abi_version());
    return 0;
  }
  fprintf(stderr, "usage: vrag_host single [pages] [tokens] | sharded N [pages] [tokens] | abi\n");
  return 2;
}
