/*
 * vrag_b200.h — C ABI of libvrag_b200.so: a B200 (sm_100a) GPU-resident corpus store and the
 * multi-stage retrieval scoring / indexing-time pooling kernels of visual-rag-toolkit.
 *
 * The reference (pure Python) has no FFI; its plug-in seam for this path is the duck-typed
 * `qdrant_client` object handed to the retrievers plus the module-level pooling functions.  Every
 * entry point below names the reference interface it replaces (paths relative to the reference
 * repository root).  The Python host layer in visual-rag-toolkit_b200/visual_rag_b200 binds these
 * with ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; vrag_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - plain pointers and sizes only.  "host" pointers are ordinary CPU memory borrowed for the
 *     duration of the call; "dev" pointers are CUDA device pointers on the corpus' device.
 *   - all device memory behind a vrag_corpus_t is owned by the library.
 *   - embedding dim is 128 everywhere (visual_rag/indexing/qdrant_indexer.py:133).
 *   - page ids are int64 "global page ids": shard-local page index + page_base of the shard.
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef VRAG_B200_H
#define VRAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vrag_corpus vrag_corpus_t;

/* dtype codes for row data handed to the library */
#define VRAG_F16 0
#define VRAG_F32 1

/* query flags */
#define VRAG_Q_NORMALIZE 1u  /* L2-normalise query rows and document rows (cosine); pooling.py:495-503 */
#define VRAG_Q_POOL 2u       /* mean-pool the query tokens to one row first; two_stage.py:142,148,154 */
#define VRAG_Q_FP16 4u       /* opt-in: contract the query as plain fp16 instead of the exact fp16 hi/lo pair (scores then carry
                                the fp16 rounding of the query, ~1e-4 relative, still inside the 1e-3 parity gate); halves the
                                tensor work of scans over stores with > 128 rows per page. Default (flag clear): fp32-exact query. */

const char* vrag_last_error(void);
int vrag_abi_version(void);

/* ------------------------------------------------------------------ corpus store
 * Replaces the Qdrant collection with named vectors created by
 * QdrantIndexer.create_collection (visual_rag/indexing/qdrant_indexer.py:200-239): one named
 * multi-vector store per name ("initial", "mean_pooling", "experimental_pooling*", "global_pooling"),
 * fp16 rows resident in HBM, variable rows per page.                                              */
int vrag_corpus_create(int device, int64_t page_base, vrag_corpus_t** out);
int vrag_corpus_destroy(vrag_corpus_t* c);

/* Add (or replace) a named store.  rows: [total_rows,128] of `dtype` (fp32 is cast to fp16 exactly like
 * QdrantIndexer._build_qdrant_points, qdrant_indexer.py:423-441).  Either fixed_rows > 0 (every page has
 * that many rows) or page_offsets[n_pages+1] (host, int64 row offsets) describes the pages.
 * rows_on_device != 0: `rows` is a device pointer (copied device-to-device).                      */
int vrag_store_add(vrag_corpus_t* c, const char* name, const void* rows, int dtype, int rows_on_device,
                   const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows);

/* Append pages behind the existing pages of a named store (the store is created on first use). This is the ingest path
 * of QdrantIndexer.upload_batch (visual_rag/indexing/qdrant_indexer.py:341-507): one call per uploaded batch and named
 * vector; page index = upload order. Same arguments as vrag_store_add (page_offsets are relative to this batch).  */
int vrag_store_append(vrag_corpus_t* c, const char* name, const void* rows, int dtype, int rows_on_device,
                      const int64_t* page_offsets, int64_t n_pages, int64_t fixed_rows);

/* Overwrite existing pages — the "upsert of a point that already exists" half of QdrantIndexer.upload_batch
 * (client.upsert, visual_rag/indexing/qdrant_indexer.py:459-507; ids are deterministic, 602-613, so re-indexing a document
 * re-sends the same ids). local_pages[n_pages]: shard-local page indices; rows / page_offsets describe the new pages
 * back to back as in vrag_store_append. Pages that keep their row count are overwritten in place and the store stays on
 * the dense layout. A page whose row count changed switches the store to a PAGE TABLE (page -> row range anywhere in the
 * row buffer): a page that shrank is overwritten in place, a page that grew moves behind the last row and its old rows
 * become garbage until vrag_store_compact. Scans of a store with a page table fetch every page on its own (the gather
 * path); everything is validated before anything is written.                                                          */
int vrag_store_replace_pages(vrag_corpus_t* c, const char* name, const int64_t* local_pages, int64_t n_pages,
                             const void* rows, int dtype, int rows_on_device, const int64_t* page_offsets, int64_t fixed_rows);

/* Delete pages (qdrant `client.delete(points_selector=ids)`; the reference itself only drops whole collections,
 * qdrant_indexer.py:175). A deleted page keeps its index — the host's id tables stay valid — but owns no rows: it scores
 * -inf in every scan and is never returned. Switches the store to the page table.                                     */
int vrag_store_delete_pages(vrag_corpus_t* c, const char* name, const int64_t* local_pages, int64_t n_pages);

/* Drop the LAST pages of a store, keeping n_pages: the rollback of a batch upload that appended to several named stores
 * and failed on a later one (client.upsert is atomic per batch).                                                       */
int vrag_store_truncate(vrag_corpus_t* c, const char* name, int64_t n_pages);

/* Rewrite a store with a page table contiguously in page order (one device pass): garbage rows are reclaimed and the
 * dense fast paths apply again. Page indices do not change (deleted pages stay as zero-row pages). No-op on a dense
 * store. vrag_store_pool compacts its source automatically.                                                           */
int vrag_store_compact(vrag_corpus_t* c, const char* name);

/* Fill a named store with the seeded synthetic corpus of SURVEY.md 8(d) directly on the device
 * (gaussian rows, L2-normalised, rounded to fp16).  Row r of the store depends only on (seed, row_seed_base + r). */
int vrag_store_add_synthetic(vrag_corpus_t* c, const char* name, const int64_t* page_offsets, int64_t n_pages,
                             int64_t fixed_rows, uint64_t seed, int64_t row_seed_base);

int vrag_store_info(vrag_corpus_t* c, const char* name, int64_t* n_pages, int64_t* total_rows,
                    int64_t* fixed_rows, int64_t* max_rows);
/* Copy rows [row0,row0+n_rows) of a store to host as fp16 (qdrant `retrieve(with_vectors=[name])`,
 * two_stage.py:383-390) */
int vrag_store_read_rows(vrag_corpus_t* c, const char* name, int64_t row0, int64_t n_rows, void* out_f16_host);
int vrag_store_page_range(vrag_corpus_t* c, const char* name, int64_t local_page, int64_t* row0, int64_t* n_rows);
/* Row counts of pages [first_page, first_page + n) in one call (the token counts the bulk re-pooling script reads point by
 * point to infer each page's patch grid, scripts/qdrant_recompute_colqwen_pooling_from_initial.py:292-300). */
int vrag_store_page_rows(vrag_corpus_t* c, const char* name, int64_t first_page, int64_t n, int64_t* out_rows);
int vrag_store_drop(vrag_corpus_t* c, const char* name);

/* ------------------------------------------------------------------ scoring: one stage
 * score(page) = sum_q max_t <qhat_q, dhat_t>  — compute_maxsim_score, pooling.py:468-514, and the
 * pooled-query stage-1 score max_t cos(qbar, d_t) of quick_test.search_two_stage (benchmarks/quick_test.py:182-191)
 * when VRAG_Q_POOL is set.  Replaces client.query_points(query, using=name, limit=k)
 * (two_stage.py:349-358, single_stage.py:123-132, three_stage.py:103-157).
 * cand_ids == NULL: score every page of the store.  Otherwise only the listed global page ids
 * (HasIdCondition restriction of three_stage.py:75-81 / rerank list of two_stage.py:380-426); ids that
 * are not in this shard score -inf.  Results are sorted by score descending, ties by lower id.
 * out_count receives the number of valid results (<= k).                                          */
int vrag_search(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, uint32_t flags,
                const int64_t* cand_ids, int64_t n_cand, int k, float* out_scores, int64_t* out_ids,
                int* out_count);

/* Score only (no top-k): out_scores[i] for page i of the store (cand_ids == NULL) or candidate i.
 * Twin of compute_maxsim_batch (pooling.py:517-552).                                              */
int vrag_score(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, uint32_t flags,
               const int64_t* cand_ids, int64_t n_cand, float* out_scores);

/* The per-call twins themselves — compute_maxsim_score(query, doc) (pooling.py:468-514) and
 * compute_maxsim_batch(query, [docs]) (pooling.py:517-552) — over documents in ordinary HOST memory, one pointer per
 * document (pages[i]: page_rows[i] x 128 values of `dtype`, contiguous; no concatenation needed): the host worker pool
 * casts fp32 documents to the fp16 store dtype (numpy astype(float16), qdrant_indexer.py:423-441) into pinned staging
 * chunks while the previous chunk is in flight, one scan scores them all. out_scores[i] = score of pages[i]; an empty
 * page scores -inf (the reference raises on one). Uses a scratch store of the handle.                                   */
int vrag_score_pages(vrag_corpus_t* c, const float* query, int n_query_rows, uint32_t flags, const void* const* pages,
                     const int64_t* page_rows, int64_t n_pages, int dtype, float* out_scores);

/* The ingest cast on the host (what vrag_store_add / vrag_store_append / vrag_score_pages apply to fp32 rows that
 * arrive in host memory): dst[i] = fp16(src[i]), round to nearest even, bit-identical to numpy's astype(float16)
 * (qdrant_indexer.py:423-441). threads <= 1: the calling thread; otherwise the library's worker pool. force_scalar != 0
 * selects the portable conversion instead of F16C (both give the same bits). Needs no GPU.                            */
int vrag_host_f32_to_f16(const float* src, uint16_t* dst, int64_t n, int threads, int force_scalar);

/* ------------------------------------------------------------------ scoring: fused multi-stage
 * n_stages stages executed back to back on the device with ONE host synchronisation: stage s scores
 * store names[s] with flags[s], restricted to the survivors of stage s-1, and keeps ks[s] pages.
 *   two-stage  (TwoStageRetriever.search_server_side, two_stage.py:102-191): {pooled store, prefetch_k}, {"initial", top_k}
 *   three-stage (ThreeStageRetriever.search_server_side, three_stage.py:83-173): {global, stage1_k}, {experimental, stage2_k}, {"initial", top_k}
 * q_offsets == NULL: every stage uses all n_query_rows rows (VRAG_Q_POOL in flags[s] mean-pools them for
 * that stage). Otherwise stage s uses query rows [q_offsets[s], q_offsets[s+1]) — e.g. a client that sends the
 * mean-pooled prefetch vector and the token matrix as two separate queries (two_stage.py:142,159).
 * cand_ids != NULL restricts stage 0 to the listed global page ids (a payload / HasId filter).
 * Outputs are per stage, concatenated: stage s occupies [sum(ks[:s]), sum(ks[:s+1])) of out_scores/out_ids;
 * out_counts[s] valid entries each. Any stage size 1 <= ks[s] <= 2^20 (the reference accepts any prefetch_k / limit): up
 * to 4096 results are sorted in shared memory, longer lists by a global bitonic sort of the selected keys; the BATCHED
 * calls below keep at most 4096 results per query and stage.                                       */
int vrag_search_multistage(vrag_corpus_t* c, int n_stages, const char* const* names, const uint32_t* flags,
                           const int* ks, const float* query, int n_query_rows, const int* q_offsets,
                           const int64_t* cand_ids, int64_t n_cand, float* out_scores, int64_t* out_ids,
                           int* out_counts);

/* ------------------------------------------------------------------ payload filters in the scan
 * A filter is a bitmask over the shard's pages (bit p of word p/32 set = page p passes), built by the host from the payload
 * conditions of TwoStageRetriever.build_filter (two_stage.py:436-480) or the per_dataset scope filter of the benchmark
 * (run_qdrant_beir.py:1987-1997), uploaded once and reused by every query that carries the same filter. A page that does not
 * pass is treated as an empty page INSIDE the scan: over full-token stores none of its tiles is fetched or multiplied (a
 * 50 %-selective filter halves the scan), over pooled stores its score is masked in the epilogue; either way it scores
 * -inf and is never returned. vrag_search_multistage_filtered is vrag_search_multistage with the mask on stage 0 (later
 * stages only see stage 0's survivors); collective on a sharded handle (every rank passes the mask of its own pages).
 * For very selective filters (a few thousand pages) the candidate-list form (cand_ids) reads less.                       */
int vrag_filter_create(vrag_corpus_t* c, const uint32_t* bits, int64_t n_pages, int* out_filter);
int vrag_filter_destroy(vrag_corpus_t* c, int filter);
int vrag_search_multistage_filtered(vrag_corpus_t* c, int filter, int n_stages, const char* const* names,
                                    const uint32_t* flags, const int* ks, const float* query, int n_query_rows,
                                    const int* q_offsets, float* out_scores, int64_t* out_ids, int* out_counts);

/* ------------------------------------------------------------------ scoring: batched queries
 * n_queries independent multi-stage searches in one call and one host synchronisation (the evaluation loop of
 * benchmarks/vidore_beir_qdrant/run_qdrant_beir.py:378-402 issues them one by one; BASELINE configs[2] batches 256).
 * Stages, names, flags and ks as in vrag_search_multistage. query_rows: all query matrices concatenated, [rows,128] fp32.
 * per_stage_queries == 0: q_offsets[n_queries+1], query b = rows [q_offsets[b], q_offsets[b+1]) for every stage.
 * per_stage_queries != 0: q_offsets[n_queries*n_stages+1], (query b, stage s) = rows [q_offsets[b*n_stages+s], q_offsets[b*n_stages+s+1]).
 * Outputs are stage-major: stage s occupies [n_queries*sum(ks[:s]), +n_queries*ks[s]) of out_scores/out_ids as
 * [n_queries][ks[s]]; out_counts[s*n_queries + b] entries of row b are valid (the rest are (-inf, -1)).
 * Stage >= 1 of ALL queries runs as one kernel launch (each query on its own candidate list).        */
int vrag_search_multistage_batch(vrag_corpus_t* c, int n_stages, const char* const* names, const uint32_t* flags,
                                 const int* ks, int n_queries, const float* query_rows, const int* q_offsets,
                                 int per_stage_queries, float* out_scores, int64_t* out_ids, int* out_counts);

/* Same search, compact results: only the LAST stage's lists travel to the host (out_scores / out_ids [n_queries][ks[last]],
 * out_counts [n_queries]) plus, for every final result, the score its page had in each earlier stage
 * (out_stage_scores [n_queries][ks[last]][n_stages-1], NaN if absent) — exactly what the result dictionaries of
 * ThreeStageRetriever.search_server_side carry (score_stage1 / score_stage2 / score_stage3, three_stage.py:160-173).   */
int vrag_search_multistage_batch_final(vrag_corpus_t* c, int n_stages, const char* const* names, const uint32_t* flags,
                                       const int* ks, int n_queries, const float* query_rows, const int* q_offsets,
                                       int per_stage_queries, float* out_scores, int64_t* out_ids, float* out_stage_scores,
                                       int* out_counts);

/* Device-level form of the batched search for the sharded multi-GPU path (the caller all-gathers the per-shard lists with
 * NCCL between the stages): upload the batch once, then per stage score it on this shard — cand_ids_dev == NULL: every
 * page, with the fused top-k prefilter when allow_prefilter != 0; else per-query candidate lists [n_queries][n_cand] of
 * global page ids (ids of other shards score -inf) — and write the LOCAL top-k as [n_queries][k] device arrays; with
 * k == 0 a candidate stage writes its raw [n_queries][n_cand] scores instead (max-all-reduced across shards by the caller,
 * so that the final top-k breaks ties by candidate order exactly like a single shard).
 * vrag_batch_prefilter_failed reports (after synchronising `stream`) whether a prefiltered stage must be repeated with
 * allow_prefilter = 0; vrag_topk_batch_dev merges gathered [n_queries][n] lists (ties -> lower position).          */
int vrag_batch_upload(vrag_corpus_t* c, int n_stages, int n_queries, const float* query_rows, const int* q_offsets,
                      int per_stage_queries);
int vrag_batch_stage_dev(vrag_corpus_t* c, int stage, const char* name, uint32_t flags, int k, const int64_t* cand_ids_dev,
                         int64_t n_cand, int allow_prefilter, float* out_scores_dev, int64_t* out_ids_dev, void* stream);
int vrag_batch_prefilter_failed(vrag_corpus_t* c, void* stream, int* failed);
int vrag_topk_batch_dev(vrag_corpus_t* c, const float* scores_dev, const int64_t* ids_dev, int64_t n, int k, int n_queries,
                        float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* ------------------------------------------------------------------ saliency
 * Per-token relevance of one page for a query: out_scores[t] = max_q <qhat_q, dhat_t>, the `patch_scores` of
 * generate_saliency_map (visual_rag/visualization/saliency.py:69-79) — the column-max twin of MaxSim, computed for the
 * pages a search returned. out_rows receives the page's token count (<= capacity).                               */
int vrag_saliency(vrag_corpus_t* c, const char* name, const float* query, int n_query_rows, int64_t page_id,
                  float* out_scores, int64_t capacity, int64_t* out_rows);

/* ------------------------------------------------------------------ device-pointer variants
 * Same kernels, caller-provided device buffers and stream (cudaStream_t passed as void*): used by the
 * sharded multi-GPU path, which all-gathers per-shard top-k lists with NCCL between stages.        */
int vrag_score_dev(vrag_corpus_t* c, const char* name, const float* query_dev, int n_query_rows, uint32_t flags,
                   const int64_t* cand_ids_dev, int64_t n_cand, float* out_scores_dev, void* stream);
int vrag_topk_dev(vrag_corpus_t* c, const float* scores_dev, const int64_t* ids_dev, int64_t id_base, int64_t n,
                  int k, float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* ------------------------------------------------------------------ indexing-time pooling
 * The pooling arithmetic of visual_rag/embedding/pooling.py as CUDA kernels. One spec = one pooling
 * function call; `kind` selects it:                                                               */
#define VRAG_POOL_TILE_MEAN 0             /* tile_level_mean_pooling, pooling.py:35-98 (patches_per_tile) */
#define VRAG_POOL_ADAPTIVE_ROWS 2         /* colpali_row_mean_pooling / adaptive_row_mean_pooling_from_grid,
                                             pooling.py:101-185 (grid_h, grid_w, target_rows, clamp_to_h) */
#define VRAG_POOL_COLSMOL_EXPERIMENTAL 3  /* colsmol_experimental_pooling, pooling.py:188-232 (num_tiles, patches_per_tile) */
#define VRAG_POOL_LEGACY_CONV 4           /* colpali_experimental_pooling_from_rows, pooling.py:235-286 (window) */
#define VRAG_POOL_SMOOTH 5                /* weighted_row_smoothing_same_length, pooling.py:289-375 (window, weights) */
#define VRAG_POOL_TILE_4N 6               /* colsmol_tile_4n_pooling_from_tiles, pooling.py:378-436 (n_rows, n_cols, ...) */
#define VRAG_POOL_GLOBAL_MEAN 7           /* global_mean_pooling, pooling.py:439-465; global_pool_from_mean_pool, visual_embedder.py:837-840 */
#define VRAG_POOL_SEQ_CHUNKS 8            /* last-resort sequence chunk pooling, visual_embedder.py:824-835 (target_rows) */

typedef struct vrag_pool_spec {
  int kind;
  int patches_per_tile;
  int grid_h, grid_w;     /* ADAPTIVE_ROWS: fixed grid (ignored when a per-page grid array is given) */
  int target_rows;        /* ADAPTIVE_ROWS / SEQ_CHUNKS; <= 0: all grid rows */
  int clamp_to_h;         /* ADAPTIVE_ROWS: rows = min(target_rows, grid_h) — the ColQwen2.5 cap, visual_embedder.py:791-793 */
  int num_tiles;          /* COLSMOL_EXPERIMENTAL: requested num_tiles (<= 0: ceil(T / patches_per_tile)) */
  int window;             /* LEGACY_CONV / SMOOTH */
  int n_weights;          /* SMOOTH: window normalised fp32 taps (pooling.py:329-355) */
  float weights[16];
  int n_rows, n_cols, has_global, include_self; /* TILE_4N */
  int via_f16;            /* GLOBAL_MEAN: round the fp32 mean to fp16 first (numpy's fp16 mean, pooling.py:463) */
  int derive_from_f32;    /* on a spec that others derive from (input_spec): hand them its fp32 rows BEFORE the store-dtype
                             rounding — the arithmetic of scripts/qdrant_recompute_colqwen_pooling_from_initial.py:292-327,
                             which pools and smooths in fp32 and lets the collection round on write. Default 0: derived specs
                             see the rows as stored (the pipeline's dtype chain, visual_embedder.py:776-799).            */
  int input_spec;         /* vrag_store_pool only. 0: pool the source store. k > 0: pool the OUTPUT of spec k-1 of the same
                             call (a token-level spec): the pipeline's "experimental / global pooling of the mean-pooled
                             rows" (pipeline.py:452-507), computed inside that spec's pass from shared memory — the pooled
                             rows are used as stored (rounded to the store dtype), never re-read from HBM.            */
  int in_row_skip;        /* token-level kinds of vrag_store_pool (input_spec == 0): pool only rows [in_row_skip, in_row_skip +  */
  int in_row_count;       /* in_row_count) of every source page (count <= 0: to the page end) — the pipeline pools the VISUAL
                             tokens of a page (visual_token_indices, pipeline.py:400-430) while `initial` also keeps the
                             instruction tokens: ColPali-v1.3 pages are 1024 visual + 6 text tokens.   */
} vrag_pool_spec_t;

/* Rows this spec produces for a page of in_rows rows (host arithmetic only; validates the arguments with the
 * reference's error conditions, pooling.py:118,158,168,205-216,262-268,317-321,405-410).            */
int vrag_pool_out_rows(const vrag_pool_spec_t* spec, int64_t in_rows, int64_t* out_rows);

/* One pooling call on one page: in [in_rows,128] (host, VRAG_F16/VRAG_F32) -> out (host, VRAG_F16/VRAG_F32).
 * The drop-in for calling a pooling.py function on one numpy array.                                 */
int vrag_pool_page(int device, const vrag_pool_spec_t* spec, const void* in, int in_dtype, int64_t in_rows,
                   void* out, int out_dtype, int64_t out_capacity_rows, int64_t* out_rows);

/* Bulk pooling on the device: derive n_specs named stores from store `src` without leaving HBM (the
 * per-page orchestration of ProcessingPipeline._process_single_page, pipeline.py:400-507, and
 * scripts/qdrant_recompute_colqwen_pooling_from_initial.py:292-327, for a whole collection).
 * Token-level kinds read `src` once each; SMOOTH / TILE_4N / LEGACY_CONV / GLOBAL_MEAN specs are all
 * produced in ONE pass over `src`, or — chained with input_spec — inside the pass of the token-level spec they derive from. grid_hw: optional host [n_pages][2] per-page (grid_h, grid_w) /
 * (n_rows, n_cols). Outputs are rounded to the fp16 store dtype (qdrant_indexer.py:423-441).        */
int vrag_store_pool(vrag_corpus_t* c, const char* src, int n_specs, const vrag_pool_spec_t* specs,
                    const char* const* dst_names, const int32_t* grid_hw);

/* ------------------------------------------------------------------ multi-GPU: page-sharded corpus
 * One process per GPU. Every rank owns a contiguous page range of every named store: its vrag_corpus_t is created with
 * page_base = the first global page id of its range. The reference itself is single-process; its scale-out is Qdrant's
 * server-side sharding behind the SAME query_points call (two_stage.py:162-178, three_stage.py:103-157) — these entry points
 * give the GPU backend that property.
 *
 * After vrag_comm_init the search entry points of the handle — vrag_search, vrag_search_multistage,
 * vrag_search_multistage_batch(_final), vrag_search_multistage_dev — are COLLECTIVE: all ranks call them with the same
 * queries and stage arguments and every rank receives the same merged GLOBAL lists (score descending, ties -> lower global
 * page id, i.e. exactly the single-shard order). The exchange is ONE collective per stage:
 *   - a stage that scans the store (or a rank-local candidate list, see below): each rank's local top-k travels as packed
 *     16-byte vrag_hit_t entries, written by the top-k kernel straight into the send buffer, in one all-gather
 *     (k = 10: 160 B per rank), followed by the same deterministic merge on every rank;
 *   - a stage restricted to the previous stage's (replicated) survivors: every candidate is owned by exactly one rank
 *     (-inf elsewhere), so one max-all-reduce of the candidate score vector completes it and ties keep candidate order.
 * cand_ids of a collective search is RANK-LOCAL: each rank lists (in ascending id order) the pages of its own range that
 * pass the filter; ids of other ranks are ignored. Reference semantics are kept: the global top-prefetch_k is formed
 * before the rerank (two_stage.py:161-178).
 * Transport: NVLink / NVSwitch peer memory where every rank can map every other rank's exchange window (CUDA IPC; see
 * vrag_comm_transport) — one kernel per collective — else NCCL; NCCL also bootstraps the windows and carries messages above
 * 4 MB per rank. NCCL is resolved at run time with dlopen("libnccl.so.2"), so that a process that already loaded a NCCL —
 * e.g. through torch — shares it; the library has no link-time NCCL dependency. The collectives of one handle must be issued
 * from one thread at a time and, for the *_dev entry points, on ONE stream (they are numbered in issue order).
 * A host with no Python in it that does all of this: tests/c_host/vrag_host.c (`vrag_host sharded N`).                 */
typedef struct vrag_hit {
  float score;
  uint32_t aux;   /* bit 0: this rank's list came from a top-k estimate that missed; the search is repeated exactly */
  int64_t id;     /* global page id; < 0: padding (the rank had fewer than k results) */
} vrag_hit_t;

#define VRAG_UNIQUE_ID_BYTES 128
/* Rank 0 creates the communicator id (ncclGetUniqueId); the host distributes the 128 bytes to the other ranks through any
 * channel it has (MPI, a TCP store, a file) and every rank calls vrag_comm_init. Needs no corpus handle.               */
int vrag_comm_unique_id(void* out_id128);
/* Join the communicator (collective: all nranks ranks call it). The handle's device is the rank's GPU.              */
int vrag_comm_init(vrag_corpus_t* c, int rank, int nranks, const void* unique_id128);
int vrag_comm_info(vrag_corpus_t* c, int* rank, int* nranks);
/* Which transport the handle's collectives use: *peer_memory = 1 when every rank could map every other rank's exchange
 * window (CUDA IPC over NVLink / NVSwitch peer memory): messages up to 4 MB per rank then travel as direct stores into the
 * peers' windows followed by a flag, in ONE kernel per collective (store to all peers -> publish -> wait -> consume);
 * 0: NCCL (also for larger messages, and when VRAG_P2P=0 or the mapping failed on any rank — agreed by all ranks at
 * vrag_comm_init).                                                                                                */
int vrag_comm_transport(vrag_corpus_t* c, int* peer_memory);
int vrag_comm_destroy(vrag_corpus_t* c);

/* The building blocks of a collective stage, for hosts that drive the stages themselves (device pointers, caller's stream):
 * local scan + local top-k as packed entries -> all-gather -> merge.
 * vrag_stage_hits_dev: score store `name` (every page, or cand_ids_dev) with the device query and write this rank's top-k
 *   as k packed entries. vrag_allgather_topk: gathered_dev[r][list][k] <- rank r's local_dev[list][k], one collective for
 *   n_lists lists (batched queries). vrag_merge_hits_dev: merge n_src gathered lists of k_src entries per list into the
 *   global top-k (scores/ids [n_lists][k]); *flag_dev is OR-ed with 1 if any rank flagged a missed estimate.          */
int vrag_stage_hits_dev(vrag_corpus_t* c, const char* name, const float* query_dev, int n_query_rows, uint32_t flags,
                        const int64_t* cand_ids_dev, int64_t n_cand, int k, vrag_hit_t* out_hits_dev, void* stream);
int vrag_allgather_topk(vrag_corpus_t* c, const vrag_hit_t* local_dev, int n_lists, int k, vrag_hit_t* gathered_dev,
                        void* stream);
int vrag_merge_hits_dev(vrag_corpus_t* c, const vrag_hit_t* gathered_dev, int n_src, int n_lists, int k_src, int k,
                        float* out_scores_dev, int64_t* out_ids_dev, int* flag_dev, void* stream);
/* In-place max-all-reduce of a device fp32 vector (the candidate-stage exchange).                                   */
int vrag_allreduce_max_dev(vrag_corpus_t* c, float* scores_dev, int64_t n, void* stream);

/* Device-resident multi-stage search: query and outputs are device pointers, everything is enqueued on `stream` and the
 * call returns without synchronising (collective when the handle has a communicator). Stage s writes ks[s] entries at
 * [sum(ks[:s]), ...) of out_scores_dev / out_ids_dev; unused slots are (-inf, -1). Exact: no sampled top-k.          */
int vrag_search_multistage_dev(vrag_corpus_t* c, int n_stages, const char* const* names, const uint32_t* flags,
                               const int* ks, const float* query_dev, int n_query_rows, const int* q_offsets,
                               float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* Device time (microseconds, CUDA events) of each collective of the most recent host-facing search on this handle, in
 * issue order; *n receives how many there were (<= capacity written).                                               */
int vrag_last_comm_timing(vrag_corpus_t* c, float* out_us, int capacity, int* n);
/* Start of each of those collectives, in microseconds after the start of the search (same events): together with the
 * durations this is the device timeline of a collective search — local stage work | exchange | local stage work | ...  */
int vrag_last_comm_offsets(vrag_corpus_t* c, float* out_begin_us, int capacity, int* n);

/* ------------------------------------------------------------------ measurement helpers */
/* Device-side time (ms, CUDA events on the library stream) of the most recent vrag_search /
 * vrag_search_multistage on this corpus: [0] whole call, [1] dominant scan kernel only.            */
int vrag_last_timing(vrag_corpus_t* c, float* out_ms, int n);
/* Number of kernels the library has launched on this corpus since creation.                        */
int64_t vrag_launch_count(vrag_corpus_t* c);

#ifdef __cplusplus
}
#endif
#endif /* VRAG_B200_H */
