/* vettore_b200.h — C ABI of the B200-native scan path of Vettore.
 *
 * This is the drop-in boundary: every entry point below is what an erl_nif (or any
 * FFI) binding for the reference's scan path would call instead of the Rust code in
 * native/vettore/src/{flat,search,multi_vector,distances}.rs. Plain pointers and
 * sizes only. Citations are relative to /root/reference/.
 *
 * Conventions (nifs.rs: `Result<T, String>` => {:ok, T} | {:error, msg}):
 *   - every call returns VB_OK (0) or a non-zero status; on failure vb_last_error()
 *     (thread-local) holds the message. VB_ERR messages are the reference's own strings,
 *     byte for byte ("dimension mismatch", "vector must not be empty",
 *     "vector contains a non-finite value", "metric overflow", "score overflow",
 *     "unknown metric", "invalid prefix dimensions", "dimensions must be positive",
 *     "vectors must not be empty"). VB_ERR_CUDA messages start with "cuda: ".
 *   - the caller owns every input buffer; nothing is borrowed past the call.
 *   - results are returned as vb_hits objects owned by the caller (vb_hits_free).
 *   - ids are arbitrary byte strings (UTF-8 binaries on the BEAM), passed as one blob
 *     plus n+1 offsets; ragged float / u64 lists are passed as values plus n+1 offsets
 *     (in elements), so validation errors the reference raises for ragged input
 *     ("dimension mismatch") are reproducible.
 *   - all handles are safe for concurrent use: searches share, mutations exclude
 *     (nifs.rs:266-309 RwLock discipline).
 *   - the library has no CPU fallback: without a CUDA device every compute call fails
 *     with VB_ERR_CUDA.
 */
#ifndef VETTORE_B200_H
#define VETTORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_OK 0
#define VB_ERR 1      /* reference-visible error string */
#define VB_ERR_CUDA 2 /* device/runtime failure, "cuda: ..." */

/* Metric codes, distances.rs:25-38 / collection.ex:1306-1315. */
#define VB_METRIC_L2 0
#define VB_METRIC_L2_SQUARED 1
#define VB_METRIC_COSINE 2
#define VB_METRIC_INNER_PRODUCT 3
#define VB_METRIC_NEGATIVE_INNER_PRODUCT 4
#define VB_METRIC_MANHATTAN 5
#define VB_METRIC_CHEBYSHEV 6
#define VB_METRIC_HAMMING 7
#define VB_METRIC_JACCARD 8

typedef struct vb_flat vb_flat; /* replaces FlatResource, flat.rs:131-134 */
typedef struct vb_hits vb_hits; /* a sorted Vec<(String, f32)> */

const char* vb_last_error(void);
const char* vb_version(void);
/* Number of visible CUDA devices (0 when none; never fails). */
int vb_device_count(void);

/* ---- hits: Vec<(String, f32)> as returned by flat_search / *_top_k ------------- */
size_t vb_hits_len(const vb_hits* h);
/* Id bytes of hit i (not NUL-terminated); *len receives the byte length. */
const char* vb_hits_id(const vb_hits* h, size_t i, size_t* len);
/* Raw metric value (flat/vector/binary) or MaxSim score (multi-vector) of hit i. */
float vb_hits_value(const vb_hits* h, size_t i);
/* Position of hit i in the caller's input batch (by-value calls) or its device row. */
uint64_t vb_hits_index(const vb_hits* h, size_t i);
/* Whole result in one call: id blob + len+1 offsets, values, indexes (valid until
 * vb_hits_free). Returns the number of hits. */
size_t vb_hits_export(const vb_hits* h, const char** id_blob, const uint64_t** id_off,
                      const float** values, const uint64_t** index);
void vb_hits_free(vb_hits* h);

/* ---- resident flat index: Nifs.flat_* ------------------------------------------ */
/* flat_new_<metric>/0, nifs.rs:200-257. The index lives on the current CUDA device. */
int vb_flat_new(int metric_code, vb_flat** out);
/* Additive: the same index spread over several GPUs INSIDE ONE PROCESS (the reference surface is one BEAM
 * process holding one FlatResource, nifs.rs:297-309 — an erl_nif caller cannot fork a process per GPU).
 * Shard s lives on CUDA device devices[s] (devices == NULL: s % device_count, so several shards may share
 * a device); an id is owned by shard fnv1a(id) % n_shards. Each shard has its own host thread and stream;
 * vb_flat_search / _search_batch run the fused scan + top-k on every GPU concurrently and merge the
 * n_shards sorted lists on the calling thread by (rank.total_cmp, id bytes) — flat.rs:34-40 exactly.
 * The returned handle works with vb_flat_insert / _insert_many / _reserve / _delete / _search /
 * _search_batch / _info / _free AND the resident pipelines vb_flat_prefix_top_k / _funnel_search /
 * _quantized_search (every stage runs on all shards at once, the sorted lists are merged on the calling
 * thread, survivors are re-scored on the shard that owns them); the stream-ordered device-level entries
 * below (one process per GPU) answer VB_ERR_CUDA "not available on a sharded (multi-GPU) index handle".
 * vb_hits_index = shard << 32 | row. */
int vb_flat_new_sharded(int metric_code, int n_shards, const int* devices, vb_flat** out);
/* Resource destructor (BEAM GC of the last reference). Frees the HBM matrix. */
void vb_flat_free(vb_flat* index);
/* flat_insert/3, nifs.rs:259-271 -> FlatIndex::insert, flat.rs:59-66. */
int vb_flat_insert(vb_flat* index, const char* id, size_t id_len, const float* vector, size_t len);
/* flat_insert_many/2, nifs.rs:273-284 -> FlatIndex::insert_many, flat.rs:69-85.
 * All-or-nothing validation; duplicate ids in a batch: last wins. */
int vb_flat_insert_many(vb_flat* index, size_t n, const char* ids, const uint64_t* id_off,
                        const float* values, const uint64_t* value_off);
/* Additive: capacity hint for a bulk load (rebuild_index, collection.ex:426-433) — the HBM matrix is
 * allocated once for `rows` rows instead of growing by doubling. Does not change the contents. */
int vb_flat_reserve(vb_flat* index, size_t rows);
/* Additive (the ingest step before the scan: Index.put_many /
 * rebuild_index, index/flat.ex:35-39, collection.ex:426-433): same insert_many, rows already in
 * device memory as one contiguous [n, dimension] fp32 matrix. Same validation (finite check
 * runs on the device), all-or-nothing, duplicate ids: last wins. */
int vb_flat_insert_many_device(vb_flat* index, size_t n, const char* ids, const uint64_t* id_off,
                               const float* d_values, size_t dimension);
/* flat_delete/2, nifs.rs:286-295 -> FlatIndex::delete, flat.rs:88-93. */
int vb_flat_delete(vb_flat* index, const char* id, size_t id_len);
/* flat_search/3, nifs.rs:297-309 -> FlatIndex::search, flat.rs:96-124.
 * Hits ascend by (rank.total_cmp, id bytes); values are the raw metric. */
int vb_flat_search(vb_flat* index, const float* query, size_t len, size_t limit, vb_hits** out);
/* Additive: nq searches in one call (same semantics per query); out[nq]. */
int vb_flat_search_batch(vb_flat* index, const float* queries, size_t nq, size_t len, size_t limit,
                         vb_hits** out);
/* Additive: row count and dimension (dimension 0 == None, flat.rs:16). */
int vb_flat_info(vb_flat* index, size_t* rows, size_t* dimension);
/* Additive (funnel_search stage over the resident matrix instead of store.all + by-value
 * vector_top_k, collection.ex:674-691): vector_top_k semantics (search.rs:38-73: prefix
 * `dimensions`, true f64 cosine when metric_code is cosine) over the resident rows whose
 * ids are listed (n_ids == SIZE_MAX: every row). Unknown ids are skipped. */
int vb_flat_prefix_top_k(vb_flat* index, size_t n_ids, const char* ids, const uint64_t* id_off,
                         const float* query, size_t len, int metric_code, size_t dimensions,
                         size_t limit, vb_hits** out);

/* Additive: funnel_search (collection.ex:244-260, 674-691) run entirely on the resident
 * matrix. Stage s keeps the best `candidates` rows of the previous stage's survivors under
 * vector_top_k semantics at prefix stages[s]; the last step is the exact rerank at the full
 * query length with `limit`. One host synchronisation in total. Any `candidates` / `limit`
 * (collection.ex:509-510 defaults candidates to 10 x limit): up to 1024 survivors per stage stay in
 * the fused collector, beyond that the stage scores every row and radix-sorts on the device. */
int vb_flat_funnel_search(vb_flat* index, const float* query, size_t len, int metric_code,
                          const size_t* stages, size_t n_stages, size_t candidates, size_t limit,
                          vb_hits** out);
/* Additive: quantized_search (collection.ex:266-295) on the resident index: the sign-bit
 * codes of the rows (packed on the device, distances.rs:413-423) are scanned by Hamming
 * distance for the best `candidates` (distance, then id; bit-exact with binary_top_k,
 * search.rs:76-92), which are then reranked exactly (vector_top_k at full length).
 * candidates <= 1024. */
int vb_flat_quantized_search(vb_flat* index, const float* query, size_t len, int metric_code,
                             size_t candidates, size_t limit, vb_hits** out);

/* Device-level entry (inputs already in HBM; used for kernel-only timing and by the
 * row-sharded multi-GPU path). Queries: nq rows of `q_stride` floats (q_stride % 4 == 0,
 * zero padded) in device memory. Writes, per query, k = min(limit, rows) sorted entries:
 * keys (rank-order key << 32 | id rank), raw values and device rows. `stream` is a cudaStream_t.
 * Single queries and small batches are stream-ordered with no host synchronisation; an unrecoverable
 * overflow ("metric overflow", flat.rs:105) cannot be returned from there and raises a sticky bit read by
 * vb_flat_device_status. Batches that take the tensor-core path (>= 16 queries, dot-product metrics)
 * synchronise `stream` once: queries whose candidate set the exact re-scoring stage could not prove
 * complete are redone by the single-query kernel before the call returns, and "metric overflow" is
 * returned as VB_ERR. Any limit (beyond 1024: dump + radix sort per query). */
int vb_flat_search_device(vb_flat* index, const float* d_queries, size_t nq, size_t q_stride,
                          size_t limit, uint64_t* d_keys, float* d_values, uint32_t* d_rows,
                          uint32_t* d_counts, void* stream);
/* Row-sharded quantized_search (collection.ex:699-713 over a corpus split by rows), stage 1:
 * sign-packs the device queries (distances.rs:413-423) and scans THIS shard's code mirror for
 * its best `candidates` by (Hamming distance, id rank) — binary_top_k, search.rs:76-92. Output
 * convention of vb_flat_search_device. The shards' lists are all-gathered and merged with
 * vb_topk_merge_device into the global candidate set (which takes lists of up to 1024 entries). */
int vb_flat_hamming_device(vb_flat* index, const float* d_queries, size_t nq, size_t q_stride,
                           size_t candidates, uint64_t* d_keys, float* d_values, uint32_t* d_rows,
                           uint32_t* d_counts, void* stream);
/* Stage 2: exact rerank (vector_top_k at full length, search.rs:38-73; cosine = the f64 true
 * cosine, distances.rs:160-177) of those global candidates that live on `shard`.
 * d_global_rows: the merged candidate rows (`shard << 32 | row`, vb_topk_merge_device's rows
 * output), *d_global_count of them, at most max_candidates. Writes this shard's best
 * min(limit, owned) (possibly 0) in the vb_flat_search_device convention. One query.
 * Synchronises `stream` once (the owned-candidate count sizes the scan). */
int vb_flat_rerank_owned_device(vb_flat* index, const float* d_query, size_t q_stride, int metric_code,
                                const uint64_t* d_global_rows, const uint32_t* d_global_count,
                                size_t max_candidates, uint32_t shard, size_t limit, uint64_t* d_keys,
                                float* d_values, uint32_t* d_rows, uint32_t* d_counts, void* stream);
/* Sticky status of the device-level entries above since the last call: bit 0 = a scan met an
 * unrecoverable overflow (the reference's "metric overflow"). Synchronises the device, clears the word. */
int vb_flat_device_status(vb_flat* index, uint32_t* status);
/* Overrides the id tie-break ranks of the resident rows (row-sharded corpora: ranks must
 * be comparable across shards). ranks[row] for row < rows, in insertion (device row) order. */
int vb_flat_set_id_ranks(vb_flat* index, const uint32_t* ranks, size_t n);
/* K7: merges `lists` sorted top-k lists per query into the best k_out per query (the
 * final select after the shards' lists were all-gathered over NVLink). All pointers are
 * device memory. List l holds arrays [nq][k_in] of keys / values / rows and [nq] counts
 * starting `l * list_stride_bytes` bytes after the given base pointers (an all-gather of
 * one packed record per shard has exactly this shape). Outputs are [nq][k_out];
 * d_rows_out receives (list << 32 | row). */
int vb_topk_merge_device(const uint64_t* d_keys, const float* d_values, const uint32_t* d_rows,
                         const uint32_t* d_counts, size_t list_stride_bytes, size_t nq, size_t lists,
                         size_t k_in, size_t k_out, uint64_t* d_keys_out, float* d_values_out,
                         uint64_t* d_rows_out, uint32_t* d_counts_out, void* stream);

/* ---- exchange step of the row-sharded searches over NVLink peer memory ----------
 * Replaces "all-gather the per-GPU top-k lists, then select" by ONE kernel per rank: the rank's packed
 * record (keys[nq][k_in] u64 | values f32 | rows u32 | counts[nq] u32 at the given byte offsets,
 * `record_bytes` in all, a multiple of 16) is stored straight into every peer's gather buffer through
 * NVLink, a flag is published, and the same launch waits for the peers' flags and runs the K7 select.
 * Every rank makes the same sequence of calls. One process per GPU: vb_peer_new -> exchange the 64-byte
 * CUDA IPC handles over the process group -> vb_peer_connect_ipc. One process driving several GPUs (or
 * several shards of one GPU): vb_peer_connect_local with all the objects. */
typedef struct vb_peer vb_peer;
int vb_peer_new(int world, int rank, size_t record_bytes, vb_peer** out, unsigned char ipc_handle_out[64]);
void vb_peer_free(vb_peer* px);
int vb_peer_connect_ipc(vb_peer* px, const unsigned char* handles /* [world][64] */);
int vb_peer_connect_local(vb_peer* const* peers, int world);
/* Outputs as vb_topk_merge_device ([nq][k_out]; d_rows_out = shard << 32 | row). `stream`: cudaStream_t. */
int vb_peer_exchange_merge(vb_peer* px, const void* d_record, size_t nq, size_t k_in, size_t k_out, size_t off_keys,
                           size_t off_values, size_t off_rows, size_t off_counts, uint64_t* d_keys_out,
                           float* d_values_out, uint64_t* d_rows_out, uint32_t* d_counts_out, void* stream);
/* The two halves of the above (store + flag; wait + select), for a caller that drives several ranks from
 * one host thread: every rank's push must be enqueued before any rank's wait_merge. */
int vb_peer_push(vb_peer* px, const void* d_record, size_t nq, size_t k_in, size_t k_out, size_t off_keys,
                 size_t off_values, size_t off_rows, size_t off_counts, void* stream);
int vb_peer_wait_merge(vb_peer* px, size_t nq, size_t k_in, size_t k_out, size_t off_keys, size_t off_values,
                       size_t off_rows, size_t off_counts, uint64_t* d_keys_out, float* d_values_out,
                       uint64_t* d_rows_out, uint32_t* d_counts_out, void* stream);
/* *error = 1 when a wait timed out (a peer never published its record). Synchronises the device. */
int vb_peer_error(vb_peer* px, uint32_t* error);

/* ---- by-value batched helpers: Nifs.vector_top_k / binary_top_k / multi_vector_* -- */
/* vector_top_k/5, nifs.rs:151-162 -> search::vector_top_k, search.rs:38-73. */
int vb_vector_top_k(size_t n, const char* ids, const uint64_t* id_off, const float* values,
                    const uint64_t* value_off, const float* query, size_t len, int metric_code,
                    size_t dimensions, size_t limit, vb_hits** out);
/* binary_top_k/4, nifs.rs:164-175 -> search::binary_top_k, search.rs:76-92. */
int vb_binary_top_k(size_t n, const char* ids, const uint64_t* id_off, const uint64_t* words,
                    const uint64_t* word_off, const uint64_t* query, size_t query_words,
                    size_t dimensions, size_t limit, vb_hits** out);
/* multi_vector_top_k/4, nifs.rs:188-198 -> multi_vector::top_k, multi_vector.rs:90-132.
 * Tokens are one ragged list (tok_vals, tok_off[ntok+1]); document i owns tokens
 * [doc_tok[i], doc_tok[i+1]). Hits descend by score (total order), ids ascending on ties. */
int vb_multi_vector_top_k(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                          const uint64_t* tok_off, const uint64_t* doc_tok, const float* q_vals,
                          const uint64_t* q_off, size_t tq, int metric_code, size_t limit, vb_hits** out);
/* multi_vector_score/3, nifs.rs:177-186 -> multi_vector::score, multi_vector.rs:40-63. */
int vb_multi_vector_score(const float* q_vals, const uint64_t* q_off, size_t tq, const float* d_vals,
                          const uint64_t* d_off, size_t td, int metric_code, float* out);

/* ---- additive: HBM-resident multi-vector collection (multi_vector_search without
 * store.all + by-value marshalling, collection.ex:313-323). Same scoring and ordering as
 * vb_multi_vector_top_k; documents are upserted by id like the flat index. ---------------- */
typedef struct vb_mv vb_mv;
int vb_mv_new(int metric_code, vb_mv** out);
/* Additive: the same collection spread over several GPUs inside ONE process (see vb_flat_new_sharded): a document
 * lives on shard fnv1a(id) % n_shards, vb_mv_search scores every shard's documents on its own GPU concurrently and
 * merges the sorted lists by (score descending, id bytes), multi_vector.rs:22-31. Works with vb_mv_insert_many /
 * _delete / _search / _info / _free; the device-ingest and stream-ordered entries answer VB_ERR_CUDA. */
int vb_mv_new_sharded(int metric_code, int n_shards, const int* devices, vb_mv** out);
void vb_mv_free(vb_mv* index);
int vb_mv_insert_many(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off,
                      const float* tok_vals, const uint64_t* tok_off, const uint64_t* doc_tok);
/* Pre-sizes the HBM arrays so a large corpus can be streamed in without realloc + copy. */
int vb_mv_reserve(vb_mv* index, size_t docs, size_t tokens, size_t dimension);
/* Uniform documents (tokens_per_doc each) whose tokens already sit in device memory as a
 * row-major [ndocs * tokens_per_doc, dimension] fp32 matrix (e.g. produced by an encoder on
 * the same GPU). Validation (finite values) and the per-token norms run on the device. */
int vb_mv_insert_many_device(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off,
                             const float* d_tokens, size_t tokens_per_doc, size_t dimension);
/* Same for ragged documents: document i owns rows [doc_tok[i], doc_tok[i + 1]) of the device matrix
 * (doc_tok: HOST array of ndocs + 1 ascending row offsets; a document may be empty). */
int vb_mv_insert_ragged_device(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off,
                               const float* d_tokens, const uint64_t* doc_tok, size_t dimension);
int vb_mv_delete(vb_mv* index, const char* id, size_t id_len);
int vb_mv_search(vb_mv* index, const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit,
                 vb_hits** out);
/* Document-sharded MaxSim (multi_vector.rs:90-132 over a corpus split by documents): same
 * search, and the shard's sorted top-k is ALSO left in device memory in the
 * vb_flat_search_device convention (keys = score-order key << 32 | id rank, scores, document
 * slots, count) so the shards' lists can be all-gathered and merged with
 * vb_topk_merge_device. An empty shard writes count 0. limit >= 1, tq >= 1. */
int vb_mv_search_packed_device(vb_mv* index, const float* q_vals, const uint64_t* q_off, size_t tq,
                               size_t limit, uint64_t* d_keys, float* d_values, uint32_t* d_rows,
                               uint32_t* d_counts, vb_hits** out);
/* Overrides the id tie-break ranks, one per document slot in insertion order (sharded corpora:
 * ranks must compare across shards). Valid until the next insert/delete relabels. */
int vb_mv_set_id_ranks(vb_mv* index, const uint32_t* ranks, size_t n);
int vb_mv_info(vb_mv* index, size_t* docs, size_t* tokens, size_t* dimension);

/* Additive (the step after the scan): Distance.result_values/3
 * (lib/vettore_distance.ex:100-102, 525-543) for a whole hit list in one call — raw f32 metric values ->
 * the (score, distance) pairs of Vettore.Result, in f64 exactly as the BEAM computes them (`raw / 1`,
 * `1.0 - raw`, `1.0 / (1.0 + raw)`, `(raw + 1.0) / 2.0` on the widened raw). score_mode 0 = :raw,
 * 1 = :similarity. Host arithmetic (k values): callers are left with the ETS lookups only
 * (index/flat.ex:72-91). "unknown metric" for codes outside 0..8. */
int vb_result_values(int metric_code, int score_mode, const float* raw, size_t n, double* score, double* distance);

/* Additive: muvera_encode_query/7 and muvera_encode_document/7 (nifs.rs:430-476 ->
 * muvera::encode, muvera.rs:26-74), batched over documents on the device. Document i owns the vectors
 * [doc_vec[i], doc_vec[i+1]) of the ragged list (vals, vec_off[nvec + 1] in elements); ndocs = 1 is exactly one
 * reference NIF call. has_final / final_projection_dimension spell Option<usize>; mode 0 = query (sum per SimHash
 * partition), 1 = document (running average, rounded to f32 after every vector like the reference). out: host
 * buffer of out_capacity floats receiving [ndocs][*fde_dimension]; out == NULL validates the input the way
 * the reference does and only reports *fde_dimension (so a binding can size its buffer). The device keeps the reference's
 * accumulation orders, so the encoding is bit-identical to the reference arithmetic. Errors are the
 * reference's strings: "empty vectors", "dimension must be positive", "num_repetitions must be positive",
 * "num_simhash_projections must be < 31", "projection_dimension must be positive",
 * "final_projection_dimension must be positive", "dimension mismatch", "vector contains a non-finite value",
 * "fde dimension overflow", "fde dimension exceeds safety limit", "encoding overflow". */
int vb_muvera_encode(size_t ndocs, const float* vals, const uint64_t* vec_off, const uint64_t* doc_vec, size_t dimension,
                     size_t num_repetitions, size_t num_simhash_projections, uint64_t seed, size_t projection_dimension,
                     int has_final, size_t final_projection_dimension, int mode, float* out, size_t out_capacity,
                     size_t* fde_dimension);

/* compress_sign_bits/1, nifs.rs:125-129 -> distances.rs:413-423. words[ceil(len/64)]. */
int vb_compress_sign_bits(const float* vector, size_t len, uint64_t* words);

#ifdef __cplusplus
}
#endif
#endif /* VETTORE_B200_H */
