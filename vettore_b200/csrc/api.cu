// api.cu — the extern "C" boundary declared in include/vettore_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "../../include/vettore_b200.h"
#include "flat_index.h"
#include "hamming.h"
#include "maxsim.h"
#include "muvera.h"
#include "peer_exchange.h"
#include "runtime.h"
#include "scan_driver.h"
#include "select.h"
#include "sharded_index.h"

namespace {

thread_local std::string g_last_error;

int finish(const vb::Status& s) {
    if (s.ok()) return VB_OK;
    g_last_error = s.msg;
    return s.code;
}

bool all_finite(const float* v, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) return false;
    return true;
}

// Ranks of a by-value id batch: position in byte-lexicographic order (stable, so duplicate
// ids still get distinct keys).
std::vector<uint32_t> id_ranks(size_t n, const char* ids, const uint64_t* id_off) {
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        size_t la = id_off[a + 1] - id_off[a], lb = id_off[b + 1] - id_off[b];
        int c = std::memcmp(ids + id_off[a], ids + id_off[b], std::min(la, lb));
        return c != 0 ? c < 0 : la < lb;
    });
    std::vector<uint32_t> rank(n);
    for (size_t i = 0; i < n; ++i) rank[order[i]] = (uint32_t)i;
    return rank;
}

void emit_hits(vb::Hits* h, const char* ids, const uint64_t* id_off, const uint32_t* rows, const float* vals,
               size_t cnt) {
    for (size_t i = 0; i < cnt; ++i) {
        uint32_t r = rows[i];
        h->add(ids + id_off[r], id_off[r + 1] - id_off[r], vals[i], r);
    }
}

int no_device() {
    g_last_error = "cuda: no CUDA device available (vettore_b200 has no CPU fallback)";
    return VB_ERR_CUDA;
}

}  // namespace

extern "C" {

const char* vb_last_error(void) { return g_last_error.c_str(); }
const char* vb_version(void) { return "vettore_b200 0.1.0 (sm_100a)"; }

int vb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

size_t vb_hits_len(const vb_hits* h) { return h ? h->size() : 0; }
const char* vb_hits_id(const vb_hits* h, size_t i, size_t* len) {
    *len = h->off[i + 1] - h->off[i];
    return h->blob.data() + h->off[i];
}
size_t vb_hits_export(const vb_hits* h, const char** id_blob, const uint64_t** id_off, const float** values,
                      const uint64_t** index) {
    if (!h) return 0;
    *id_blob = h->blob.data();
    *id_off = h->off.data();
    *values = h->values.data();
    *index = h->index.data();
    return h->size();
}
float vb_hits_value(const vb_hits* h, size_t i) { return h->values[i]; }
uint64_t vb_hits_index(const vb_hits* h, size_t i) { return h->index[i]; }
void vb_hits_free(vb_hits* h) { delete h; }

// ---------------------------------------------------------------- resident flat index
int vb_flat_new(int metric_code, vb_flat** out) {
    *out = nullptr;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    if (vb_device_count() <= 0) return no_device();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return no_device();
    *out = new vb_flat{new vb::FlatIndex(metric_code, dev), nullptr};
    return VB_OK;
}

int vb_flat_new_sharded(int metric_code, int n_shards, const int* devices, vb_flat** out) {
    *out = nullptr;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    const int ndev = vb_device_count();
    if (ndev <= 0) return no_device();
    if (n_shards < 1 || n_shards > 64) return finish(vb::Status::Cuda("sharded index: shard count must be in 1..64"));
    std::vector<int> devs((size_t)n_shards);
    for (int s = 0; s < n_shards; ++s) {
        devs[s] = devices ? devices[s] : s % ndev;
        if (devs[s] < 0 || devs[s] >= ndev) return finish(vb::Status::Cuda("sharded index: no such CUDA device"));
    }
    *out = new vb_flat{nullptr, new vb::ShardedFlatIndex(metric_code, devs)};
    return VB_OK;
}

void vb_flat_free(vb_flat* index) {
    if (!index) return;
    delete index->impl;
    delete index->sharded;
    delete index;
}

#define VB_SINGLE_GPU_ONLY(index)                                                                          \
    if (!(index)->impl) return finish(vb::Status::Cuda("not available on a sharded (multi-GPU) index handle"))

int vb_flat_insert(vb_flat* index, const char* id, size_t id_len, const float* vector, size_t len) {
    const uint64_t id_off[2] = {0, id_len}, val_off[2] = {0, len};
    if (index->sharded) return finish(index->sharded->insert_many(1, id, id_off, vector, val_off));
    return finish(index->impl->insert_many(1, id, id_off, vector, val_off, true));
}

int vb_flat_insert_many(vb_flat* index, size_t n, const char* ids, const uint64_t* id_off, const float* values,
                        const uint64_t* value_off) {
    if (index->sharded) return finish(index->sharded->insert_many(n, ids, id_off, values, value_off));
    return finish(index->impl->insert_many(n, ids, id_off, values, value_off, false));
}

int vb_flat_reserve(vb_flat* index, size_t rows) {
    return finish(index->sharded ? index->sharded->reserve(rows) : index->impl->reserve(rows));
}
int vb_flat_insert_many_device(vb_flat* index, size_t n, const char* ids, const uint64_t* id_off,
                               const float* d_values, size_t dimension) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->insert_many_device(n, ids, id_off, d_values, dimension));
}
int vb_flat_delete(vb_flat* index, const char* id, size_t id_len) {
    return finish(index->sharded ? index->sharded->remove(id, id_len) : index->impl->remove(id, id_len));
}

int vb_flat_search(vb_flat* index, const float* query, size_t len, size_t limit, vb_hits** out) {
    *out = nullptr;
    std::vector<vb::Hits> hits;
    vb::Status s = index->sharded ? index->sharded->search(query, 1, len, limit, &hits)
                                  : index->impl->search(query, 1, len, limit, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits[0])};
    return VB_OK;
}

int vb_flat_search_batch(vb_flat* index, const float* queries, size_t nq, size_t len, size_t limit,
                         vb_hits** out) {
    for (size_t q = 0; q < nq; ++q) out[q] = nullptr;
    std::vector<vb::Hits> hits;
    vb::Status s = index->sharded ? index->sharded->search(queries, nq, len, limit, &hits)
                                  : index->impl->search(queries, nq, len, limit, &hits);
    if (!s.ok()) return finish(s);
    for (size_t q = 0; q < nq; ++q) out[q] = new vb_hits{std::move(hits[q])};
    return VB_OK;
}

int vb_flat_info(vb_flat* index, size_t* rows, size_t* dimension) {
    if (index->sharded) index->sharded->info(rows, dimension);
    else index->impl->info(rows, dimension);
    return VB_OK;
}

int vb_flat_prefix_top_k(vb_flat* index, size_t n_ids, const char* ids, const uint64_t* id_off,
                         const float* query, size_t len, int metric_code, size_t dimensions, size_t limit,
                         vb_hits** out) {
    *out = nullptr;
    vb::Hits hits;
    const bool all = n_ids == SIZE_MAX;
    vb::Status s = index->sharded ? index->sharded->prefix_top_k(all, all ? 0 : n_ids, ids, id_off, query, len, metric_code,
                                                                 dimensions, limit, &hits)
                                  : index->impl->prefix_top_k(all, all ? 0 : n_ids, ids, id_off, query, len, metric_code,
                                                              dimensions, limit, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

int vb_flat_funnel_search(vb_flat* index, const float* query, size_t len, int metric_code, const size_t* stages,
                          size_t n_stages, size_t candidates, size_t limit, vb_hits** out) {
    *out = nullptr;
    vb::Hits hits;
    vb::Status s = index->sharded ? index->sharded->funnel_search(query, len, metric_code, stages, n_stages, candidates, limit, &hits)
                                  : index->impl->funnel_search(query, len, metric_code, stages, n_stages, candidates, limit, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

int vb_flat_quantized_search(vb_flat* index, const float* query, size_t len, int metric_code, size_t candidates,
                             size_t limit, vb_hits** out) {
    *out = nullptr;
    vb::Hits hits;
    vb::Status s = index->sharded ? index->sharded->quantized_search(query, len, metric_code, candidates, limit, &hits)
                                  : index->impl->quantized_search(query, len, metric_code, candidates, limit, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

int vb_flat_search_device(vb_flat* index, const float* d_queries, size_t nq, size_t q_stride, size_t limit,
                          uint64_t* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts, void* stream) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->search_device(d_queries, nq, q_stride, limit,
                                             reinterpret_cast<vb::u64*>(d_keys), d_values, d_rows, d_counts,
                                             static_cast<cudaStream_t>(stream)));
}

int vb_flat_hamming_device(vb_flat* index, const float* d_queries, size_t nq, size_t q_stride, size_t candidates,
                           uint64_t* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts, void* stream) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->hamming_device(d_queries, nq, q_stride, candidates, reinterpret_cast<vb::u64*>(d_keys),
                                              d_values, d_rows, d_counts, static_cast<cudaStream_t>(stream)));
}
int vb_flat_rerank_owned_device(vb_flat* index, const float* d_query, size_t q_stride, int metric_code,
                                const uint64_t* d_global_rows, const uint32_t* d_global_count, size_t max_candidates,
                                uint32_t shard, size_t limit, uint64_t* d_keys, float* d_values, uint32_t* d_rows,
                                uint32_t* d_counts, void* stream) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->rerank_owned_device(d_query, q_stride, metric_code,
                                                   reinterpret_cast<const vb::u64*>(d_global_rows), d_global_count,
                                                   max_candidates, shard, limit, reinterpret_cast<vb::u64*>(d_keys),
                                                   d_values, d_rows, d_counts, static_cast<cudaStream_t>(stream)));
}
int vb_flat_device_status(vb_flat* index, uint32_t* status) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->device_status(status));
}
int vb_flat_set_id_ranks(vb_flat* index, const uint32_t* ranks, size_t n) {
    VB_SINGLE_GPU_ONLY(index);
    return finish(index->impl->set_id_ranks(ranks, n));
}

int vb_topk_merge_device(const uint64_t* d_keys, const float* d_values, const uint32_t* d_rows,
                         const uint32_t* d_counts, size_t list_stride_bytes, size_t nq, size_t lists, size_t k_in,
                         size_t k_out, uint64_t* d_keys_out, float* d_values_out, uint64_t* d_rows_out, uint32_t* d_counts_out,
                         void* stream) {
    return finish(vb::topk_merge_device(reinterpret_cast<const vb::u64*>(d_keys), d_values, d_rows, d_counts,
                                        list_stride_bytes, nq, lists, k_in, k_out, reinterpret_cast<vb::u64*>(d_keys_out), d_values_out,
                                        reinterpret_cast<vb::u64*>(d_rows_out), d_counts_out,
                                        static_cast<cudaStream_t>(stream)));
}

// ---------------------------------------------------------------- NVLink peer exchange
namespace {
vb::PeerRecord peer_record(vb_peer* px, size_t nq, size_t k_in, size_t k_out, size_t off_keys, size_t off_values,
                           size_t off_rows, size_t off_counts) {
    vb::PeerRecord r;
    r.nq = (uint32_t)nq; r.k_in = (uint32_t)k_in; r.k_out = (uint32_t)k_out;
    r.off_keys = (uint32_t)off_keys; r.off_values = (uint32_t)off_values; r.off_rows = (uint32_t)off_rows;
    r.off_counts = (uint32_t)off_counts;
    r.bytes = px->impl->record_bytes();
    return r;
}
}  // namespace

int vb_peer_new(int world, int rank, size_t record_bytes, vb_peer** out, unsigned char ipc_handle_out[64]) {
    *out = nullptr;
    if (vb_device_count() <= 0) return no_device();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return no_device();
    auto* px = new vb::PeerExchange(world, rank, record_bytes, dev);
    vb::Status s = px->allocate();
    if (s.ok() && ipc_handle_out) s = px->export_handle(ipc_handle_out);
    if (!s.ok()) { delete px; return finish(s); }
    *out = new vb_peer{px};
    return VB_OK;
}
void vb_peer_free(vb_peer* px) {
    if (!px) return;
    delete px->impl;
    delete px;
}
int vb_peer_connect_ipc(vb_peer* px, const unsigned char* handles) { return finish(px->impl->connect_ipc(handles)); }
int vb_peer_connect_local(vb_peer* const* peers, int world) {
    if (world < 1 || world > vb::kMaxPeers) return finish(vb::Status::Cuda("peer exchange: bad world"));
    vb::PeerExchange* impls[vb::kMaxPeers];
    for (int r = 0; r < world; ++r) impls[r] = peers[r]->impl;
    for (int r = 0; r < world; ++r) {
        vb::Status s = impls[r]->connect_local(impls);
        if (!s.ok()) return finish(s);
    }
    return VB_OK;
}
int vb_peer_exchange_merge(vb_peer* px, const void* d_record, size_t nq, size_t k_in, size_t k_out, size_t off_keys,
                           size_t off_values, size_t off_rows, size_t off_counts, uint64_t* d_keys_out,
                           float* d_values_out, uint64_t* d_rows_out, uint32_t* d_counts_out, void* stream) {
    return finish(px->impl->exchange_merge(d_record, peer_record(px, nq, k_in, k_out, off_keys, off_values, off_rows, off_counts),
                                           reinterpret_cast<vb::u64*>(d_keys_out), d_values_out,
                                           reinterpret_cast<vb::u64*>(d_rows_out), d_counts_out,
                                           static_cast<cudaStream_t>(stream)));
}
int vb_peer_push(vb_peer* px, const void* d_record, size_t nq, size_t k_in, size_t k_out, size_t off_keys,
                 size_t off_values, size_t off_rows, size_t off_counts, void* stream) {
    return finish(px->impl->push(d_record, peer_record(px, nq, k_in, k_out, off_keys, off_values, off_rows, off_counts),
                                 static_cast<cudaStream_t>(stream)));
}
int vb_peer_wait_merge(vb_peer* px, size_t nq, size_t k_in, size_t k_out, size_t off_keys, size_t off_values,
                       size_t off_rows, size_t off_counts, uint64_t* d_keys_out, float* d_values_out,
                       uint64_t* d_rows_out, uint32_t* d_counts_out, void* stream) {
    return finish(px->impl->wait_merge(peer_record(px, nq, k_in, k_out, off_keys, off_values, off_rows, off_counts),
                                       reinterpret_cast<vb::u64*>(d_keys_out), d_values_out,
                                       reinterpret_cast<vb::u64*>(d_rows_out), d_counts_out,
                                       static_cast<cudaStream_t>(stream)));
}
int vb_peer_error(vb_peer* px, uint32_t* error) { return finish(px->impl->error_state(error)); }

// ---------------------------------------------------------------- by-value helpers
int vb_vector_top_k(size_t n, const char* ids, const uint64_t* id_off, const float* values,
                    const uint64_t* value_off, const float* query, size_t len, int metric_code, size_t dimensions,
                    size_t limit, vb_hits** out) {
    *out = nullptr;
    // search.rs:38-55 validation order.
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    if (dimensions == 0 || dimensions > len) return finish(vb::Status::Ref("invalid prefix dimensions"));
    if (!all_finite(query, dimensions)) return finish(vb::Status::Ref("vector contains a non-finite value"));
    size_t good = n;  // rows before the first one the reference would reject
    const char* bad_msg = nullptr;
    for (size_t r = 0; r < n; ++r) {
        const size_t rl = value_off[r + 1] - value_off[r];
        if (dimensions > rl) { good = r; bad_msg = "dimension mismatch"; break; }
        if (!all_finite(values + value_off[r], dimensions)) {
            good = r;
            bad_msg = "vector contains a non-finite value";
            break;
        }
    }
    vb::Hits hits;
    if (good > 0) {
        if (good >= (1ull << 32) - 1) return finish(vb::Status::Cuda("batch too large"));
        if (vb_device_count() <= 0) return no_device();
        vb::CtxLease ctx;
        vb::Status s = ctx.get();
        if (!s.ok()) return finish(s);
        // Stage only the scored prefix: [good, stride] fp32, zero padded to 16-byte rows.
        const size_t stride = (dimensions + 3) & ~(size_t)3;
        vb::PinnedBuf hb;
        s = hb.reserve(good * stride * sizeof(float) + good * sizeof(uint32_t));
        if (!s.ok()) return finish(s);
        float* hrows = hb.as<float>();
        uint32_t* hrank = reinterpret_cast<uint32_t*>(hrows + good * stride);
        for (size_t r = 0; r < good; ++r) {
            std::memcpy(hrows + r * stride, values + value_off[r], dimensions * sizeof(float));
            for (size_t c = dimensions; c < stride; ++c) hrows[r * stride + c] = 0.0f;
        }
        std::vector<uint32_t> rank = id_ranks(good, ids, id_off);
        std::memcpy(hrank, rank.data(), good * sizeof(uint32_t));
        s = ctx->staging.reserve(good * stride * sizeof(float));
        if (s.ok()) s = ctx->staging_rank.reserve(good * sizeof(uint32_t));
        if (s.ok()) {
            cudaError_t e = cudaMemcpyAsync(ctx->staging.p, hrows, good * stride * sizeof(float),
                                            cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(ctx->staging_rank.p, hrank, good * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                    ctx->stream);
            if (e != cudaSuccess) s = vb::Status::Cuda(cudaGetErrorString(e));
        }
        vb::ScanResult res;
        if (s.ok()) {
            vb::ScanJob job;
            job.metric = metric_code == vb::kCosine ? vb::kCosineTrue : metric_code;   // search.rs:56-60
            job.d_rows = ctx->staging.as<float>();
            job.row_stride = stride;
            job.d_id_rank = ctx->staging_rank.as<uint32_t>();
            job.n = (uint32_t)good;
            job.dims = (uint32_t)dimensions;
            job.whole_rows = true;  // staged rows hold the prefix only, zero padded
            job.h_queries = query;
            job.nq = 1;
            job.q_len = len;
            job.k = std::max<size_t>(1, std::min(limit, good));
            s = vb::run_scan(*ctx.ctx, job, &res);
        }
        hb.release();
        if (!s.ok()) return finish(s);
        if (res.err_rows[0] != vb::kNoError) return finish(vb::Status::Ref("metric overflow"));
        if (!bad_msg && limit > 0) emit_hits(&hits, ids, id_off, res.rows.data(), res.raws.data(), res.counts[0]);
    }
    if (bad_msg) return finish(vb::Status::Ref(bad_msg));
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

int vb_binary_top_k(size_t n, const char* ids, const uint64_t* id_off, const uint64_t* words,
                    const uint64_t* word_off, const uint64_t* query, size_t query_words, size_t dimensions,
                    size_t limit, vb_hits** out) {
    *out = nullptr;
    // search.rs:84 -> distances.rs:459-470: the query is validated even for an empty batch.
    const size_t nw = (dimensions + 63) / 64;
    if (dimensions == 0) return finish(vb::Status::Ref("dimensions must be positive"));
    if (query_words != nw) return finish(vb::Status::Ref("dimension mismatch"));
    size_t good = n;
    for (size_t r = 0; r < n; ++r)
        if (word_off[r + 1] - word_off[r] != nw) { good = r; break; }
    vb::Hits hits;
    if (good > 0 && limit > 0 && good == n) {
        if (good >= (1ull << 32) - 1) return finish(vb::Status::Cuda("batch too large"));
        if (vb_device_count() <= 0) return no_device();
        vb::CtxLease ctx;
        vb::Status s = ctx.get();
        if (!s.ok()) return finish(s);
        std::vector<uint32_t> rank = id_ranks(good, ids, id_off);
        std::vector<uint32_t> rows;
        std::vector<float> vals;
        // Codes of a well-formed batch are contiguous: words[word_off[0] ..].
        s = vb::hamming_top_k_host(*ctx.ctx, words + word_off[0], good, nw, dimensions, rank.data(), query,
                                   std::min(limit, good), &rows, &vals);
        if (!s.ok()) return finish(s);
        emit_hits(&hits, ids, id_off, rows.data(), vals.data(), rows.size());
    }
    if (good < n) return finish(vb::Status::Ref("dimension mismatch"));
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

// ---- multi-vector (MaxSim) ---------------------------------------------------------------
namespace {

// multi_vector.rs:144-152 over tokens [t0, t1) of a ragged token list
const char* validate_tokens(const float* vals, const uint64_t* off, size_t t0, size_t t1, size_t dim) {
    for (size_t t = t0; t < t1; ++t) {
        const size_t len = off[t + 1] - off[t];
        if (len != dim) return "dimension mismatch";
        if (!all_finite(vals + off[t], len)) return "vector contains a non-finite value";
    }
    return nullptr;
}
// multi_vector.rs:134-142
const char* validate_standalone(const float* vals, const uint64_t* off, size_t t0, size_t t1) {
    if (t0 == t1) return nullptr;
    const size_t first = off[t0 + 1] - off[t0];
    if (first == 0) return "vectors must not be empty";
    return validate_tokens(vals, off, t0, t1, first);
}

// Scores documents [0, ndocs) (already validated, every token `dim` long) on the device.
vb::Status maxsim_by_value(size_t ndocs, const float* tok_vals, const uint64_t* tok_off, const uint64_t* doc_tok,
                           const uint32_t* ranks, const float* q_vals, size_t tq, size_t dim, int metric_code,
                           size_t k, vb::MaxSimResult* res) {
    vb::CtxLease ctx;
    VB_TRY(ctx.get());
    const size_t stride = (dim + 3) & ~(size_t)3;
    const size_t ntok = doc_tok[ndocs] - doc_tok[0];
    if (ntok >= 0xFFFFFFFFull || ndocs >= 0xFFFFFFFEull) return vb::Status::Cuda("batch too large");
    // Side arrays next to the token rows: doc_off[ndocs + 1] | doc_rank[ndocs] | tok_doc[ntok] | 1/|token| [ntok]
    // (the last two feed the ragged tensor-core kernel, maxsim_tcr.cu).
    const size_t side_words = 2 * ndocs + 1 + 2 * ntok;
    vb::PinnedBuf hb;
    VB_TRY(hb.reserve(std::max<size_t>(ntok, 1) * stride * sizeof(float) + side_words * sizeof(uint32_t)));
    float* hrows = hb.as<float>();
    uint32_t* hoff = reinterpret_cast<uint32_t*>(hrows + std::max<size_t>(ntok, 1) * stride);
    uint32_t* hrank = hoff + ndocs + 1;
    uint32_t* howner = hrank + ndocs;
    float* hinv = reinterpret_cast<float*>(howner + ntok);
    for (size_t t = 0; t < ntok; ++t) {
        const float* src = tok_vals + tok_off[doc_tok[0] + t];
        std::memcpy(hrows + t * stride, src, dim * sizeof(float));
        for (size_t c = dim; c < stride; ++c) hrows[t * stride + c] = 0.0f;
        double nn = 0.0;   // distances.rs:166: f64_dot(right, right).sqrt()
        for (size_t c = 0; c < dim; ++c) nn += (double)src[c] * (double)src[c];
        hinv[t] = nn > 0.0 ? (float)(1.0 / std::sqrt(nn)) : 0.0f;
    }
    uint32_t min_td = 0;
    bool has_empty = false;
    for (size_t d = 0; d <= ndocs; ++d) hoff[d] = (uint32_t)(doc_tok[d] - doc_tok[0]);
    for (size_t d = 0; d < ndocs; ++d) {
        const uint32_t cnt = hoff[d + 1] - hoff[d];
        if (cnt == 0) has_empty = true;
        else if (min_td == 0 || cnt < min_td) min_td = cnt;
        for (uint32_t t = hoff[d]; t < hoff[d + 1]; ++t) howner[t] = (uint32_t)d;
    }
    std::memcpy(hrank, ranks, ndocs * sizeof(uint32_t));
    VB_TRY(ctx->staging.reserve(std::max<size_t>(ntok, 1) * stride * sizeof(float)));
    VB_TRY(ctx->staging_rank.reserve(side_words * sizeof(uint32_t)));
    VB_CUDA(cudaMemcpyAsync(ctx->staging.p, hrows, ntok * stride * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    VB_CUDA(cudaMemcpyAsync(ctx->staging_rank.p, hoff, side_words * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    vb::MaxSimJob job;
    job.metric = metric_code == vb::kCosine ? vb::kCosineTrue : metric_code;   // multi_vector.rs:74-75
    job.d_tokens = ctx->staging.as<float>();
    job.stride = stride;
    job.d_doc_off = ctx->staging_rank.as<uint32_t>();
    job.d_doc_rank = job.d_doc_off + ndocs + 1;
    job.d_tok_doc = job.d_doc_rank + ndocs;
    job.d_inv_dnorm = reinterpret_cast<const float*>(job.d_tok_doc + ntok);
    job.ntok = ntok;
    job.min_td = min_td;
    job.has_empty = has_empty;
    job.ndocs = ndocs;
    job.dims = (uint32_t)dim;
    job.h_query = q_vals;
    job.tq = (uint32_t)tq;
    job.k = k;
    vb::Status s = vb::maxsim_top_k(*ctx.ctx, job, res);
    hb.release();
    return s;
}

}  // namespace

int vb_multi_vector_top_k(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                          const uint64_t* tok_off, const uint64_t* doc_tok, const float* q_vals,
                          const uint64_t* q_off, size_t tq, int metric_code, size_t limit, vb_hits** out) {
    *out = nullptr;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));   // nifs.rs:196
    if (const char* e = validate_standalone(q_vals, q_off, 0, tq)) return finish(vb::Status::Ref(e));   // :96
    const size_t qdim = tq ? (size_t)(q_off[1] - q_off[0]) : 0;
    size_t good = ndocs;
    const char* bad_msg = nullptr;
    for (size_t d = 0; d < ndocs && !bad_msg; ++d) {                      // multi_vector.rs:100-111
        const size_t t0 = doc_tok[d], t1 = doc_tok[d + 1];
        const char* e = nullptr;
        if (tq == 0) e = validate_standalone(tok_vals, tok_off, t0, t1);
        else if (t0 != t1) e = validate_tokens(tok_vals, tok_off, t0, t1, qdim);
        if (e) { good = d; bad_msg = e; }
    }
    vb::Hits hits;
    if (good > 0) {
        std::vector<uint32_t> rank = id_ranks(good, ids, id_off);
        const size_t ntok = doc_tok[good] - doc_tok[0];
        if (tq == 0 || ntok == 0) {
            // every score is 0.0 (:102-106): the best `limit` are simply the smallest ids
            if (!bad_msg && limit > 0) {
                std::vector<uint32_t> by_rank(good);
                for (size_t d = 0; d < good; ++d) by_rank[rank[d]] = (uint32_t)d;
                for (size_t i = 0; i < std::min(limit, good); ++i) {
                    const uint32_t d = by_rank[i];
                    hits.add(ids + id_off[d], id_off[d + 1] - id_off[d], 0.0f, d);
                }
            }
        } else {
            if (vb_device_count() <= 0) return no_device();
            vb::MaxSimResult res;
            vb::Status s = maxsim_by_value(good, tok_vals, tok_off, doc_tok, rank.data(), q_vals + q_off[0], tq, qdim,
                                           metric_code, std::max<size_t>(1, std::min(limit, good)), &res);
            if (!s.ok()) return finish(s);
            if (res.err != vb::kNoError)
                return finish(vb::Status::Ref((res.err & 1u) ? "score overflow" : "metric overflow"));
            if (!bad_msg && limit > 0) emit_hits(&hits, ids, id_off, res.rows.data(), res.scores.data(), res.rows.size());
        }
    }
    if (bad_msg) return finish(vb::Status::Ref(bad_msg));
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}

int vb_multi_vector_score(const float* q_vals, const uint64_t* q_off, size_t tq, const float* d_vals,
                          const uint64_t* d_off, size_t td, int metric_code, float* out) {
    *out = 0.0f;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));   // nifs.rs:184
    if (tq == 0) {                                                       // multi_vector.rs:45-48
        if (const char* e = validate_standalone(d_vals, d_off, 0, td)) return finish(vb::Status::Ref(e));
        return VB_OK;
    }
    const size_t dim = q_off[1] - q_off[0];
    if (dim == 0) return finish(vb::Status::Ref("vectors must not be empty"));
    if (const char* e = validate_tokens(q_vals, q_off, 0, tq, dim)) return finish(vb::Status::Ref(e));
    if (td == 0) return VB_OK;
    if (const char* e = validate_tokens(d_vals, d_off, 0, td, dim)) return finish(vb::Status::Ref(e));
    if (vb_device_count() <= 0) return no_device();
    const uint64_t doc_tok[2] = {0, td};
    const uint32_t rank0 = 0;
    vb::MaxSimResult res;
    vb::Status s = maxsim_by_value(1, d_vals, d_off, doc_tok, &rank0, q_vals + q_off[0], tq, dim, metric_code, 1, &res);
    if (!s.ok()) return finish(s);
    if (res.err != vb::kNoError) return finish(vb::Status::Ref((res.err & 1u) ? "score overflow" : "metric overflow"));
    *out = res.scores.empty() ? 0.0f : res.scores[0];
    return VB_OK;
}

int vb_mv_new(int metric_code, vb_mv** out) {
    *out = nullptr;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    if (vb_device_count() <= 0) return no_device();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return no_device();
    *out = new vb_mv{new vb::MvIndex(metric_code, dev), nullptr};
    return VB_OK;
}
int vb_mv_new_sharded(int metric_code, int n_shards, const int* devices, vb_mv** out) {
    *out = nullptr;
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    const int ndev = vb_device_count();
    if (ndev <= 0) return no_device();
    if (n_shards < 1 || n_shards > 64) return finish(vb::Status::Cuda("sharded collection: shard count must be in 1..64"));
    std::vector<int> devs((size_t)n_shards);
    for (int s = 0; s < n_shards; ++s) {
        devs[s] = devices ? devices[s] : s % ndev;
        if (devs[s] < 0 || devs[s] >= ndev) return finish(vb::Status::Cuda("sharded collection: no such CUDA device"));
    }
    *out = new vb_mv{nullptr, new vb::ShardedMvIndex(metric_code, devs)};
    return VB_OK;
}
void vb_mv_free(vb_mv* index) {
    if (!index) return;
    delete index->impl;
    delete index->sharded;
    delete index;
}
#define VB_MV_SINGLE_GPU_ONLY(index)                                                                          \
    if (!(index)->impl) return finish(vb::Status::Cuda("not available on a sharded (multi-GPU) collection handle"))
int vb_mv_insert_many(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                      const uint64_t* tok_off, const uint64_t* doc_tok) {
    if (index->sharded) return finish(index->sharded->insert_many(ndocs, ids, id_off, tok_vals, tok_off, doc_tok));
    return finish(index->impl->insert_many(ndocs, ids, id_off, tok_vals, tok_off, doc_tok));
}
int vb_mv_reserve(vb_mv* index, size_t docs, size_t tokens, size_t dimension) {
    VB_MV_SINGLE_GPU_ONLY(index);
    return finish(index->impl->reserve(docs, tokens, dimension));
}
int vb_mv_insert_many_device(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off,
                             const float* d_tokens, size_t tokens_per_doc, size_t dimension) {
    VB_MV_SINGLE_GPU_ONLY(index);
    return finish(index->impl->insert_many_device(ndocs, ids, id_off, d_tokens, tokens_per_doc, dimension));
}
int vb_mv_insert_ragged_device(vb_mv* index, size_t ndocs, const char* ids, const uint64_t* id_off,
                               const float* d_tokens, const uint64_t* doc_tok, size_t dimension) {
    VB_MV_SINGLE_GPU_ONLY(index);
    if (!doc_tok) return finish(vb::Status::Cuda("document token offsets required"));
    return finish(index->impl->insert_many_device(ndocs, ids, id_off, d_tokens, 0, dimension, doc_tok));
}
int vb_mv_delete(vb_mv* index, const char* id, size_t id_len) {
    return finish(index->sharded ? index->sharded->remove(id, id_len) : index->impl->remove(id, id_len));
}
int vb_mv_search(vb_mv* index, const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, vb_hits** out) {
    *out = nullptr;
    vb::Hits hits;
    vb::Status s = index->sharded ? index->sharded->search(q_vals, q_off, tq, limit, &hits)
                                  : index->impl->search(q_vals, q_off, tq, limit, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}
int vb_mv_search_packed_device(vb_mv* index, const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit,
                               uint64_t* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts, vb_hits** out) {
    *out = nullptr;
    VB_MV_SINGLE_GPU_ONLY(index);
    vb::Hits hits;
    vb::Status s = index->impl->search_packed_device(q_vals, q_off, tq, limit, reinterpret_cast<vb::u64*>(d_keys), d_values,
                                                     d_rows, d_counts, &hits);
    if (!s.ok()) return finish(s);
    *out = new vb_hits{std::move(hits)};
    return VB_OK;
}
int vb_mv_set_id_ranks(vb_mv* index, const uint32_t* ranks, size_t n) {
    VB_MV_SINGLE_GPU_ONLY(index);
    return finish(index->impl->set_id_ranks(ranks, n));
}
int vb_mv_info(vb_mv* index, size_t* docs, size_t* tokens, size_t* dimension) {
    if (index->sharded) index->sharded->info(docs, tokens, dimension);
    else index->impl->info(docs, tokens, dimension);
    return VB_OK;
}

int vb_muvera_encode(size_t ndocs, const float* vals, const uint64_t* vec_off, const uint64_t* doc_vec, size_t dimension,
                     size_t num_repetitions, size_t num_simhash_projections, uint64_t seed, size_t projection_dimension,
                     int has_final, size_t final_projection_dimension, int mode, float* out, size_t out_capacity,
                     size_t* fde_dimension) {
    *fde_dimension = 0;
    if (mode != 0 && mode != 1) return finish(vb::Status::Cuda("muvera: mode must be 0 (query) or 1 (document)"));
    vb::MuveraConfig c;
    c.dimension = dimension;
    c.num_repetitions = num_repetitions;
    c.num_simhash_projections = num_simhash_projections;
    c.seed = seed;
    c.projection_dimension = projection_dimension;
    c.has_final = has_final != 0;
    c.final_projection_dimension = final_projection_dimension;
    // validation that needs no device runs first (muvera.rs:77-108), so error strings come back on any host
    for (size_t d = 0; d < ndocs; ++d) {
        if (doc_vec[d + 1] == doc_vec[d]) return finish(vb::Status::Ref("empty vectors"));
        size_t a = 0, b = 0;
        vb::Status cs = vb::muvera_output_dimension(c, &a, &b);
        const bool size_error = !cs.ok() && (cs.msg == "fde dimension overflow" || cs.msg == "fde dimension exceeds safety limit");
        if (!cs.ok() && !size_error) return finish(cs);
        for (size_t v = doc_vec[d]; v < doc_vec[d + 1]; ++v)
            if (vec_off[v + 1] - vec_off[v] != dimension) return finish(vb::Status::Ref("dimension mismatch"));
        for (size_t v = doc_vec[d]; v < doc_vec[d + 1]; ++v)
            if (!all_finite(vals + vec_off[v], dimension)) return finish(vb::Status::Ref("vector contains a non-finite value"));
        if (size_error) return finish(cs);
    }
    size_t full = 0, fin = 0;
    vb::Status cs = vb::muvera_output_dimension(c, &full, &fin);
    if (!cs.ok()) return finish(cs);
    *fde_dimension = fin;
    if (ndocs == 0 || out == nullptr) return VB_OK;   // out == NULL: validate and size only
    if (vb_device_count() <= 0) return no_device();
    vb::CtxLease ctx;
    vb::Status s = ctx.get();
    if (!s.ok()) return finish(s);
    return finish(vb::muvera_encode_batch(*ctx.ctx, c, ndocs, vals, vec_off, doc_vec, mode, out, out_capacity));
}

int vb_result_values(int metric_code, int score_mode, const float* raw, size_t n, double* score, double* distance) {
    if (metric_code < 0 || metric_code > 8) return finish(vb::Status::Ref("unknown metric"));
    if (score_mode != 0 && score_mode != 1) return finish(vb::Status::Ref("unknown score mode"));
    const bool similarity_metric = metric_code == vb::kCosine || metric_code == vb::kInnerProduct;   // @similarity_metrics
    for (size_t i = 0; i < n; ++i) {
        const double r = (double)raw[i];                     // the BEAM holds the NIF's f32 as an f64
        if (metric_code == vb::kNegativeInnerProduct) {      // vettore_distance.ex:527-529 (both modes)
            score[i] = -r;
            distance[i] = r;
        } else if (similarity_metric) {                      // :531-532, :537-538
            distance[i] = metric_code == vb::kCosine ? 1.0 - r : -r;
            score[i] = (score_mode == 1 && metric_code == vb::kCosine) ? (r + 1.0) / 2.0 : r;
        } else {                                             // :534-535, :540-541
            score[i] = score_mode == 1 ? 1.0 / (1.0 + r) : -r;
            distance[i] = r;
        }
    }
    return VB_OK;
}

int vb_compress_sign_bits(const float* vector, size_t len, uint64_t* words) {
    const size_t nw = (len + 63) / 64;
    for (size_t w = 0; w < nw; ++w) words[w] = 0;
    for (size_t i = 0; i < len; ++i)
        if (vector[i] >= 0.0f) words[i / 64] |= 1ull << (i % 64);   // distances.rs:416-420
    return VB_OK;
}

}  // extern "C"
