// mv_index.cu — HBM-resident multi-vector index (see maxsim.h) and the by-value entry.
#include <algorithm>
#include <cmath>
#include <mutex>

#include "maxsim.h"

namespace vb {

namespace {
bool all_finite(const float* v, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) return false;
    return true;
}
constexpr uint32_t kDead = 0xFFFFFFFFu;
}  // namespace

MvIndex::~MvIndex() {
    cudaSetDevice(device_);
    if (d_tokens_) cudaFree(d_tokens_);
    if (d_inv_norm_) cudaFree(d_inv_norm_);
    if (d_tok_doc_) cudaFree(d_tok_doc_);
    if (d_doc_off_) cudaFree(d_doc_off_);
    if (d_doc_rank_) cudaFree(d_doc_rank_);
}

void MvIndex::info(size_t* docs, size_t* tokens, size_t* dim) {
    std::shared_lock<std::shared_mutex> g(mu_);
    *docs = id_doc_.size();
    *tokens = ntok_ - dead_tok_;
    *dim = dim_;
}

Status MvIndex::reserve_tokens(size_t need) {
    if (need <= tok_cap_) return Status::Ok();
    size_t cap = std::max<size_t>(need, std::max<size_t>(tok_cap_ * 2, 4096));
    float* t = nullptr;
    cudaError_t e = cudaMalloc(&t, cap * stride_ * sizeof(float));
    if (e != cudaSuccess && cap > need) {
        cudaGetLastError();
        cap = need;
        e = cudaMalloc(&t, cap * stride_ * sizeof(float));
    }
    if (e != cudaSuccess) return Status::Cuda(cudaGetErrorString(e));
    float* inv = nullptr;
    uint32_t* owner = nullptr;
    VB_CUDA(cudaMalloc(&inv, cap * sizeof(float)));
    VB_CUDA(cudaMalloc(&owner, cap * sizeof(uint32_t)));
    if (ntok_) {
        VB_CUDA(cudaMemcpy(t, d_tokens_, ntok_ * stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
        VB_CUDA(cudaMemcpy(inv, d_inv_norm_, ntok_ * sizeof(float), cudaMemcpyDeviceToDevice));
        VB_CUDA(cudaMemcpy(owner, d_tok_doc_, ntok_ * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
    }
    if (d_tokens_) cudaFree(d_tokens_);
    if (d_inv_norm_) cudaFree(d_inv_norm_);
    if (d_tok_doc_) cudaFree(d_tok_doc_);
    d_tokens_ = t;
    d_inv_norm_ = inv;
    d_tok_doc_ = owner;
    tok_cap_ = cap;
    return Status::Ok();
}

Status MvIndex::reserve_docs(size_t need) {
    if (need <= doc_cap_) return Status::Ok();
    const size_t cap = std::max<size_t>(need, std::max<size_t>(doc_cap_ * 2, 1024));
    uint32_t *off = nullptr, *rank = nullptr;
    VB_CUDA(cudaMalloc(&off, (cap + 1) * sizeof(uint32_t)));
    VB_CUDA(cudaMalloc(&rank, cap * sizeof(uint32_t)));
    if (d_doc_off_) cudaFree(d_doc_off_);
    if (d_doc_rank_) cudaFree(d_doc_rank_);
    d_doc_off_ = off;
    d_doc_rank_ = rank;
    doc_cap_ = cap;
    return Status::Ok();  // contents are re-uploaded by the caller
}

// Ranks follow the byte order of the live ids (map order); tombstones keep kDead.
Status MvIndex::relabel() {
    uint32_t r = 0;
    for (auto& kv : id_doc_) h_rank_[kv.second] = r++;
    return Status::Ok();
}

Status MvIndex::compact() {
    // Rebuild the token matrix without tombstoned documents (device-side gather per document).
    float* fresh = nullptr;
    float* fresh_inv = nullptr;
    const size_t live = ntok_ - dead_tok_;
    VB_CUDA(cudaMalloc(&fresh, std::max<size_t>(live, 1) * stride_ * sizeof(float)));
    VB_CUDA(cudaMalloc(&fresh_inv, std::max<size_t>(live, 1) * sizeof(float)));
    uint32_t* fresh_owner = nullptr;
    VB_CUDA(cudaMalloc(&fresh_owner, std::max<size_t>(live, 1) * sizeof(uint32_t)));
    std::vector<uint32_t> owner(live);
    std::vector<uint32_t> off{0};
    std::vector<uint32_t> rank;
    std::vector<std::string> ids;
    size_t cursor = 0;
    for (size_t d = 0; d < ndocs_; ++d) {
        if (h_rank_[d] == kDead) continue;
        const size_t t0 = h_doc_off_[d], cnt = h_doc_off_[d + 1] - t0;
        if (cnt) {
            VB_CUDA(cudaMemcpy(fresh + cursor * stride_, d_tokens_ + t0 * stride_, cnt * stride_ * sizeof(float),
                               cudaMemcpyDeviceToDevice));
            VB_CUDA(cudaMemcpy(fresh_inv + cursor, d_inv_norm_ + t0, cnt * sizeof(float), cudaMemcpyDeviceToDevice));
        }
        for (size_t t = 0; t < cnt; ++t) owner[cursor + t] = (uint32_t)ids.size();
        cursor += cnt;
        id_doc_[doc_id_[d]] = (uint32_t)ids.size();
        ids.push_back(std::move(doc_id_[d]));
        rank.push_back(h_rank_[d]);
        off.push_back((uint32_t)cursor);
    }
    if (live) VB_CUDA(cudaMemcpy(fresh_owner, owner.data(), live * sizeof(uint32_t), cudaMemcpyHostToDevice));
    cudaFree(d_tokens_);
    cudaFree(d_inv_norm_);
    cudaFree(d_tok_doc_);
    d_tokens_ = fresh;
    d_inv_norm_ = fresh_inv;
    d_tok_doc_ = fresh_owner;
    uniform_known_ = false;   // recomputed from the surviving documents below
    min_td_ = 0;
    has_empty_ = false;
    tok_cap_ = std::max<size_t>(live, 1);
    ntok_ = live;
    dead_tok_ = 0;
    ndocs_ = ids.size();
    doc_id_ = std::move(ids);
    h_rank_ = std::move(rank);
    h_doc_off_ = std::move(off);
    for (size_t d = 0; d < ndocs_; ++d) {
        const uint32_t cnt = h_doc_off_[d + 1] - h_doc_off_[d];
        if (!uniform_known_) { uniform_td_ = cnt; uniform_known_ = true; }
        else if (uniform_td_ != cnt) uniform_td_ = 0;
        note_doc_length(cnt);
    }
    return Status::Ok();
}

Status MvIndex::insert_many(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                            const uint64_t* tok_off, const uint64_t* doc_tok) {
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_CUDA(cudaSetDevice(device_));
    // All-or-nothing validation (the same rules multi_vector.rs:134-152 applies at scoring time).
    size_t expected = dim_;
    for (size_t d = 0; d < ndocs; ++d) {
        for (size_t t = doc_tok[d]; t < doc_tok[d + 1]; ++t) {
            const size_t len = tok_off[t + 1] - tok_off[t];
            if (len == 0) return Status::Ref("vectors must not be empty");
            if (expected == 0) expected = len;
            if (len != expected) return Status::Ref("dimension mismatch");
            if (!all_finite(tok_vals + tok_off[t], len)) return Status::Ref("vector contains a non-finite value");
        }
    }
    if (ndocs == 0) return Status::Ok();
    if (dim_ == 0 && expected != 0) {
        dim_ = expected;
        stride_ = (dim_ + 3) & ~(size_t)3;
    }
    const size_t new_tok = doc_tok[ndocs] - doc_tok[0];
    if (ntok_ + new_tok >= 0xFFFFFFFFull || ndocs_ + ndocs >= 0xFFFFFFFEull)
        return Status::Cuda("multi-vector index limit (2^32 tokens) exceeded");
    if (stride_) VB_TRY(reserve_tokens(ntok_ + new_tok));

    // stage the new tokens contiguously (zero padded rows) and upload in one copy
    if (new_tok) {
        PinnedBuf stage;
        const size_t chunk_rows = std::max<size_t>(1, std::min<size_t>(new_tok, (64u << 20) / (stride_ * sizeof(float))));
        VB_TRY(stage.reserve(chunk_rows * stride_ * sizeof(float)));
        float* sb = stage.as<float>();
        std::vector<float> inv(chunk_rows);
        {   // owning document slot of every new token
            std::vector<uint32_t> owner(new_tok);
            for (size_t d = 0; d < ndocs; ++d)
                for (size_t t = doc_tok[d]; t < doc_tok[d + 1]; ++t) owner[t - doc_tok[0]] = (uint32_t)(ndocs_ + d);
            VB_CUDA(cudaMemcpy(d_tok_doc_ + ntok_, owner.data(), new_tok * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        size_t done = 0;
        while (done < new_tok) {
            const size_t n = std::min(chunk_rows, new_tok - done);
            for (size_t i = 0; i < n; ++i) {
                const size_t t = doc_tok[0] + done + i;
                const float* src = tok_vals + tok_off[t];
                std::memcpy(sb + i * stride_, src, dim_ * sizeof(float));
                for (size_t c = dim_; c < stride_; ++c) sb[i * stride_ + c] = 0.0f;
                double s = 0.0;   // distances.rs:166: f64_dot(right, right).sqrt()
                for (size_t c = 0; c < dim_; ++c) s += (double)src[c] * (double)src[c];
                inv[i] = s > 0.0 ? (float)(1.0 / std::sqrt(s)) : 0.0f;
            }
            VB_CUDA(cudaMemcpy(d_tokens_ + (ntok_ + done) * stride_, sb, n * stride_ * sizeof(float),
                               cudaMemcpyHostToDevice));
            VB_CUDA(cudaMemcpy(d_inv_norm_ + ntok_ + done, inv.data(), n * sizeof(float), cudaMemcpyHostToDevice));
            done += n;
        }
        stage.release();
    }
    for (size_t d = 0; d < ndocs; ++d) {
        std::string id(ids + id_off[d], ids + id_off[d + 1]);
        auto it = id_doc_.find(id);
        if (it != id_doc_.end()) {  // upsert: tombstone the previous copy
            const uint32_t old = it->second;
            dead_tok_ += h_doc_off_[old + 1] - h_doc_off_[old];
            h_rank_[old] = kDead;
            doc_id_[old].clear();
            it->second = (uint32_t)ndocs_;
        } else {
            id_doc_.emplace(id, (uint32_t)ndocs_);
        }
        doc_id_.push_back(std::move(id));
        h_rank_.push_back(0);
        {
            const uint32_t cnt = (uint32_t)(doc_tok[d + 1] - doc_tok[d]);
            if (!uniform_known_) { uniform_td_ = cnt; uniform_known_ = true; }
            else if (uniform_td_ != cnt) uniform_td_ = 0;
            note_doc_length(cnt);
        }
        ntok_ += doc_tok[d + 1] - doc_tok[d];
        h_doc_off_.push_back((uint32_t)ntok_);
        ++ndocs_;
    }
    if (dead_tok_ > ntok_ - dead_tok_ && dead_tok_ > 4096) VB_TRY(compact());
    VB_TRY(relabel());
    VB_TRY(reserve_docs(ndocs_));
    VB_CUDA(cudaMemcpy(d_doc_off_, h_doc_off_.data(), (ndocs_ + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
    VB_CUDA(cudaMemcpy(d_doc_rank_, h_rank_.data(), ndocs_ * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return Status::Ok();
}

// 1/|token| in f32 from an f64 norm (distances.rs:166), one warp per token; flags non-finite input.
__global__ void token_inv_norm_kernel(const float* tokens, size_t stride, uint32_t dim, uint32_t ntok, float* inv,
                                      uint32_t* bad) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t t = warp; t < ntok; t += warps) {
        double s = 0.0;
        bool finite = true;
        for (uint32_t c = lane; c < dim; c += 32) {
            const float v = tokens[t * stride + c];
            finite &= isfinite(v);
            s = fma((double)v, (double)v, s);
        }
        s = warp_sum(s);
        if (!__all_sync(0xffffffffu, finite)) { if (lane == 0) atomicOr(bad, 1u); }
        if (lane == 0) inv[t] = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.0f;
    }
}

// tok_doc for `ndocs` uniform documents of `td` tokens appended at document slot `doc0`.
__global__ void fill_tok_doc_kernel(uint32_t* tok_doc, size_t ntok, uint32_t td, uint32_t doc0) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntok; t += (size_t)gridDim.x * blockDim.x)
        tok_doc[t] = doc0 + (uint32_t)(t / td);
}

void MvIndex::note_doc_length(uint32_t cnt) {
    if (cnt == 0) has_empty_ = true;
    else if (min_td_ == 0 || cnt < min_td_) min_td_ = cnt;
}

Status MvIndex::reserve(size_t docs, size_t tokens, size_t dim) {
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_CUDA(cudaSetDevice(device_));
    if (dim == 0) return Status::Ref("vectors must not be empty");
    if (dim_ != 0 && dim != dim_) return Status::Ref("dimension mismatch");
    if (dim_ == 0) {
        dim_ = dim;
        stride_ = (dim + 3) & ~(size_t)3;
    }
    VB_TRY(reserve_tokens(tokens));
    if (docs > doc_cap_) {
        VB_TRY(reserve_docs(docs));
        if (ndocs_) {
            VB_CUDA(cudaMemcpy(d_doc_off_, h_doc_off_.data(), (ndocs_ + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
            VB_CUDA(cudaMemcpy(d_doc_rank_, h_rank_.data(), ndocs_ * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    return Status::Ok();
}

Status MvIndex::insert_many_device(size_t ndocs, const char* ids, const uint64_t* id_off, const float* d_tokens,
                                   size_t td, size_t dim, const uint64_t* doc_tok) {
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_CUDA(cudaSetDevice(device_));
    if (ndocs == 0) return Status::Ok();
    if (dim == 0 || (!doc_tok && td == 0)) return Status::Ref("vectors must not be empty");
    if (dim_ != 0 && dim != dim_) return Status::Ref("dimension mismatch");
    if (dim % 4 != 0) return Status::Cuda("device ingest needs a dimension that is a multiple of 4");
    // doc_tok (host, [ndocs + 1], ascending) makes the batch ragged: document i owns rows [doc_tok[i], doc_tok[i+1]).
    if (doc_tok)
        for (size_t d = 0; d < ndocs; ++d)
            if (doc_tok[d + 1] < doc_tok[d]) return Status::Cuda("document token offsets must ascend");
    const size_t new_tok = doc_tok ? (size_t)(doc_tok[ndocs] - doc_tok[0]) : ndocs * td;
    if (doc_tok) d_tokens += (size_t)doc_tok[0] * dim;
    if (ntok_ + new_tok >= 0xFFFFFFFFull || ndocs_ + ndocs >= 0xFFFFFFFEull)
        return Status::Cuda("multi-vector index limit (2^32 tokens) exceeded");
    if (dim_ == 0) {
        dim_ = dim;
        stride_ = dim;
    }
    VB_TRY(reserve_tokens(ntok_ + new_tok));
    // validate + norms on the device, then one D2D copy
    uint32_t* d_bad = nullptr;
    VB_CUDA(cudaMalloc(&d_bad, sizeof(uint32_t)));
    VB_CUDA(cudaMemset(d_bad, 0, sizeof(uint32_t)));
    token_inv_norm_kernel<<<148 * 8, 256>>>(d_tokens, dim, (uint32_t)dim, (uint32_t)new_tok, d_inv_norm_ + ntok_, d_bad);
    if (!doc_tok) {
        fill_tok_doc_kernel<<<148 * 4, 256>>>(d_tok_doc_ + ntok_, new_tok, (uint32_t)td, (uint32_t)ndocs_);
    } else if (new_tok) {
        std::vector<uint32_t> owner(new_tok);
        for (size_t d = 0; d < ndocs; ++d)
            for (uint64_t t = doc_tok[d]; t < doc_tok[d + 1]; ++t) owner[t - doc_tok[0]] = (uint32_t)(ndocs_ + d);
        VB_CUDA(cudaMemcpy(d_tok_doc_ + ntok_, owner.data(), new_tok * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    uint32_t bad = 0;
    cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess) return Status::Cuda(cudaGetErrorString(e));
    if (bad) return Status::Ref("vector contains a non-finite value");
    if (new_tok)
        VB_CUDA(cudaMemcpy(d_tokens_ + ntok_ * stride_, d_tokens, new_tok * stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
    for (size_t d = 0; d < ndocs; ++d) {
        std::string id(ids + id_off[d], ids + id_off[d + 1]);
        auto hint = id_doc_.end();
        if (id_doc_.empty() || std::prev(hint)->first < id) {
            id_doc_.emplace_hint(hint, id, (uint32_t)ndocs_);
        } else {
            auto it = id_doc_.find(id);
            if (it != id_doc_.end()) {
                const uint32_t old = it->second;
                dead_tok_ += h_doc_off_[old + 1] - h_doc_off_[old];
                h_rank_[old] = kDead;
                doc_id_[old].clear();
                it->second = (uint32_t)ndocs_;
            } else {
                id_doc_.emplace(id, (uint32_t)ndocs_);
            }
        }
        doc_id_.push_back(std::move(id));
        h_rank_.push_back(0);
        const size_t cnt = doc_tok ? (size_t)(doc_tok[d + 1] - doc_tok[d]) : td;
        if (!uniform_known_) { uniform_td_ = (uint32_t)cnt; uniform_known_ = true; }
        else if (uniform_td_ != cnt) uniform_td_ = 0;
        note_doc_length((uint32_t)cnt);
        ntok_ += cnt;
        h_doc_off_.push_back((uint32_t)ntok_);
        ++ndocs_;
    }
    VB_TRY(relabel());
    if (ndocs_ > doc_cap_) VB_TRY(reserve_docs(ndocs_));
    VB_CUDA(cudaMemcpy(d_doc_off_, h_doc_off_.data(), (ndocs_ + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
    VB_CUDA(cudaMemcpy(d_doc_rank_, h_rank_.data(), ndocs_ * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return Status::Ok();
}

Status MvIndex::remove(const char* id, size_t id_len) {
    std::unique_lock<std::shared_mutex> g(mu_);
    auto it = id_doc_.find(std::string(id, id + id_len));
    if (it == id_doc_.end()) return Status::Ok();
    VB_CUDA(cudaSetDevice(device_));
    const uint32_t slot = it->second;
    id_doc_.erase(it);
    dead_tok_ += h_doc_off_[slot + 1] - h_doc_off_[slot];
    h_rank_[slot] = kDead;
    doc_id_[slot].clear();
    VB_CUDA(cudaMemcpy(d_doc_rank_ + slot, &h_rank_[slot], sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (id_doc_.empty()) {  // like FlatIndex: an empty index forgets its dimension
        dim_ = stride_ = 0;
        ntok_ = dead_tok_ = tok_cap_ = 0;
        ndocs_ = 0;
        if (d_tokens_) cudaFree(d_tokens_);
        if (d_inv_norm_) cudaFree(d_inv_norm_);
        if (d_tok_doc_) cudaFree(d_tok_doc_);
        d_tokens_ = nullptr;
        d_inv_norm_ = nullptr;
        d_tok_doc_ = nullptr;
        uniform_known_ = false;
        uniform_td_ = 0;
        min_td_ = 0;
        has_empty_ = false;
        h_doc_off_.assign(1, 0);
        h_rank_.clear();
        doc_id_.clear();
    }
    return Status::Ok();
}

Status MvIndex::search(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, Hits* out) {
    return search_impl(q_vals, q_off, tq, limit, nullptr, nullptr, nullptr, nullptr, out);
}

Status MvIndex::search_packed_device(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, u64* d_keys,
                                     float* d_values, uint32_t* d_rows, uint32_t* d_counts, Hits* out) {
    if (!d_keys || !d_values || !d_rows || !d_counts) return Status::Cuda("device outputs required");
    if (limit == 0 || tq == 0) return Status::Cuda("sharded multi-vector search needs limit >= 1 and a non-empty query");
    VB_CUDA(cudaSetDevice(device_));
    VB_CUDA(cudaMemset(d_counts, 0, sizeof(uint32_t)));   // an empty shard contributes an empty list
    return search_impl(q_vals, q_off, tq, limit, d_keys, d_values, d_rows, d_counts, out);
}

Status MvIndex::set_id_ranks(const uint32_t* ranks, size_t n) {
    std::unique_lock<std::shared_mutex> g(mu_);
    if (n != ndocs_) return Status::Ref("dimension mismatch");
    VB_CUDA(cudaSetDevice(device_));
    for (size_t d = 0; d < n; ++d)
        if (h_rank_[d] != kDead) h_rank_[d] = ranks[d];
    if (n > 0) VB_CUDA(cudaMemcpy(d_doc_rank_, h_rank_.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return Status::Ok();
}

Status MvIndex::search_impl(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, u64* d_keys,
                            float* d_values, uint32_t* d_rows, uint32_t* d_counts, Hits* out) {
    *out = Hits{};
    // multi_vector.rs:96-97: the query is validated on its own first.
    size_t qdim = 0;
    for (size_t t = 0; t < tq; ++t) {
        const size_t len = q_off[t + 1] - q_off[t];
        if (t == 0) {
            if (len == 0) return Status::Ref("vectors must not be empty");
            qdim = len;
        }
        if (len != qdim) return Status::Ref("dimension mismatch");
        if (!all_finite(q_vals + q_off[t], len)) return Status::Ref("vector contains a non-finite value");
    }
    std::shared_lock<std::shared_mutex> g(mu_);
    if (id_doc_.empty()) return Status::Ok();
    if (tq > 0 && ntok_ - dead_tok_ > 0 && qdim != dim_) return Status::Ref("dimension mismatch");  // :108
    if (limit == 0) return Status::Ok();
    const size_t k = std::min(limit, id_doc_.size());
    if (tq == 0 || ntok_ - dead_tok_ == 0) {
        // empty query (or only empty documents): every score is 0.0, ties resolve by id (:102-106)
        size_t n = 0;
        for (auto& kv : id_doc_) {
            if (n++ == k) break;
            out->add(kv.first.data(), kv.first.size(), 0.0f, kv.second);
        }
        return Status::Ok();
    }
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    // the ragged query is contiguous when every token has qdim elements
    MaxSimJob job;
    job.metric = metric_ == kCosine ? kCosineTrue : metric_;   // multi_vector.rs:74-75
    job.d_tokens = d_tokens_;
    job.stride = stride_;
    job.d_doc_off = d_doc_off_;
    job.d_doc_rank = d_doc_rank_;
    job.ndocs = ndocs_;
    job.dims = (uint32_t)dim_;
    job.h_query = q_vals + q_off[0];
    job.tq = (uint32_t)tq;
    job.k = k;
    job.uniform_td = uniform_known_ ? uniform_td_ : 0;
    job.d_inv_dnorm = d_inv_norm_;
    job.d_tok_doc = d_tok_doc_;
    job.ntok = ntok_;
    job.min_td = min_td_;
    job.has_empty = has_empty_;
    job.d_keys_out = d_keys;
    job.d_values_out = d_values;
    job.d_rows_out = d_rows;
    job.d_counts_out = d_counts;
    MaxSimResult res;
    VB_TRY(maxsim_top_k(*ctx.ctx, job, &res));
    if (res.err != kNoError) return Status::Ref((res.err & 1u) ? "score overflow" : "metric overflow");
    for (size_t i = 0; i < res.rows.size(); ++i) {
        const std::string& id = doc_id_[res.rows[i]];
        out->add(id.data(), id.size(), res.scores[i], res.rows[i]);
    }
    return Status::Ok();
}

}  // namespace vb
