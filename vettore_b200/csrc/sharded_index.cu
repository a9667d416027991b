// sharded_index.cu — see sharded_index.h.
#include "sharded_index.h"

#include <algorithm>
#include <cmath>

namespace vb {

ShardWorker::ShardWorker() : thread_([this] { loop(); }) {}

ShardWorker::~ShardWorker() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    thread_.join();
}

void ShardWorker::post(std::function<void()> task) {
    {
        std::lock_guard<std::mutex> g(mu_);
        tasks_.push_back(std::move(task));
    }
    cv_.notify_one();
}

void ShardWorker::loop() {
    for (;;) {
        std::function<void()> task;
        {
            std::unique_lock<std::mutex> g(mu_);
            cv_.wait(g, [this] { return stop_ || !tasks_.empty(); });
            if (tasks_.empty()) return;   // stop requested and nothing left to do
            task = std::move(tasks_.front());
            tasks_.pop_front();
        }
        task();
    }
}

ShardedFlatIndex::ShardedFlatIndex(int metric, const std::vector<int>& devices) : metric_(metric) {
    for (size_t s = 0; s < devices.size(); ++s) {
        shards_.emplace_back(new FlatIndex(metric, devices[s]));
        if (s > 0) workers_.emplace_back(new ShardWorker());
    }
}

ShardedFlatIndex::~ShardedFlatIndex() {
    workers_.clear();   // joins the threads before the shards go away
    shards_.clear();
}

size_t ShardedFlatIndex::shard_of(const char* id, size_t len) const {
    uint64_t h = 1469598103934665603ull;   // FNV-1a over the id bytes
    for (size_t i = 0; i < len; ++i) { h ^= (unsigned char)id[i]; h *= 1099511628211ull; }
    h ^= h >> 32;
    return (size_t)(h % shards_.size());
}

void ShardedFlatIndex::for_each_shard(const std::function<void(size_t)>& fn) {
    const size_t g = shards_.size();
    std::mutex mu;
    std::condition_variable cv;
    size_t pending = g - 1;
    for (size_t s = 1; s < g; ++s) {
        workers_[s - 1]->post([&, s] {
            fn(s);
            std::lock_guard<std::mutex> lk(mu);
            if (--pending == 0) cv.notify_one();
        });
    }
    fn(0);
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return pending == 0; });
}

void ShardedFlatIndex::info(size_t* rows, size_t* dim) {
    std::shared_lock<std::shared_mutex> g(mu_);
    *rows = rows_;
    *dim = dim_;
}

Status ShardedFlatIndex::reserve(size_t rows) {
    std::unique_lock<std::shared_mutex> g(mu_);
    const size_t per = rows / shards_.size() + rows / (8 * shards_.size()) + 1024;   // hash routing is not perfectly even
    std::vector<Status> st(shards_.size());
    for_each_shard([&](size_t s) { st[s] = shards_[s]->reserve(per); });
    for (auto& x : st) VB_TRY(x);
    return Status::Ok();
}

Status ShardedFlatIndex::insert_many(size_t n, const char* ids, const uint64_t* id_off, const float* values,
                                     const uint64_t* value_off) {
    std::unique_lock<std::shared_mutex> g(mu_);
    // flat.rs:70-76 over the WHOLE batch before any shard is touched (all-or-nothing): every row against the
    // index dimension, or the first row's when the index is empty.
    size_t expected = dim_;
    if (expected == 0 && n > 0) expected = value_off[1] - value_off[0];
    for (size_t i = 0; i < n; ++i) {
        const size_t len = value_off[i + 1] - value_off[i];
        if (len == 0) return Status::Ref("vector must not be empty");
        if (len != expected) return Status::Ref("dimension mismatch");
        const float* v = values + value_off[i];
        for (size_t c = 0; c < len; ++c)
            if (!std::isfinite(v[c])) return Status::Ref("vector contains a non-finite value");
    }
    if (n == 0) return Status::Ok();
    // split by owner shard (batch order is kept inside a shard: duplicate ids, last wins)
    const size_t G = shards_.size();
    struct Part { std::string ids; std::vector<uint64_t> id_off{0}; std::vector<float> vals; std::vector<uint64_t> val_off{0}; size_t n = 0; };
    std::vector<Part> parts(G);
    for (size_t i = 0; i < n; ++i) {
        const char* id = ids + id_off[i];
        const size_t il = id_off[i + 1] - id_off[i];
        Part& p = parts[shard_of(id, il)];
        p.ids.append(id, il);
        p.id_off.push_back(p.ids.size());
        p.vals.insert(p.vals.end(), values + value_off[i], values + value_off[i + 1]);
        p.val_off.push_back(p.vals.size());
        ++p.n;
    }
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) {
        Part& p = parts[s];
        if (p.n == 0) return;
        st[s] = shards_[s]->insert_many(p.n, p.ids.data(), p.id_off.data(), p.vals.data(), p.val_off.data(), false);
    });
    dim_ = expected;
    size_t rows = 0;
    for (size_t s = 0; s < G; ++s) {
        size_t r = 0, d = 0;
        shards_[s]->info(&r, &d);
        rows += r;
    }
    rows_ = rows;
    for (auto& x : st) VB_TRY(x);   // device failures only: the reference-visible validation already passed
    return Status::Ok();
}

Status ShardedFlatIndex::remove(const char* id, size_t id_len) {
    std::unique_lock<std::shared_mutex> g(mu_);
    FlatIndex& sh = *shards_[shard_of(id, id_len)];
    size_t before = 0, after = 0, d = 0;
    sh.info(&before, &d);
    VB_TRY(sh.remove(id, id_len));
    sh.info(&after, &d);
    rows_ -= before - after;
    if (rows_ == 0) dim_ = 0;   // flat.rs:90-92: dimension resets to None
    return Status::Ok();
}

namespace {

// Merges per-shard sorted lists into the best `limit` by (order key of the rank, id bytes) — FlatHit's /
// SearchHit's order (flat.rs:34-40, search.rs:22-31). `rank_of` maps a hit's value to its rank.
template <typename RankFn>
void merge_parts(const std::vector<Hits>& part, size_t limit, RankFn rank_of, Hits* dst) {
    struct Ref { uint32_t key; uint32_t shard; uint32_t pos; };
    std::vector<Ref> pool;
    for (size_t s = 0; s < part.size(); ++s)
        for (size_t i = 0; i < part[s].size(); ++i)
            pool.push_back(Ref{order_key(rank_of(part[s].values[i])), (uint32_t)s, (uint32_t)i});
    auto id_of = [&](const Ref& r, size_t* n) {
        const Hits& h = part[r.shard];
        *n = (size_t)(h.off[r.pos + 1] - h.off[r.pos]);
        return h.blob.data() + h.off[r.pos];
    };
    std::sort(pool.begin(), pool.end(), [&](const Ref& a, const Ref& b) {
        if (a.key != b.key) return a.key < b.key;
        size_t la, lb;
        const char* ia = id_of(a, &la);
        const char* ib = id_of(b, &lb);
        const int c = std::memcmp(ia, ib, std::min(la, lb));
        return c != 0 ? c < 0 : la < lb;
    });
    const size_t take = std::min(limit, pool.size());
    for (size_t i = 0; i < take; ++i) {
        size_t il;
        const char* id = id_of(pool[i], &il);
        const Hits& h = part[pool[i].shard];
        dst->add(id, il, h.values[pool[i].pos], ((uint64_t)pool[i].shard << 32) | (uint32_t)h.index[pool[i].pos]);
    }
}

float rank_of_metric(int metric_code, float raw) {   // distances.rs:113-119
    if (metric_code == kCosine) return 1.0f - raw;
    if (metric_code == kInnerProduct) return -raw;
    return raw;
}

bool finite_prefix(const float* v, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) return false;
    return true;
}

}  // namespace

Status ShardedFlatIndex::search(const float* queries, size_t nq, size_t len, size_t limit, std::vector<Hits>* out) {
    out->assign(nq, Hits{});
    if (limit == 0 || nq == 0) return Status::Ok();   // flat.rs:97-99: before any validation
    std::shared_lock<std::shared_mutex> g(mu_);
    for (size_t q = 0; q < nq; ++q) {                  // flat.rs:101 against the dimension of the whole index
        if (len == 0) return Status::Ref("vector must not be empty");
        if (dim_ != 0 && len != dim_) return Status::Ref("dimension mismatch");
        const float* v = queries + q * len;
        for (size_t c = 0; c < len; ++c)
            if (!std::isfinite(v[c])) return Status::Ref("vector contains a non-finite value");
    }
    if (rows_ == 0) return Status::Ok();
    const size_t G = shards_.size();
    std::vector<std::vector<Hits>> part(G);
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) {
        size_t r = 0, d = 0;
        shards_[s]->info(&r, &d);
        if (r == 0) { part[s].assign(nq, Hits{}); return; }   // an empty shard has no dimension to check against
        st[s] = shards_[s]->search(queries, nq, len, limit, &part[s]);
    });
    for (auto& x : st) VB_TRY(x);   // "metric overflow" of any shard aborts the search (flat.rs:105)
    // merge: FlatHit order = (rank.total_cmp, id bytes), flat.rs:34-40; rank per distances.rs:113-119
    std::vector<Hits> lists(G);
    for (size_t q = 0; q < nq; ++q) {
        for (size_t s = 0; s < G; ++s) lists[s] = std::move(part[s][q]);
        const int metric = metric_;
        merge_parts(lists, limit, [metric](float raw) { return rank_of_metric(metric, raw); }, &(*out)[q]);
    }
    return Status::Ok();
}


Status ShardedFlatIndex::stage_top_k(const Hits* from, const float* query, size_t len, int metric_code, size_t dimensions,
                                     size_t limit, Hits* out) {
    const size_t G = shards_.size();
    // survivors of the previous stage, regrouped by owner shard (the shard is in the upper half of a hit's index)
    struct IdList { std::string blob; std::vector<uint64_t> off{0}; };
    std::vector<IdList> own(G);
    if (from) {
        for (size_t i = 0; i < from->size(); ++i) {
            IdList& l = own[(size_t)(from->index[i] >> 32)];
            l.blob.append(from->blob.data() + from->off[i], (size_t)(from->off[i + 1] - from->off[i]));
            l.off.push_back(l.blob.size());
        }
    }
    std::vector<Hits> part(G);
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) {
        if (from && own[s].off.size() == 1) return;   // owns none of the survivors
        st[s] = shards_[s]->prefix_top_k(from == nullptr, from ? own[s].off.size() - 1 : 0, own[s].blob.data(),
                                         own[s].off.data(), query, len, metric_code, dimensions, limit, &part[s]);
    });
    for (auto& x : st) VB_TRY(x);
    *out = Hits{};
    merge_parts(part, limit, [metric_code](float raw) { return rank_of_metric(metric_code, raw); }, out);
    return Status::Ok();
}

Status ShardedFlatIndex::prefix_top_k(bool all_rows, size_t n_ids, const char* ids, const uint64_t* id_off,
                                      const float* query, size_t len, int metric_code, size_t dimensions, size_t limit,
                                      Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");              // nifs.rs:160
    if (dimensions == 0 || dimensions > len) return Status::Ref("invalid prefix dimensions");    // search.rs:45-47
    if (!finite_prefix(query, dimensions)) return Status::Ref("vector contains a non-finite value");
    std::shared_lock<std::shared_mutex> g(mu_);
    if (rows_ == 0) return Status::Ok();
    if (all_rows) {
        if (dimensions > dim_) return Status::Ref("dimension mismatch");                       // search.rs:52-54
        return stage_top_k(nullptr, query, len, metric_code, dimensions, limit, out);
    }
    // the listed ids, each on its owner shard: reuse the stage path with a synthetic survivor list
    Hits listed;
    for (size_t i = 0; i < n_ids; ++i) {
        const char* id = ids + id_off[i];
        const size_t il = (size_t)(id_off[i + 1] - id_off[i]);
        listed.add(id, il, 0.0f, (uint64_t)shard_of(id, il) << 32);
    }
    if (listed.size() == 0) return Status::Ok();
    return stage_top_k(&listed, query, len, metric_code, dimensions, limit, out);
}

Status ShardedFlatIndex::funnel_search(const float* query, size_t len, int metric_code, const size_t* stages,
                                       size_t nstages, size_t candidates, size_t limit, Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");
    for (size_t s = 0; s <= nstages; ++s) {   // every stage is a vector_top_k call: validated in the order the reference meets them
        const size_t d = s < nstages ? stages[s] : len;
        if (d == 0 || d > len) return Status::Ref("invalid prefix dimensions");
        if (!finite_prefix(query, d)) return Status::Ref("vector contains a non-finite value");
    }
    std::shared_lock<std::shared_mutex> g(mu_);
    if (rows_ == 0) return Status::Ok();
    for (size_t s = 0; s <= nstages; ++s)
        if ((s < nstages ? stages[s] : len) > dim_) return Status::Ref("dimension mismatch");
    if (candidates == 0) return Status::Ok();   // vector_top_k(limit 0) -> [] at the first stage (search.rs:48)
    Hits cur, next;
    for (size_t s = 0; s < nstages; ++s) {
        VB_TRY(stage_top_k(s == 0 ? nullptr : &cur, query, len, metric_code, stages[s], candidates, &next));
        std::swap(cur, next);
        if (cur.size() == 0) return Status::Ok();
    }
    return stage_top_k(nstages == 0 ? nullptr : &cur, query, len, metric_code, len, limit, out);   // exact rerank
}

Status ShardedFlatIndex::quantized_search(const float* query, size_t len, int metric_code, size_t candidates, size_t limit,
                                          Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");
    if (len == 0) return Status::Ref("vector must not be empty");
    if (!finite_prefix(query, len)) return Status::Ref("vector contains a non-finite value");
    std::shared_lock<std::shared_mutex> g(mu_);
    if (rows_ == 0) return Status::Ok();
    if (len != dim_) return Status::Ref("dimension mismatch");
    if (std::min(candidates, rows_) == 0) return Status::Ok();   // binary_top_k(limit 0) -> [] (search.rs:95-97)
    const size_t G = shards_.size();
    std::vector<Hits> part(G);
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) { st[s] = shards_[s]->hamming_candidates(query, len, candidates, &part[s]); });
    for (auto& x : st) VB_TRY(x);
    Hits cand;   // global candidate set: (distance, id bytes) ascending, search.rs:87-90
    merge_parts(part, candidates, [](float d) { return d; }, &cand);
    if (cand.size() == 0) return Status::Ok();
    return stage_top_k(&cand, query, len, metric_code, len, limit, out);   // exact rerank on the owners
}

// ---------------------------------------------------------------------------------------------------
// Multi-vector collection over several GPUs (see sharded_index.h)
ShardedMvIndex::ShardedMvIndex(int metric, const std::vector<int>& devices) {
    for (size_t s = 0; s < devices.size(); ++s) {
        shards_.emplace_back(new MvIndex(metric, devices[s]));
        if (s > 0) workers_.emplace_back(new ShardWorker());
    }
}

ShardedMvIndex::~ShardedMvIndex() {
    workers_.clear();
    shards_.clear();
}

size_t ShardedMvIndex::shard_of(const char* id, size_t len) const {
    uint64_t h = 1469598103934665603ull;   // FNV-1a over the id bytes (as ShardedFlatIndex)
    for (size_t i = 0; i < len; ++i) { h ^= (unsigned char)id[i]; h *= 1099511628211ull; }
    h ^= h >> 32;
    return (size_t)(h % shards_.size());
}

void ShardedMvIndex::for_each_shard(const std::function<void(size_t)>& fn) {
    const size_t g = shards_.size();
    std::mutex mu;
    std::condition_variable cv;
    size_t pending = g - 1;
    for (size_t s = 1; s < g; ++s) {
        workers_[s - 1]->post([&, s] {
            fn(s);
            std::lock_guard<std::mutex> lk(mu);
            if (--pending == 0) cv.notify_one();
        });
    }
    fn(0);
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return pending == 0; });
}

void ShardedMvIndex::refresh_totals() {
    docs_ = tokens_ = 0;
    for (auto& sh : shards_) {
        size_t d = 0, t = 0, dm = 0;
        sh->info(&d, &t, &dm);
        docs_ += d;
        tokens_ += t;
    }
    if (docs_ == 0) dim_ = 0;   // like MvIndex: an empty collection forgets its dimension
}

void ShardedMvIndex::info(size_t* docs, size_t* tokens, size_t* dim) {
    std::shared_lock<std::shared_mutex> g(mu_);
    *docs = docs_;
    *tokens = tokens_;
    *dim = dim_;
}

Status ShardedMvIndex::insert_many(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                                   const uint64_t* tok_off, const uint64_t* doc_tok) {
    std::unique_lock<std::shared_mutex> g(mu_);
    // multi_vector.rs:134-152 over the WHOLE batch against the collection's dimension before any shard is touched
    size_t expected = dim_;
    for (size_t d = 0; d < ndocs; ++d) {
        for (size_t t = doc_tok[d]; t < doc_tok[d + 1]; ++t) {
            const size_t len = tok_off[t + 1] - tok_off[t];
            if (len == 0) return Status::Ref("vectors must not be empty");
            if (expected == 0) expected = len;
            if (len != expected) return Status::Ref("dimension mismatch");
            for (size_t c = 0; c < len; ++c)
                if (!std::isfinite(tok_vals[tok_off[t] + c])) return Status::Ref("vector contains a non-finite value");
        }
    }
    if (ndocs == 0) return Status::Ok();
    const size_t G = shards_.size();
    struct Part { std::string ids; std::vector<uint64_t> id_off{0}, tok_off{0}, doc_tok{0}; std::vector<float> vals; size_t n = 0; };
    std::vector<Part> parts(G);
    for (size_t d = 0; d < ndocs; ++d) {
        const char* id = ids + id_off[d];
        const size_t il = id_off[d + 1] - id_off[d];
        Part& p = parts[shard_of(id, il)];
        p.ids.append(id, il);
        p.id_off.push_back(p.ids.size());
        for (size_t t = doc_tok[d]; t < doc_tok[d + 1]; ++t) {
            p.vals.insert(p.vals.end(), tok_vals + tok_off[t], tok_vals + tok_off[t + 1]);
            p.tok_off.push_back(p.vals.size());
        }
        p.doc_tok.push_back(p.tok_off.size() - 1);
        ++p.n;
    }
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) {
        Part& p = parts[s];
        if (p.n == 0) return;
        st[s] = shards_[s]->insert_many(p.n, p.ids.data(), p.id_off.data(), p.vals.data(), p.tok_off.data(), p.doc_tok.data());
    });
    if (expected != 0) dim_ = expected;
    refresh_totals();
    for (auto& x : st) VB_TRY(x);
    return Status::Ok();
}

Status ShardedMvIndex::remove(const char* id, size_t id_len) {
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_TRY(shards_[shard_of(id, id_len)]->remove(id, id_len));
    refresh_totals();
    return Status::Ok();
}

Status ShardedMvIndex::search(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, Hits* out) {
    *out = Hits{};
    // multi_vector.rs:96-97: the query is validated on its own first
    size_t qdim = 0;
    for (size_t t = 0; t < tq; ++t) {
        const size_t len = q_off[t + 1] - q_off[t];
        if (t == 0) {
            if (len == 0) return Status::Ref("vectors must not be empty");
            qdim = len;
        }
        if (len != qdim) return Status::Ref("dimension mismatch");
        for (size_t c = 0; c < len; ++c)
            if (!std::isfinite(q_vals[q_off[t] + c])) return Status::Ref("vector contains a non-finite value");
    }
    std::shared_lock<std::shared_mutex> g(mu_);
    if (docs_ == 0) return Status::Ok();
    if (tq > 0 && tokens_ > 0 && qdim != dim_) return Status::Ref("dimension mismatch");   // :108
    if (limit == 0) return Status::Ok();
    const size_t G = shards_.size();
    std::vector<Hits> part(G);
    std::vector<Status> st(G);
    for_each_shard([&](size_t s) {
        size_t d = 0, t = 0, dm = 0;
        shards_[s]->info(&d, &t, &dm);
        if (d == 0) return;
        st[s] = shards_[s]->search(q_vals, q_off, tq, limit, &part[s]);   // a shard of empty documents scores 0.0 for all (:102-106)
    });
    for (auto& x : st) VB_TRY(x);   // "score overflow" / "metric overflow" of any shard aborts the search
    merge_parts(part, limit, [](float score) { return -score; }, out);   // descending score, ascending id (:22-31)
    return Status::Ok();
}

}  // namespace vb