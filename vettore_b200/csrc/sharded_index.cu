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
    struct Ref { uint32_t key; uint32_t shard; uint32_t pos; };
    std::vector<Ref> pool;
    for (size_t q = 0; q < nq; ++q) {
        pool.clear();
        for (size_t s = 0; s < G; ++s) {
            const Hits& h = part[s][q];
            for (size_t i = 0; i < h.size(); ++i) {
                const float raw = h.values[i];
                float rank = raw;
                if (metric_ == kCosine) rank = 1.0f - raw;
                else if (metric_ == kInnerProduct) rank = -raw;
                pool.push_back(Ref{order_key(rank), (uint32_t)s, (uint32_t)i});
            }
        }
        auto id_of = [&](const Ref& r, size_t* n) {
            const Hits& h = part[r.shard][q];
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
        Hits& dst = (*out)[q];
        const size_t take = std::min(limit, pool.size());
        for (size_t i = 0; i < take; ++i) {
            size_t il;
            const char* id = id_of(pool[i], &il);
            const Hits& h = part[pool[i].shard][q];
            dst.add(id, il, h.values[pool[i].pos], ((uint64_t)pool[i].shard << 32) | h.index[pool[i].pos]);
        }
    }
    return Status::Ok();
}

}  // namespace vb
