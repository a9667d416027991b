// sharded_index.h — one flat index spread over several GPUs INSIDE ONE PROCESS, behind the same C ABI
// handle as the single-GPU index. The reference surface is one BEAM process holding one FlatResource
// (flat.rs:131-134, nifs.rs:297-309): an erl_nif caller cannot fork one OS process per GPU, so the G-GPU
// path must be reachable through vb_flat_new_sharded + the ordinary vb_flat_insert / _search calls.
//
// Layout: shard s is a complete FlatIndex (HBM matrix + id table + workspaces) on device devices[s]; an id
// lives on shard fnv1a(id) % G, so upserts and deletes need no directory and shards stay balanced. Every
// shard has its own host thread (CUDA calls of different devices never serialise behind one another); a
// search posts the query to all of them, each runs the fused scan + top-k on its own GPU and stream, and
// the G sorted lists (G x k entries, ~100 bytes each way over PCIe) are merged on the calling thread by
// (rank.total_cmp, id bytes) — exactly FlatHit's order (flat.rs:34-40), comparing the real id bytes, so no
// cross-shard id-rank bookkeeping exists. For results that must stay on the device (one process per GPU
// under torchrun) the exchange is the NVLink peer-memory kernel of peer_exchange.cu instead.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <vector>

#include "flat_index.h"
#include "maxsim.h"

namespace vb {

class ShardWorker {
  public:
    ShardWorker();
    ~ShardWorker();
    void post(std::function<void()> task);
  private:
    void loop();
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> tasks_;
    bool stop_ = false;
    std::thread thread_;
};

class ShardedFlatIndex {
  public:
    ShardedFlatIndex(int metric, const std::vector<int>& devices);
    ~ShardedFlatIndex();
    size_t shards() const { return shards_.size(); }

    Status insert_many(size_t n, const char* ids, const uint64_t* id_off, const float* values, const uint64_t* value_off);
    Status reserve(size_t rows);
    Status remove(const char* id, size_t id_len);
    Status search(const float* queries, size_t nq, size_t len, size_t limit, std::vector<Hits>* out);
    // The resident pipelines over the whole sharded corpus (same semantics as the single-GPU index): every stage
    // runs on all shards at once, the shards' sorted lists are merged on the calling thread in the reference's
    // order, and the next stage re-scores each survivor on the shard that owns it.
    Status prefix_top_k(bool all_rows, size_t n_ids, const char* ids, const uint64_t* id_off, const float* query,
                        size_t len, int metric_code, size_t dimensions, size_t limit, Hits* out);
    Status funnel_search(const float* query, size_t len, int metric_code, const size_t* stages, size_t nstages,
                         size_t candidates, size_t limit, Hits* out);
    Status quantized_search(const float* query, size_t len, int metric_code, size_t candidates, size_t limit, Hits* out);
    void info(size_t* rows, size_t* dim);

  private:
    // One vector_top_k (search.rs:38-73) over all rows, or over the ids of `from` (each on its owner shard).
    Status stage_top_k(const Hits* from, const float* query, size_t len, int metric_code, size_t dimensions, size_t limit,
                       Hits* out);
    size_t shard_of(const char* id, size_t len) const;
    // Runs fn(shard) for every shard: shard 0 on the calling thread, the others on their workers.
    void for_each_shard(const std::function<void(size_t)>& fn);

    const int metric_;
    std::vector<std::unique_ptr<FlatIndex>> shards_;
    std::vector<std::unique_ptr<ShardWorker>> workers_;   // workers_[s - 1] drives shard s
    std::shared_mutex mu_;      // searches share, mutations exclude (nifs.rs:266-309) — across ALL shards
    size_t dim_ = 0;            // 0 == None (flat.rs:16): the dimension of the whole index
    size_t rows_ = 0;
};

// The multi-vector (MaxSim) collection spread over several GPUs in one process, behind the same vb_mv handle: a
// document lives on shard fnv1a(id) % G, every shard scores its own documents (K5) on its own GPU and thread, and
// the G sorted lists are merged on the calling thread by (score descending, id bytes) — multi_vector.rs:22-31.
class ShardedMvIndex {
  public:
    ShardedMvIndex(int metric, const std::vector<int>& devices);
    ~ShardedMvIndex();
    Status insert_many(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                       const uint64_t* tok_off, const uint64_t* doc_tok);
    Status remove(const char* id, size_t id_len);
    Status search(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, Hits* out);
    void info(size_t* docs, size_t* tokens, size_t* dim);

  private:
    size_t shard_of(const char* id, size_t len) const;
    void for_each_shard(const std::function<void(size_t)>& fn);
    void refresh_totals();

    std::vector<std::unique_ptr<MvIndex>> shards_;
    std::vector<std::unique_ptr<ShardWorker>> workers_;
    std::shared_mutex mu_;
    size_t dim_ = 0, docs_ = 0, tokens_ = 0;   // of the whole collection (dimension 0 == none yet)
};

}  // namespace vb
