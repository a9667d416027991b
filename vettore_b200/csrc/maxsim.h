// maxsim.h — K5: MaxSim scoring + top-k over a device-resident token matrix
// (reference multi_vector.rs:65-132), and the HBM-resident multi-vector index built on it.
#pragma once
#include <map>
#include <shared_mutex>
#include <string>
#include <vector>

#include "runtime.h"

namespace vb {

struct MaxSimJob {
    int metric = 0;                        // kernel metric (kCosineTrue when the NIF metric is cosine)
    const float* d_tokens = nullptr;       // [ntok, stride] device
    size_t stride = 0;                     // floats, multiple of 4
    const uint32_t* d_doc_off = nullptr;   // [ndocs + 1]
    const uint32_t* d_doc_rank = nullptr;  // [ndocs] id ranks (0xFFFFFFFF = deleted) or null (= doc index)
    size_t ndocs = 0;
    uint32_t dims = 0;
    const float* h_query = nullptr;        // host [tq, dims]
    uint32_t tq = 0;
    size_t k = 0;
    uint32_t uniform_td = 0;               // every document slot has exactly this many tokens (0 = ragged)
    const float* d_inv_dnorm = nullptr;    // [ntok] 1/|token| (device), enables the tensor-core cosine path
    // ragged tensor-core path (maxsim_tcr.cu): per-token owning document and what bounds its collector cadence
    const uint32_t* d_tok_doc = nullptr;   // [ntok] or null
    size_t ntok = 0;                       // rows of d_tokens in use
    uint32_t min_td = 0;                   // no non-empty document is shorter than this (0 = unknown)
    bool has_empty = false;                // some document has no token at all
    // optional: the sorted list is ALSO left on the device in the vb_flat_search_device convention
    // (keys = order key << 32 | doc rank, scores, doc slots, count) for the sharded merge
    u64* d_keys_out = nullptr;
    float* d_values_out = nullptr;
    uint32_t* d_rows_out = nullptr;
    uint32_t* d_counts_out = nullptr;
};

struct MaxSimResult {
    std::vector<uint32_t> rows;   // document indices, best first
    std::vector<float> scores;
    uint32_t err = 0xFFFFFFFFu;   // kNoError or (doc << 1 | kind): 0 metric overflow, 1 score overflow
};

Status maxsim_top_k(SearchCtx& ctx, const MaxSimJob& job, MaxSimResult* out);

// Tensor-core path (maxsim_tc.cu): inner-product family, uniform documents of 32/64/128
// tokens, dims a multiple of 32 up to 128, at most 32 query tokens.
bool maxsim_tc_eligible(const MaxSimJob& job, uint32_t uniform_td);
Status maxsim_tc_top_k(SearchCtx& ctx, const MaxSimJob& job, uint32_t td, const float* d_inv_dnorm, MaxSimResult* out);
// Tensor-core path for ragged documents (maxsim_tcr.cu): inner-product family, any document lengths, dims <= 128,
// at most 64 query tokens; needs d_tok_doc.
bool maxsim_tcr_eligible(const MaxSimJob& job);
Status maxsim_tcr_top_k(SearchCtx& ctx, const MaxSimJob& job, MaxSimResult* out);
Status maxsim_prepare_dump(SearchCtx& ctx, size_t ndocs);
Status maxsim_collect_dump(SearchCtx& ctx, size_t ndocs, uint32_t k, cudaError_t launch, MaxSimResult* out);
struct TopkWorkspace;
// Common tail of the three launchers: optional device-side unpack, D2H of the sorted list, stream sync, decode.
Status maxsim_collect_result(SearchCtx& ctx, const MaxSimJob& job, const TopkWorkspace& ws, uint32_t k, cudaError_t launch,
                             MaxSimResult* out);

// HBM-resident multi-vector collection: token matrix + document offsets + id ranks.
// Upserts append a fresh copy and tombstone the old one; the matrix is compacted when
// tombstones outweigh live tokens.
class MvIndex {
  public:
    MvIndex(int metric, int device) : metric_(metric), device_(device) {}
    ~MvIndex();
    // Documents: doc i owns tokens [doc_tok[i], doc_tok[i+1]) of the ragged token list.
    Status insert_many(size_t ndocs, const char* ids, const uint64_t* id_off, const float* tok_vals,
                       const uint64_t* tok_off, const uint64_t* doc_tok);
    Status remove(const char* id, size_t id_len);
    // Pre-sizes the HBM arrays (no realloc + copy while a large corpus is streamed in).
    Status reserve(size_t docs, size_t tokens, size_t dim);
    // Uniform documents whose tokens are already in device memory: [ndocs * td, dim] fp32.
    // With doc_tok (host, [ndocs + 1]) the batch is ragged: document i owns rows [doc_tok[i], doc_tok[i + 1]).
    Status insert_many_device(size_t ndocs, const char* ids, const uint64_t* id_off, const float* d_tokens,
                              size_t td, size_t dim, const uint64_t* doc_tok = nullptr);
    Status search(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, Hits* out);
    // Document-sharded search: the shard's sorted top-k also stays on the device (see MaxSimJob).
    Status search_packed_device(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, u64* d_keys,
                                float* d_values, uint32_t* d_rows, uint32_t* d_counts, Hits* out);
    // Overrides the id tie-break ranks (per document slot) so they compare across shards.
    Status set_id_ranks(const uint32_t* ranks, size_t n);
    void info(size_t* docs, size_t* tokens, size_t* dim);

  private:
    Status reserve_tokens(size_t need);
    Status reserve_docs(size_t need);
    Status relabel();
    Status compact();
    void note_doc_length(uint32_t tokens);
    Status search_impl(const float* q_vals, const uint64_t* q_off, size_t tq, size_t limit, u64* d_keys,
                       float* d_values, uint32_t* d_rows, uint32_t* d_counts, Hits* out);

    const int metric_;
    const int device_;
    std::shared_mutex mu_;
    size_t dim_ = 0, stride_ = 0;
    size_t ntok_ = 0, tok_cap_ = 0, dead_tok_ = 0;
    size_t ndocs_ = 0, doc_cap_ = 0;       // doc slots used (live + tombstoned)
    float* d_tokens_ = nullptr;
    float* d_inv_norm_ = nullptr;          // [tok_cap] 1/|token| in f32 (0 for zero tokens)
    uint32_t* d_tok_doc_ = nullptr;        // [tok_cap] owning document slot of every token (ragged tensor-core kernel)
    uint32_t min_td_ = 0;                  // shortest non-empty document ever inserted (0 = none yet)
    bool has_empty_ = false;               // some slot holds a document without tokens
    uint32_t uniform_td_ = 0;              // tokens per document while all slots agree, else 0
    bool uniform_known_ = false;
    uint32_t* d_doc_off_ = nullptr;        // [doc_cap + 1]
    uint32_t* d_doc_rank_ = nullptr;       // [doc_cap]
    std::vector<uint32_t> h_doc_off_{0};
    std::vector<uint32_t> h_rank_;
    std::vector<std::string> doc_id_;      // slot -> id ("" for tombstones)
    std::map<std::string, uint32_t> id_doc_;
};

}  // namespace vb

namespace vb { class ShardedMvIndex; }

// The C ABI handle: a single-GPU collection (impl) or a multi-GPU one in the same process (sharded).
struct vb_mv {
    vb::MvIndex* impl;
    vb::ShardedMvIndex* sharded;
};
