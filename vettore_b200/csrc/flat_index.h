// flat_index.h — DeviceFlatIndex: the HBM-resident replacement of the reference's
// FlatIndex (flat.rs:13-129, a HashMap<String, Vec<f32>> walked row by row).
//
// Layout in HBM: one row-major [capacity, stride] fp32 matrix (stride = dimension rounded
// up to 4 floats so every row is 16-byte aligned, padding zero) plus one u32 id-rank per
// row. Host side: an ordered id -> row map (byte-lexicographic, the reference's id.cmp)
// and the row -> id table. The id rank is an order-maintenance label: rank order == id
// byte order, so the device can break rank ties exactly like flat.rs:34-40 without ever
// seeing a string. Deletes move the last row into the hole (ranks stay valid), ascending
// appends take the next label, out-of-order inserts bisect the neighbouring labels and
// fall back to an even relabel when a gap is exhausted.
#pragma once
#include <map>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "runtime.h"
#include "scan_driver.h"

namespace vb {

class FlatIndex {
  public:
    explicit FlatIndex(int metric, int device) : metric_(metric), device_(device) {}
    ~FlatIndex();

    // flat.rs:59-85 (single insert == batch of one). Ragged input: row i is
    // values[value_off[i] .. value_off[i+1]).
    Status insert_many(size_t n, const char* ids, const uint64_t* id_off, const float* values,
                       const uint64_t* value_off, bool single);
    // Capacity hint: the next growth allocates room for `rows` rows at once (no realloc + copy while a
    // large corpus streams in; doubling would need old + new matrix resident together).
    Status reserve(size_t rows);
    // Bulk ingest of rows that already sit in device memory: [n, dim] fp32, contiguous (rebuild_index
    // from a device-side snapshot, device-generated corpora). Same validation and upsert rules.
    Status insert_many_device(size_t n, const char* ids, const uint64_t* id_off, const float* d_values, size_t dim);
    Status remove(const char* id, size_t id_len);                                    // flat.rs:88-93
    Status search(const float* queries, size_t nq, size_t len, size_t limit, std::vector<Hits>* out);  // flat.rs:96-124
    // search.rs:38-73 over resident rows (all rows, or the listed ids).
    Status prefix_top_k(bool all_rows, size_t n_ids, const char* ids, const uint64_t* id_off, const float* query,
                        size_t len, int metric_code, size_t dimensions, size_t limit, Hits* out);
    // funnel_search (collection.ex:244-260) on the resident matrix: every stage is a
    // vector_top_k over the survivors of the previous one, then the exact rerank; one host
    // synchronisation for the whole pipeline.
    Status funnel_search(const float* query, size_t len, int metric_code, const size_t* stages, size_t nstages,
                         size_t candidates, size_t limit, Hits* out);
    // quantized_search (collection.ex:266-295): sign-code Hamming candidates over the code
    // mirror of the rows (K3), then the exact rerank of those rows (K4).
    Status quantized_search(const float* query, size_t len, int metric_code, size_t candidates, size_t limit,
                            Hits* out);
    // Stage 1 of quantized_search on its own (search.rs:76-92 over the resident code mirror): the best
    // `candidates` rows by (Hamming distance of the sign codes, id), distances as f32. The multi-GPU handle
    // merges the shards' candidate lists before the owners rerank (sharded_index.cu).
    Status hamming_candidates(const float* query, size_t len, size_t candidates, Hits* out);
    Status search_device(const float* d_queries, size_t nq, size_t q_stride, size_t limit, u64* d_keys,
                         float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream);
    // Row-sharded quantized_search, stage 1: sign-packs the device queries and scans this shard's
    // code mirror for the best `candidates` (same output convention as search_device).
    Status hamming_device(const float* d_queries, size_t nq, size_t q_stride, size_t candidates, u64* d_keys,
                          float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream);
    // Stage 2: exact rerank (vector_top_k semantics at full length) of those of the globally selected
    // candidates (`shard << 32 | row`, *d_global_count of them) that live on `shard`. One stream
    // synchronisation (the number of owned candidates sizes the launch).
    Status rerank_owned_device(const float* d_query, size_t q_stride, int metric_code, const u64* d_global_rows,
                               const uint32_t* d_global_count, size_t max_candidates, uint32_t shard, size_t limit,
                               u64* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts,
                               cudaStream_t stream);
    Status set_id_ranks(const uint32_t* ranks, size_t n);
    void info(size_t* rows, size_t* dim);
    // Sticky status of the stream-ordered device-level entries since the last call (they cannot return
    // the reference's error strings): bit 0 = a scan met an unrecoverable overflow ("metric overflow").
    // Synchronises the device; clears the word.
    Status device_status(uint32_t* out);

    // Resident matrix view for sibling indexes (valid under the caller's own lock discipline).
    const float* device_rows() const { return d_rows_; }
    size_t stride() const { return stride_; }

  private:
    Status grow(size_t need_rows);
    Status ensure_codes();                       // builds the sign-code mirror on first use (K6)
    Status pack_rows(size_t row0, size_t rows);  // refreshes the mirrors (sign codes, dense prefix) of rows [row0, row0 + rows)
    // Dense mirror of the first `dims` columns (funnel stage 1, search.rs:51-70 over every row): a row-major
    // [cap, round4(dims)] matrix kept in sync like the sign codes, so the stage is a whole-row stream through
    // the TMA ring instead of 4*dims-byte pieces at a 4*stride-byte pitch (DRAM page locality).
    Status ensure_prefix(size_t dims);
    bool prefix_wanted(size_t dims) const;
    Status relabel_all();
    Status assign_rank(std::map<std::string, uint32_t>::iterator it, uint32_t row, bool* relabel_needed);
    void reset_if_empty();
    // Id table. While every insert so far was an ascending append (rebuild_index feeds ids sorted,
    // collection.ex:426-433) the rows themselves are in id order: lookups are a binary search over
    // row_id_ and no ordered map exists (`sorted_`). The first out-of-order insert or hole-filling
    // delete materialises id_row_ once (O(n)) and the general path takes over.
    int64_t find_row(const std::string& id) const;
    void leave_sorted_mode();
    // Registers a new id at row n_ (bookkeeping only); false = the id exists (row in *existing).
    bool add_id(std::string&& id, uint32_t* existing, bool* relabel_needed);
    Status ensure_dev_ctx();
    Status finish_mutation();   // orders the mutation's default-stream work before the lock is released

    const int metric_;
    const int device_;
    std::shared_mutex mu_;
    size_t dim_ = 0;        // 0 == None (flat.rs:16)
    size_t stride_ = 0;     // floats per device row
    size_t n_ = 0, cap_ = 0, reserve_hint_ = 0;
    float* d_rows_ = nullptr;
    uint32_t* d_rank_ = nullptr;
    u64* d_codes_ = nullptr;   // [cap, code_words_] sign codes of the rows, kept in sync once built
    size_t code_words_ = 0;
    float* d_prefix_ = nullptr;   // [cap, prefix_stride_] first prefix_dims_ columns of the rows, kept in sync once built
    size_t prefix_dims_ = 0, prefix_stride_ = 0;
    std::vector<std::string> row_id_;
    std::vector<uint32_t> h_rank_;
    std::map<std::string, uint32_t> id_row_;   // general mode only (see sorted_)
    bool sorted_ = true;
    bool external_ranks_ = false;
    std::mutex norm_mu_;
    float max_norm_ = -1.0f;   // max |row| (device reduction, lazily), < 0 = unknown
    float* d_norm2_ = nullptr; // [norm2_cap_] |row|^2 (L2-family indexes: batched searches), valid while max_norm_ >= 0
    size_t norm2_cap_ = 0;
    Status ensure_norms(SearchCtx& ctx);   // (re)computes max_norm_ (and the row-norm mirror for the L2 family)
    SearchCtx* dev_ctx_ = nullptr;  // workspace of the stream-ordered device-level entry
    uint32_t* d_status_ = nullptr;  // sticky status word of the device-level entries
};

}  // namespace vb

namespace vb { class ShardedFlatIndex; }

// The C ABI handle: a single-GPU index (impl) or a multi-GPU one in the same process (sharded).
struct vb_flat {
    vb::FlatIndex* impl;
    vb::ShardedFlatIndex* sharded;
};
