// runtime.h — small host runtime shared by the resident indexes and the by-value calls:
// growable device / pinned buffers, per-call search contexts (stream + workspace) drawn
// from a pool so concurrent searches (many BEAM dirty schedulers, nifs.rs:297-309 read
// lock) never share scratch memory, and the hit list returned over the C ABI.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace vb {

struct DeviceBuf {
    void* p = nullptr;
    size_t cap = 0;
    Status reserve(size_t bytes);   // contents are NOT preserved on growth
    void release();
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    Status reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// Everything one in-flight search needs. Control words are armed once and re-armed by the
// kernels themselves, so a search is H2D(query) -> 1 launch -> D2H(result block).
struct SearchCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DeviceBuf queries, q_norms, cand_keys, cand_pays, cand_counts, ctrl, out_keys, result, row_sel, row_sel2;
    DeviceBuf staging, staging_rank, dump_keys, dump_pays, dump_keys2, dump_pays2, sort_tmp, hist;
    DeviceBuf misc;                 // a few scratch words (overflow flag of a dump scan, owned-candidate count)
    PinnedBuf h_queries, h_result, h_misc;
    uint32_t ctrl_queries = 0;      // query slots armed in `ctrl`

    // ctrl layout for nq slots: g_thresh[nq] u64 | done[nq] u32 | err_row[nq] u32
    u64* g_thresh() const { return ctrl.as<u64>(); }
    uint32_t* done() const { return reinterpret_cast<uint32_t*>(ctrl.as<u64>() + ctrl_queries); }
    uint32_t* err_row() const { return done() + ctrl_queries; }

    Status arm_ctrl(uint32_t nq);   // (re)allocates and initialises control words
    void poison() { ctrl_queries = 0; }
    void destroy();
};

class CtxPool {
  public:
    ~CtxPool();
    Status acquire(SearchCtx** out);
    void release(SearchCtx* ctx);
  private:
    std::mutex mu_;
    std::vector<SearchCtx*> free_;
};

CtxPool& ctx_pool();

// Raises a kernel's dynamic shared-memory limit once per (device, kernel, size): function attributes are
// per device, so a process-wide `static bool` would leave the second GPU of a process unconfigured.
Status ensure_dynamic_smem(const void* kernel, size_t bytes);
template <typename K>
inline Status ensure_dynamic_smem_for(K kernel, size_t bytes) { return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes); }

struct CtxLease {
    SearchCtx* ctx = nullptr;
    ~CtxLease() { if (ctx) ctx_pool().release(ctx); }
    Status get() { return ctx_pool().acquire(&ctx); }
    SearchCtx* operator->() const { return ctx; }
};

// Sorted Vec<(String, f32)> handed to the caller (opaque vb_hits in the C ABI).
// Ids are one blob plus n+1 offsets so a binding can take the whole result in one read.
struct Hits {
    std::string blob;
    std::vector<uint64_t> off{0};
    std::vector<float> values;
    std::vector<uint64_t> index;
    size_t size() const { return values.size(); }
    void add(const char* id, size_t len, float value, uint64_t idx) {
        blob.append(id, len);
        off.push_back(blob.size());
        values.push_back(value);
        index.push_back(idx);
    }
};

}  // namespace vb

struct vb_hits : vb::Hits {};
