// select.h — K7: merge of per-shard (or per-GPU, after an all-gather) sorted top-k lists.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace vb {

// List l: arrays [nq][k_in] (keys ascending) and counts [nq], `l * list_stride` bytes after
// the base pointers; outputs [nq][k_out]: keys, values, (list << 32 | row), counts.
Status topk_merge_device(const u64* d_keys, const float* d_values, const uint32_t* d_rows, const uint32_t* d_counts,
                         size_t list_stride, size_t nq, size_t lists, size_t k_in, size_t k_out, u64* d_keys_out, float* d_values_out,
                         u64* d_rows_out, uint32_t* d_counts_out, cudaStream_t stream);

struct TopkWorkspace;
// Merge tree for a scan whose CTAs only published their per-CTA lists (ws.defer_merge): groups of
// lists are merged by independent CTAs, level by level, and the last level writes ws.out_* (and
// snapshots / re-arms the error word and the grid threshold). `scratch` holds the intermediate lists.
struct DeviceBuf;
Status run_merge_tree(const TopkWorkspace& ws, uint32_t nq, uint32_t lists, DeviceBuf& scratch, cudaStream_t stream);
// True when the per-CTA lists of a scan are better merged by the tree than by the last CTA. The last CTA walks
// lists * k slots alone while every other SM idles: at 148 lists it cost 85 us for k = 50-100 against ~40 us through
// the tree (1M x 768, K1: 0.537 / 0.548 ms -> 0.498 / 0.503 ms); k = 10 (1 480 slots) stays in the kernel.
// (VB_MERGE_TREE_MIN overrides the lists * k threshold for tuning runs.)
inline bool merge_tree_wanted(size_t lists, size_t k) {
    static const size_t limit = [] {
        const char* e = std::getenv("VB_MERGE_TREE_MIN");
        return e && *e ? (size_t)std::atol(e) : (size_t)4096;
    }();
    return lists * k > limit;
}

}  // namespace vb
