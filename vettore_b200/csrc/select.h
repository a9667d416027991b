// select.h — K7: merge of per-shard (or per-GPU, after an all-gather) sorted top-k lists.
#pragma once
#include "common.cuh"

namespace vb {

// List l: arrays [nq][k_in] (keys ascending) and counts [nq], `l * list_stride` bytes after
// the base pointers; outputs [nq][k_out]: keys, values, (list << 32 | row), counts.
Status topk_merge_device(const u64* d_keys, const float* d_values, const uint32_t* d_rows, const uint32_t* d_counts,
                         size_t list_stride, size_t nq, size_t lists, size_t k_in, size_t k_out, u64* d_keys_out, float* d_values_out,
                         u64* d_rows_out, uint32_t* d_counts_out, cudaStream_t stream);

}  // namespace vb
