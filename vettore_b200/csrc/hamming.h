// hamming.h — K3: packed sign-code Hamming scan + fused top-k (reference search.rs:76-92,
// distances.rs:426-437) and K6: sign-bit packing of an fp32 matrix (distances.rs:413-423).
#pragma once
#include <vector>

#include "runtime.h"

namespace vb {

// Scans a device-resident [n, nw] u64 code matrix against one query per slot and keeps the
// best k (distance ascending, then id rank). Results land in ctx.result (pays|counts) and
// ctx.out_keys like the float scans. k <= kMaxFusedK.
Status hamming_scan_device(SearchCtx& ctx, const u64* d_codes, uint32_t n, uint32_t nw, uint32_t dims,
                           const uint32_t* d_id_rank, const u64* d_queries, uint32_t nq, uint32_t k,
                           cudaStream_t stream);

// Candidate counts beyond the fused collector: every row's key to HBM, one radix sort; the n sorted
// (key, payload) pairs are left in ctx.dump_keys2 / ctx.dump_pays2. One query, stream-ordered, no host sync.
Status hamming_dump_sorted(SearchCtx& ctx, const u64* d_codes, uint32_t n, uint32_t nw, uint32_t dims,
                           const uint32_t* d_rank, const u64* d_query, cudaStream_t stream);

// By-value helper: uploads host codes/ranks/query, scans, returns sorted (row, distance).
// Any k (beyond the fused collector the scan dumps every key and radix-sorts).
Status hamming_top_k_host(SearchCtx& ctx, const uint64_t* h_codes, size_t n, size_t nw, size_t dims,
                          const uint32_t* h_rank, const uint64_t* h_query, size_t k, std::vector<uint32_t>* rows,
                          std::vector<float>* values);

// Resident variant of the above: codes and ranks already on the device.
Status hamming_top_k_resident(SearchCtx& ctx, const u64* d_codes, size_t n, size_t nw, size_t dims,
                              const uint32_t* d_rank, const uint64_t* h_query, size_t k,
                              std::vector<uint32_t>* rows, std::vector<float>* values);

// K6: codes[r][w] bit b = (rows[r][64 w + b] >= 0.0f). One warp ballot per 32 coordinates.
Status sign_pack_device(const float* d_rows, size_t row_stride, uint32_t n, uint32_t dims, u64* d_codes,
                        cudaStream_t stream);

}  // namespace vb
