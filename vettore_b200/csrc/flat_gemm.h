// flat_gemm.h — K2: batched exact flat scan on tcgen05 (3xTF32) with fused per-query top-k.
#pragma once
#include <vector>

#include "runtime.h"

namespace vb {

struct GemmResult {
    size_t k = 0;
    int terms = 3;                    // TF32 passes the filter ran with (1 or 3)
    bool non_finite = false;          // a tensor-core score overflowed: redo the whole batch on the K1 path
    std::vector<uint8_t> flags;       // [nq] 0 ok, 1 redo this query on the K1 path, 2 metric overflow
    std::vector<uint32_t> counts;     // [nq]
    std::vector<uint32_t> rows;       // [nq][k] device rows
    std::vector<float> raws;          // [nq][k]
};

// Dot-product family (flat cosine == dot of stored vectors, inner product, negative inner product) and the L2 family
// (L2, L2 squared, as |x|^2 - 2 q.x with a row-norm mirror),
// dims a multiple of 32 with no row padding, k <= 128, batches of >= 16 queries (VB_FLAT_GEMM_MIN_BATCH).
bool flat_gemm_eligible(int metric, size_t dims, size_t stride, size_t nq, size_t k, size_t n);

// max_row_norm: upper bound of |row| over the index (flat_gemm_max_row_norm), used by the
// completeness check of the exact re-scoring stage.
// d_row_norm2: [n] |row|^2 (the row-norm mirror; required for the L2 family, ignored otherwise).
Status flat_gemm_search(SearchCtx& ctx, int metric, const float* d_rows, size_t stride, const uint32_t* d_id_rank,
                        size_t n, size_t dims, float max_row_norm, const float* d_row_norm2, const float* h_queries, size_t nq,
                        size_t k, GemmResult* out, int force_terms = 0);
Status flat_gemm_search_device(SearchCtx& ctx, int metric, const float* d_rows, size_t stride,
                               const uint32_t* d_id_rank, size_t n, size_t dims, float max_row_norm,
                               const float* d_row_norm2, const float* d_queries, size_t nq, size_t k, u64* d_out_keys, u64* d_out_pays,
                               uint32_t* d_out_counts, uint32_t* d_out_flags, uint32_t* d_bad, cudaStream_t stream,
                               int force_terms = 0, int* terms_used = nullptr);
// A single-pass batch in which this many queries could not be proven complete is redone as ONE 3xTF32 batch
// (dense score distributions — clustered real embeddings — put many rows within the single-pass error bound of the
// k-th score) instead of one single-query scan per flagged query; what is still flagged after that goes to K1.
constexpr size_t kGemmRedoAsBatch = 16;
// max |row| over the index; when d_norm2_out is given, also |row|^2 per row (the L2 family's row-norm mirror).
Status flat_gemm_max_row_norm(SearchCtx& ctx, const float* d_rows, size_t stride, size_t n, size_t dims, float* out,
                              float* d_norm2_out = nullptr);

}  // namespace vb
