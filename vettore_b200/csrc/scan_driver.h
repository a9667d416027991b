// scan_driver.h — host driver for one scan + top-k job: stages host queries, plans and
// launches the K1/K4 kernel (or the dump + radix-sort path for k beyond the fused
// collector), and brings the sorted (row, raw) pairs back.
#pragma once
#include <vector>

#include "runtime.h"

namespace vb {

struct ScanJob {
    int metric = 0;                     // kernel metric (kCosineTrue for vector_top_k cosine)
    const float* d_rows = nullptr;      // device matrix
    size_t row_stride = 0;              // floats
    const uint32_t* d_row_sel = nullptr;
    const uint32_t* d_id_rank = nullptr;
    uint32_t n = 0;                     // logical rows
    uint32_t dims = 0;                  // scored prefix
    bool whole_rows = false;            // dims covers every non-zero element of a row (no data beyond it)
    const float* h_queries = nullptr;   // host, nq rows of q_len floats (q_len >= dims)
    uint32_t nq = 1;
    size_t q_len = 0;
    size_t k = 0;                       // results per query (1..n)
};

struct ScanResult {
    size_t k = 0;
    std::vector<uint32_t> counts;       // [nq]
    std::vector<uint32_t> rows;         // [nq][k] device rows
    std::vector<float> raws;            // [nq][k]
    std::vector<uint32_t> err_rows;     // [nq] kNoError or the first overflowing logical row
};

Status run_scan(SearchCtx& ctx, const ScanJob& job, ScanResult* out);

// Device-resident variant: queries already on the device, sorted results stay on the device.
Status run_scan_device(SearchCtx& ctx, const ScanJob& job, const float* d_queries, size_t q_stride,
                       const double* d_q_norms, u64* d_keys, float* d_values, uint32_t* d_rows,
                       uint32_t* d_counts, cudaStream_t stream);

}  // namespace vb
