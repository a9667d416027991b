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

// Pipeline stage: same scan, but the k = min(job.k, job.n) winning device rows stay on the
// device (ctx.row_sel[0..k)) to drive the next stage as its row list; the stage's overflow
// word is copied asynchronously into *h_err (pinned). No host synchronisation.
Status run_scan_to_rows(SearchCtx& ctx, const ScanJob& job, uint32_t slot, uint32_t nslots, uint32_t* h_err);
// Last stage: results come back to the host (one synchronisation for the whole pipeline).
Status run_scan_final(SearchCtx& ctx, const ScanJob& job, uint32_t slot, uint32_t nslots, ScanResult* out);

// Copies the rows (low 32 bits) of `n` payloads into ctx.row_sel on the stream.
Status extract_rows(SearchCtx& ctx, const u64* d_pays, uint32_t n);

// Device-resident variant: queries already on the device, sorted results stay on the device.
// `d_status` (optional device word): bit 0 is raised when a query's scan hit an unrecoverable overflow
// (the reference's "metric overflow"), since nothing can be returned to the host from here.
Status run_scan_device(SearchCtx& ctx, const ScanJob& job, const float* d_queries, size_t q_stride,
                       const double* d_q_norms, u64* d_keys, float* d_values, uint32_t* d_rows,
                       uint32_t* d_counts, uint32_t* d_status, cudaStream_t stream);

// (key, payload) pairs -> separate keys / values / rows arrays (device, on `stream`).
// Sorts the n (key, payload) pairs a dump-mode kernel left in ctx.dump_keys / ctx.dump_pays and fetches the best k
// payloads + the error word into ctx.h_result (fused-collector result layout). Synchronises ctx.stream.
Status sort_dump_and_fetch(SearchCtx& ctx, size_t n, size_t k);
Status unpack_device_results(const u64* d_keys_in, const u64* d_pays, const uint32_t* d_counts_in, uint32_t nq, uint32_t k,
                             u64* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream);

// The first k entries of a sorted (key, payload) dump -> keys / values / rows (+ count = k).
Status unpack_sorted_device(const u64* d_keys_in, const u64* d_pays, uint32_t k, u64* d_keys, float* d_values,
                            uint32_t* d_rows, uint32_t* d_count, cudaStream_t stream);

}  // namespace vb
