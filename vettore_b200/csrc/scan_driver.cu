// scan_driver.cu — see scan_driver.h.
#include "scan_driver.h"

#include <cmath>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "flat_scan.cuh"
#include "flat_scan.h"
#include "select.h"

namespace vb {

static int scan_layout(const ScanJob& job) {
    if (job.d_row_sel != nullptr) return kScanRowList;
    return job.whole_rows ? kScanWholeRows : kScanPrefixAllRows;
}

static Status prepare_workspace(SearchCtx& ctx, const ScanPlan& plan, uint32_t nq, size_t k) {
    VB_TRY(ctx.arm_ctrl(nq));
    const size_t lists = (size_t)nq * plan.grid_x;
    VB_TRY(ctx.cand_keys.reserve(lists * k * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve(lists * k * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve(lists * sizeof(uint32_t)));
    VB_TRY(ctx.out_keys.reserve((size_t)nq * k * sizeof(u64)));
    // result block: pays[nq][k] u64 | counts[nq] u32 | err[nq] u32
    VB_TRY(ctx.result.reserve((size_t)nq * k * sizeof(u64) + (size_t)nq * 8));
    return Status::Ok();
}

static void fill_params(const SearchCtx& ctx, const ScanJob& job, const float* d_queries, size_t q_stride,
                        const double* d_q_norms, size_t k, ScanParams* p) {
    p->rows = job.d_rows;
    p->row_stride = job.row_stride;
    p->row_sel = job.d_row_sel;
    p->id_rank = job.d_id_rank;
    p->n = job.n;
    p->dims = job.dims;
    p->queries = d_queries;
    p->q_stride = (uint32_t)q_stride;
    p->q_norms = d_q_norms;
    p->cap = 0;
    { const char* dbg = std::getenv("VB_SCAN_DEBUG"); p->debug = dbg ? (uint32_t)std::atoi(dbg) : 0u; }
    p->ws.k = (uint32_t)k;
    p->ws.cand_keys = ctx.cand_keys.as<u64>();
    p->ws.cand_pays = ctx.cand_pays.as<u64>();
    p->ws.cand_counts = ctx.cand_counts.as<uint32_t>();
    p->ws.done = ctx.done();
    p->ws.g_thresh = ctx.g_thresh();
    p->err_row = ctx.err_row();
    p->ws.out_keys = ctx.out_keys.as<u64>();
    p->ws.out_pays = ctx.result.as<u64>();
    p->ws.out_counts = reinterpret_cast<uint32_t*>(ctx.result.as<u64>() + (size_t)job.nq * k);
    p->ws.err_row = ctx.err_row();
    p->ws.out_err = p->ws.out_counts + job.nq;
    p->ws.defer_merge = 0;
    p->ws.piv_keys = nullptr;
    p->ws.piv_counts = nullptr;
    p->ws.piv_state = nullptr;
    p->dump_keys = nullptr;
    p->dump_pays = nullptr;
}

// Launch-wide pivot ladder for large k (topk.cuh, collector_pivot_step): per query slot 16 keys, 16 counters
// and a state word, zeroed on the launch stream before every scan that uses it.
constexpr uint32_t kPivotMinK = 32;
static Status arm_pivots(SearchCtx& ctx, ScanParams* p, uint32_t nq, size_t k, cudaStream_t stream) {
    p->ws.piv_keys = nullptr;
    p->ws.piv_counts = nullptr;
    p->ws.piv_state = nullptr;
    if (k < kPivotMinK || std::getenv("VB_NO_PIVOTS")) return Status::Ok();
    const size_t keys_b = (size_t)nq * kPivots * sizeof(u64), cnt_b = (size_t)nq * kPivots * sizeof(uint32_t);
    const size_t bytes = keys_b + cnt_b + (size_t)nq * sizeof(uint32_t);
    VB_TRY(ctx.hist.reserve(bytes));
    VB_CUDA(cudaMemsetAsync(ctx.hist.p, 0, bytes, stream));
    unsigned char* base = ctx.hist.as<unsigned char>();
    p->ws.piv_keys = reinterpret_cast<u64*>(base);
    p->ws.piv_counts = reinterpret_cast<uint32_t*>(base + keys_b);
    p->ws.piv_state = reinterpret_cast<uint32_t*>(base + keys_b + cnt_b);
    return Status::Ok();
}

// Copies err_row into the result block and re-arms it (one thread per query).
__global__ void collect_err_kernel(uint32_t* err_row, uint32_t* out_err, uint32_t nq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { out_err[i] = err_row[i]; err_row[i] = kNoError; }
}

// `err` (optional): per-query overflow words of the scan; any set word raises bit 0 of *status (sticky,
// read by vb_flat_device_status: the device-level entries cannot return "metric overflow" themselves).
__global__ void unpack_results_kernel(const u64* keys, const u64* pays, const uint32_t* counts, uint32_t nq,
                                      uint32_t k, u64* out_keys, float* out_values, uint32_t* out_rows,
                                      uint32_t* out_counts, const uint32_t* err, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && err != nullptr && status != nullptr && err[i] != kNoError) *(volatile uint32_t*)status = 1u;
    if (i < nq * k) {
        uint32_t q = i / k, j = i - q * k;
        bool valid = j < counts[q];
        u64 pay = valid ? pays[i] : 0;
        if (out_keys) out_keys[i] = valid ? keys[i] : kKeyMax;
        if (out_values) out_values[i] = __uint_as_float((uint32_t)(pay >> 32));
        if (out_rows) out_rows[i] = (uint32_t)pay;
    }
    if (i < nq && out_counts) out_counts[i] = counts[i];
}

static Status stage_queries(SearchCtx& ctx, const ScanJob& job, size_t* q_stride_out) {
    const size_t q_stride = ((size_t)job.dims + 3) & ~(size_t)3;
    const bool need_norm = job.metric == kCosineTrue;
    VB_TRY(ctx.h_queries.reserve(job.nq * q_stride * sizeof(float) + job.nq * sizeof(double)));
    VB_TRY(ctx.queries.reserve(job.nq * q_stride * sizeof(float)));
    float* hq = ctx.h_queries.as<float>();
    double* hn = reinterpret_cast<double*>(hq + job.nq * q_stride);
    for (uint32_t q = 0; q < job.nq; ++q) {
        const float* src = job.h_queries + (size_t)q * job.q_len;
        float* dst = hq + (size_t)q * q_stride;
        std::memcpy(dst, src, job.dims * sizeof(float));
        for (size_t i = job.dims; i < q_stride; ++i) dst[i] = 0.0f;
        if (need_norm) {  // reference distances.rs:165: f64_dot(left, left).sqrt()
            double s = 0.0;
            for (uint32_t i = 0; i < job.dims; ++i) s += (double)src[i] * (double)src[i];
            hn[q] = std::sqrt(s);
        }
    }
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, hq, job.nq * q_stride * sizeof(float), cudaMemcpyHostToDevice,
                            ctx.stream));
    if (need_norm) {
        VB_TRY(ctx.q_norms.reserve(job.nq * sizeof(double)));
        VB_CUDA(cudaMemcpyAsync(ctx.q_norms.p, hn, job.nq * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    *q_stride_out = q_stride;
    return Status::Ok();
}

// k beyond the fused collector: every key/payload of ONE query to HBM, then a device radix sort.
// Leaves the n sorted (key, payload) pairs in ctx.dump_keys2 / ctx.dump_pays2 and the query's
// overflow word in *d_err_dst (device memory; the control word is re-armed). Stream-ordered, no
// host synchronisation.
static Status dump_and_sort(SearchCtx& ctx, const ScanJob& job, const float* d_q, size_t q_stride,
                            const double* d_norm, uint32_t* d_err_dst, cudaStream_t stream) {
    ScanPlan plan;
    VB_TRY(plan_flat_scan(job.metric, job.dims, job.row_stride, scan_layout(job), job.n, 1, /*dump=*/true, &plan));
    cudaStream_t saved = ctx.stream;
    ctx.stream = stream;   // workspace arming must be ordered on the launch stream
    Status ws = prepare_workspace(ctx, plan, 1, 1);
    ctx.stream = saved;
    VB_TRY(ws);
    const size_t n = job.n;
    VB_TRY(ctx.dump_keys.reserve(n * sizeof(u64)));
    VB_TRY(ctx.dump_pays.reserve(n * sizeof(u64)));
    VB_TRY(ctx.dump_keys2.reserve(n * sizeof(u64)));
    VB_TRY(ctx.dump_pays2.reserve(n * sizeof(u64)));
    size_t tmp_bytes = 0;
    VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx.dump_keys.as<u64>(), ctx.dump_keys2.as<u64>(),
                                            ctx.dump_pays.as<u64>(), ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64,
                                            stream));
    VB_TRY(ctx.sort_tmp.reserve(tmp_bytes));
    ScanJob one = job;
    one.nq = 1;
    ScanParams p;
    fill_params(ctx, one, d_q, q_stride, d_norm, 1, &p);
    p.dump_keys = ctx.dump_keys.as<u64>();
    p.dump_pays = ctx.dump_pays.as<u64>();
    VB_TRY(run_flat_scan(plan, p, 1, stream));
    VB_CUDA(cub::DeviceRadixSort::SortPairs(ctx.sort_tmp.p, tmp_bytes, ctx.dump_keys.as<u64>(),
                                            ctx.dump_keys2.as<u64>(), ctx.dump_pays.as<u64>(),
                                            ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64, stream));
    collect_err_kernel<<<1, 32, 0, stream>>>(ctx.err_row(), d_err_dst, 1);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

// Dump mode of kernels outside this file (MaxSim with a limit beyond the fused collector): sorts the n pairs a
// kernel left in ctx.dump_keys / ctx.dump_pays (absent entries hold kKeyMax) and brings the best k payloads to
// ctx.h_result in the fused collector's result layout (k payloads | count | error word). Synchronises the stream.
Status sort_dump_and_fetch(SearchCtx& ctx, size_t n, size_t k) {
    size_t tmp_bytes = 0;
    VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx.dump_keys.as<u64>(), ctx.dump_keys2.as<u64>(),
                                            ctx.dump_pays.as<u64>(), ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64,
                                            ctx.stream));
    VB_TRY(ctx.sort_tmp.reserve(tmp_bytes));
    VB_CUDA(cub::DeviceRadixSort::SortPairs(ctx.sort_tmp.p, tmp_bytes, ctx.dump_keys.as<u64>(), ctx.dump_keys2.as<u64>(),
                                            ctx.dump_pays.as<u64>(), ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64,
                                            ctx.stream));
    VB_TRY(ctx.misc.reserve(17 * sizeof(uint32_t)));
    collect_err_kernel<<<1, 32, 0, ctx.stream>>>(ctx.err_row(), ctx.misc.as<uint32_t>(), 1);
    VB_CUDA(cudaGetLastError());
    VB_TRY(ctx.h_result.reserve(k * sizeof(u64) + 8));
    uint32_t* tail = reinterpret_cast<uint32_t*>(ctx.h_result.as<u64>() + k);
    VB_CUDA(cudaMemcpyAsync(ctx.h_result.p, ctx.dump_pays2.p, k * sizeof(u64), cudaMemcpyDeviceToHost, ctx.stream));
    VB_CUDA(cudaMemcpyAsync(tail + 1, ctx.misc.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    VB_CUDA(cudaStreamSynchronize(ctx.stream));
    tail[0] = (uint32_t)k;
    return Status::Ok();
}

// Scratch word receiving the overflow flag of a dump scan.
static Status dump_err_word(SearchCtx& ctx, uint32_t** out, uint32_t nslots = 1) {
    VB_TRY(ctx.misc.reserve(((size_t)nslots + 16) * sizeof(uint32_t)));   // same size for every stage of a pipeline
    *out = ctx.misc.as<uint32_t>();
    return Status::Ok();
}

static Status run_scan_dump(SearchCtx& ctx, const ScanJob& job, size_t q_stride, ScanResult* out) {
    const size_t k = job.k;
    uint32_t* d_err = nullptr;
    VB_TRY(dump_err_word(ctx, &d_err));
    VB_TRY(ctx.h_result.reserve(k * sizeof(u64) + 8));
    out->k = k;
    out->counts.assign(job.nq, 0);
    out->rows.assign((size_t)job.nq * k, 0);
    out->raws.assign((size_t)job.nq * k, 0.0f);
    out->err_rows.assign(job.nq, kNoError);
    for (uint32_t q = 0; q < job.nq; ++q) {
        VB_TRY(dump_and_sort(ctx, job, ctx.queries.as<float>() + (size_t)q * q_stride, q_stride,
                             job.metric == kCosineTrue ? ctx.q_norms.as<double>() + q : nullptr, d_err, ctx.stream));
        uint32_t* h_err = reinterpret_cast<uint32_t*>(ctx.h_result.as<u64>() + k);
        VB_CUDA(cudaMemcpyAsync(ctx.h_result.p, ctx.dump_pays2.p, k * sizeof(u64), cudaMemcpyDeviceToHost,
                                ctx.stream));
        VB_CUDA(cudaMemcpyAsync(h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
        VB_CUDA(cudaStreamSynchronize(ctx.stream));
        const u64* pays = ctx.h_result.as<u64>();
        out->counts[q] = (uint32_t)k;
        out->err_rows[q] = *h_err;
        for (size_t i = 0; i < k; ++i) {
            uint32_t bits = (uint32_t)(pays[i] >> 32);
            std::memcpy(&out->raws[(size_t)q * k + i], &bits, 4);
            out->rows[(size_t)q * k + i] = (uint32_t)pays[i];
        }
    }
    return Status::Ok();
}

Status run_scan(SearchCtx& ctx, const ScanJob& job, ScanResult* out) {
    if (job.n == 0 || job.k == 0 || job.nq == 0) return Status::Cuda("empty scan job");
    size_t q_stride = 0;
    VB_TRY(stage_queries(ctx, job, &q_stride));
    const size_t k = std::min<size_t>(job.k, job.n);
    if (k > (size_t)kMaxFusedK) {
        ScanJob j2 = job;
        j2.k = k;
        Status s = run_scan_dump(ctx, j2, q_stride, out);
        if (!s.ok()) ctx.poison();
        return s;
    }

    ScanPlan plan;
    VB_TRY(plan_flat_scan(job.metric, job.dims, job.row_stride, scan_layout(job), job.n, (uint32_t)k, false, &plan));
    VB_TRY(prepare_workspace(ctx, plan, job.nq, k));
    ScanParams p;
    fill_params(ctx, job, ctx.queries.as<float>(), q_stride,
                job.metric == kCosineTrue ? ctx.q_norms.as<double>() : nullptr, k, &p);
    p.ws.defer_merge = merge_tree_wanted(plan.grid_x, k) ? 1u : 0u;
    VB_TRY(arm_pivots(ctx, &p, job.nq, k, ctx.stream));
    Status s = run_flat_scan(plan, p, job.nq, ctx.stream);
    if (s.ok() && p.ws.defer_merge) s = run_merge_tree(p.ws, job.nq, plan.grid_x, ctx.sort_tmp, ctx.stream);
    if (!s.ok()) { ctx.poison(); return s; }
    const size_t bytes = (size_t)job.nq * k * sizeof(u64) + (size_t)job.nq * 8;
    VB_TRY(ctx.h_result.reserve(bytes));
    cudaError_t e = cudaMemcpyAsync(ctx.h_result.p, ctx.result.p, bytes, cudaMemcpyDeviceToHost, ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
    if (e != cudaSuccess) {
        ctx.poison();
        return Status::Cuda(cudaGetErrorString(e));
    }
    const u64* pays = ctx.h_result.as<u64>();
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(pays + (size_t)job.nq * k);
    const uint32_t* errs = counts + job.nq;
    out->k = k;
    out->counts.assign(counts, counts + job.nq);
    out->err_rows.assign(errs, errs + job.nq);
    out->rows.resize((size_t)job.nq * k);
    out->raws.resize((size_t)job.nq * k);
    for (size_t i = 0; i < (size_t)job.nq * k; ++i) {
        uint32_t bits = (uint32_t)(pays[i] >> 32);
        std::memcpy(&out->raws[i], &bits, 4);
        out->rows[i] = (uint32_t)pays[i];
    }
    return Status::Ok();
}

__global__ void extract_rows_kernel(const u64* pays, uint32_t n, uint32_t* rows) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rows[i] = (uint32_t)pays[i];
}

Status extract_rows(SearchCtx& ctx, const u64* d_pays, uint32_t n) {
    VB_TRY(ctx.row_sel2.reserve((size_t)n * sizeof(uint32_t)));
    extract_rows_kernel<<<(n + 255) / 256, 256, 0, ctx.stream>>>(d_pays, n, ctx.row_sel2.as<uint32_t>());
    VB_CUDA(cudaGetLastError());
    std::swap(ctx.row_sel, ctx.row_sel2);   // the new list becomes the current one
    return Status::Ok();
}

// Pipeline stages run back to back without host synchronisation, so every stage gets its
// own slot of the pinned / device query buffers (an async H2D copy reads the pinned source
// when it executes, not when it is enqueued). Slot layout: round4(q_len) floats | f64 norm.
static Status stage_query_slot(SearchCtx& ctx, const ScanJob& job, uint32_t slot, uint32_t nslots,
                               const float** d_q, size_t* q_stride, const double** d_norm) {
    const size_t full = ((size_t)job.q_len + 3) & ~(size_t)3;
    const size_t slot_floats = full + 4;
    VB_TRY(ctx.h_queries.reserve(nslots * slot_floats * sizeof(float)));   // no-op after the first stage
    VB_TRY(ctx.queries.reserve(nslots * slot_floats * sizeof(float)));
    float* hq = ctx.h_queries.as<float>() + slot * slot_floats;
    float* dq = ctx.queries.as<float>() + slot * slot_floats;
    const size_t stride = ((size_t)job.dims + 3) & ~(size_t)3;
    std::memcpy(hq, job.h_queries, job.dims * sizeof(float));
    for (size_t i = job.dims; i < full; ++i) hq[i] = 0.0f;
    double s = 0.0;
    for (uint32_t i = 0; i < job.dims; ++i) s += (double)job.h_queries[i] * (double)job.h_queries[i];
    *reinterpret_cast<double*>(hq + full) = std::sqrt(s);
    VB_CUDA(cudaMemcpyAsync(dq, hq, slot_floats * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));
    *d_q = dq;
    *q_stride = stride;
    *d_norm = reinterpret_cast<const double*>(dq + full);
    return Status::Ok();
}

Status run_scan_to_rows(SearchCtx& ctx, const ScanJob& job, uint32_t slot, uint32_t nslots, uint32_t* h_err) {
    if (job.n == 0 || job.k == 0 || job.nq != 1) return Status::Cuda("bad pipeline stage");
    const size_t k = std::min<size_t>(job.k, job.n);
    size_t q_stride = 0;
    const float* d_q = nullptr;
    const double* d_norm = nullptr;
    VB_TRY(stage_query_slot(ctx, job, slot, nslots, &d_q, &q_stride, &d_norm));
    if (k > (size_t)kMaxFusedK) {
        // more survivors than the fused collector holds (collection.ex:509-510: candidates default to
        // 10 x limit, unbounded): score every row, radix-sort, keep the first k rows. Still no host sync.
        uint32_t* d_err = nullptr;
        VB_TRY(dump_err_word(ctx, &d_err, nslots));
        Status s = dump_and_sort(ctx, job, d_q, q_stride, job.metric == kCosineTrue ? d_norm : nullptr, d_err + slot,
                                 ctx.stream);
        if (!s.ok()) { ctx.poison(); return s; }
        VB_CUDA(cudaMemcpyAsync(h_err, d_err + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
        return extract_rows(ctx, ctx.dump_pays2.as<u64>(), (uint32_t)k);
    }
    ScanPlan plan;
    VB_TRY(plan_flat_scan(job.metric, job.dims, job.row_stride, scan_layout(job), job.n,
                          (uint32_t)k, false, &plan));
    VB_TRY(prepare_workspace(ctx, plan, 1, k));
    ScanParams p;
    fill_params(ctx, job, d_q, q_stride, d_norm, k, &p);
    p.ws.defer_merge = merge_tree_wanted(plan.grid_x, k) ? 1u : 0u;
    VB_TRY(arm_pivots(ctx, &p, 1, k, ctx.stream));
    Status s = run_flat_scan(plan, p, 1, ctx.stream);
    if (s.ok() && p.ws.defer_merge) s = run_merge_tree(p.ws, 1, plan.grid_x, ctx.sort_tmp, ctx.stream);
    if (!s.ok()) { ctx.poison(); return s; }
    VB_CUDA(cudaMemcpyAsync(h_err, p.ws.out_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    return extract_rows(ctx, p.ws.out_pays, (uint32_t)k);
}

// Final stage of a pipeline: like run_scan for one query, but staged in its own query slot.
Status run_scan_final(SearchCtx& ctx, const ScanJob& job, uint32_t slot, uint32_t nslots, ScanResult* out) {
    if (job.n == 0 || job.k == 0 || job.nq != 1) return Status::Cuda("bad pipeline stage");
    const size_t k = std::min<size_t>(job.k, job.n);
    size_t q_stride = 0;
    const float* d_q = nullptr;
    const double* d_norm = nullptr;
    VB_TRY(stage_query_slot(ctx, job, slot, nslots, &d_q, &q_stride, &d_norm));
    if (k > (size_t)kMaxFusedK) {
        uint32_t* d_err = nullptr;
        VB_TRY(dump_err_word(ctx, &d_err, nslots));
        Status s = dump_and_sort(ctx, job, d_q, q_stride, job.metric == kCosineTrue ? d_norm : nullptr, d_err + slot,
                                 ctx.stream);
        if (!s.ok()) { ctx.poison(); return s; }
        VB_TRY(ctx.h_result.reserve(k * sizeof(u64) + 8));
        uint32_t* h_tail = reinterpret_cast<uint32_t*>(ctx.h_result.as<u64>() + k);
        cudaError_t e = cudaMemcpyAsync(ctx.h_result.p, ctx.dump_pays2.p, k * sizeof(u64), cudaMemcpyDeviceToHost,
                                        ctx.stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h_tail, d_err + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
        if (e != cudaSuccess) { ctx.poison(); return Status::Cuda(cudaGetErrorString(e)); }
        const u64* pays = ctx.h_result.as<u64>();
        out->k = k;
        out->counts.assign(1, (uint32_t)k);
        out->err_rows.assign(1, h_tail[0]);
        out->rows.resize(k);
        out->raws.resize(k);
        for (size_t i = 0; i < k; ++i) {
            uint32_t bits = (uint32_t)(pays[i] >> 32);
            std::memcpy(&out->raws[i], &bits, 4);
            out->rows[i] = (uint32_t)pays[i];
        }
        return Status::Ok();
    }
    ScanPlan plan;
    VB_TRY(plan_flat_scan(job.metric, job.dims, job.row_stride, scan_layout(job), job.n,
                          (uint32_t)k, false, &plan));
    VB_TRY(prepare_workspace(ctx, plan, 1, k));
    ScanParams p;
    fill_params(ctx, job, d_q, q_stride, d_norm, k, &p);
    p.ws.defer_merge = merge_tree_wanted(plan.grid_x, k) ? 1u : 0u;
    VB_TRY(arm_pivots(ctx, &p, 1, k, ctx.stream));
    Status s = run_flat_scan(plan, p, 1, ctx.stream);
    if (s.ok() && p.ws.defer_merge) s = run_merge_tree(p.ws, 1, plan.grid_x, ctx.sort_tmp, ctx.stream);
    if (!s.ok()) { ctx.poison(); return s; }
    const size_t bytes = k * sizeof(u64) + 8;
    VB_TRY(ctx.h_result.reserve(bytes));
    cudaError_t e = cudaMemcpyAsync(ctx.h_result.p, ctx.result.p, bytes, cudaMemcpyDeviceToHost, ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
    if (e != cudaSuccess) { ctx.poison(); return Status::Cuda(cudaGetErrorString(e)); }
    const u64* pays = ctx.h_result.as<u64>();
    const uint32_t* tail = reinterpret_cast<const uint32_t*>(pays + k);
    out->k = k;
    out->counts.assign(1, tail[0]);
    out->err_rows.assign(1, tail[1]);
    out->rows.resize(k);
    out->raws.resize(k);
    for (size_t i = 0; i < k; ++i) {
        uint32_t bits = (uint32_t)(pays[i] >> 32);
        std::memcpy(&out->raws[i], &bits, 4);
        out->rows[i] = (uint32_t)pays[i];
    }
    return Status::Ok();
}

Status unpack_device_results(const u64* d_keys_in, const u64* d_pays, const uint32_t* d_counts_in, uint32_t nq, uint32_t k,
                             u64* d_keys, float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream) {
    const uint32_t total = nq * k;
    unpack_results_kernel<<<(std::max(total, nq) + 255) / 256, 256, 0, stream>>>(d_keys_in, d_pays, d_counts_in, nq, k, d_keys,
                                                                                  d_values, d_rows, d_counts, nullptr, nullptr);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

// Sorted dump of one query -> the caller's output arrays (k entries, all valid).
__global__ void unpack_sorted_kernel(const u64* keys, const u64* pays, uint32_t k, u64* out_keys, float* out_values,
                                     uint32_t* out_rows, uint32_t* out_count, const uint32_t* err, uint32_t* status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) {
        const u64 pay = pays[i];
        if (out_keys) out_keys[i] = keys[i];
        if (out_values) out_values[i] = __uint_as_float((uint32_t)(pay >> 32));
        if (out_rows) out_rows[i] = (uint32_t)pay;
    }
    if (i == 0) {
        if (out_count) *out_count = k;
        if (err != nullptr && status != nullptr && *err != kNoError) *(volatile uint32_t*)status = 1u;
    }
}

Status unpack_sorted_device(const u64* d_keys_in, const u64* d_pays, uint32_t k, u64* d_keys, float* d_values,
                            uint32_t* d_rows, uint32_t* d_count, cudaStream_t stream) {
    unpack_sorted_kernel<<<(k + 255) / 256, 256, 0, stream>>>(d_keys_in, d_pays, k, d_keys, d_values, d_rows, d_count,
                                                              nullptr, nullptr);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

Status run_scan_device(SearchCtx& ctx, const ScanJob& job, const float* d_queries, size_t q_stride,
                       const double* d_q_norms, u64* d_keys, float* d_values, uint32_t* d_rows,
                       uint32_t* d_counts, uint32_t* d_status, cudaStream_t stream) {
    if (job.n == 0 || job.k == 0 || job.nq == 0) return Status::Cuda("empty scan job");
    const size_t k = std::min<size_t>(job.k, job.n);
    if (k > (size_t)kMaxFusedK) {
        // beyond the fused collector: one dump + radix sort per query, sorted prefix to the outputs
        uint32_t* d_err = nullptr;
        VB_TRY(dump_err_word(ctx, &d_err));
        for (uint32_t q = 0; q < job.nq; ++q) {
            Status s = dump_and_sort(ctx, job, d_queries + (size_t)q * q_stride, q_stride,
                                     d_q_norms ? d_q_norms + q : nullptr, d_err, stream);
            if (!s.ok()) { ctx.poison(); return s; }
            unpack_sorted_kernel<<<((uint32_t)k + 255) / 256, 256, 0, stream>>>(
                ctx.dump_keys2.as<u64>(), ctx.dump_pays2.as<u64>(), (uint32_t)k, d_keys ? d_keys + (size_t)q * k : nullptr,
                d_values ? d_values + (size_t)q * k : nullptr, d_rows ? d_rows + (size_t)q * k : nullptr,
                d_counts ? d_counts + q : nullptr, d_err, d_status);
            VB_CUDA(cudaGetLastError());
        }
        return Status::Ok();
    }
    ScanPlan plan;
    VB_TRY(plan_flat_scan(job.metric, job.dims, job.row_stride, scan_layout(job), job.n, (uint32_t)k, false, &plan));
    cudaStream_t saved = ctx.stream;
    ctx.stream = stream;  // workspace arming must be ordered on the caller's stream
    Status s = prepare_workspace(ctx, plan, job.nq, k);
    ctx.stream = saved;
    VB_TRY(s);
    ScanParams p;
    fill_params(ctx, job, d_queries, q_stride, d_q_norms, k, &p);
    p.ws.defer_merge = merge_tree_wanted(plan.grid_x, k) ? 1u : 0u;
    VB_TRY(arm_pivots(ctx, &p, job.nq, k, stream));
    s = run_flat_scan(plan, p, job.nq, stream);
    if (s.ok() && p.ws.defer_merge) s = run_merge_tree(p.ws, job.nq, plan.grid_x, ctx.sort_tmp, stream);
    if (!s.ok()) { ctx.poison(); return s; }
    const uint32_t total = job.nq * (uint32_t)k;
    unpack_results_kernel<<<(std::max(total, job.nq) + 255) / 256, 256, 0, stream>>>(
        p.ws.out_keys, p.ws.out_pays, p.ws.out_counts, job.nq, (uint32_t)k, d_keys, d_values, d_rows, d_counts,
        p.ws.out_err, d_status);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

}  // namespace vb
