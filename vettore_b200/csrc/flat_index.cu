// flat_index.cu — see flat_index.h. Validation order and messages follow the reference
// (flat.rs:59-144, search.rs:38-73) so errors are indistinguishable at the NIF boundary.
#include "flat_index.h"

#include "flat_gemm.h"
#include "hamming.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace vb {

namespace {

constexpr uint64_t kRankSpace = 1ull << 32;
constexpr uint64_t kAppendStep = 16;
constexpr size_t kStageBytes = 64u << 20;

bool all_finite(const float* v, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) return false;
    return true;
}

// flat.rs:136-144
const char* validate_vector(const float* v, size_t len, size_t expected /*0 == None*/) {
    if (len == 0) return "vector must not be empty";
    if (expected != 0 && len != expected) return "dimension mismatch";
    if (!all_finite(v, len)) return "vector contains a non-finite value";
    return nullptr;
}

}  // namespace

FlatIndex::~FlatIndex() {
    cudaSetDevice(device_);
    if (dev_ctx_) {
        cudaDeviceSynchronize();
        dev_ctx_->destroy();
        delete dev_ctx_;
    }
    if (d_rows_) cudaFree(d_rows_);
    if (d_rank_) cudaFree(d_rank_);
    if (d_codes_) cudaFree(d_codes_);
    if (d_prefix_) cudaFree(d_prefix_);
    if (d_norm2_) cudaFree(d_norm2_);
    if (d_status_) cudaFreeHost(d_status_);
}

// Index mutators work on the legacy default stream; searches run on non-blocking streams that do
// not order against it, and D2D copies / memsets return before they complete. Every mutator ends
// here, still holding the write lock, so a search that starts next sees the finished state.
Status FlatIndex::finish_mutation() {
    VB_CUDA(cudaStreamSynchronize(nullptr));
    return Status::Ok();
}

Status FlatIndex::ensure_dev_ctx() {
    if (!dev_ctx_) {
        dev_ctx_ = new SearchCtx();
        dev_ctx_->device = device_;
        VB_CUDA(cudaStreamCreateWithFlags(&dev_ctx_->stream, cudaStreamNonBlocking));
    }
    if (!d_status_) {   // pinned host word (UVA: the kernels store to the same address): reading it costs no copy
        VB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&d_status_), 16, cudaHostAllocPortable));
        d_status_[0] = 0u;
    }
    return Status::Ok();
}

Status FlatIndex::device_status(uint32_t* out) {
    std::unique_lock<std::shared_mutex> g(mu_);
    *out = 0;
    if (!d_status_) return Status::Ok();
    VB_CUDA(cudaSetDevice(device_));
    VB_CUDA(cudaDeviceSynchronize());
    *out = *reinterpret_cast<volatile uint32_t*>(d_status_);
    d_status_[0] = 0u;
    return Status::Ok();
}

// Callers hold the read lock; norm_mu_ serialises the lazy (re)computation among concurrent batched searches.
Status FlatIndex::ensure_norms(SearchCtx& ctx) {
    std::lock_guard<std::mutex> ng(norm_mu_);
    if (max_norm_ >= 0.0f) return Status::Ok();
    const bool l2_family = metric_ == kL2 || metric_ == kL2Squared;
    if (l2_family && norm2_cap_ < n_) {
        if (d_norm2_) cudaFree(d_norm2_);
        d_norm2_ = nullptr;
        norm2_cap_ = 0;
        VB_CUDA(cudaMalloc(&d_norm2_, cap_ * sizeof(float)));
        norm2_cap_ = cap_;
    }
    return flat_gemm_max_row_norm(ctx, d_rows_, stride_, n_, dim_, &max_norm_, l2_family ? d_norm2_ : nullptr);
}

int64_t FlatIndex::find_row(const std::string& id) const {
    if (sorted_) {
        auto pos = std::lower_bound(row_id_.begin(), row_id_.end(), id);
        return (pos != row_id_.end() && *pos == id) ? (int64_t)(pos - row_id_.begin()) : -1;
    }
    auto it = id_row_.find(id);
    return it == id_row_.end() ? -1 : (int64_t)it->second;
}

void FlatIndex::leave_sorted_mode() {
    if (!sorted_) return;
    id_row_.clear();
    for (size_t r = 0; r < row_id_.size(); ++r) id_row_.emplace_hint(id_row_.end(), row_id_[r], (uint32_t)r);
    sorted_ = false;
}

bool FlatIndex::add_id(std::string&& id, uint32_t* existing, bool* relabel_needed) {
    const uint32_t row = (uint32_t)n_;
    if (sorted_) {
        if (n_ == 0 || row_id_.back() < id) {   // ascending append: O(1), no map
            uint64_t r = row;
            if (!external_ranks_ && !*relabel_needed) {
                r = (n_ == 0 ? 0 : (uint64_t)h_rank_.back()) + kAppendStep;
                if (r >= kRankSpace - 1) { *relabel_needed = true; r = row; }
            }
            row_id_.push_back(std::move(id));
            h_rank_.push_back((uint32_t)r);
            ++n_;
            return true;
        }
        const int64_t found = find_row(id);
        if (found >= 0) { *existing = (uint32_t)found; return false; }
        leave_sorted_mode();
    }
    auto hint = id_row_.end();
    if (!(!id_row_.empty() && std::prev(hint)->first < id)) hint = id_row_.lower_bound(id);
    if (hint != id_row_.end() && hint->first == id) { *existing = hint->second; return false; }
    auto it = id_row_.emplace_hint(hint, id, row);
    row_id_.push_back(std::move(id));
    h_rank_.push_back(0);
    ++n_;
    assign_rank(it, row, relabel_needed);
    return true;
}

void FlatIndex::info(size_t* rows, size_t* dim) {
    std::shared_lock<std::shared_mutex> g(mu_);
    *rows = n_;
    *dim = dim_;
}

Status FlatIndex::grow(size_t need_rows) {
    if (need_rows <= cap_) return Status::Ok();
    size_t new_cap = std::max<size_t>(need_rows, std::max<size_t>(cap_ * 2, 1024));
    if (reserve_hint_ >= need_rows) new_cap = reserve_hint_;   // sized by the caller: no head-room guess
    float* rows = nullptr;
    uint32_t* rank = nullptr;
    cudaError_t e = cudaMalloc(&rows, new_cap * stride_ * sizeof(float));
    if (e != cudaSuccess && new_cap > need_rows) {  // retry without head-room
        cudaGetLastError();
        new_cap = need_rows;
        e = cudaMalloc(&rows, new_cap * stride_ * sizeof(float));
    }
    if (e != cudaSuccess) return Status::Cuda(cudaGetErrorString(e));
    e = cudaMalloc(&rank, new_cap * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaFree(rows);
        return Status::Cuda(cudaGetErrorString(e));
    }
    if (n_ > 0) {
        VB_CUDA(cudaMemcpy(rows, d_rows_, n_ * stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
        VB_CUDA(cudaMemcpy(rank, d_rank_, n_ * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
    }
    if (d_codes_) {
        u64* codes = nullptr;
        VB_CUDA(cudaMalloc(&codes, new_cap * code_words_ * sizeof(u64)));
        if (n_ > 0) VB_CUDA(cudaMemcpy(codes, d_codes_, n_ * code_words_ * sizeof(u64), cudaMemcpyDeviceToDevice));
        cudaFree(d_codes_);
        d_codes_ = codes;
    }
    if (d_prefix_) {
        float* pre = nullptr;
        VB_CUDA(cudaMalloc(&pre, new_cap * prefix_stride_ * sizeof(float)));
        if (n_ > 0) VB_CUDA(cudaMemcpy(pre, d_prefix_, n_ * prefix_stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
        cudaFree(d_prefix_);
        d_prefix_ = pre;
    }
    if (d_rows_) cudaFree(d_rows_);
    if (d_rank_) cudaFree(d_rank_);
    d_rows_ = rows;
    d_rank_ = rank;
    cap_ = new_cap;
    return Status::Ok();
}

Status FlatIndex::reserve(size_t rows) {
    std::unique_lock<std::shared_mutex> g(mu_);
    if (rows >= kRankSpace - 1) return Status::Cuda("index row limit (2^32) exceeded");
    reserve_hint_ = rows;
    if (dim_ != 0 && rows > cap_) {
        VB_CUDA(cudaSetDevice(device_));
        if (dev_ctx_) VB_CUDA(cudaDeviceSynchronize());
        VB_TRY(grow(rows));
    }
    return Status::Ok();
}

Status FlatIndex::pack_rows(size_t row0, size_t rows) {
    if (rows == 0) return Status::Ok();
    if (d_prefix_) {
        float* dst = d_prefix_ + row0 * prefix_stride_;
        if (prefix_stride_ != prefix_dims_) VB_CUDA(cudaMemset(dst, 0, rows * prefix_stride_ * sizeof(float)));
        VB_CUDA(cudaMemcpy2D(dst, prefix_stride_ * sizeof(float), d_rows_ + row0 * stride_, stride_ * sizeof(float),
                             prefix_dims_ * sizeof(float), rows, cudaMemcpyDeviceToDevice));
    }
    if (d_codes_)
        VB_TRY(sign_pack_device(d_rows_ + row0 * stride_, stride_, (uint32_t)rows, (uint32_t)dim_,
                                d_codes_ + row0 * code_words_, nullptr));
    VB_CUDA(cudaStreamSynchronize(nullptr));
    return Status::Ok();
}

bool FlatIndex::prefix_wanted(size_t dims) const {
    // worth a mirror: a real prefix (at most half of the row, so the copy costs at most half the matrix again)
    // over enough rows for the stage to be bandwidth- rather than latency-bound
    return dims > 0 && 2 * dims <= dim_ && n_ >= 32768 && !std::getenv("VB_NO_PREFIX_MIRROR");
}

Status FlatIndex::ensure_prefix(size_t dims) {
    if (d_prefix_ && prefix_dims_ == dims) return Status::Ok();
    if (d_prefix_) { cudaFree(d_prefix_); d_prefix_ = nullptr; }
    prefix_dims_ = dims;
    prefix_stride_ = (dims + 3) & ~(size_t)3;
    cudaError_t e = cudaMalloc(&d_prefix_, cap_ * prefix_stride_ * sizeof(float));
    if (e != cudaSuccess) {   // no room for the mirror: the stage keeps reading the main matrix
        cudaGetLastError();
        d_prefix_ = nullptr;
        prefix_dims_ = prefix_stride_ = 0;
        return Status::Ok();
    }
    return pack_rows(0, n_);
}

Status FlatIndex::ensure_codes() {
    if (d_codes_ || n_ == 0) return Status::Ok();
    code_words_ = (dim_ + 63) / 64;
    VB_CUDA(cudaMalloc(&d_codes_, cap_ * code_words_ * sizeof(u64)));
    return pack_rows(0, n_);
}

Status FlatIndex::relabel_all() {
    const uint64_t spacing = std::max<uint64_t>(1, std::min<uint64_t>(kRankSpace / (n_ + 1), 1u << 20));
    uint64_t r = 0;
    if (sorted_) {
        for (size_t row = 0; row < n_; ++row) {
            r += spacing;
            h_rank_[row] = (uint32_t)std::min<uint64_t>(r, kRankSpace - 1);
        }
    } else {
        for (auto& kv : id_row_) {
            r += spacing;
            h_rank_[kv.second] = (uint32_t)std::min<uint64_t>(r, kRankSpace - 1);
        }
    }
    if (n_ > 0) VB_CUDA(cudaMemcpy(d_rank_, h_rank_.data(), n_ * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return Status::Ok();
}

// Order-maintenance label for a freshly inserted id: strictly between its neighbours.
Status FlatIndex::assign_rank(std::map<std::string, uint32_t>::iterator it, uint32_t row, bool* relabel_needed) {
    if (external_ranks_ || *relabel_needed) { h_rank_[row] = row; return Status::Ok(); }
    int64_t lo = -1;
    int64_t hi = (int64_t)kRankSpace;
    if (it != id_row_.begin()) lo = h_rank_[std::prev(it)->second];
    auto nx = std::next(it);
    const bool at_end = nx == id_row_.end();
    if (!at_end) hi = h_rank_[nx->second];
    int64_t r;
    if (at_end && lo + (int64_t)kAppendStep < hi) r = lo + (int64_t)kAppendStep;
    else if (hi - lo >= 2) r = lo + (hi - lo) / 2;
    else { *relabel_needed = true; r = row; }
    h_rank_[row] = (uint32_t)r;
    return Status::Ok();
}

Status FlatIndex::insert_many(size_t n, const char* ids, const uint64_t* id_off, const float* values,
                              const uint64_t* value_off, bool single) {
    (void)single;
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_CUDA(cudaSetDevice(device_));
    // flat.rs:70-76: validate the whole batch against the index dimension (or the first row's).
    size_t expected = dim_;
    if (expected == 0 && n > 0) expected = value_off[1] - value_off[0];
    for (size_t i = 0; i < n; ++i) {
        const char* e = validate_vector(values + value_off[i], value_off[i + 1] - value_off[i], expected);
        if (e) return Status::Ref(e);
    }
    if (n == 0) return Status::Ok();
    if (n_ + n >= kRankSpace - 1) return Status::Cuda("index row limit (2^32) exceeded");
    max_norm_ = -1.0f;   // recomputed on the next batched search
    if (dev_ctx_) VB_CUDA(cudaDeviceSynchronize());

    if (dim_ == 0) {
        dim_ = expected;
        stride_ = (dim_ + 3) & ~(size_t)3;
    }
    VB_TRY(grow(n_ + n));

    const size_t row_bytes = stride_ * sizeof(float);
    const size_t stage_rows = std::max<size_t>(1, std::min<size_t>(n, kStageBytes / row_bytes));
    PinnedBuf stage;
    VB_TRY(stage.reserve(stage_rows * row_bytes + row_bytes));
    float* sbuf = stage.as<float>();
    float* one = sbuf + stage_rows * stride_;  // scratch row for in-place replacement
    size_t staged = 0;                         // rows in sbuf
    size_t flushed_to = n_;                    // device row where sbuf[0] lands
    const size_t n_before = n_;
    bool relabel_needed = false;
    Status st = Status::Ok();

    auto flush = [&]() -> Status {
        if (staged == 0) return Status::Ok();
        VB_CUDA(cudaMemcpy(d_rows_ + flushed_to * stride_, sbuf, staged * row_bytes, cudaMemcpyHostToDevice));
        flushed_to += staged;
        staged = 0;
        return Status::Ok();
    };
    auto fill_row = [&](float* dst, size_t i) {
        std::memcpy(dst, values + value_off[i], dim_ * sizeof(float));
        for (size_t c = dim_; c < stride_; ++c) dst[c] = 0.0f;
    };

    for (size_t i = 0; i < n && st.ok(); ++i) {
        std::string id(ids + id_off[i], ids + id_off[i + 1]);
        uint32_t row = 0;
        if (!add_id(std::move(id), &row, &relabel_needed)) {
            // upsert of an existing id (flat.rs:64 / :79): replace the row in place
            if (row >= flushed_to) {
                fill_row(sbuf + (row - flushed_to) * stride_, i);  // still staged (duplicate id in this batch)
            } else {
                fill_row(one, i);
                cudaError_t e = cudaMemcpy(d_rows_ + (size_t)row * stride_, one, row_bytes, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) st = Status::Cuda(cudaGetErrorString(e));
                else st = pack_rows(row, 1);
            }
            continue;
        }
        if (staged == stage_rows) st = flush();
        fill_row(sbuf + staged * stride_, i);
        ++staged;
    }
    if (st.ok()) st = flush();
    if (st.ok()) st = pack_rows(n_before, n_ - n_before);
    if (st.ok()) {
        if (relabel_needed) st = relabel_all();
        else if (n_ > n_before) {
            cudaError_t e = cudaMemcpy(d_rank_ + n_before, h_rank_.data() + n_before,
                                       (n_ - n_before) * sizeof(uint32_t), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) st = Status::Cuda(cudaGetErrorString(e));
        }
    }
    stage.release();
    if (st.ok()) st = finish_mutation();
    return st;
}

// Sets *bad when any of the n * dim values is NaN or infinite.
__global__ void any_non_finite_kernel(const float* v, size_t total, uint32_t* bad) {
    bool b = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        b |= (__float_as_uint(v[i]) & 0x7f800000u) == 0x7f800000u;
    if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) *bad = 1u;
}

Status FlatIndex::insert_many_device(size_t n, const char* ids, const uint64_t* id_off, const float* d_values,
                                     size_t dim) {
    std::unique_lock<std::shared_mutex> g(mu_);
    VB_CUDA(cudaSetDevice(device_));
    if (n == 0) return Status::Ok();
    // flat.rs:70-76 / 136-144 on the device copy: all-or-nothing
    if (dim == 0) return Status::Ref("vector must not be empty");
    if (dim_ != 0 && dim != dim_) return Status::Ref("dimension mismatch");
    if (n_ + n >= kRankSpace - 1) return Status::Cuda("index row limit (2^32) exceeded");
    {
        uint32_t* d_bad = nullptr;
        uint32_t h_bad = 0;
        VB_CUDA(cudaMalloc(&d_bad, sizeof(uint32_t)));
        cudaMemset(d_bad, 0, sizeof(uint32_t));
        any_non_finite_kernel<<<1184, 256>>>(d_values, n * dim, d_bad);
        cudaError_t e = cudaMemcpy(&h_bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaFree(d_bad);
        if (e != cudaSuccess) return Status::Cuda(cudaGetErrorString(e));
        if (h_bad) return Status::Ref("vector contains a non-finite value");
    }
    max_norm_ = -1.0f;
    if (dev_ctx_) VB_CUDA(cudaDeviceSynchronize());
    if (dim_ == 0) {
        dim_ = dim;
        stride_ = (dim_ + 3) & ~(size_t)3;
    }
    VB_TRY(grow(n_ + n));
    const size_t n_before = n_;
    bool relabel_needed = false;
    Status st = Status::Ok();
    // rows [run0, i) of the batch are new ids in a row: they land at device rows [run_dst, ...)
    size_t run0 = 0, run_dst = n_;
    auto copy_rows = [&](size_t dst_row, size_t src_row, size_t rows) -> Status {
        if (rows == 0) return Status::Ok();
        float* dst = d_rows_ + dst_row * stride_;
        if (stride_ != dim_) VB_CUDA(cudaMemset(dst, 0, rows * stride_ * sizeof(float)));
        VB_CUDA(cudaMemcpy2D(dst, stride_ * sizeof(float), d_values + src_row * dim_, dim_ * sizeof(float),
                             dim_ * sizeof(float), rows, cudaMemcpyDeviceToDevice));
        return Status::Ok();
    };
    for (size_t i = 0; i < n && st.ok(); ++i) {
        std::string id(ids + id_off[i], ids + id_off[i + 1]);
        uint32_t row = 0;
        if (!add_id(std::move(id), &row, &relabel_needed)) {   // upsert: replace in place (after pending appends)
            st = copy_rows(run_dst, run0, i - run0);
            run_dst += i - run0;
            run0 = i + 1;
            if (st.ok()) st = copy_rows(row, i, 1);
            if (st.ok() && row < n_before) st = pack_rows(row, 1);
            continue;
        }
    }
    if (st.ok()) st = copy_rows(run_dst, run0, n - run0);
    if (st.ok()) st = pack_rows(n_before, n_ - n_before);
    if (st.ok()) {
        if (relabel_needed) st = relabel_all();
        else if (n_ > n_before) {
            cudaError_t e = cudaMemcpy(d_rank_ + n_before, h_rank_.data() + n_before,
                                       (n_ - n_before) * sizeof(uint32_t), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) st = Status::Cuda(cudaGetErrorString(e));
        }
    }
    if (st.ok()) st = finish_mutation();
    return st;
}

void FlatIndex::reset_if_empty() {
    if (n_ != 0) return;
    dim_ = 0;  // flat.rs:90-92: dimension resets to None
    stride_ = 0;
    cap_ = 0;
    if (d_rows_) cudaFree(d_rows_);
    if (d_rank_) cudaFree(d_rank_);
    if (d_codes_) cudaFree(d_codes_);
    if (d_prefix_) cudaFree(d_prefix_);
    if (d_norm2_) cudaFree(d_norm2_);
    d_norm2_ = nullptr;
    norm2_cap_ = 0;
    max_norm_ = -1.0f;
    d_rows_ = nullptr;
    d_rank_ = nullptr;
    d_codes_ = nullptr;
    d_prefix_ = nullptr;
    code_words_ = 0;
    prefix_dims_ = prefix_stride_ = 0;
    external_ranks_ = false;
    sorted_ = true;
    id_row_.clear();
}

Status FlatIndex::remove(const char* id, size_t id_len) {
    std::unique_lock<std::shared_mutex> g(mu_);
    const std::string key(id, id + id_len);
    const int64_t found = find_row(key);
    if (found < 0) return Status::Ok();
    VB_CUDA(cudaSetDevice(device_));
    if (dev_ctx_) VB_CUDA(cudaDeviceSynchronize());
    const uint32_t row = (uint32_t)found;
    const uint32_t last = (uint32_t)(n_ - 1);
    if (d_norm2_) max_norm_ = -1.0f;        // the row-norm mirror follows the rows: recomputed by the next batched search
    if (row != last) leave_sorted_mode();   // the hole is filled by the last row: rows leave id order
    if (!sorted_) id_row_.erase(key);
    if (row != last) {  // move the last row into the hole; its rank label travels with it
        VB_CUDA(cudaMemcpy(d_rows_ + (size_t)row * stride_, d_rows_ + (size_t)last * stride_,
                           stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
        VB_CUDA(cudaMemcpy(d_rank_ + row, d_rank_ + last, sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        if (d_codes_)
            VB_CUDA(cudaMemcpy(d_codes_ + (size_t)row * code_words_, d_codes_ + (size_t)last * code_words_,
                               code_words_ * sizeof(u64), cudaMemcpyDeviceToDevice));
        if (d_prefix_)
            VB_CUDA(cudaMemcpy(d_prefix_ + (size_t)row * prefix_stride_, d_prefix_ + (size_t)last * prefix_stride_,
                               prefix_stride_ * sizeof(float), cudaMemcpyDeviceToDevice));
        row_id_[row] = std::move(row_id_[last]);
        h_rank_[row] = h_rank_[last];
        id_row_[row_id_[row]] = row;
    }
    row_id_.pop_back();
    h_rank_.pop_back();
    --n_;
    reset_if_empty();
    return finish_mutation();
}

Status FlatIndex::search(const float* queries, size_t nq, size_t len, size_t limit, std::vector<Hits>* out) {
    out->assign(nq, Hits{});
    if (limit == 0 || nq == 0) return Status::Ok();  // flat.rs:97-99: before any validation
    std::shared_lock<std::shared_mutex> g(mu_);
    for (size_t q = 0; q < nq; ++q) {
        const char* e = validate_vector(queries + q * len, len, dim_);  // flat.rs:101
        if (e) return Status::Ref(e);
    }
    if (n_ == 0) return Status::Ok();
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    const size_t kk = std::min(limit, n_);
    if (flat_gemm_eligible(metric_, dim_, stride_, nq, kk, n_)) {
        // K2: the batch is one dense contraction on the tensor cores (+ exact re-scoring)
        VB_TRY(ensure_norms(*ctx.ctx));
        GemmResult gr;
        VB_TRY(flat_gemm_search(*ctx.ctx, metric_, d_rows_, stride_, d_rank_, n_, dim_, max_norm_, d_norm2_, queries, nq, kk, &gr));
        if (!gr.non_finite && gr.terms == 1) {
            size_t flagged = 0;
            for (size_t q = 0; q < nq; ++q) flagged += gr.flags[q] == 1;
            if (flagged >= kGemmRedoAsBatch)   // too dense for the single-pass margin: one 3xTF32 batch, not `flagged` scans
                VB_TRY(flat_gemm_search(*ctx.ctx, metric_, d_rows_, stride_, d_rank_, n_, dim_, max_norm_, d_norm2_, queries, nq,
                                        kk, &gr, 3));
        }
        if (!gr.non_finite) {
            std::vector<size_t> redo;
            for (size_t q = 0; q < nq; ++q) {
                if (gr.flags[q] == 2) return Status::Ref("metric overflow");
                if (gr.flags[q] == 1) { redo.push_back(q); continue; }
                Hits& h = (*out)[q];
                for (uint32_t i = 0; i < gr.counts[q]; ++i) {
                    const uint32_t row = gr.rows[q * gr.k + i];
                    h.add(row_id_[row].data(), row_id_[row].size(), gr.raws[q * gr.k + i], row);
                }
            }
            // queries whose candidate set could not be proven complete: single-query kernel
            for (size_t q : redo) {
                ScanJob one;
                one.metric = metric_;
                one.d_rows = d_rows_;
                one.row_stride = stride_;
                one.d_id_rank = d_rank_;
                one.n = (uint32_t)n_;
                one.dims = (uint32_t)dim_;
                one.whole_rows = true;
                one.h_queries = queries + q * len;
                one.nq = 1;
                one.q_len = len;
                one.k = kk;
                ScanResult r1;
                VB_TRY(run_scan(*ctx.ctx, one, &r1));
                if (r1.err_rows[0] != kNoError) return Status::Ref("metric overflow");
                Hits& h = (*out)[q];
                for (uint32_t i = 0; i < r1.counts[0]; ++i)
                    h.add(row_id_[r1.rows[i]].data(), row_id_[r1.rows[i]].size(), r1.raws[i], r1.rows[i]);
            }
            return Status::Ok();
        }
        // an overflowing score needs the f64 recovery of the per-query kernel: fall through
    }
    ScanJob job;
    job.metric = metric_;
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.n = (uint32_t)n_;
    job.dims = (uint32_t)dim_;
    job.whole_rows = true;
    job.h_queries = queries;
    job.nq = (uint32_t)nq;
    job.q_len = len;
    job.k = kk;
    ScanResult res;
    VB_TRY(run_scan(*ctx.ctx, job, &res));
    for (size_t q = 0; q < nq; ++q)
        if (res.err_rows[q] != kNoError) return Status::Ref("metric overflow");  // flat.rs:105 `?`
    for (size_t q = 0; q < nq; ++q) {
        Hits& h = (*out)[q];
        const uint32_t cnt = res.counts[q];
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t row = res.rows[q * res.k + i];
            h.add(row_id_[row].data(), row_id_[row].size(), res.raws[q * res.k + i], row);
        }
    }
    return Status::Ok();
}

Status FlatIndex::prefix_top_k(bool all_rows, size_t n_ids, const char* ids, const uint64_t* id_off,
                               const float* query, size_t len, int metric_code, size_t dimensions, size_t limit,
                               Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");            // nifs.rs:160
    if (dimensions == 0 || dimensions > len) return Status::Ref("invalid prefix dimensions");  // search.rs:45-47
    if (!all_finite(query, dimensions)) return Status::Ref("vector contains a non-finite value");
    std::shared_lock<std::shared_mutex> g(mu_);
    std::vector<uint32_t> sel;
    if (!all_rows) {
        sel.reserve(n_ids);
        for (size_t i = 0; i < n_ids; ++i) {
            const int64_t row = find_row(std::string(ids + id_off[i], ids + id_off[i + 1]));
            if (row >= 0) sel.push_back((uint32_t)row);
        }
    }
    const size_t cand = all_rows ? n_ : sel.size();
    if (cand == 0) return Status::Ok();
    if (dimensions > dim_) return Status::Ref("dimension mismatch");                          // search.rs:52-54
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    ScanJob job;
    job.metric = metric_code == kCosine ? kCosineTrue : metric_code;                          // search.rs:56-60
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.n = (uint32_t)cand;
    job.dims = (uint32_t)dimensions;
    job.whole_rows = dimensions == dim_;
    job.h_queries = query;
    job.nq = 1;
    job.q_len = len;
    job.k = std::max<size_t>(1, std::min(limit, cand));
    if (!all_rows) {
        VB_TRY(ctx->row_sel.reserve(sel.size() * sizeof(uint32_t)));
        VB_CUDA(cudaMemcpyAsync(ctx->row_sel.p, sel.data(), sel.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                ctx->stream));
        VB_CUDA(cudaStreamSynchronize(ctx->stream));  // `sel` is pageable: finish before it can go away
        job.d_row_sel = ctx->row_sel.as<uint32_t>();
    }
    ScanResult res;
    VB_TRY(run_scan(*ctx.ctx, job, &res));
    if (res.err_rows[0] != kNoError) return Status::Ref("metric overflow");
    if (limit == 0) return Status::Ok();
    for (uint32_t i = 0; i < res.counts[0]; ++i) {
        const uint32_t row = res.rows[i];
        out->add(row_id_[row].data(), row_id_[row].size(), res.raws[i], row);
    }
    return Status::Ok();
}

Status FlatIndex::funnel_search(const float* query, size_t len, int metric_code, const size_t* stages,
                                size_t nstages, size_t candidates, size_t limit, Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");
    // Every stage is one vector_top_k call (search.rs:38-73): validate them all up front in
    // the order the reference would meet them (collection.ex:674-691, then exact_rerank).
    for (size_t s = 0; s <= nstages; ++s) {
        const size_t d = s < nstages ? stages[s] : len;
        if (d == 0 || d > len) return Status::Ref("invalid prefix dimensions");
        if (!all_finite(query, d)) return Status::Ref("vector contains a non-finite value");
    }
    std::shared_lock<std::shared_mutex> g(mu_);
    // stage 1 scores a prefix of EVERY row: give it a dense mirror of those columns (built once, under the
    // write lock, like the sign codes; re-checked after the read lock is back)
    bool no_room = false;
    while (!no_room && nstages > 0 && prefix_wanted(stages[0]) && !(d_prefix_ && prefix_dims_ == stages[0])) {
        g.unlock();
        {
            std::unique_lock<std::shared_mutex> w(mu_);
            if (prefix_wanted(stages[0]) && !(d_prefix_ && prefix_dims_ == stages[0])) {
                VB_CUDA(cudaSetDevice(device_));
                VB_TRY(ensure_prefix(stages[0]));
                VB_TRY(finish_mutation());
                no_room = d_prefix_ == nullptr;
            }
        }
        g.lock();
    }
    if (n_ == 0) return Status::Ok();
    for (size_t s = 0; s <= nstages; ++s)
        if ((s < nstages ? stages[s] : len) > dim_) return Status::Ref("dimension mismatch");
    if (candidates == 0) return Status::Ok();   // vector_top_k(limit 0) -> [] at the first stage (search.rs:48)
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    const uint32_t nslots = (uint32_t)nstages + 1;
    VB_TRY(ctx->h_misc.reserve(nslots * sizeof(uint32_t)));
    uint32_t* h_err = ctx->h_misc.as<uint32_t>();
    for (uint32_t s = 0; s < nslots; ++s) h_err[s] = kNoError;
    ScanJob job;
    job.metric = metric_code == kCosine ? kCosineTrue : metric_code;
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.h_queries = query;
    job.nq = 1;
    job.q_len = len;
    size_t survivors = n_;
    for (size_t s = 0; s < nstages; ++s) {
        job.n = (uint32_t)survivors;
        job.dims = (uint32_t)stages[s];
        job.whole_rows = stages[s] == dim_;
        job.d_row_sel = s == 0 ? nullptr : ctx->row_sel.as<uint32_t>();
        job.d_rows = d_rows_;
        job.row_stride = stride_;
        if (s == 0 && d_prefix_ && prefix_dims_ == stages[0]) {   // same row numbers, dense columns: a whole-row stream
            job.d_rows = d_prefix_;
            job.row_stride = prefix_stride_;
            job.whole_rows = true;
        }
        job.k = std::min(candidates, survivors);
        VB_TRY(run_scan_to_rows(*ctx.ctx, job, (uint32_t)s, nslots, h_err + s));
        survivors = job.k;
    }
    job.n = (uint32_t)survivors;
    job.dims = (uint32_t)len;
    job.whole_rows = len == dim_;
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_row_sel = nstages == 0 ? nullptr : ctx->row_sel.as<uint32_t>();
    job.k = std::max<size_t>(1, std::min(limit, survivors));
    ScanResult res;
    VB_TRY(run_scan_final(*ctx.ctx, job, (uint32_t)nstages, nslots, &res));
    for (size_t s = 0; s < nstages; ++s)
        if (h_err[s] != kNoError) return Status::Ref("metric overflow");
    if (res.err_rows[0] != kNoError) return Status::Ref("metric overflow");
    if (limit == 0) return Status::Ok();
    for (uint32_t i = 0; i < res.counts[0]; ++i) {
        const uint32_t row = res.rows[i];
        out->add(row_id_[row].data(), row_id_[row].size(), res.raws[i], row);
    }
    return Status::Ok();
}

Status FlatIndex::quantized_search(const float* query, size_t len, int metric_code, size_t candidates, size_t limit,
                                   Hits* out) {
    *out = Hits{};
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");
    if (len == 0) return Status::Ref("vector must not be empty");
    if (!all_finite(query, len)) return Status::Ref("vector contains a non-finite value");
    std::shared_lock<std::shared_mutex> g(mu_);
    while (n_ > 0 && !d_codes_) {
        // first use builds the code mirror under the write lock; re-check after re-acquiring the read
        // lock (a delete-to-empty in the window frees the mirror again, reset_if_empty)
        g.unlock();
        {
            std::unique_lock<std::shared_mutex> w(mu_);
            if (n_ > 0 && !d_codes_) {
                VB_CUDA(cudaSetDevice(device_));
                VB_TRY(ensure_codes());
                VB_TRY(finish_mutation());
            }
        }
        g.lock();
    }
    if (n_ == 0) return Status::Ok();
    if (len != dim_) return Status::Ref("dimension mismatch");
    const size_t cand = std::min(candidates, n_);
    if (cand == 0) return Status::Ok();   // binary_top_k(limit 0) -> [] (search.rs:95-97)
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    // query sign code (distances.rs:413-423), host scalar
    const size_t nw = code_words_;
    VB_TRY(ctx->h_misc.reserve(nw * sizeof(u64) + 16));
    u64* hq = ctx->h_misc.as<u64>();
    for (size_t w = 0; w < nw; ++w) hq[w] = 0;
    for (size_t i = 0; i < len; ++i)
        if (query[i] >= 0.0f) hq[i / 64] |= 1ull << (i % 64);
    VB_TRY(ctx->staging.reserve(nw * sizeof(u64)));
    VB_CUDA(cudaMemcpyAsync(ctx->staging.p, hq, nw * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    if (cand <= (size_t)kMaxFusedK) {
        VB_TRY(hamming_scan_device(*ctx.ctx, d_codes_, (uint32_t)n_, (uint32_t)nw, (uint32_t)dim_, d_rank_,
                                   ctx->staging.as<u64>(), 1, (uint32_t)cand, ctx->stream));
        VB_TRY(extract_rows(*ctx.ctx, ctx->result.as<u64>(), (uint32_t)cand));
    } else {
        // more candidates than the fused collector holds (collection.ex:509-510: 10 x limit by default)
        VB_TRY(hamming_dump_sorted(*ctx.ctx, d_codes_, (uint32_t)n_, (uint32_t)nw, (uint32_t)dim_, d_rank_,
                                   ctx->staging.as<u64>(), ctx->stream));
        VB_TRY(extract_rows(*ctx.ctx, ctx->dump_pays2.as<u64>(), (uint32_t)cand));
    }
    ScanJob job;
    job.metric = metric_code == kCosine ? kCosineTrue : metric_code;   // exact_rerank -> vector_top_k
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.d_row_sel = ctx->row_sel.as<uint32_t>();
    job.n = (uint32_t)cand;
    job.dims = (uint32_t)dim_;
    job.whole_rows = true;
    job.h_queries = query;
    job.nq = 1;
    job.q_len = len;
    job.k = std::max<size_t>(1, std::min(limit, cand));
    ScanResult res;
    VB_TRY(run_scan_final(*ctx.ctx, job, 0, 1, &res));
    if (res.err_rows[0] != kNoError) return Status::Ref("metric overflow");
    if (limit == 0) return Status::Ok();
    for (uint32_t i = 0; i < res.counts[0]; ++i) {
        const uint32_t row = res.rows[i];
        out->add(row_id_[row].data(), row_id_[row].size(), res.raws[i], row);
    }
    return Status::Ok();
}

Status FlatIndex::hamming_candidates(const float* query, size_t len, size_t candidates, Hits* out) {
    *out = Hits{};
    if (len == 0) return Status::Ref("vector must not be empty");
    if (!all_finite(query, len)) return Status::Ref("vector contains a non-finite value");
    std::shared_lock<std::shared_mutex> g(mu_);
    while (n_ > 0 && !d_codes_) {   // first use builds the code mirror (see quantized_search)
        g.unlock();
        {
            std::unique_lock<std::shared_mutex> w(mu_);
            if (n_ > 0 && !d_codes_) {
                VB_CUDA(cudaSetDevice(device_));
                VB_TRY(ensure_codes());
                VB_TRY(finish_mutation());
            }
        }
        g.lock();
    }
    if (n_ == 0) return Status::Ok();
    if (len != dim_) return Status::Ref("dimension mismatch");
    const size_t cand = std::min(candidates, n_);
    if (cand == 0) return Status::Ok();
    VB_CUDA(cudaSetDevice(device_));
    CtxLease ctx;
    VB_TRY(ctx.get());
    std::vector<uint64_t> code(code_words_, 0);     // query sign code (distances.rs:413-423)
    for (size_t i = 0; i < len; ++i)
        if (query[i] >= 0.0f) code[i / 64] |= 1ull << (i % 64);
    std::vector<uint32_t> rows;
    std::vector<float> vals;
    VB_TRY(hamming_top_k_resident(*ctx.ctx, d_codes_, n_, code_words_, dim_, d_rank_, code.data(), cand, &rows, &vals));
    for (size_t i = 0; i < rows.size(); ++i) out->add(row_id_[rows[i]].data(), row_id_[rows[i]].size(), vals[i], rows[i]);
    return Status::Ok();
}

Status FlatIndex::search_device(const float* d_queries, size_t nq, size_t q_stride, size_t limit, u64* d_keys,
                                float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream) {
    if (limit == 0 || nq == 0) return Status::Cuda("device search needs limit >= 1 and nq >= 1");
    std::unique_lock<std::shared_mutex> g(mu_);  // owns dev_ctx_; device-level calls are stream-ordered
    if (n_ == 0) return Status::Cuda("device search on an empty index");
    if (q_stride < stride_ || (q_stride & 3)) return Status::Ref("dimension mismatch");
    VB_CUDA(cudaSetDevice(device_));
    VB_TRY(ensure_dev_ctx());
    const size_t kk = std::min(limit, n_);
    ScanJob job;
    job.metric = metric_;
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.n = (uint32_t)n_;
    job.dims = (uint32_t)dim_;
    job.whole_rows = true;
    job.nq = (uint32_t)nq;
    job.k = kk;
    if (q_stride == dim_ && flat_gemm_eligible(metric_, dim_, stride_, nq, kk, n_)) {
        // K2 on the caller's stream. The tensor-core scores are a filter: the exact re-scoring stage
        // flags every query whose kept set could not be proven to hold the true top-k (1), or whose
        // exact score overflowed (2), and raises `bad` when a tensor-core score was non-finite. Those
        // are resolved HERE, like FlatIndex::search does: one small D2H read of the flag words, then
        // the flagged queries are redone by the single-query kernel into the same output slots — so a
        // batch through this entry (the row-sharded path) is exactly as complete as the host-facing one.
        VB_TRY(ensure_norms(*dev_ctx_));
        SearchCtx& c = *dev_ctx_;
        VB_TRY(c.result.reserve(nq * kk * sizeof(u64) + 2 * nq * sizeof(uint32_t) + 16));
        VB_TRY(c.out_keys.reserve(nq * kk * sizeof(u64)));
        VB_TRY(c.h_misc.reserve((nq + 1) * sizeof(uint32_t)));
        u64* pays = c.result.as<u64>();
        uint32_t* counts = reinterpret_cast<uint32_t*>(pays + nq * kk);
        uint32_t* flags = counts + nq;
        uint32_t* h_flags = c.h_misc.as<uint32_t>();
        for (int force = 0;; force = 3) {
            int terms_used = 3;
            VB_TRY(flat_gemm_search_device(c, metric_, d_rows_, stride_, d_rank_, n_, dim_, max_norm_, d_norm2_, d_queries, nq, kk,
                                           c.out_keys.as<u64>(), pays, counts, flags, flags + nq, stream, force, &terms_used));
            VB_CUDA(cudaMemcpyAsync(h_flags, flags, (nq + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            VB_CUDA(cudaStreamSynchronize(stream));
            size_t flagged = 0;
            for (size_t q = 0; q < nq; ++q) flagged += h_flags[q] == 1;
            if (terms_used == 1 && h_flags[nq] == 0 && flagged >= kGemmRedoAsBatch) continue;   // redo as one 3xTF32 batch
            break;
        }
        VB_TRY(unpack_device_results(c.out_keys.as<u64>(), pays, counts, (uint32_t)nq, (uint32_t)kk, d_keys, d_values,
                                     d_rows, d_counts, stream));
        if (h_flags[nq] != 0) {
            // a non-finite tensor-core score: the f64 recovery lives in the per-query kernel, redo all
            return run_scan_device(c, job, d_queries, q_stride, nullptr, d_keys, d_values, d_rows, d_counts, d_status_,
                                   stream);
        }
        for (size_t q = 0; q < nq; ++q) {
            if (h_flags[q] == 2) return Status::Ref("metric overflow");
            if (h_flags[q] != 1) continue;
            ScanJob one = job;
            one.nq = 1;
            VB_TRY(run_scan_device(c, one, d_queries + q * q_stride, q_stride, nullptr, d_keys ? d_keys + q * kk : nullptr,
                                   d_values ? d_values + q * kk : nullptr, d_rows ? d_rows + q * kk : nullptr,
                                   d_counts ? d_counts + q : nullptr, d_status_, stream));
        }
        return Status::Ok();
    }
    return run_scan_device(*dev_ctx_, job, d_queries, q_stride, nullptr, d_keys, d_values, d_rows, d_counts, d_status_,
                           stream);
}

// ---- row-sharded quantized_search pieces (all-gather of the local candidates,
// global select, every owner reranks its own survivors) ------------------------------------------
__global__ void select_owned_rows_kernel(const u64* global_rows, const uint32_t* global_count, uint32_t max_candidates,
                                         uint32_t shard, uint32_t* row_sel, uint32_t* owned) {
    const uint32_t n = min(*global_count, max_candidates);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const u64 r = global_rows[i];
        if ((uint32_t)(r >> 32) == shard) row_sel[atomicAdd(owned, 1u)] = (uint32_t)r;
    }
}

// f64 norm of the query, accumulated in index order like distances.rs:160-177 does.
__global__ void query_norm_f64_kernel(const float* q, uint32_t dims, double* out) {
    double s = 0.0;
    for (uint32_t i = 0; i < dims; ++i) s += (double)q[i] * (double)q[i];
    *out = sqrt(s);
}

Status FlatIndex::hamming_device(const float* d_queries, size_t nq, size_t q_stride, size_t candidates, u64* d_keys,
                                 float* d_values, uint32_t* d_rows, uint32_t* d_counts, cudaStream_t stream) {
    if (candidates == 0 || nq == 0) return Status::Cuda("device candidate pass needs candidates >= 1 and nq >= 1");
    std::unique_lock<std::shared_mutex> g(mu_);  // owns dev_ctx_; may build the code mirror
    if (n_ == 0) return Status::Cuda("device candidate pass on an empty index");
    if (q_stride < dim_) return Status::Ref("dimension mismatch");
    VB_CUDA(cudaSetDevice(device_));
    if (!d_codes_) { VB_TRY(ensure_codes()); VB_TRY(finish_mutation()); }
    VB_TRY(ensure_dev_ctx());
    SearchCtx& c = *dev_ctx_;
    const size_t cand = std::min(candidates, n_);
    const size_t nw = code_words_;
    VB_TRY(c.staging.reserve(nq * nw * sizeof(u64)));
    VB_TRY(sign_pack_device(d_queries, q_stride, (uint32_t)nq, (uint32_t)dim_, c.staging.as<u64>(), stream));
    if (cand > (size_t)kMaxFusedK) {   // beyond the fused collector: dump + radix sort per query
        for (size_t q = 0; q < nq; ++q) {
            VB_TRY(hamming_dump_sorted(c, d_codes_, (uint32_t)n_, (uint32_t)nw, (uint32_t)dim_, d_rank_,
                                       c.staging.as<u64>() + q * nw, stream));
            VB_TRY(unpack_sorted_device(c.dump_keys2.as<u64>(), c.dump_pays2.as<u64>(), (uint32_t)cand,
                                        d_keys ? d_keys + q * cand : nullptr, d_values ? d_values + q * cand : nullptr,
                                        d_rows ? d_rows + q * cand : nullptr, d_counts ? d_counts + q : nullptr, stream));
        }
        return Status::Ok();
    }
    cudaStream_t saved = c.stream;
    c.stream = stream;   // workspace arming must be ordered on the caller's stream
    Status s = hamming_scan_device(c, d_codes_, (uint32_t)n_, (uint32_t)nw, (uint32_t)dim_, d_rank_, c.staging.as<u64>(),
                                   (uint32_t)nq, (uint32_t)cand, stream);
    c.stream = saved;
    VB_TRY(s);
    const u64* pays = c.result.as<u64>();
    return unpack_device_results(c.out_keys.as<u64>(), pays, reinterpret_cast<const uint32_t*>(pays + nq * cand),
                                 (uint32_t)nq, (uint32_t)cand, d_keys, d_values, d_rows, d_counts, stream);
}

Status FlatIndex::rerank_owned_device(const float* d_query, size_t q_stride, int metric_code, const u64* d_global_rows,
                                      const uint32_t* d_global_count, size_t max_candidates, uint32_t shard,
                                      size_t limit, u64* d_keys, float* d_values, uint32_t* d_rows,
                                      uint32_t* d_counts, cudaStream_t stream) {
    if (metric_code < 0 || metric_code > 8) return Status::Ref("unknown metric");
    if (limit == 0 || max_candidates == 0) return Status::Cuda("device rerank needs limit >= 1 and candidates >= 1");
    std::unique_lock<std::shared_mutex> g(mu_);
    if (q_stride < stride_ || (q_stride & 3)) return Status::Ref("dimension mismatch");
    VB_CUDA(cudaSetDevice(device_));
    VB_TRY(ensure_dev_ctx());
    SearchCtx& c = *dev_ctx_;
    VB_TRY(c.row_sel.reserve(max_candidates * sizeof(uint32_t)));
    VB_TRY(c.row_sel2.reserve(16));
    VB_TRY(c.h_misc.reserve(16));
    VB_TRY(c.q_norms.reserve(sizeof(double)));
    uint32_t* d_owned = c.row_sel2.as<uint32_t>();
    uint32_t* h_owned = c.h_misc.as<uint32_t>();
    VB_CUDA(cudaMemsetAsync(d_owned, 0, sizeof(uint32_t), stream));
    select_owned_rows_kernel<<<1, 256, 0, stream>>>(d_global_rows, d_global_count, (uint32_t)max_candidates, shard,
                                                    c.row_sel.as<uint32_t>(), d_owned);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(h_owned, d_owned, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));
    const uint32_t owned = *h_owned;
    if (owned == 0 || n_ == 0) {   // nothing of the global candidate set lives here: an empty list
        VB_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t), stream));
        return Status::Ok();
    }
    ScanJob job;
    job.metric = metric_code == kCosine ? kCosineTrue : metric_code;   // exact_rerank -> vector_top_k
    job.d_rows = d_rows_;
    job.row_stride = stride_;
    job.d_id_rank = d_rank_;
    job.d_row_sel = c.row_sel.as<uint32_t>();
    job.n = owned;
    job.dims = (uint32_t)dim_;
    job.whole_rows = true;
    job.nq = 1;
    job.k = std::min<size_t>(limit, owned);
    const double* d_norm = nullptr;
    if (job.metric == kCosineTrue) {
        query_norm_f64_kernel<<<1, 1, 0, stream>>>(d_query, (uint32_t)dim_, c.q_norms.as<double>());
        VB_CUDA(cudaGetLastError());
        d_norm = c.q_norms.as<double>();
    }
    return run_scan_device(c, job, d_query, q_stride, d_norm, d_keys, d_values, d_rows, d_counts, d_status_, stream);
}

Status FlatIndex::set_id_ranks(const uint32_t* ranks, size_t n) {
    std::unique_lock<std::shared_mutex> g(mu_);
    if (n != n_) return Status::Ref("dimension mismatch");
    VB_CUDA(cudaSetDevice(device_));
    if (dev_ctx_) VB_CUDA(cudaDeviceSynchronize());
    std::copy(ranks, ranks + n, h_rank_.begin());
    if (n > 0) VB_CUDA(cudaMemcpy(d_rank_, h_rank_.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    external_ranks_ = true;
    return finish_mutation();
}

}  // namespace vb
