// select.cu — K7 top-k list merge (see select.h). One CTA per query streams every list
// through the shared-memory collector; payload = position of the entry in the input, so
// values and rows are gathered once for the k_out survivors.
#include "select.h"

#include <cstdlib>

#include "runtime.h"
#include "topk.cuh"

namespace vb {

__global__ void __launch_bounds__(256)
topk_merge_kernel(const u64* keys, const float* values, const uint32_t* rows, const uint32_t* counts,
                  size_t list_stride, uint32_t lists, uint32_t k_in, uint32_t k_out, uint32_t cap, u64* keys_out, float* values_out,
                  u64* rows_out, uint32_t* counts_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    const uint32_t qi = blockIdx.x;
    Collector col;
    col.init(smem, &s_thresh, &s_count, cap, k_out);
    __syncthreads();
    auto at = [list_stride](const void* base, uint32_t l) {
        return reinterpret_cast<const unsigned char*>(base) + (size_t)l * list_stride;
    };
    // A list that holds >= k_out entries bounds the global k_out-th key from above.
    for (uint32_t l = threadIdx.x; l < lists; l += blockDim.x) {
        const uint32_t cnt = min(reinterpret_cast<const uint32_t*>(at(counts, l))[qi], k_in);
        if (cnt >= k_out) {
            const u64 kth = reinterpret_cast<const u64*>(at(keys, l))[(size_t)qi * k_in + k_out - 1];
            if (kth != kKeyMax) atomicMin(col.thresh, kth + 1);
        }
    }
    __syncthreads();
    collector_merge_lists(
        col, lists, k_in,
        [&](uint32_t l) { return reinterpret_cast<const uint32_t*>(at(counts, l))[qi]; },
        [&](uint32_t l, uint32_t i) { return reinterpret_cast<const u64*>(at(keys, l))[(size_t)qi * k_in + i]; },
        [&](uint32_t l, uint32_t i) { return ((u64)l << 32) | i; });
    const uint32_t total = *col.count;
    for (uint32_t i = threadIdx.x; i < k_out; i += blockDim.x) {
        const size_t o = (size_t)qi * k_out + i;
        if (i < total) {
            const u64 pos = col.pays[i];
            const uint32_t l = (uint32_t)(pos >> 32);
            const size_t src = (size_t)qi * k_in + (uint32_t)pos;
            keys_out[o] = col.keys[i];
            if (values_out) values_out[o] = reinterpret_cast<const float*>(at(values, l))[src];
            if (rows_out) rows_out[o] = ((u64)l << 32) | reinterpret_cast<const uint32_t*>(at(rows, l))[src];
        } else {
            keys_out[o] = kKeyMax;
            if (values_out) values_out[o] = 0.0f;
            if (rows_out) rows_out[o] = 0;
        }
    }
    if (threadIdx.x == 0) counts_out[qi] = total;
}

Status topk_merge_device(const u64* d_keys, const float* d_values, const uint32_t* d_rows, const uint32_t* d_counts,
                         size_t list_stride, size_t nq, size_t lists, size_t k_in, size_t k_out, u64* d_keys_out, float* d_values_out,
                         u64* d_rows_out, uint32_t* d_counts_out, cudaStream_t stream) {
    if (nq == 0 || lists == 0 || k_in == 0 || k_out == 0) return Status::Cuda("empty merge");
    if (k_out > (size_t)kMaxFusedK || k_in > (size_t)kMaxFusedK) return Status::Cuda("merge k beyond 1024");
    uint32_t cap = 256;
    while (cap < 2 * k_out || cap < k_out + k_in) cap <<= 1;
    const size_t smem = (size_t)cap * 16;
    VB_TRY(ensure_dynamic_smem_for(topk_merge_kernel, smem));
    topk_merge_kernel<<<(unsigned)nq, 256, smem, stream>>>(d_keys, d_values, d_rows, d_counts, list_stride, (uint32_t)lists,
                                                          (uint32_t)k_in, (uint32_t)k_out, cap, d_keys_out,
                                                          d_values_out, d_rows_out, d_counts_out);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

constexpr uint32_t kTreeGroup = 16;

// One CTA merges `group` consecutive lists of query blockIdx.y into list blockIdx.x of the output.
__global__ void __launch_bounds__(256)
topk_tree_merge_kernel(const u64* keys_in, const u64* pays_in, const uint32_t* counts_in, uint32_t lists_in,
                       uint32_t k, uint32_t cap, u64* keys_out, u64* pays_out, uint32_t* counts_out,
                       uint32_t lists_out, uint32_t* err_row, uint32_t* out_err, u64* g_thresh, const u64* g_bound) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    const uint32_t qi = blockIdx.y, og = blockIdx.x;
    const uint32_t l0 = og * kTreeGroup, nl = min(kTreeGroup, lists_in - l0);
    Collector col;
    col.init(smem, &s_thresh, &s_count, cap, k);
    __syncthreads();
    // the scan's launch-wide bound (k-th key some CTA proved, or a pivot below which k rows were counted)
    if (threadIdx.x == 0 && g_bound != nullptr) {
        const u64 g = g_bound[qi];
        if (g != kKeyMax) atomicMin(col.thresh, g + 1);
    }
    const u64* qk = keys_in + ((size_t)qi * lists_in + l0) * k;
    const u64* qp = pays_in + ((size_t)qi * lists_in + l0) * k;
    const uint32_t* qc = counts_in + (size_t)qi * lists_in + l0;
    // a full list bounds the k-th key of the merged result from above
    for (uint32_t l = threadIdx.x; l < nl; l += blockDim.x)
        if (qc[l] >= k) {
            const u64 kth = qk[(size_t)l * k + k - 1];
            if (kth != kKeyMax) atomicMin(col.thresh, kth + 1);
        }
    __syncthreads();
    collector_merge_lists(
        col, nl, k, [qc](uint32_t l) { return qc[l]; },
        [qk, k](uint32_t l, uint32_t i) { return qk[(size_t)l * k + i]; },
        [qp, k](uint32_t l, uint32_t i) { return qp[(size_t)l * k + i]; });
    const uint32_t total = *col.count;
    const size_t o = ((size_t)qi * lists_out + og) * k;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        keys_out[o + i] = col.keys[i];
        pays_out[o + i] = col.pays[i];
    }
    if (threadIdx.x == 0) {
        counts_out[(size_t)qi * lists_out + og] = total;
        if (lists_out == 1) {   // final level: finish what the last CTA of the scan would have done
            if (out_err) { out_err[qi] = err_row[qi]; err_row[qi] = kNoError; }
            if (g_thresh) g_thresh[qi] = kKeyMax;
        }
    }
}

// Final level for LONG lists (k in the hundreds: the 1000 Hamming candidates of quantized_search, funnel stages):
// one CTA pushing 10 x 1000 entries through the collector's sorts cost 43 us on a 265 us scan. The lists are
// sorted and keys are unique (the low word is the id rank), so an entry's place in the merged order is its own
// index plus, per other list, the number of entries below it — a binary search each, no sort, no atomics,
// every entry independent. Each of a few CTAs stages all keys in shared memory and ranks its share.
constexpr uint32_t kRankMergeThreads = 512;
constexpr uint32_t kRankMergeCtas = 8;
constexpr uint32_t kRankMergeMaxEntries = 24576;   // lists * k keys of 8 bytes: 192 KB of shared memory

__global__ void __launch_bounds__(kRankMergeThreads)
topk_rank_merge_kernel(const u64* keys_in, const u64* pays_in, const uint32_t* counts_in, uint32_t lists_in, uint32_t k,
                       u64* keys_out, u64* pays_out, uint32_t* counts_out, uint32_t* err_row, uint32_t* out_err,
                       u64* g_thresh) {
    extern __shared__ __align__(16) unsigned char smem[];
    u64* s_keys = reinterpret_cast<u64*>(smem);                       // [lists_in][k]
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_keys + (size_t)lists_in * k);
    const uint32_t qi = blockIdx.y;
    const u64* qk = keys_in + (size_t)qi * lists_in * k;
    const u64* qp = pays_in + (size_t)qi * lists_in * k;
    const uint32_t* qc = counts_in + (size_t)qi * lists_in;
    for (uint32_t l = threadIdx.x; l < lists_in; l += blockDim.x) s_cnt[l] = min(qc[l], k);
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < lists_in * k; s += blockDim.x) {
        const uint32_t l = s / k, i = s - l * k;
        if (i < s_cnt[l]) s_keys[s] = qk[s];
    }
    __syncthreads();
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < lists_in * k; s += gridDim.x * blockDim.x) {
        const uint32_t l = s / k, i = s - l * k;
        if (i >= s_cnt[l]) continue;
        const u64 key = s_keys[s];
        uint32_t pos = i;
        for (uint32_t m = 0; m < lists_in && pos < k; ++m) {
            if (m == l) continue;
            const u64* lk = s_keys + (size_t)m * k;
            uint32_t lo = 0, hi = s_cnt[m];
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lk[mid] < key) lo = mid + 1; else hi = mid;
            }
            pos += lo;
        }
        if (pos < k) {
            keys_out[(size_t)qi * k + pos] = key;
            pays_out[(size_t)qi * k + pos] = qp[s];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t total = 0;
        for (uint32_t l = 0; l < lists_in; ++l) total += s_cnt[l];
        counts_out[qi] = min(total, k);
        if (out_err) { out_err[qi] = err_row[qi]; err_row[qi] = kNoError; }   // what the last CTA of the scan would have done
        if (g_thresh) g_thresh[qi] = kKeyMax;
    }
}

Status run_merge_tree(const TopkWorkspace& ws, uint32_t nq, uint32_t lists, DeviceBuf& scratch, cudaStream_t stream) {
    const uint32_t k = ws.k;
    uint32_t cap = 256;
    while (cap < 2 * k + 64) cap <<= 1;
    if (k > 64) cap = std::max<uint32_t>(cap, 4096);   // sparse lists under a tight bound: one pass, one sort
    const size_t smem = (size_t)cap * 16;
    VB_TRY(ensure_dynamic_smem_for(topk_tree_merge_kernel, smem));
    // scratch: two ping-pong levels of at most ceil(lists / group) lists each
    const uint32_t l1 = (lists + kTreeGroup - 1) / kTreeGroup;
    const size_t level_entries = (size_t)nq * l1 * k;
    const size_t level_bytes = level_entries * 16 + (size_t)nq * l1 * 4;
    VB_TRY(scratch.reserve(2 * ((level_bytes + 255) & ~(size_t)255)));
    unsigned char* base[2] = {scratch.as<unsigned char>(),
                              scratch.as<unsigned char>() + ((level_bytes + 255) & ~(size_t)255)};
    const u64* kin = ws.cand_keys;
    const u64* pin = ws.cand_pays;
    const uint32_t* cin = ws.cand_counts;
    uint32_t lin = lists;
    int level = 0;
    const bool rank_final = k >= 128 && !std::getenv("VB_NO_RANK_MERGE");
    for (;;) {
        if (rank_final && level > 0 && (size_t)lin * k <= kRankMergeMaxEntries) {
            // the remaining lists are sorted outputs of the previous level: rank-merge them straight into the result
            const size_t smem_r = (size_t)lin * k * sizeof(u64) + (size_t)lin * sizeof(uint32_t) + 16;
            VB_TRY(ensure_dynamic_smem_for(topk_rank_merge_kernel, smem_r));
            topk_rank_merge_kernel<<<dim3(kRankMergeCtas, nq), kRankMergeThreads, smem_r, stream>>>(
                kin, pin, cin, lin, k, ws.out_keys, ws.out_pays, ws.out_counts, ws.err_row, ws.out_err, ws.g_thresh);
            VB_CUDA(cudaGetLastError());
            break;
        }
        const uint32_t lout = (lin + kTreeGroup - 1) / kTreeGroup;
        u64 *kout, *pout;
        uint32_t* cout;
        if (lout == 1) {
            kout = ws.out_keys;
            pout = ws.out_pays;
            cout = ws.out_counts;
        } else {
            unsigned char* b = base[level & 1];
            kout = reinterpret_cast<u64*>(b);
            pout = kout + (size_t)nq * lout * k;
            cout = reinterpret_cast<uint32_t*>(pout + (size_t)nq * lout * k);
        }
        topk_tree_merge_kernel<<<dim3(lout, nq), 256, smem, stream>>>(kin, pin, cin, lin, k, cap, kout, pout, cout, lout,
                                                                      ws.err_row, lout == 1 ? ws.out_err : nullptr,
                                                                      lout == 1 ? ws.g_thresh : nullptr, ws.g_thresh);
        VB_CUDA(cudaGetLastError());
        if (lout == 1) break;
        kin = kout;
        pin = pout;
        cin = cout;
        lin = lout;
        ++level;
    }
    return Status::Ok();
}

}  // namespace vb
