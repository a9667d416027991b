// select.cu — K7 top-k list merge (see select.h). One CTA per query streams every list
// through the shared-memory collector; payload = position of the entry in the input, so
// values and rows are gathered once for the k_out survivors.
#include "select.h"
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
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        VB_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    topk_merge_kernel<<<(unsigned)nq, 256, smem, stream>>>(d_keys, d_values, d_rows, d_counts, list_stride, (uint32_t)lists,
                                                          (uint32_t)k_in, (uint32_t)k_out, cap, d_keys_out,
                                                          d_values_out, d_rows_out, d_counts_out);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

}  // namespace vb
