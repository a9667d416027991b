// topk.cuh — CTA-level bounded top-k collector (shared memory) used by every scan kernel.
//
// Replaces the reference's BinaryHeap of capacity `limit` (flat.rs:103-118,
// search.rs:94-104, multi_vector.rs:112-121). Entries are (key, payload) pairs of u64:
//   key     = rank-order key << 32 | id rank   (ascending == the reference's
//             (rank.total_cmp, id.cmp) order; keys are unique because id ranks are)
//   payload = raw value bits << 32 | device row
// Rows whose key is below the running threshold are appended with one shared-memory
// atomic; when the buffer nears capacity the CTA sorts it (bitonic), keeps the best k
// and tightens the threshold. A grid-wide threshold (global atomicMin) lets every CTA
// filter with the best k-th key any CTA has proven so far.
#pragma once
#include "common.cuh"

namespace vb {

struct Collector {
    u64* keys;          // smem [cap]
    u64* pays;          // smem [cap]
    u64* thresh;        // smem: entries must be < *thresh to matter
    uint32_t* count;    // smem
    uint32_t cap;       // power of two
    uint32_t k;

    __device__ __forceinline__ void init(unsigned char* smem, u64* s_thresh, uint32_t* s_count,
                                         uint32_t cap_, uint32_t k_) {
        keys = reinterpret_cast<u64*>(smem);
        pays = keys + cap_;
        thresh = s_thresh;
        count = s_count;
        cap = cap_;
        k = k_;
        if (threadIdx.x == 0) { *thresh = kKeyMax; *count = 0; }
    }

    __device__ __forceinline__ u64 threshold() const { return *reinterpret_cast<volatile u64*>(thresh); }

    // Any thread. The caller guarantees (by its sync cadence) that cap is not exceeded.
    __device__ __forceinline__ void push(u64 key, u64 pay) {
        uint32_t slot = atomicAdd(count, 1u);
        if (slot < cap) { keys[slot] = key; pays[slot] = pay; }
    }

    // Block-wide (every thread of the CTA must call). Sorts the buffer ascending, keeps the
    // best k and tightens the threshold. Returns the number of retained entries.
    __device__ uint32_t compact() {
        __syncthreads();
        uint32_t n = min(*count, cap);
        uint32_t p2 = 32;
        while (p2 < n) p2 <<= 1;
        for (uint32_t i = n + threadIdx.x; i < p2; i += blockDim.x) { keys[i] = kKeyMax; pays[i] = 0; }
        __syncthreads();
        for (uint32_t size = 2; size <= p2; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
                    uint32_t lo = 2 * t - (t & (stride - 1));
                    uint32_t hi = lo + stride;
                    bool up = (lo & size) == 0;
                    u64 a = keys[lo], b = keys[hi];
                    if ((a > b) == up) {
                        keys[lo] = b; keys[hi] = a;
                        u64 pa = pays[lo]; pays[lo] = pays[hi]; pays[hi] = pa;
                    }
                }
                __syncthreads();
            }
        }
        uint32_t kept = min(n, k);
        if (threadIdx.x == 0) {
            *count = kept;
            if (n >= k) atomicMin(thresh, keys[k - 1]);
        }
        __syncthreads();
        return kept;
    }
};

// Merges `lists` candidate lists (global memory, written by other CTAs or other GPUs) into
// the collector. Block-wide. List l holds list_len(l) <= k entries; key_at(l, i) / pay_at(l, i)
// fetch entry i of list l.
template <typename CountFn, typename KeyFn, typename PayFn>
__device__ void collector_merge_lists(Collector& c, uint32_t lists, CountFn list_len, KeyFn key_at, PayFn pay_at) {
    // Each round appends at most (cap - k) surviving candidates, so push never overflows.
    const uint32_t room = c.cap - c.k;
    for (uint32_t l0 = 0; l0 < lists;) {
        // take as many whole lists as fit in `room` candidates (a list holds <= room entries)
        uint32_t l1 = l0, total = 0;
        while (l1 < lists) {
            uint32_t len = list_len(l1);
            if (total + len > room) break;
            total += len;
            ++l1;
        }
        if (l1 == l0) ++l1;  // unreachable while list_len <= room; keeps the loop total
        u64 T = c.threshold();
        for (uint32_t l = l0; l < l1; ++l) {
            uint32_t len = list_len(l);
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                u64 key = key_at(l, i);
                if (key < T) c.push(key, pay_at(l, i));
            }
        }
        __syncthreads();
        const uint32_t filled = *c.count;
        __syncthreads();  // everyone has read `filled` before the next round's pushes
        if (filled + room > c.cap || l1 == lists) c.compact();
        l0 = l1;
    }
}

}  // namespace vb

namespace vb {

// Global-memory side of the fused top-k: per-CTA candidate lists, the grid-wide
// threshold, the completion counter and the final sorted output, per query slot.
struct TopkWorkspace {
    u64* cand_keys;         // [nq][grid][k]
    u64* cand_pays;
    uint32_t* cand_counts;  // [nq][grid]
    uint32_t* done;         // [nq] CTAs finished (re-armed by the last CTA)
    u64* g_thresh;          // [nq] best k-th key any CTA has proven (re-armed by the last CTA)
    u64* out_keys;          // [nq][k] sorted ascending
    u64* out_pays;
    uint32_t* out_counts;   // [nq]
    uint32_t k;
};

// Block-wide, called at a CTA-uniform cadence: adopts the grid-wide threshold, and when
// the buffer could overflow within the next `slack` pushes compacts it and publishes the
// CTA's own k-th key.
__device__ __forceinline__ void collector_checkpoint(Collector& col, const TopkWorkspace& ws, uint32_t qi,
                                                     uint32_t slack) {
    if (threadIdx.x == 0) {
        u64 g = ld_volatile_u64(ws.g_thresh + qi);
        if (g < col.threshold()) atomicMin(col.thresh, g);
    }
    const bool need = __syncthreads_or(
        (threadIdx.x & 31) == 0 && *reinterpret_cast<volatile uint32_t*>(col.count) + slack > col.cap);
    if (need) {
        col.compact();
        if (threadIdx.x == 0 && *col.thresh != kKeyMax) atomicMin(ws.g_thresh + qi, *col.thresh);
    }
}

// Block-wide epilogue of a fused scan: publishes this CTA's best k; the last CTA of the
// grid (per query slot) merges every list and writes the sorted result. `s_last` is a
// shared-memory flag owned by the caller.
__device__ __forceinline__ void collector_publish_and_merge(Collector& col, const TopkWorkspace& ws, uint32_t qi,
                                                            int* s_last) {
    const uint32_t kept = col.compact();
    const size_t slot = (size_t)qi * gridDim.x + blockIdx.x;
    for (uint32_t i = threadIdx.x; i < kept; i += blockDim.x) {
        ws.cand_keys[slot * ws.k + i] = col.keys[i];
        ws.cand_pays[slot * ws.k + i] = col.pays[i];
    }
    if (threadIdx.x == 0) ws.cand_counts[slot] = kept;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ticket = atomicAdd(ws.done + qi, 1u);
        *s_last = (ticket == gridDim.x - 1u);
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();

    if (threadIdx.x == 0) { *col.thresh = kKeyMax; *col.count = 0; }
    __syncthreads();
    const uint32_t* counts = ws.cand_counts + (size_t)qi * gridDim.x;
    const u64* gk = ws.cand_keys + (size_t)qi * gridDim.x * ws.k;
    const u64* gp = ws.cand_pays + (size_t)qi * gridDim.x * ws.k;
    const uint32_t stride = ws.k;
    collector_merge_lists(
        col, gridDim.x, [counts](uint32_t l) { return __ldcg(counts + l); },
        [gk, stride](uint32_t l, uint32_t i) { return __ldcg(gk + (size_t)l * stride + i); },
        [gp, stride](uint32_t l, uint32_t i) { return __ldcg(gp + (size_t)l * stride + i); });
    const uint32_t total = *col.count;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        ws.out_keys[(size_t)qi * ws.k + i] = col.keys[i];
        ws.out_pays[(size_t)qi * ws.k + i] = col.pays[i];
    }
    if (threadIdx.x == 0) {
        ws.out_counts[qi] = total;
        ws.done[qi] = 0u;
        ws.g_thresh[qi] = kKeyMax;
    }
}

}  // namespace vb
