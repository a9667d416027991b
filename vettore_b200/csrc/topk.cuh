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

constexpr uint32_t kPivots = 16;   // rungs of the launch-wide pivot ladder

struct Collector {
    u64* keys;          // smem [cap]
    u64* pays;          // smem [cap]
    u64* thresh;        // smem: entries must be < *thresh to matter
    uint32_t* count;    // smem
    uint32_t cap;       // power of two
    uint32_t k;
    uint32_t nthreads;  // threads taking part (threadIdx.x < nthreads), multiple of 32
    uint32_t bar_id;    // hardware barrier they synchronise on (0 == __syncthreads when all do)
    // launch-wide pivot ladder (large k, see collector_pivot_step): shared-memory copies, null when unused
    u64* piv;           // [kPivots] ascending pivot keys
    uint32_t* piv_bins; // [kPivots] pushes since the last flush, by first pivot >= key
    uint32_t* piv_on;   // 1 once the ladder was adopted
    uint32_t* piv_early;// CTA 0 only: 1 once it has sorted its first rows to publish the ladder

    __device__ __forceinline__ void init(unsigned char* smem, u64* s_thresh, uint32_t* s_count,
                                         uint32_t cap_, uint32_t k_, uint32_t nthreads_ = 0,
                                         uint32_t bar_id_ = 0) {
        keys = reinterpret_cast<u64*>(smem);
        pays = keys + cap_;
        thresh = s_thresh;
        count = s_count;
        cap = cap_;
        k = k_;
        nthreads = nthreads_ ? nthreads_ : blockDim.x;
        bar_id = bar_id_;
        piv = nullptr;
        piv_bins = nullptr;
        piv_on = nullptr;
        piv_early = nullptr;
        if (threadIdx.x == 0) { *thresh = kKeyMax; *count = 0; }
    }

    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    }
    // Barrier + OR-reduction of a predicate over the participating threads.
    __device__ __forceinline__ bool sync_or(bool pred) const {
        uint32_t res;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\t"
            "selp.u32 %0, 1, 0, q;\n\t}"
            : "=r"(res) : "r"((uint32_t)pred), "r"(bar_id), "r"(nthreads) : "memory");
        return res != 0;
    }

    __device__ __forceinline__ u64 threshold() const { return *reinterpret_cast<volatile u64*>(thresh); }

    // Counts `key` into the bin of the first pivot that is >= key (nothing if it is beyond the ladder).
    __device__ __forceinline__ void count_pivot(u64 key) {
        uint32_t j = 0;
#pragma unroll
        for (uint32_t step = kPivots / 2; step > 0; step >>= 1)
            if (piv[j + step - 1] < key) j += step;
        if (piv[j] >= key) atomicAdd(&piv_bins[j], 1u);
    }

    // Any thread. The caller guarantees (by its sync cadence) that cap is not exceeded.
    __device__ __forceinline__ void push(u64 key, u64 pay) {
        uint32_t slot = atomicAdd(count, 1u);
        if (slot < cap) { keys[slot] = key; pays[slot] = pay; }
        if (piv_on != nullptr && *reinterpret_cast<volatile uint32_t*>(piv_on)) count_pivot(key);
    }

    // Block-wide (every thread of the CTA must call). Sorts the buffer ascending, keeps the
    // best k and tightens the threshold. Returns the number of retained entries.
    // Bitonic network in shared memory. A power-of-two number of warps each own an aligned segment of the array:
    // every step whose stride is shorter than a segment only touches pairs inside one segment, so it needs a
    // __syncwarp, not a block barrier — 1024 entries on 8 warps take 9 block barriers instead of 55 (a barrier round
    // of 12-16 warps costs ~0.2 us, and with k = 100 a short scan pays for several sorts: measured on the 1M-row
    // prefix scan).
    __device__ uint32_t compact() {
        sync();
        uint32_t n = min(*count, cap);
        uint32_t p2 = 32;
        while (p2 < n) p2 <<= 1;
        for (uint32_t i = n + threadIdx.x; i < p2; i += nthreads) { keys[i] = kKeyMax; pays[i] = 0; }
        sync();
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        uint32_t aw = 1;                                   // warps that own a segment: power of two, >= 64 entries each
        while (aw * 2u <= (nthreads >> 5) && p2 / (aw * 2u) >= 64u) aw <<= 1;
        const uint32_t seg = p2 / aw;
        auto exchange = [this](uint32_t t, uint32_t stride, uint32_t size) {
            const uint32_t lo = 2 * t - (t & (stride - 1));
            const uint32_t hi = lo + stride;
            const bool up = (lo & size) == 0;
            const u64 a = keys[lo], b = keys[hi];
            if ((a > b) == up) {
                keys[lo] = b; keys[hi] = a;
                const u64 pa = pays[lo]; pays[lo] = pays[hi]; pays[hi] = pa;
            }
        };
        bool local = false;                                // the previous step was warp-local
        for (uint32_t size = 2; size <= p2; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride < seg) {
                    if (warp < aw) {
                        for (uint32_t t = lane; t < (seg >> 1); t += 32u) exchange(warp * (seg >> 1) + t, stride, size);
                        __syncwarp();
                    }
                    local = true;
                } else {
                    if (local) sync();
                    for (uint32_t t = threadIdx.x; t < (p2 >> 1); t += nthreads) exchange(t, stride, size);
                    sync();
                    local = false;
                }
            }
        }
        if (local) sync();
        uint32_t kept = min(n, k);
        if (threadIdx.x == 0) {
            *count = kept;
            if (n >= k) atomicMin(thresh, keys[k - 1]);
        }
        sync();
        return kept;
    }
};

constexpr uint32_t kMergeChunk = 1024;  // lists whose counts are staged in shared memory at once

// Merges `lists` candidate lists (global memory, written by other CTAs or other GPUs) into
// the collector. Block-wide. List l holds list_len(l) <= list_cap entries; key_at(l, i) /
// pay_at(l, i) fetch entry i of list l. The collector's threshold may be pre-tightened by
// the caller (entries must be < threshold to matter).
//
// Counts are staged in shared memory and the candidates are walked as one flat index space,
// so no thread ever chases a dependent chain of global loads. First an optimistic single
// pass (with a tight threshold almost nothing survives the filter); only if that would
// overflow the buffer it is redone in windows that cannot overflow.
template <typename CountFn, typename KeyFn, typename PayFn>
__device__ void collector_merge_lists(Collector& c, uint32_t lists, uint32_t list_cap, CountFn list_len,
                                      KeyFn key_at, PayFn pay_at) {
    __shared__ uint32_t s_cnt[kMergeChunk];
    __shared__ int s_overflow;
    const uint32_t room = c.cap - c.k;
    for (uint32_t chunk0 = 0; chunk0 < lists; chunk0 += kMergeChunk) {
        const uint32_t nl = min(kMergeChunk, lists - chunk0);
        for (uint32_t l = threadIdx.x; l < nl; l += c.nthreads) s_cnt[l] = min(list_len(chunk0 + l), list_cap);
        if (threadIdx.x == 0) s_overflow = 0;
        c.sync();
        const uint32_t base = *c.count;  // <= k: left by the previous chunk's compaction
        const uint32_t slots = nl * list_cap;
        const u64 T = c.threshold();
        c.sync();
        for (uint32_t s = threadIdx.x; s < slots; s += c.nthreads) {
            const uint32_t l = s / list_cap, i = s - l * list_cap;
            if (i >= s_cnt[l]) continue;
            const u64 key = key_at(chunk0 + l, i);
            if (key >= T) continue;
            const uint32_t slot = atomicAdd(c.count, 1u);
            if (slot < c.cap) { c.keys[slot] = key; c.pays[slot] = pay_at(chunk0 + l, i); }
            else s_overflow = 1;
        }
        c.sync();
        const bool overflow = s_overflow != 0;
        c.sync();
        if (!overflow) {
            c.compact();
            continue;
        }
        // safe path: drop the partial pass, then windows of `room` slots (cannot overflow)
        if (threadIdx.x == 0) *c.count = base;
        c.sync();
        for (uint32_t w0 = 0; w0 < slots; w0 += room) {
            const u64 Tw = c.threshold();
            const uint32_t w1 = min(slots, w0 + room);
            for (uint32_t s = w0 + threadIdx.x; s < w1; s += c.nthreads) {
                const uint32_t l = s / list_cap, i = s - l * list_cap;
                if (i >= s_cnt[l]) continue;
                const u64 key = key_at(chunk0 + l, i);
                if (key < Tw) c.push(key, pay_at(chunk0 + l, i));
            }
            c.sync();
            const uint32_t filled = *c.count;
            c.sync();  // everyone has read `filled` before the next window's pushes
            if (filled + room > c.cap || w1 == slots) c.compact();
        }
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
    uint32_t* err_row;      // [nq] optional: first overflowing row (atomicMin by any CTA)
    uint32_t* out_err;      // [nq] optional: err_row snapshot taken by the last CTA (err_row re-armed)
    uint32_t k;
    uint32_t defer_merge;   // 1: CTAs only publish their lists; the host runs the merge tree (large grid * k)
    // launch-wide pivot ladder (optional, large k; zeroed before every launch): per query slot
    // kPivots ascending keys, kPivots counters and a state word (0 none, 1 being written, 2 published)
    u64* piv_keys;
    uint32_t* piv_counts;
    uint32_t* piv_state;
};

// Block-wide, called at a CTA-uniform cadence: adopts the grid-wide threshold, and when
// the buffer could overflow within the next `slack` pushes compacts it and publishes the
// CTA's own k-th key. `g_prefetch` (meaningful in thread 0) carries the grid-wide threshold
// loaded one cadence earlier, so the global-memory round trip overlaps the scan instead of
// stalling the CTA at the barrier; it is re-issued here for the next call.
// Gives the collector its shared-memory copy of the pivot ladder (call right after init, before the CTA's
// first barrier, by every thread of the CTA; kernels whose workspace carries no ladder skip it).
__device__ __forceinline__ void collector_attach_pivots(Collector& col, const TopkWorkspace& ws) {
    __shared__ u64 s_piv[kPivots];
    __shared__ uint32_t s_piv_bins[kPivots];
    __shared__ uint32_t s_piv_on, s_piv_early;
    if (ws.piv_state == nullptr) return;
    col.piv = s_piv;
    col.piv_bins = s_piv_bins;
    col.piv_on = &s_piv_on;
    col.piv_early = &s_piv_early;
    if (threadIdx.x < kPivots) { s_piv[threadIdx.x] = kKeyMax; s_piv_bins[threadIdx.x] = 0u; }
    if (threadIdx.x == 0) { s_piv_on = 0u; s_piv_early = 0u; }
}

// Launch-wide pivot ladder (float keys, large k): a CTA's own k-th key only bounds the top-k of ITS rows
// (with 148 CTAs and k = 100 that lets ~1.5 % of all rows through). The first CTA that has sorted its buffer
// publishes kPivots of its keys (ranks cnt-1, (cnt-1)/2, (cnt-1)/4, ...) as a ladder; from then on every CTA
// counts the rows it pushes into the ladder's bins, folds the counts into launch-wide counters at its
// checkpoints, and adopts as threshold the smallest pivot below which the WHOLE launch has already seen k
// rows. Every row is counted at most once and late counts only loosen the bound, so it is always valid.
// Block-wide, CTA-uniform. `sorted` = the buffer was just compacted (sorted, *count entries).
__device__ __forceinline__ void collector_pivot_step(Collector& col, const TopkWorkspace& ws, uint32_t qi, bool sorted) {
    __shared__ uint32_t s_state;
    u64* g_piv = ws.piv_keys + (size_t)qi * kPivots;
    uint32_t* g_cnt = ws.piv_counts + (size_t)qi * kPivots;
    uint32_t* g_state = ws.piv_state + qi;
    if (*reinterpret_cast<volatile uint32_t*>(col.piv_on) == 0u) {
        if (threadIdx.x == 0) {
            s_state = *reinterpret_cast<volatile uint32_t*>(g_state);
            __threadfence();          // the ladder is read only after the flag that publishes it
        }
        col.sync();
        const uint32_t state = s_state;
        if (state == 2u) {            // adopt the published ladder and count what the buffer already holds
            if (threadIdx.x < kPivots) col.piv[threadIdx.x] = __ldcg(g_piv + threadIdx.x);
            col.sync();
            const uint32_t n = min(*col.count, col.cap);
            for (uint32_t i = threadIdx.x; i < n; i += col.nthreads) col.count_pivot(col.keys[i]);
            if (threadIdx.x == 0) *col.piv_on = 1u;
            col.sync();
        } else {
            if (state == 0u && sorted && threadIdx.x == 0) {
                const uint32_t cnt = *col.count;
                if (cnt >= 2u && atomicCAS(g_state, 0u, 1u) == 0u) {
                    for (uint32_t j = 0; j < kPivots; ++j) g_piv[j] = col.keys[(cnt - 1u) >> (kPivots - 1u - j)];
                    __threadfence();
                    atomicExch(g_state, 2u);
                }
            }
            return;                   // adopted at the next checkpoint
        }
    }
    // fold this CTA's bins into the launch-wide counters, then read them back
    if (threadIdx.x < kPivots) {
        const uint32_t v = col.piv_bins[threadIdx.x];
        if (v) {
            atomicAdd(g_cnt + threadIdx.x, v);
            col.piv_bins[threadIdx.x] = 0u;
        }
    }
    col.sync();
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        uint32_t c = lane < kPivots ? __ldcg(g_cnt + lane) : 0u;
#pragma unroll
        for (int o = 1; o < (int)kPivots; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, c, o);
            if ((int)lane >= o) c += v;
        }
        const uint32_t ok = __ballot_sync(0xffffffffu, lane < kPivots && c >= ws.k);
        if (ok != 0u && lane == (uint32_t)(__ffs(ok) - 1)) {
            const u64 bound = col.piv[lane] + 1ull;      // k rows with key <= this pivot exist
            if (bound < *col.thresh) {
                atomicMin(col.thresh, bound);
                atomicMin(ws.g_thresh + qi, bound);
            }
        }
    }
    col.sync();
}

__device__ __forceinline__ void collector_checkpoint(Collector& col, const TopkWorkspace& ws, uint32_t qi,
                                                     uint32_t slack, u64& g_prefetch) {
    if (threadIdx.x == 0 && g_prefetch < col.threshold()) atomicMin(col.thresh, g_prefetch);
    // CTA 0 of a query sorts its first rows early (once) so the pivot ladder exists long before any buffer fills:
    // with k large against the rows a CTA sees, no CTA would otherwise compact before the end of the scan.
    const bool early = col.piv_early != nullptr && blockIdx.x == 0 &&
                       *reinterpret_cast<volatile uint32_t*>(col.piv_early) == 0u;
    const bool need = col.sync_or(
        (threadIdx.x & 31) == 0 && (*reinterpret_cast<volatile uint32_t*>(col.count) + slack > col.cap ||
                                    (early && *reinterpret_cast<volatile uint32_t*>(col.count) >= 64u)));
    if (need) {
        col.compact();
        if (threadIdx.x == 0) {
            if (*col.thresh != kKeyMax) atomicMin(ws.g_thresh + qi, *col.thresh);
            if (col.piv_early != nullptr) *col.piv_early = 1u;
        }
    }
    if (col.piv_on != nullptr) collector_pivot_step(col, ws, qi, need);
    if (threadIdx.x == 0) g_prefetch = ld_volatile_u64(ws.g_thresh + qi);
}

// Block-wide epilogue of a fused scan: publishes this CTA's best k; the last CTA of the
// grid (per query slot) merges every list and writes the sorted result. `s_last` is a
// shared-memory flag owned by the caller.
__device__ __forceinline__ void collector_publish_and_merge(Collector& col, const TopkWorkspace& ws, uint32_t qi,
                                                            int* s_last, unsigned char* big_mem = nullptr,
                                                            uint32_t big_cap = 0) {
    __shared__ uint32_t s_publish;
    const uint32_t kept0 = col.compact();
    // Adopt the launch-wide bound once more and publish only what it still admits (the buffer is sorted):
    // with a tight bound a CTA hands over a few entries instead of k, which is what keeps the merge short.
    if (threadIdx.x == 0) {
        if (*col.thresh != kKeyMax) atomicMin(ws.g_thresh + qi, *col.thresh);
        const u64 g = ld_volatile_u64(ws.g_thresh + qi);
        const u64 T = g == kKeyMax ? kKeyMax : g + 1;      // keys are unique: "<= g" is "< g + 1"
        uint32_t lo = 0, hi = kept0;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (col.keys[mid] < T) lo = mid + 1; else hi = mid;
        }
        s_publish = lo;
    }
    col.sync();
    const uint32_t kept = s_publish;
    const size_t slot = (size_t)qi * gridDim.x + blockIdx.x;
    for (uint32_t i = threadIdx.x; i < kept; i += col.nthreads) {
        ws.cand_keys[slot * ws.k + i] = col.keys[i];
        ws.cand_pays[slot * ws.k + i] = col.pays[i];
    }
    if (threadIdx.x == 0) ws.cand_counts[slot] = kept;
    if (ws.defer_merge) return;   // merged by topk_tree_merge_kernel launches (select.cu)
    __threadfence();
    col.sync();
    if (threadIdx.x == 0) {
        uint32_t ticket = atomicAdd(ws.done + qi, 1u);
        *s_last = (ticket == gridDim.x - 1u);
    }
    col.sync();
    if (!*s_last) return;
    __threadfence();

    // Every list's k-th key was folded into g_thresh: nothing above the smallest of them can be
    // in the global top-k. Keys are unique, so "<= g" is "< g + 1".
    // The last CTA may merge in a larger buffer (e.g. the drained TMA ring): with up to big_cap candidates
    // below the bound, one pass and one sort finish the merge instead of many small windows.
    if (big_mem != nullptr && big_cap > col.cap) {
        col.keys = reinterpret_cast<u64*>(big_mem);
        col.pays = col.keys + big_cap;
        col.cap = big_cap;
    }
    if (threadIdx.x == 0) {
        const u64 g = ld_volatile_u64(ws.g_thresh + qi);
        *col.thresh = g == kKeyMax ? kKeyMax : g + 1;
        *col.count = 0;
    }
    col.sync();
    const uint32_t* counts = ws.cand_counts + (size_t)qi * gridDim.x;
    const u64* gk = ws.cand_keys + (size_t)qi * gridDim.x * ws.k;
    const u64* gp = ws.cand_pays + (size_t)qi * gridDim.x * ws.k;
    const uint32_t stride = ws.k;
    collector_merge_lists(
        col, gridDim.x, ws.k, [counts](uint32_t l) { return __ldcg(counts + l); },
        [gk, stride](uint32_t l, uint32_t i) { return __ldcg(gk + (size_t)l * stride + i); },
        [gp, stride](uint32_t l, uint32_t i) { return __ldcg(gp + (size_t)l * stride + i); });
    const uint32_t total = *col.count;
    for (uint32_t i = threadIdx.x; i < total; i += col.nthreads) {
        ws.out_keys[(size_t)qi * ws.k + i] = col.keys[i];
        ws.out_pays[(size_t)qi * ws.k + i] = col.pays[i];
    }
    if (threadIdx.x == 0) {
        ws.out_counts[qi] = total;
        if (ws.out_err) {
            ws.out_err[qi] = __ldcg(ws.err_row + qi);
            ws.err_row[qi] = kNoError;
        }
        ws.done[qi] = 0u;
        ws.g_thresh[qi] = kKeyMax;
    }
}

}  // namespace vb
