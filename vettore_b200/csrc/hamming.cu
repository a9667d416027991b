// hamming.cu — K3 Hamming scan and K6 sign packing (see hamming.h).
//
// K3 is a pure HBM stream: N * ceil(D/64) * 8 algorithmic bytes per query, one XOR+POPC
// per 8 bytes. Rows are split into 16-byte chunks (8-byte when the word count is odd); a
// group of g = pow2 >= chunks-per-row lanes owns one row, so a warp-level load covers
// 32/g whole consecutive rows (fully coalesced), 8 of them in flight per lane. The 8 group
// popcounts are summed by a transposing butterfly (7 shuffles for g = 8) that leaves one row
// total per lane, so the collector test runs once per 8 loads with every lane busy.
#include "hamming.h"

#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "flat_scan.cuh"
#include "select.h"

namespace vb {

constexpr int kHamThreads = 256;
constexpr int kHamWarps = kHamThreads / 32;
constexpr int kHamUnroll = 8;
#ifndef VB_HAM_CTAS_PER_SM
#define VB_HAM_CTAS_PER_SM 2
#endif
constexpr int kHamCtasPerSm = VB_HAM_CTAS_PER_SM;   // register double-buffered loads: 16 x 16 B in flight per lane
constexpr uint32_t kHamSlackRows = 1024;  // rows scanned per CTA between collector checks

struct HammingParams {
    const u64* codes;          // [n, nw]
    uint32_t n, nw, dims;
    const uint32_t* id_rank;   // optional
    const u64* queries;        // [nq, nw]
    uint32_t g;                // lanes per row (power of two)
    uint32_t sync_every;       // steps between collector checks (power of two)
    uint32_t use_hist;         // 1: thresholds from the distance histograms (large k)
    uint32_t* g_hist;          // [nq, dims + 1] launch-wide histograms, zeroed per launch
    uint32_t cap;
    TopkWorkspace ws;
    u64* dump_keys;            // dump mode: [n] keys / pays, no collector
    u64* dump_pays;
};


// ---- histogram thresholds (large k): distances are integers in [0, dims]. Every CTA counts the rows
// it scans in a shared-memory histogram and folds it, from time to time, into one launch-wide histogram
// per query in global memory; the smallest distance whose launch-wide cumulative count reaches k bounds
// the k-th best of the WHOLE scan (each row is counted once, late counts only make the bound looser),
// so a CTA keeps little more than its share of the final answer and nothing is sorted on the way.
__device__ __forceinline__ void hamming_hist_tighten(Collector& col, uint32_t* hist, uint32_t* g_hist, uint32_t dims,
                                                     uint32_t k, uint32_t* s_td) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t top = min(*s_td, dims);           // bins above the current bound can no longer matter
    for (uint32_t b = threadIdx.x; b <= top; b += col.nthreads) {
        const uint32_t v = hist[b];
        if (v) {
            atomicAdd(g_hist + b, v);
            hist[b] = 0;
        }
    }
    col.sync();
    if (warp == 0) {
        const uint32_t bins = top + 1u, per = (bins + 31u) / 32u;
        const uint32_t b0 = lane * per, b1 = min(bins, (lane + 1u) * per);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; ++b) sum += __ldcg(g_hist + b);
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const uint32_t excl = incl - sum;
        if (excl < k && incl >= k) {     // the k-th smallest distance falls into this lane's bins
            uint32_t cum = excl, b = b0;
            for (; b < b1; ++b) {
                cum += __ldcg(g_hist + b);   // counts only grow: cum reaches k no later than before
                if (cum >= k) break;
            }
            *s_td = min(b, top);
        }
    }
    col.sync();
    const uint32_t td = *s_td;
    if (td <= dims) {
        // keys are (order_key(float(d)) << 32 | id rank): everything above distance td is out
        const u64 bound = (((u64)order_key((float)td)) << 32 | 0xFFFFFFFFull) + 1ull;
        if (threadIdx.x == 0 && bound < *col.thresh) atomicMin(col.thresh, bound);
    }
    col.sync();
}

// Drops buffered entries that no longer beat the threshold (stream compaction, no sort).
template <int PER_THREAD>
__device__ __forceinline__ void collector_filter(Collector& col) {
    const uint32_t n = min(*col.count, col.cap);
    const u64 T = col.threshold();
    u64 kk[PER_THREAD], pp[PER_THREAD];
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
        const uint32_t i = threadIdx.x + j * col.nthreads;
        if (i < n) { kk[j] = col.keys[i]; pp[j] = col.pays[i]; m = j + 1; }
    }
    col.sync();
    if (threadIdx.x == 0) *col.count = 0;
    col.sync();
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j)
        if ((uint32_t)j < m && kk[j] < T) col.push(kk[j], pp[j]);
    col.sync();
}

// Checkpoint in histogram mode: adopt the grid threshold; when the buffer runs full, or on the schedule
// `due`, tighten from the launch-wide histogram and filter; sort only when filtering cannot make room
// (a huge tie at the bound distance).
template <int PER_THREAD>
__device__ __forceinline__ void hamming_hist_checkpoint(Collector& col, const TopkWorkspace& ws, uint32_t qi, uint32_t slack,
                                                        uint32_t* hist, uint32_t* g_hist, uint32_t dims, uint32_t* s_td,
                                                        bool due, u64& g_prefetch) {
    if (threadIdx.x == 0 && g_prefetch < col.threshold()) atomicMin(col.thresh, g_prefetch);
    const bool need = col.sync_or((threadIdx.x & 31) == 0 &&
                                  *reinterpret_cast<volatile uint32_t*>(col.count) + slack > col.cap);
    if (need || due) {
        hamming_hist_tighten(col, hist, g_hist, dims, ws.k, s_td);
        if (need) {
            collector_filter<PER_THREAD>(col);
            if (*reinterpret_cast<volatile uint32_t*>(col.count) + slack > col.cap) col.compact();
        }
        if (threadIdx.x == 0 && *col.thresh != kKeyMax) atomicMin(ws.g_thresh + qi, *col.thresh);
    }
    if (threadIdx.x == 0) g_prefetch = ld_volatile_u64(ws.g_thresh + qi);
}
// Tighten-schedule: checkpoints 1, 2, 4, 8, ... and then every 16th.
__device__ __forceinline__ bool hamming_hist_due(uint32_t ck) { return (ck & (ck - 1u)) == 0u || (ck & 15u) == 0u; }

// Sums v[0..8) over the G = 2^LOGG lanes of a row group, transposing as it goes: each stage halves
// the values a lane still carries (the lane keeps the half its group bit selects and ships the other),
// so 8 row totals cost 4+2+1 shuffles instead of 8 x 3, and they end up spread over the lanes:
// lane c0 of a group holds max(1, 8/G) consecutive totals starting at slot hamming_slot0<LOGG>(c0).
template <int LOGG>
__device__ __forceinline__ void group_transpose_sum(uint32_t (&v)[8], uint32_t c0) {
    int live = 8;
#pragma unroll
    for (int b = LOGG - 1; b >= 0; --b) {
        const uint32_t o = 1u << b;
        if (live > 1) {
            const bool hi = (c0 & o) != 0;
            const int half = live / 2;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < half) {
                    const uint32_t send = hi ? v[i] : v[i + half], keep = hi ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            live = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
}
template <int LOGG>
__device__ __forceinline__ uint32_t hamming_slot0(uint32_t c0) {
    if constexpr (LOGG >= 3) return c0 >> (LOGG - 3);
    else return c0 << (3 - LOGG);
}

// WIDE: 16-byte chunks (nw even); else 8-byte chunks. LOGG: log2 of the lanes that share a row.
template <bool WIDE, int LOGG, int CPS>
__global__ void __launch_bounds__(kHamThreads, CPS) hamming_scan_kernel(const HammingParams p) {
    static_assert(kHamUnroll == 8, "group_transpose_sum carries 8 rows per lane");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;
    __shared__ uint32_t s_td;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    const bool dump = p.dump_keys != nullptr;
    Collector col;
    col.init(smem, &s_thresh, &s_count, p.cap, p.ws.k);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)p.cap * 16);
    if (p.use_hist)
        for (uint32_t b = threadIdx.x; b <= p.dims; b += blockDim.x) hist[b] = 0;
    if (threadIdx.x == 0) s_td = p.dims + 1;
    __syncthreads();

    constexpr uint32_t WPC = WIDE ? 2 : 1;              // words per chunk
    constexpr uint32_t G = 1u << LOGG;                  // lanes per row
    constexpr uint32_t RPW = 32u / G;                   // rows per warp-level load
    constexpr uint32_t OWN = G >= 8 ? 1 : 8 / G;        // row totals a lane ends up with
    const uint32_t cpr = p.nw / WPC;                    // chunks per row
    const uint32_t sub = lane / G, c0 = lane % G;       // row within the load, first chunk
    const uint32_t rem = p.dims & 63u;
    const u64 last_mask = rem ? ((1ull << rem) - 1ull) : ~0ull;   // distances.rs:472-481
    const u64* q = p.queries + (size_t)qi * p.nw;

    // query chunk of this lane (first pass over the row); later passes re-read through L1.
    // Lanes past the row's last chunk keep a zero mask and never load.
    const bool has_chunk = c0 < cpr;
    u64 qa = 0, qb = 0, ma = 0, mb = 0;
    if (has_chunk) {
        qa = q[c0 * WPC];
        ma = (c0 * WPC == p.nw - 1u) ? last_mask : ~0ull;
        if (WIDE) {
            qb = q[c0 * WPC + 1];
            mb = (c0 * WPC + 1u == p.nw - 1u) ? last_mask : ~0ull;
        }
    }
    const uint32_t slot0 = hamming_slot0<LOGG>(c0);
    const bool owner = G > 8 ? (c0 & ((G >> 3) - 1u)) == 0u : true;

    constexpr uint32_t rows_per_step = kHamWarps * RPW * kHamUnroll;
    const uint32_t steps = (p.n + rows_per_step - 1u) / rows_per_step;
    const size_t row_step = (size_t)RPW * p.nw;
    u64 g_prefetch = kKeyMax;
    uint32_t it = 0, ck = 0, td_seen = p.dims + 1u;
    uint32_t* g_hist = p.use_hist ? p.g_hist + (size_t)qi * (p.dims + 1u) : nullptr;

    // All 8 loads of a step are issued back to back, one step AHEAD of the one being counted, so the
    // memory pipe stays full through the popcounts and the collector barriers.
    auto load_step = [&](uint32_t step, u64 (&wa)[kHamUnroll], u64 (&wb)[kHamUnroll]) {
        const uint32_t base = step * rows_per_step + warp * RPW * kHamUnroll + sub;
        const u64* src = p.codes + (size_t)base * p.nw + c0 * WPC;
        const bool full = (step + 1u) * rows_per_step <= p.n;   // CTA-uniform: no per-row bound checks
#pragma unroll
        for (int u = 0; u < kHamUnroll; ++u) {
            wa[u] = 0;       // lanes without a chunk and rows past n: masked / skipped below
            wb[u] = 0;
        }
        if (has_chunk) {
            if (full) {
#pragma unroll
                for (int u = 0; u < kHamUnroll; ++u) {
                    if (WIDE) {
                        const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(src + u * row_step));
                        wa[u] = ((u64)v.y << 32) | v.x;
                        wb[u] = ((u64)v.w << 32) | v.z;
                    } else {
                        wa[u] = __ldg(src + u * row_step);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < kHamUnroll; ++u) {
                    if (base + u * RPW < p.n) {
                        if (WIDE) {
                            const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(src + u * row_step));
                            wa[u] = ((u64)v.y << 32) | v.x;
                            wb[u] = ((u64)v.w << 32) | v.z;
                        } else {
                            wa[u] = __ldg(src + u * row_step);
                        }
                    }
                }
            }
        }
    };

    auto count_step = [&](uint32_t step, const u64 (&wa)[kHamUnroll], const u64 (&wb)[kHamUnroll]) {
        const uint32_t base = step * rows_per_step + warp * RPW * kHamUnroll + sub;
        uint32_t dist[kHamUnroll];
#pragma unroll
        for (int u = 0; u < kHamUnroll; ++u) {
            dist[u] = __popcll((wa[u] ^ qa) & ma);
            if (WIDE) dist[u] += __popcll((wb[u] ^ qb) & mb);
        }
        // rows longer than G chunks (only when cpr > 32): remaining passes
        if (LOGG == 5) {
            for (uint32_t c = c0 + G; c < cpr; c += G) {
#pragma unroll
                for (int u = 0; u < kHamUnroll; ++u) {
                    const uint32_t row = base + u * RPW;
                    if (row >= p.n) continue;
                    const u64* src = p.codes + (size_t)row * p.nw + c * WPC;
                    u64 xa = __ldg(src) ^ __ldg(q + c * WPC), xb = 0;
                    if (WIDE) xb = __ldg(src + 1) ^ __ldg(q + c * WPC + 1);
                    if (c * WPC == p.nw - 1u) xa &= last_mask;
                    if (WIDE && c * WPC + 1u == p.nw - 1u) xb &= last_mask;
                    dist[u] += __popcll(xa) + (WIDE ? __popcll(xb) : 0);
                }
            }
        }
        group_transpose_sum<LOGG>(dist, c0);

        if (owner) {
            const u64 T = dump ? kKeyMax : col.threshold();
#pragma unroll
            for (uint32_t j = 0; j < OWN; ++j) {
                const uint32_t row = base + (slot0 + j) * RPW;
                if (row >= p.n) continue;
                const uint32_t d = dist[j];
                if (p.use_hist && d <= td_seen) atomicAdd(&hist[d], 1u);
                const float raw = (float)d;                             // distances.rs:436
                const uint32_t rk = order_key(raw);
                if (rk > (uint32_t)(T >> 32)) continue;
                const uint32_t idr = p.id_rank ? __ldg(p.id_rank + row) : row;
                const u64 key = ((u64)rk << 32) | idr;
                const u64 pay = ((u64)__float_as_uint(raw) << 32) | row;
                if (dump) {
                    p.dump_keys[row] = key;
                    p.dump_pays[row] = pay;
                } else if (key < T) {
                    col.push(key, pay);
                }
            }
        }
        if (!dump && (it & (p.sync_every - 1)) == p.sync_every - 1) {
            if (p.use_hist) {
                ++ck;
                hamming_hist_checkpoint<16>(col, p.ws, qi, p.sync_every * rows_per_step, hist, g_hist, p.dims, &s_td,
                                            hamming_hist_due(ck), g_prefetch);
                td_seen = s_td;
            } else {
                collector_checkpoint(col, p.ws, qi, p.sync_every * rows_per_step, g_prefetch);
            }
        }
        ++it;
    };

    u64 wa0[kHamUnroll], wb0[kHamUnroll], wa1[kHamUnroll], wb1[kHamUnroll];
    uint32_t step = blockIdx.x;
    if (step < steps) load_step(step, wa0, wb0);
    while (step < steps) {
        uint32_t next = step + gridDim.x;
        if (next < steps) load_step(next, wa1, wb1);
        count_step(step, wa0, wb0);
        step = next;
        if (step >= steps) break;
        next = step + gridDim.x;
        if (next < steps) load_step(next, wa0, wb0);
        count_step(step, wa1, wb1);
        step = next;
    }
    if (dump) return;
    if (p.use_hist) {   // leave only what the launch-wide bound still admits
        __syncthreads();
        hamming_hist_tighten(col, hist, g_hist, p.dims, p.ws.k, &s_td);
        collector_filter<16>(col);
    }
    collector_publish_and_merge(col, p.ws, qi, &s_last);
}

using HammingKernel = void (*)(const HammingParams);
template <bool WIDE, int CPS>
static HammingKernel hamming_kernel_for(uint32_t logg) {
    switch (logg) {
        case 0: return hamming_scan_kernel<WIDE, 0, CPS>;
        case 1: return hamming_scan_kernel<WIDE, 1, CPS>;
        case 2: return hamming_scan_kernel<WIDE, 2, CPS>;
        case 3: return hamming_scan_kernel<WIDE, 3, CPS>;
        case 4: return hamming_scan_kernel<WIDE, 4, CPS>;
        default: return hamming_scan_kernel<WIDE, 5, CPS>;
    }
}

// ---------------------------------------------------------------------------------------
// K3 stream variant: the code matrix is contiguous, so one producer lane streams 32 KB tiles
// HBM -> shared memory with TMA bulk copies through a multi-stage ring (as flat_stream_kernel),
// which keeps far more bytes in flight per SM than registers can, and 16 consumer warps XOR/POPC
// them from shared memory with the same lane <-> chunk mapping as above. A warp takes 4 warp-level
// loads from each tile and sums the 8 row groups of a PAIR of tiles with one transposing butterfly.
constexpr int kHsWarps = 16;
constexpr int kHsThreads = kHsWarps * 32 + 32;
constexpr int kHsLoads = 4;              // warp-level loads per warp per tile
constexpr int kHsMaxStages = 6;

struct HammingStreamGeom {
    uint32_t stages, tile_bytes, tile_rows;
    uint32_t use_hist;                   // 1: histogram thresholds (large k)
};

template <int LOGG>   // 16-byte chunks only (even word count): bulk copies move 16-byte multiples
__global__ void __launch_bounds__(kHsThreads, 1)
hamming_stream_kernel(const HammingParams p, const HammingStreamGeom geom) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kHsMaxStages], empty_bar[kHsMaxStages];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;
    __shared__ uint32_t s_td;            // histogram-derived distance bound

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    unsigned char* ring = smem;
    unsigned char* col_mem = smem + (size_t)geom.stages * geom.tile_bytes;
    uint32_t* hist = reinterpret_cast<uint32_t*>(col_mem + (size_t)p.cap * 16);

    Collector col;
    col.init(col_mem, &s_thresh, &s_count, p.cap, p.ws.k, kHsWarps * 32, 1);
    if (geom.use_hist)
        for (uint32_t b = threadIdx.x; b <= p.dims; b += blockDim.x) hist[b] = 0;
    if (threadIdx.x == 0) {
        s_td = p.dims + 1;
        for (uint32_t s = 0; s < geom.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kHsWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t row_bytes = p.nw * 8u;
    const uint32_t num_tiles = (p.n + geom.tile_rows - 1u) / geom.tile_rows;

    if (warp == kHsWarps) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const uint32_t row0 = tile * geom.tile_rows;
                const uint32_t bytes = min(geom.tile_rows, p.n - row0) * row_bytes;
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                tma_bulk_g2s(ring + (size_t)s * geom.tile_bytes, p.codes + (size_t)row0 * p.nw, bytes, &full_bar[s]);
                if (++s == geom.stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    constexpr uint32_t G = 1u << LOGG, RPW = 32u / G;          // lanes per row, rows per warp-level load
    constexpr uint32_t OWN = G >= 8 ? 1 : 8 / G;
    const uint32_t cpr = p.nw / 2u;
    const uint32_t sub = lane / G, c0 = lane % G;
    const uint32_t rem = p.dims & 63u;
    const u64 last_mask = rem ? ((1ull << rem) - 1ull) : ~0ull;
    const u64* q = p.queries + (size_t)qi * p.nw;
    // Lanes past the row's last chunk (chunk counts that are not a power of two) read chunk 0 with a
    // zero mask, so the shared-memory loads below need no predicate.
    const bool has_chunk = c0 < cpr;
    const uint32_t ce = has_chunk ? c0 : 0u;
    const u64 qa = q[ce * 2], qb = q[ce * 2 + 1];
    const u64 ma = !has_chunk ? 0ull : (ce * 2 == p.nw - 1u) ? last_mask : ~0ull;
    const u64 mb = !has_chunk ? 0ull : (ce * 2 + 1u == p.nw - 1u) ? last_mask : ~0ull;
    const uint32_t qa_lo = (uint32_t)qa, qa_hi = (uint32_t)(qa >> 32), qb_lo = (uint32_t)qb, qb_hi = (uint32_t)(qb >> 32);
    const uint32_t ma_lo = (uint32_t)ma, ma_hi = (uint32_t)(ma >> 32), mb_lo = (uint32_t)mb, mb_hi = (uint32_t)(mb >> 32);
    const uint32_t slot0 = hamming_slot0<LOGG>(c0);
    const bool owner = G > 8 ? (c0 & ((G >> 3) - 1u)) == 0u : true;
    const uint32_t trow0 = warp * (RPW * kHsLoads) + sub;      // this lane's first row within a tile
    const uint32_t lane_off = trow0 * row_bytes + ce * 16u;    // ... and its byte offset
    const uint32_t load_step = RPW * row_bytes;

    uint32_t stage = 0, phase = 0, ck = 0, pairs = 0, td_seen = p.dims + 1u;
    uint32_t* g_hist = geom.use_hist ? p.g_hist + (size_t)qi * (p.dims + 1u) : nullptr;
    u64 g_prefetch = kKeyMax;

    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += 2u * gridDim.x) {
        uint32_t dist[8];
        const uint32_t tile_b = tile + gridDim.x;
        const bool has_b = tile_b < num_tiles;
#pragma unroll
        for (int H = 0; H < 2; ++H) {
            if (H == 0 || has_b) {
                mbar_wait(&full_bar[stage], phase);
                const unsigned char* src = ring + stage * geom.tile_bytes + lane_off;
                uint4 v[kHsLoads];
#pragma unroll
                for (int u = 0; u < kHsLoads; ++u) v[u] = *reinterpret_cast<const uint4*>(src + u * load_step);
#pragma unroll
                for (int u = 0; u < kHsLoads; ++u)
                    dist[H * 4 + u] = __popc((v[u].x ^ qa_lo) & ma_lo) + __popc((v[u].y ^ qa_hi) & ma_hi) +
                                      __popc((v[u].z ^ qb_lo) & mb_lo) + __popc((v[u].w ^ qb_hi) & mb_hi);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == geom.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            } else {
#pragma unroll
                for (int u = 0; u < kHsLoads; ++u) dist[H * 4 + u] = 0;
            }
        }
        group_transpose_sum<LOGG>(dist, c0);
        if (owner) {
            const u64 T = col.threshold();
#pragma unroll
            for (uint32_t j = 0; j < OWN; ++j) {
                const uint32_t slot = slot0 + j;                 // (tile of the pair) * 4 + load
                const uint32_t t = (slot & 4u) ? tile_b : tile;
                const uint32_t row = t * geom.tile_rows + trow0 + (slot & 3u) * RPW;
                if (t >= num_tiles || row >= p.n) continue;
                const uint32_t d = dist[j];
                if (geom.use_hist && d <= td_seen) atomicAdd(&hist[d], 1u);
                const float raw = (float)d;
                const uint32_t rk = order_key(raw);
                if (rk > (uint32_t)(T >> 32)) continue;
                const uint32_t idr = p.id_rank ? __ldg(p.id_rank + row) : row;
                const u64 key = ((u64)rk << 32) | idr;
                if (key < T) col.push(key, ((u64)__float_as_uint(raw) << 32) | row);
            }
        }
        if ((pairs & (p.sync_every - 1)) == p.sync_every - 1) {
            if (geom.use_hist) {
                ++ck;
                hamming_hist_checkpoint<8>(col, p.ws, qi, p.sync_every * 2u * geom.tile_rows, hist, g_hist, p.dims, &s_td,
                                           hamming_hist_due(ck), g_prefetch);
                td_seen = s_td;
            } else {
                collector_checkpoint(col, p.ws, qi, p.sync_every * 2u * geom.tile_rows, g_prefetch);
            }
        }
        ++pairs;
    }
    if (geom.use_hist) {   // leave only what the launch-wide bound still admits
        col.sync();
        hamming_hist_tighten(col, hist, g_hist, p.dims, p.ws.k, &s_td);
        collector_filter<8>(col);
    }
    collector_publish_and_merge(col, p.ws, qi, &s_last);
}

using HammingStreamKernel = void (*)(const HammingParams, const HammingStreamGeom);
static HammingStreamKernel hamming_stream_kernel_for(uint32_t logg) {
    switch (logg) {
        case 1: return hamming_stream_kernel<1>;
        case 2: return hamming_stream_kernel<2>;
        case 3: return hamming_stream_kernel<3>;
        case 4: return hamming_stream_kernel<4>;
        default: return hamming_stream_kernel<5>;
    }
}

static uint32_t pow2_at_least(uint32_t v, uint32_t lo) {
    uint32_t p = lo;
    while (p < v) p <<= 1;
    return p;
}

static Status prepare_topk_ws(SearchCtx& ctx, HammingParams& p, uint32_t nq, uint32_t k, uint32_t grid_x,
                              cudaStream_t stream) {
    VB_TRY(ctx.arm_ctrl(nq));
    const size_t lists = (size_t)nq * grid_x;
    VB_TRY(ctx.cand_keys.reserve(lists * k * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve(lists * k * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve(lists * sizeof(uint32_t)));
    VB_TRY(ctx.out_keys.reserve((size_t)nq * k * sizeof(u64)));
    VB_TRY(ctx.result.reserve((size_t)nq * k * sizeof(u64) + (size_t)nq * 8));
    p.ws.k = k;
    p.ws.cand_keys = ctx.cand_keys.as<u64>();
    p.ws.cand_pays = ctx.cand_pays.as<u64>();
    p.ws.cand_counts = ctx.cand_counts.as<uint32_t>();
    p.ws.done = ctx.done();
    p.ws.g_thresh = ctx.g_thresh();
    p.ws.out_keys = ctx.out_keys.as<u64>();
    p.ws.out_pays = ctx.result.as<u64>();
    p.ws.out_counts = reinterpret_cast<uint32_t*>(ctx.result.as<u64>() + (size_t)nq * k);
    p.ws.err_row = nullptr;
    p.ws.out_err = nullptr;
    p.ws.defer_merge = merge_tree_wanted(grid_x, k) ? 1u : 0u;
    if (p.use_hist) {
        const size_t bytes = (size_t)nq * (p.dims + 1u) * sizeof(uint32_t);
        VB_TRY(ctx.hist.reserve(bytes));
        p.g_hist = ctx.hist.as<uint32_t>();
        VB_CUDA(cudaMemsetAsync(p.g_hist, 0, bytes, stream));
    }
    return Status::Ok();
}

static Status launch_hamming(SearchCtx& ctx, HammingParams p, uint32_t nq, uint32_t k, bool dump,
                             cudaStream_t stream) {
    const bool wide = (p.nw % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.codes) & 15) == 0);
    const uint32_t cpr = wide ? p.nw / 2 : p.nw;
    p.g = std::min<uint32_t>(32, pow2_at_least(cpr, 1));
    int dev = 0, sms = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    // ---- TMA-staged stream kernel: whole rows fit one lane group, 16-byte chunks (bulk copies move
    // 16-byte multiples, so an even word count), at least 2 lanes per row (a pair of tiles of
    // 128-bit codes would be 4096 rows: too much collector slack). VB_HAMMING_NO_STREAM=1 disables it.
    if (!dump && wide && cpr <= 32 && p.g >= 2 && p.dims <= 4096 && !std::getenv("VB_HAMMING_NO_STREAM")) {
        HammingStreamGeom geom{};
        const uint32_t rpl = 32 / p.g;
        geom.tile_rows = kHsWarps * kHsLoads * rpl;
        geom.tile_bytes = (uint32_t)(((size_t)geom.tile_rows * p.nw * 8 + 127) & ~(size_t)127);
        const uint32_t pair_rows = 2 * geom.tile_rows;
        p.sync_every = std::max<uint32_t>(1, kHamSlackRows / pair_rows);      // powers of two
        const uint32_t slack = p.sync_every * pair_rows;
        p.cap = pow2_at_least(2 * k + slack, 256);
        geom.use_hist = (k > 64 && p.cap <= 8u * kHsWarps * 32u) ? 1u : 0u;   // collector_filter<8> holds the buffer in registers
        p.use_hist = geom.use_hist;
        const size_t col_bytes = (size_t)p.cap * 16 + (geom.use_hist ? (size_t)(p.dims + 1) * 4 + 16 : 0);
        const size_t budget = 200 * 1024;
        if (col_bytes + 3 * (size_t)geom.tile_bytes <= budget) {
            geom.stages = (uint32_t)std::min<size_t>(kHsMaxStages, (budget - col_bytes) / geom.tile_bytes);
            const size_t smem = (size_t)geom.stages * geom.tile_bytes + col_bytes;
            uint32_t logg = 0;
            while ((1u << logg) < p.g) ++logg;
            HammingStreamKernel kernel = hamming_stream_kernel_for(logg);
            VB_TRY(ensure_dynamic_smem_for(kernel, budget));
            const uint32_t tiles = (p.n + geom.tile_rows - 1) / geom.tile_rows;
            uint32_t grid_x = std::min<uint32_t>(tiles, (uint32_t)sms);
            if (const char* e = std::getenv("VB_HAMMING_MAX_GRID"))
                grid_x = std::max<uint32_t>(1, std::min<uint32_t>(grid_x, (uint32_t)std::atoi(e)));
            VB_TRY(prepare_topk_ws(ctx, p, nq, k, grid_x, stream));
            kernel<<<dim3(grid_x, nq), kHsThreads, smem, stream>>>(p, geom);
            VB_CUDA(cudaGetLastError());
            if (p.ws.defer_merge) VB_TRY(run_merge_tree(p.ws, nq, grid_x, ctx.sort_tmp, stream));
            return Status::Ok();
        }
        p.use_hist = 0;
    }

    const uint32_t rows_per_step = kHamWarps * (32 / p.g) * kHamUnroll;
    const uint32_t kk = dump ? 1 : k;
    p.sync_every = std::max<uint32_t>(1, kHamSlackRows / rows_per_step);  // both are powers of two
    p.cap = pow2_at_least(2 * kk + p.sync_every * rows_per_step, 256);
    p.use_hist = (!dump && k > 64 && p.dims <= 4096 && p.cap <= 16 * kHamThreads) ? 1u : 0u;
    const size_t smem = (size_t)p.cap * 16 + (p.use_hist ? (size_t)(p.dims + 1) * 4 + 16 : 0);
    uint32_t logg = 0;
    while ((1u << logg) < p.g) ++logg;
    int cps = kHamCtasPerSm;
    if (const char* e = std::getenv("VB_HAM_CTAS")) cps = std::atoi(e);
    HammingKernel kernel = cps == 3 ? (wide ? hamming_kernel_for<true, 3>(logg) : hamming_kernel_for<false, 3>(logg))
                                    : (wide ? hamming_kernel_for<true, 2>(logg) : hamming_kernel_for<false, 2>(logg));
    VB_TRY(ensure_dynamic_smem_for(kernel, smem));
    int per_sm = 0;
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kHamThreads, smem));
    if (per_sm < 1) return Status::Cuda("hamming kernel does not fit on an SM");
    const uint32_t steps = (p.n + rows_per_step - 1) / rows_per_step;
    uint32_t grid_x = std::min<uint32_t>(steps, (uint32_t)(sms * per_sm));
    if (const char* e = std::getenv("VB_HAMMING_MAX_GRID"))   // tests: few CTAs, many prune rounds each
        grid_x = std::max<uint32_t>(1, std::min<uint32_t>(grid_x, (uint32_t)std::atoi(e)));
    if (!dump) {
        VB_TRY(prepare_topk_ws(ctx, p, nq, k, grid_x, stream));
    } else {
        p.ws = TopkWorkspace{};
        p.ws.k = 1;
    }
    kernel<<<dim3(grid_x, nq), kHamThreads, smem, stream>>>(p);
    VB_CUDA(cudaGetLastError());
    if (!dump && p.ws.defer_merge) VB_TRY(run_merge_tree(p.ws, nq, grid_x, ctx.sort_tmp, stream));
    return Status::Ok();
}

Status hamming_scan_device(SearchCtx& ctx, const u64* d_codes, uint32_t n, uint32_t nw, uint32_t dims,
                           const uint32_t* d_id_rank, const u64* d_queries, uint32_t nq, uint32_t k,
                           cudaStream_t stream) {
    if (n == 0 || k == 0 || nq == 0) return Status::Cuda("empty hamming scan");
    if (k > (uint32_t)kMaxFusedK) return Status::Cuda("k beyond fused collector");
    HammingParams p{};
    p.codes = d_codes;
    p.n = n;
    p.nw = nw;
    p.dims = dims;
    p.id_rank = d_id_rank;
    p.queries = d_queries;
    cudaStream_t saved = ctx.stream;
    ctx.stream = stream;
    Status s = launch_hamming(ctx, p, nq, std::min(k, n), false, stream);
    ctx.stream = saved;
    if (!s.ok()) ctx.poison();
    return s;
}

Status hamming_dump_sorted(SearchCtx& ctx, const u64* d_codes, uint32_t n, uint32_t nw, uint32_t dims,
                           const uint32_t* d_rank, const u64* d_query, cudaStream_t stream) {
    VB_TRY(ctx.dump_keys.reserve((size_t)n * sizeof(u64)));
    VB_TRY(ctx.dump_pays.reserve((size_t)n * sizeof(u64)));
    VB_TRY(ctx.dump_keys2.reserve((size_t)n * sizeof(u64)));
    VB_TRY(ctx.dump_pays2.reserve((size_t)n * sizeof(u64)));
    HammingParams p{};
    p.codes = d_codes;
    p.n = n;
    p.nw = nw;
    p.dims = dims;
    p.id_rank = d_rank;
    p.queries = d_query;
    p.dump_keys = ctx.dump_keys.as<u64>();
    p.dump_pays = ctx.dump_pays.as<u64>();
    size_t tmp_bytes = 0;
    VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx.dump_keys.as<u64>(), ctx.dump_keys2.as<u64>(),
                                            ctx.dump_pays.as<u64>(), ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64,
                                            stream));
    VB_TRY(ctx.sort_tmp.reserve(tmp_bytes));
    VB_TRY(launch_hamming(ctx, p, 1, 1, true, stream));
    // keys are distance << 32 | id rank: distances need at most 32 bits, so the sort covers bits 0..63
    VB_CUDA(cub::DeviceRadixSort::SortPairs(ctx.sort_tmp.p, tmp_bytes, ctx.dump_keys.as<u64>(),
                                            ctx.dump_keys2.as<u64>(), ctx.dump_pays.as<u64>(),
                                            ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64, stream));
    return Status::Ok();
}

Status hamming_top_k_resident(SearchCtx& ctx, const u64* d_codes, size_t n, size_t nw, size_t dims,
                              const uint32_t* d_rank, const uint64_t* h_query, size_t k,
                              std::vector<uint32_t>* rows, std::vector<float>* values) {
    rows->clear();
    values->clear();
    k = std::min(k, n);
    if (k == 0) return Status::Ok();
    VB_TRY(ctx.h_queries.reserve(nw * sizeof(u64)));
    VB_TRY(ctx.queries.reserve(nw * sizeof(u64)));
    std::memcpy(ctx.h_queries.p, h_query, nw * sizeof(u64));
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, ctx.h_queries.p, nw * sizeof(u64), cudaMemcpyHostToDevice, ctx.stream));
    VB_TRY(ctx.h_result.reserve(k * sizeof(u64) + 8));
    const u64* h_pays = ctx.h_result.as<u64>();
    size_t count = k;
    if (k <= (size_t)kMaxFusedK) {
        VB_TRY(hamming_scan_device(ctx, d_codes, (uint32_t)n, (uint32_t)nw, (uint32_t)dims, d_rank,
                                   ctx.queries.as<u64>(), 1, (uint32_t)k, ctx.stream));
        VB_CUDA(cudaMemcpyAsync(ctx.h_result.p, ctx.result.p, k * sizeof(u64) + 4, cudaMemcpyDeviceToHost,
                                ctx.stream));
        VB_CUDA(cudaStreamSynchronize(ctx.stream));
        count = *reinterpret_cast<const uint32_t*>(h_pays + k);
    } else {
        VB_TRY(hamming_dump_sorted(ctx, d_codes, (uint32_t)n, (uint32_t)nw, (uint32_t)dims, d_rank,
                                   ctx.queries.as<u64>(), ctx.stream));
        VB_CUDA(cudaMemcpyAsync(ctx.h_result.p, ctx.dump_pays2.p, k * sizeof(u64), cudaMemcpyDeviceToHost,
                                ctx.stream));
        VB_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    rows->resize(count);
    values->resize(count);
    for (size_t i = 0; i < count; ++i) {
        uint32_t bits = (uint32_t)(h_pays[i] >> 32);
        std::memcpy(&(*values)[i], &bits, 4);
        (*rows)[i] = (uint32_t)h_pays[i];
    }
    return Status::Ok();
}

Status hamming_top_k_host(SearchCtx& ctx, const uint64_t* h_codes, size_t n, size_t nw, size_t dims,
                          const uint32_t* h_rank, const uint64_t* h_query, size_t k, std::vector<uint32_t>* rows,
                          std::vector<float>* values) {
    VB_TRY(ctx.staging.reserve(n * nw * sizeof(u64)));
    VB_TRY(ctx.staging_rank.reserve(n * sizeof(uint32_t)));
    VB_CUDA(cudaMemcpyAsync(ctx.staging.p, h_codes, n * nw * sizeof(u64), cudaMemcpyHostToDevice, ctx.stream));
    VB_CUDA(cudaMemcpyAsync(ctx.staging_rank.p, h_rank, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx.stream));
    VB_CUDA(cudaStreamSynchronize(ctx.stream));  // sources are pageable caller memory
    return hamming_top_k_resident(ctx, ctx.staging.as<u64>(), n, nw, dims, ctx.staging_rank.as<uint32_t>(), h_query,
                                  k, rows, values);
}

// ---------------------------------------------------------------------------------- K6
__global__ void __launch_bounds__(256) sign_pack_kernel(const float* rows, size_t row_stride, uint32_t n,
                                                        uint32_t dims, uint32_t nw, u64* codes) {
    // One warp per (row, 64-coordinate word) pair; two ballots build the word.
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t total = (size_t)n * nw;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t item = warp_global; item < total; item += warps) {
        const uint32_t row = (uint32_t)(item / nw), w = (uint32_t)(item % nw);
        const float* src = rows + (size_t)row * row_stride + (size_t)w * 64;
        const uint32_t i0 = w * 64 + lane, i1 = i0 + 32;
        // padding coordinates (>= dims) must not set bits: distances.rs:414-421
        const bool b0 = i0 < dims && src[lane] >= 0.0f;
        const bool b1 = i1 < dims && src[lane + 32] >= 0.0f;
        const uint32_t lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
        if (lane == 0) codes[item] = ((u64)hi << 32) | lo;
    }
}

Status sign_pack_device(const float* d_rows, size_t row_stride, uint32_t n, uint32_t dims, u64* d_codes,
                        cudaStream_t stream) {
    if (n == 0 || dims == 0) return Status::Ok();
    const uint32_t nw = (dims + 63) / 64;
    const size_t total_warps = (size_t)n * nw;
    const unsigned blocks = (unsigned)std::min<size_t>((total_warps + 7) / 8, 148 * 16);
    sign_pack_kernel<<<blocks, 256, 0, stream>>>(d_rows, row_stride, n, dims, nw, d_codes);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

}  // namespace vb

// ---------------------------------------------------------------------------------------
// Measurement entry (not part of the drop-in ABI): K3 over n synthetic codes generated on the
// device (splitmix64 of the word index), `iters` timed scans with CUDA events. Used by
// tools/bench_hamming.py for the C4 shape (100M x 1024 bits) where no host copy can exist.
namespace vb {
__global__ void fill_codes_kernel(u64* codes, size_t total, u64 seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        u64 z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        codes[i] = z ^ (z >> 31);
    }
}
}  // namespace vb

extern "C" int vb_debug_hamming_bench(size_t n, size_t dims, size_t k, int iters, float* ms_per_scan,
                                      uint32_t* top_rows, float* top_dist) {
    using namespace vb;
    const size_t nw = (dims + 63) / 64;
    u64 *d_codes = nullptr, *d_q = nullptr;
    if (cudaMalloc(&d_codes, n * nw * sizeof(u64)) != cudaSuccess) return -1;
    cudaMalloc(&d_q, nw * sizeof(u64));
    fill_codes_kernel<<<148 * 8, 256>>>(d_codes, n * nw, 1);
    fill_codes_kernel<<<1, 32>>>(d_q, nw, 0xABCDEF);
    SearchCtx ctx;
    cudaGetDevice(&ctx.device);
    cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking);
    cudaDeviceSynchronize();
    int rc = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3 && rc == 0; ++i)
        if (!hamming_scan_device(ctx, d_codes, (uint32_t)n, (uint32_t)nw, (uint32_t)dims, nullptr, d_q, 1, (uint32_t)k,
                                 ctx.stream).ok()) rc = -2;
    cudaEventRecord(e0, ctx.stream);
    for (int i = 0; i < iters && rc == 0; ++i)
        if (!hamming_scan_device(ctx, d_codes, (uint32_t)n, (uint32_t)nw, (uint32_t)dims, nullptr, d_q, 1, (uint32_t)k,
                                 ctx.stream).ok()) rc = -2;
    cudaEventRecord(e1, ctx.stream);
    if (cudaStreamSynchronize(ctx.stream) != cudaSuccess) rc = -3;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_scan = ms / (iters > 0 ? iters : 1);
    if (rc == 0) {
        std::vector<u64> pays(k);
        cudaMemcpy(pays.data(), ctx.result.p, k * sizeof(u64), cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < k; ++i) {
            uint32_t bits = (uint32_t)(pays[i] >> 32);
            std::memcpy(&top_dist[i], &bits, 4);
            top_rows[i] = (uint32_t)pays[i];
        }
    }
    ctx.destroy();
    cudaFree(d_codes);
    cudaFree(d_q);
    return rc;
}
