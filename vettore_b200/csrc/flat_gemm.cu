// flat_gemm.cu — K2: batched exact flat scan on the tensor cores. For a batch of queries the
// scan is a dense contraction S[row][query] = rows . queries^T (reference flat.rs:104-118 run
// once per query), so it goes to tcgen05 as 3xTF32 (fp32-accurate, tc.cuh) with the per-query
// top-k fused into the epilogue: the N x Q score matrix never exists in memory.
//
// Geometry: UMMA M = 128 rows x N = 256 queries x K = 8, fp32 accumulator in TMEM (256 columns).
// CTA c serves query block (c mod QB) over a contiguous range of row tiles. Per 32-dim K chunk:
//   producer warp   TMA: A chunk [128 rows x 32] (16 KB, tensor map, 128-byte swizzle) + the query block's
//                   hi|lo chunk as ONE 64 KB bulk copy of an image pre-swizzled by split_queries_kernel
//                   (512 separate 128-byte tensor-map rows per chunk made the TMA unit the bottleneck)
//   split warps     A chunk -> (hi, lo) -> TMEM (tcgen05.st), double buffered
//   MMA warp        12 tcgen05.mma kind::tf32 (hi.hi, hi.lo, lo.hi x 4 k-steps), B from shared memory
//   epilogue warps  per finished tile: tcgen05.ld 8 x 32 columns, score -> (rank key, id rank) ->
//                   compare with the query's running threshold (shared memory) -> append to the
//                   query's candidate list; lists are compacted to their best k by a warp when they
//                   could overflow, which also tightens the threshold.
// A final kernel merges, per query, the lists of the CTAs that shared its query block.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "flat_gemm.h"
#include "flat_scan.cuh"
#include "tc.cuh"

namespace vb {

constexpr int kGmEpiWarps = 8;           // two per TMEM lane quarter, half of the query columns each
constexpr int kGmSplitWarp0 = kGmEpiWarps, kGmProducerWarp = kGmEpiWarps + 4, kGmMmaWarp = kGmEpiWarps + 5;
constexpr int kGmThreads = (kGmMmaWarp + 1) * 32;   // warps 0-7 epilogue, 8-11 split, 12 producer, 13 MMA
constexpr int kGmTile = 128;             // rows per tile (UMMA M)
constexpr int kGmN = 256;                // queries per block (UMMA N)
constexpr int kGmStages = 2;
constexpr uint32_t kGmStageBytes = 16384 + 2 * 32768;   // A chunk + B hi + B lo
constexpr uint32_t kGmListSmall = 256;   // candidate list capacity per (CTA, query), k' <= 64
constexpr uint32_t kGmListLarge = 512;   // ... larger k': a cut back to k' must leave room for many tiles
constexpr uint32_t kGmACol = 256;        // TMEM: D at [0,256), A buffer u at 256 + 64 u: hi [0,32) lo [32,64)

struct GemmParams {
    uint32_t n, dims, nq, k;
    int metric;                    // kCosine / kInnerProduct / kNegativeInnerProduct / kL2 / kL2Squared
    // rank of a score: fma(dot, rank_scale, row bias). Dot family: scale -1, bias 1 (cosine: 1 - dot) or 0. L2 family:
    // scale -2, bias |row|^2 from the row-norm mirror, i.e. |x|^2 - 2 q.x = |x - q|^2 - |q|^2 — the squared distance up
    // to a per-query constant, which is all a per-query filter needs (distances.rs:140-152 via the exact re-scoring).
    float rank_scale;
    const float* row_norm2;        // [n] or null
    uint32_t qblocks, ranges;      // query blocks, row ranges (qblocks * ranges CTAs — or CTA pairs — do work)
    uint32_t pair;                 // 1: the single-pass kernel runs as CTA pairs (cta_group::2)
    uint32_t nblk;                 // queries per block: 256, or 128 for the CTA-pair form
    const uint32_t* id_rank;       // [n] or null
    uint32_t list_cap;             // kGmListSmall / kGmListLarge
    u64* list_keys;                // [cta][256][list_cap]
    u64* list_pays;
    uint32_t* list_counts;         // [cta][256]
    uint32_t* bad;                 // set when a non-finite score shows up (caller falls back)
    // optional: per query of this launch, the merged best-k list of a PRE-PASS over a sample of the rows. Its
    // k-th key bounds the final k-th key from above, so every CTA starts with a tight filter instead of an open one.
    const u64* init_keys;          // [nq][k] ascending, or null
    const uint32_t* init_counts;   // [nq]
    // single-pass kernel: the pre-pass as a dense dump. score_dump != null: the rank of every (query, sample row) goes
    // to score_dump[query][dump_stride] (coalesced, no lists, no filter) and gemm1_sample_select_kernel finds each
    // query's k-th smallest; init_rank != null: the main pass starts from those rank bounds.
    float* score_dump;
    uint32_t dump_stride;
    const float* init_rank;        // [nq] or null; +inf = no bound
    uint32_t debug;                // timing experiments only (VB_GEMM_DEBUG): 1 / 2 skip operand loads, 4 skip split, 8 skip the filter, 16 skip MMA
};

__device__ __forceinline__ float rank_from_key(u64 key) {
    const uint32_t kbits = (uint32_t)(key >> 32);
    const uint32_t bits = (kbits & 0x80000000u) ? (kbits ^ 0x80000000u) : ~kbits;
    return __uint_as_float(bits);
}

// Cuts one list (<= 32 * PL entries, PL per lane in registers) back to its best k WITHOUT sorting:
// the k-th smallest key is found by bisection on the key VALUE (32-bit rank word first, then the id
// word among rank ties; one predicated compare per entry and one REDUX per step), then the entries
// up to it are stream-compacted to the front of the list. Returns the k-th key (kKeyMax when the
// list holds fewer than k entries). The merge that consumes the lists does not need them sorted.
template <int PL>
__device__ __noinline__ u64 warp_select_list(u64* keys, u64* pays, uint32_t count, uint32_t k, int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    if (count < k) return kKeyMax;
    u64 lk[PL], lp[PL];
    uint32_t hmin = 0xffffffffu, hmax = 0u;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const uint32_t e = lane + 32u * i;
        const bool valid = e < count;
        lk[i] = valid ? keys[e] : kKeyMax;
        lp[i] = valid ? pays[e] : 0ull;
        if (valid) {
            hmin = min(hmin, (uint32_t)(lk[i] >> 32));
            hmax = max(hmax, (uint32_t)(lk[i] >> 32));
        }
    }
    __syncwarp();
    uint32_t lo = __reduce_min_sync(kFull, hmin), hi = __reduce_max_sync(kFull, hmax);
    while (lo < hi) {                 // smallest rank word H with #(word <= H) >= k
        const uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < PL; ++i) c += ((uint32_t)(lk[i] >> 32) <= mid && lane + 32u * i < count) ? 1u : 0u;
        c = __reduce_add_sync(kFull, c);
        if (c >= k) hi = mid; else lo = mid + 1u;
    }
    const uint32_t H = lo;
    uint32_t c_less = 0, c_eq = 0;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const bool valid = lane + 32u * i < count;
        c_less += (valid && (uint32_t)(lk[i] >> 32) < H) ? 1u : 0u;
        c_eq += (valid && (uint32_t)(lk[i] >> 32) == H) ? 1u : 0u;
    }
    c_less = __reduce_add_sync(kFull, c_less);
    c_eq = __reduce_add_sync(kFull, c_eq);
    const uint32_t need = k - c_less;          // 1 <= need <= c_eq entries of the tie group survive
    uint32_t lcut = 0xffffffffu;
    if (c_eq > need) {                         // rank ties at the cut: the id word decides
        uint32_t l0 = 0u, l1 = 0xffffffffu;
        while (l0 < l1) {
            const uint32_t mid = l0 + ((l1 - l0) >> 1);
            uint32_t c = 0;
#pragma unroll
            for (int i = 0; i < PL; ++i)
                c += (lane + 32u * i < count && (uint32_t)(lk[i] >> 32) == H && (uint32_t)lk[i] <= mid) ? 1u : 0u;
            c = __reduce_add_sync(kFull, c);
            if (c >= need) l1 = mid; else l0 = mid + 1u;
        }
        lcut = l0;
    }
    const u64 T = ((u64)H << 32) | lcut;
    uint32_t base = 0, kth_lo = 0;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const bool keep = lane + 32u * i < count && lk[i] <= T;
        const uint32_t m = __ballot_sync(kFull, keep);
        if (keep) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            keys[pos] = lk[i];
            pays[pos] = lp[i];
            if ((uint32_t)(lk[i] >> 32) == H) kth_lo = max(kth_lo, (uint32_t)lk[i]);
        }
        base += __popc(m);
    }
    __syncwarp();
    return ((u64)H << 32) | __reduce_max_sync(kFull, kth_lo);
}

__global__ void __launch_bounds__(kGmThreads, 1)
flat_gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_a, const unsigned char* __restrict__ q_blobs,
                      const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char gsmem[];
    __shared__ __align__(8) uint64_t full_bar[kGmStages], empty_bar[kGmStages], a_ready[2], a_free[2], d_full, d_free;
    __shared__ uint32_t tmem_slot;
    __shared__ u64 s_thr[kGmN];          // exact threshold key per query of this block
    __shared__ float s_thr_rank[kGmN];   // its rank value: the cheap first-level filter
    __shared__ uint32_t s_cnt[kGmN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qb = blockIdx.x % p.qblocks, rr = blockIdx.x / p.qblocks;
    const bool active = rr < p.ranges;
    const uint32_t tiles_total = (p.n + kGmTile - 1) / kGmTile;
    const uint32_t tile0 = active ? (uint32_t)((uint64_t)tiles_total * rr / p.ranges) : 0;
    const uint32_t tile1 = active ? (uint32_t)((uint64_t)tiles_total * (rr + 1) / p.ranges) : 0;
    const uint32_t chunks = p.dims / 32;

    for (int q = tid; q < kGmN; q += kGmThreads) {
        u64 thr = kKeyMax;
        const uint32_t qg = qb * kGmN + q;
        if (p.init_keys != nullptr && qg < p.nq && p.init_counts[qg] >= p.k) thr = p.init_keys[(size_t)qg * p.k + p.k - 1u];
        s_thr[q] = thr == kKeyMax ? kKeyMax : thr + 1u;             // keys are unique: "<= k-th" is "< k-th + 1"
        s_thr_rank[q] = thr == kKeyMax ? INFINITY : rank_from_key(thr);
        s_cnt[q] = 0;
    }
    if (tid == 0) {
        for (int s = 0; s < kGmStages; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 5);      // 4 split warps (A consumed) + MMA commit (B consumed)
            tc::mbar_init(&a_ready[s], 4);
            tc::mbar_init(&a_free[s], 1);
        }
        tc::mbar_init(&d_full, 1);
        tc::mbar_init(&d_free, kGmEpiWarps);
        tc::mbar_fence_init();
    }
    if (warp == kGmMmaWarp) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (warp == kGmProducerWarp) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t cc = 0;
            for (uint32_t tile = tile0; tile < tile1; ++tile) {
                for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                    const uint32_t s = cc % kGmStages, ph = (cc / kGmStages) & 1u;
                    unsigned char* st = gsmem + (size_t)s * kGmStageBytes;
                    tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t tx = ((p.debug & 1u) ? 0u : 16384u) + ((p.debug & 2u) ? 0u : 65536u);
                    if (tx) tc::mbar_arrive_expect_tx(&full_bar[s], tx);
                    else tc::mbar_arrive(&full_bar[s]);
                    if (!(p.debug & 1u)) tc::tma_load_2d(st, &tmap_a, kc * 32, tile * kGmTile, &full_bar[s]);
                    // the query block's (hi | lo) chunk is one contiguous, pre-swizzled 64 KB image
                    if (!(p.debug & 2u))
                        tma_bulk_g2s(st + 16384, q_blobs + ((size_t)qb * chunks + kc) * 65536u, 65536u, &full_bar[s]);
                }
            }
        }
    } else if (warp == kGmMmaWarp) {
        // ===== MMA issuer (warp-uniform control flow, one elected lane issues) =====
        const uint32_t idesc = tc::umma_idesc_tf32(kGmTile, kGmN);
        uint32_t cc = 0, it = 0;
        for (uint32_t tile = tile0; tile < tile1; ++tile, ++it) {
            for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                const uint32_t s = cc % kGmStages, ph = (cc / kGmStages) & 1u;
                tc::mbar_wait(&full_bar[s], ph);          // B chunk landed
                tc::mbar_wait(&a_ready[s], ph);           // A chunk split into TMEM
                if (kc == 0) tc::mbar_wait(&d_free, (it & 1u) ^ 1u);   // epilogue drained the accumulator
                tc::fence_after_sync();
                const uint32_t st_addr = tc::smem_addr(gsmem + (size_t)s * kGmStageBytes);
                const uint64_t bh0 = tc::umma_smem_desc_sw128(st_addr + 16384);
                const uint64_t bl0 = tc::umma_smem_desc_sw128(st_addr + 16384 + 32768);
                const uint32_t a0 = tbase + kGmACol + s * 64u;
                if (tc::elect_one()) {
                    if (!(p.debug & 16u))
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks) {
#pragma unroll
                        for (uint32_t term = 0; term < 3; ++term) {
                            const uint64_t bdesc = (term == 1 ? bl0 : bh0) + (uint64_t)(ks * 2u);
                            const uint32_t a_addr = a0 + ks * 8u + (term == 2 ? 32u : 0u);
                            tc::umma_tf32_ts(tbase, a_addr, bdesc, idesc, (kc | ks | term) != 0u);
                        }
                    }
                    tc::umma_commit(&a_free[s]);
                    tc::umma_commit(&empty_bar[s]);
                    if (kc + 1 == chunks) tc::umma_commit(&d_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= kGmSplitWarp0) {
        // ===== split warps: A chunk -> (hi, lo) -> TMEM =====
        const uint32_t quarter = warp & 3u;
        const uint32_t row = quarter * 32u + lane;
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        uint32_t cc = 0;
        for (uint32_t tile = tile0; tile < tile1; ++tile) {
            for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                const uint32_t s = cc % kGmStages, ph = (cc / kGmStages) & 1u;
                tc::mbar_wait(&full_bar[s], ph);
                tc::mbar_wait(&a_free[s], ph ^ 1u);
                tc::fence_after_sync();
                const unsigned char* blk = gsmem + (size_t)s * kGmStageBytes + row * 128u;
                if (!(p.debug & 4u)) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(blk + ((c ^ (row & 7u)) << 4));
                    const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t hbits = __float_as_uint(xs[e]) & 0xFFFFE000u;
                        hi[c * 4 + e] = hbits;
                        lo[c * 4 + e] = __float_as_uint(xs[e] - __uint_as_float(hbits));
                    }
                }
                tc::tmem_st32(lane_addr + kGmACol + s * 64u, hi);
                tc::tmem_st32(lane_addr + kGmACol + s * 64u + 32u, lo);
                tc::tmem_st_wait();
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&a_ready[s]);
                    tc::mbar_arrive(&empty_bar[s]);
                }
            }
        }
    } else {
        // ===== epilogue warps 0-7: warp w reads TMEM lanes of quarter (w & 3), query columns of half (w >> 2) =====
        const uint32_t quarter = warp & 3u, half = warp >> 2;
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        const size_t list_base = (size_t)blockIdx.x * kGmN;
        // rank = fma(dot, -1, bias): cosine 1 - dot (one rounding, == 1.0f - raw), inner product / negative
        // inner product -dot (distances.rs:113-119 with raw = dot resp. -dot)
        uint32_t it = 0;
        for (uint32_t tile = tile0; tile < tile1; ++tile, ++it) {
            const uint32_t row = tile * kGmTile + quarter * 32u + lane;
            const bool valid = row < p.n;
            const uint32_t idr = valid ? (p.id_rank ? __ldg(p.id_rank + row) : row) : 0u;
            const float bias = p.row_norm2 ? (valid ? __ldg(p.row_norm2 + row) : 0.0f) : (p.metric == kCosine ? 1.0f : 0.0f);
            tc::mbar_wait(&d_full, it & 1u);
            tc::fence_after_sync();
            uint32_t worst_bits = 0;   // max |score| bits: >= 0x7f800000 means a non-finite score
            for (uint32_t cg = half * 4u; cg < half * 4u + 4u; ++cg) {
                if ((p.debug & 8u) || qb * kGmN + cg * 32u >= p.nq) break;       // padded query columns (uniform)
                uint32_t r[32];
                tc::tmem_ld32(lane_addr + cg * 32u, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t q = cg * 32u + j;
                    const float dot = __uint_as_float(r[j]);
                    worst_bits = max(worst_bits, r[j] & 0x7fffffffu);
                    const float rankv = fmaf(dot, p.rank_scale, bias);
                    if (rankv <= s_thr_rank[q] && valid && qb * kGmN + q < p.nq) {   // first-level filter: one compare
                        const u64 key = ((u64)order_key(rankv) << 32) | idr;
                        if (key < s_thr[q]) {
                            const float raw = p.row_norm2 ? rankv : (p.metric == kNegativeInnerProduct ? -dot : dot);
                            const uint32_t slot = atomicAdd(&s_cnt[q], 1u);
                            if (slot < p.list_cap) {
                                p.list_keys[(list_base + q) * p.list_cap + slot] = key;
                                p.list_pays[(list_base + q) * p.list_cap + slot] = ((u64)__float_as_uint(raw) << 32) | row;
                            }
                        }
                    }
                }
            }
            if (valid && worst_bits >= 0x7f800000u) *p.bad = 1u;
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d_free);
            // lists that could overflow during the next tile are cut back to their best k
            asm volatile("bar.sync 2, 256;" ::: "memory");
            for (uint32_t q = warp; q < kGmN; q += kGmEpiWarps) {
                const uint32_t cnt = min(s_cnt[q], p.list_cap);
                if (cnt + kGmTile > p.list_cap) {
                    u64* lkeys = p.list_keys + (list_base + q) * p.list_cap;
                    u64* lpays = p.list_pays + (list_base + q) * p.list_cap;
                    const u64 kth = p.list_cap == kGmListSmall ? warp_select_list<8>(lkeys, lpays, cnt, p.k, lane)
                                                               : warp_select_list<16>(lkeys, lpays, cnt, p.k, lane);
                    if (lane == 0) {
                        s_cnt[q] = min(cnt, p.k);
                        s_thr[q] = kth;
                        s_thr_rank[q] = kth == kKeyMax ? INFINITY : rank_from_key(kth);
                    }
                }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        // final cut of every list of this CTA
        for (uint32_t q = warp; q < kGmN; q += kGmEpiWarps) {
            const uint32_t cnt = min(s_cnt[q], p.list_cap);
            u64* lkeys = p.list_keys + (list_base + q) * p.list_cap;
            u64* lpays = p.list_pays + (list_base + q) * p.list_cap;
            if (cnt > 0) {
                if (p.list_cap == kGmListSmall) warp_select_list<8>(lkeys, lpays, cnt, p.k, lane);
                else warp_select_list<16>(lkeys, lpays, cnt, p.k, lane);
            }
            if (lane == 0) p.list_counts[list_base + q] = min(cnt, p.k);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kGmMmaWarp) tc::tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------
// K2 single-pass variant. The tensor-core scores are only a FILTER (every kept candidate is re-scored exactly
// by flat_gemm_rescore_kernel and the kept set is proven complete against an error bound), so fp32 accuracy is
// not needed HERE: one TF32 pass with a wider candidate margin returns bit-identical final results for a third
// of the MMA work. That also removes what throttled the 3xTF32 kernel (ncu, 1024 queries: tensor pipe 58 %
// active, every warp role waiting): no hi/lo split warps, one 32 KB query chunk instead of 64 KB per K chunk
// (the ring holds 4 stages instead of 2 — the 80 KB stages left ~1 stage in flight against a ~2-3k-cycle L2
// latency), both operands straight from shared memory (SS mode, the hardware reads fp32 words as TF32), and
// the whole of TMEM for TWO accumulators, released as soon as a tile's scores sit in registers.
// Error bound: operands lose < 2^-10 relative each, so |approx - exact| <= 2^-9 * |q| * |row| (+ fp32
// accumulation): 2e-3 for unit vectors. The candidate margin k' - k covers the rows that can sit that close to
// the k-th score: k' = k + 32 up to k = 32, k' = min(3k, 192) >= k + 64 above (flat_gemm_search_device); queries whose
// kept set cannot be proven complete are redone by the caller on the 3xTF32 kernel above (second tier).
// Sixteen epilogue warps (four per TMEM lane quarter, a quarter of the query columns each): with eight, each warp had
// eight 32-column groups to pull out of TMEM and filter per tile pair, two warps per scheduler could not hide the
// latencies, and the ~8 us before both accumulators were released again were added to every pair's MMA time.
constexpr int kG1EpiWarps = 16, kG1ProducerWarp = 16, kG1MmaWarp = 17;
constexpr int kG1ColParts = kG1EpiWarps / 4;             // column parts per accumulator (one per epilogue warp of a lane quarter)
constexpr int kG1Threads = (kG1MmaWarp + 1) * 32;        // warps 0-15 epilogue, 16 producer, 17 MMA
constexpr int kG1Stages = 3;
// One stage = the same 32-dim chunk of TWO consecutive row tiles + the query chunk: the query operand is what
// the kernel streams most (it is re-read from L2 for every row tile, ~7-8 TB/s measured = the L2->SM limit), so
// every query chunk is used for 256 rows — one tile per TMEM accumulator — instead of 128.
constexpr uint32_t kG1StageBytes = 2 * 16384 + 32768;     // A chunks of tile 2t, 2t+1 [128 x 32] + query chunk [256 x 32], fp32, SW128

constexpr int kG2N = 128;      // CTA-pair form: queries per block (UMMA N), two accumulator sets in TMEM
constexpr int kG2Stages = 4;   // ... and its ring: A chunks of the CTA's two tiles (32 KB) + its half of the query chunk (8 KB)

// (Tried and dropped: warp-aggregated appends — one vote + ballot + a single shared-memory atomic per (warp, query
// column). The vote costs an instruction on EVERY score: the main pass went from 2.67 to 5.25 ms, and even the
// pre-pass, where nearly every score passes, got slower: 736 vs 618 us.)
// Second-level filter + append of one score that passed the rank compare (rare: a fraction of a percent of the scores).
// The payload carries the approximate rank next to the row; only the row is read downstream (the candidates are
// re-scored exactly).
__device__ __noinline__ void gemm1_append(const u64* s_thr, uint32_t* s_cnt, u64* list_keys, u64* list_pays, uint32_t list_cap,
                                          uint32_t list_base, uint32_t q, float rankv, uint32_t idr, uint32_t row) {
    const u64 key = ((u64)order_key(rankv) << 32) | idr;
    if (key < s_thr[q]) {
        const uint32_t slot = atomicAdd(&s_cnt[q], 1u);
        if (slot < list_cap) {
            const size_t at = ((size_t)list_base + q) * list_cap + slot;
            list_keys[at] = key;
            list_pays[at] = ((u64)__float_as_uint(rankv) << 32) | row;
        }
    }
}

// kPair: two CTAs of a cluster work as one (tcgen05 cta_group::2, UMMA M = 256). The pair serves one query block; each
// CTA streams the rows of ITS two tiles and only HALF of the query chunk (the MMA reads the other half from the
// peer's shared memory), so a stage is 48 KB instead of 64 and the ring holds four of them: three stages in flight
// behind the one the tensor cores read, where the single-CTA form had two and ran the MMA stream at 0.73 of the
// pipe. The leader (cluster rank 0) issues every MMA; its commits arrive on both CTAs' barriers; the peer's idle
// MMA warp relays "my stage landed" to the leader, and the peer's epilogue warps release the accumulators there.
// First-level bound of the single-pass filter for a rank threshold `thr`. L2 family: the rank itself (the kernel
// computes |x|^2 - 2 q.x per score). Dot family (rank = bias - dot): the smallest dot that can still reach the
// threshold, lowered by a few ulps so that rounding of bias - dot can never exclude a row the exact key test would
// keep (that test, in gemm1_append, decides).
__device__ __forceinline__ float filter_bound(float thr, bool dot_family, float bias) {
    if (!dot_family) return thr;
    if (thr == INFINITY) return -INFINITY;
    if (thr == -INFINITY) return INFINITY;
    const float t = bias - thr;
    return t - 4.8e-7f * fmaxf(1.0f, fabsf(t));
}

template <bool kPair>
__global__ void __launch_bounds__(kG1Threads, 1)
flat_gemm1_topk_kernel(const __grid_constant__ CUtensorMap tmap_a, const unsigned char* __restrict__ q_blobs,
                       const GemmParams p) {
    constexpr int kStages = kPair ? kG2Stages : kG1Stages;
    constexpr uint32_t kN = kPair ? (uint32_t)kG2N : (uint32_t)kGmN;   // queries per block = UMMA N
    constexpr uint32_t kSets = kPair ? 2u : 1u;                        // accumulator sets (2 x kN columns each) in TMEM
    constexpr uint32_t kGroups = kN / (32u * kG1ColParts);              // 32-column groups per epilogue warp and accumulator
    constexpr uint32_t kBBytes = kPair ? kN * 128u / 2u : kN * 128u;   // this CTA's part of a query chunk
    constexpr uint32_t kStageBytes = 2u * 16384u + kBBytes;
    extern __shared__ __align__(1024) unsigned char gsmem[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], peer_full[kStages], d_full[kSets], d_free[kSets][2];
    __shared__ uint32_t tmem_slot;
    __shared__ u64 s_thr[kN];
    __shared__ __align__(16) float s_thr_rank[kN];   // first-level bound, see filter_bound()
    __shared__ uint32_t s_cnt[kN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Work units: a tile pair per CTA; a CTA pair takes two consecutive tile pairs per unit (leader the even one).
    const uint32_t crank = kPair ? tc::cluster_ctarank() : 0u;
    const uint32_t wid = kPair ? blockIdx.x / 2u : blockIdx.x;           // work-group index: CTA, or CTA pair
    const uint32_t qb = wid % p.qblocks, rr = wid / p.qblocks;
    const bool active = rr < p.ranges;
    const uint32_t tiles_total = (p.n + kGmTile - 1) / kGmTile;
    const uint32_t pairs_total = (tiles_total + 1) / 2;
    const uint32_t units_total = kPair ? (pairs_total + 1) / 2 : pairs_total;
    const uint32_t unit0 = active ? (uint32_t)((uint64_t)units_total * rr / p.ranges) : 0;
    const uint32_t unit1 = active ? (uint32_t)((uint64_t)units_total * (rr + 1) / p.ranges) : 0;
    auto pair_of = [crank](uint32_t unit) { return kPair ? 2u * unit + crank : unit; };
    const uint32_t chunks = p.dims / 32;
    const bool dot_family = p.row_norm2 == nullptr;            // rank = bias - dot: the filter compares the dot itself
    const float dot_bias = p.metric == kCosine ? 1.0f : 0.0f;

    for (int q = tid; q < (int)kN; q += kG1Threads) {
        u64 thr = kKeyMax;
        const uint32_t qg = qb * kN + q;
        if (p.init_keys != nullptr && qg < p.nq && p.init_counts[qg] >= p.k) thr = p.init_keys[(size_t)qg * p.k + p.k - 1u];
        s_thr[q] = thr == kKeyMax ? kKeyMax : thr + 1u;             // keys are unique: "<= k-th" is "< k-th + 1"
        s_thr_rank[q] = thr == kKeyMax ? INFINITY : rank_from_key(thr);
        if (p.init_rank != nullptr && qg < p.nq) {                  // a rank bound: every key of that rank or better passes
            const float r = p.init_rank[qg];
            const uint32_t rk = order_key(r);
            s_thr_rank[q] = r;
            s_thr[q] = (r == INFINITY || rk == 0xFFFFFFFFu) ? kKeyMax : ((u64)(rk + 1u) << 32);
            if (s_thr[q] == kKeyMax) s_thr_rank[q] = INFINITY;
        }
        if (qg >= p.nq) { s_thr[q] = 0ull; s_thr_rank[q] = -INFINITY; }   // padded query column: nothing ever passes
        s_thr_rank[q] = filter_bound(s_thr_rank[q], dot_family, dot_bias);
        s_cnt[q] = 0;
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);          // the MMA commit frees the three operands of the stage
            tc::mbar_init(&peer_full[s], 1);          // pair, leader: the peer's stage has landed
        }
        for (uint32_t st = 0; st < kSets; ++st) {
            tc::mbar_init(&d_full[st], 1);
            tc::mbar_init(&d_free[st][0], kPair ? 2 * kG1EpiWarps : kG1EpiWarps);   // pair, leader: both CTAs' epilogue warps
            tc::mbar_init(&d_free[st][1], kPair ? 2 * kG1EpiWarps : kG1EpiWarps);
        }
        tc::mbar_fence_init();
    }
    if (warp == kG1MmaWarp) {
        if constexpr (kPair) tc::tmem_alloc2(&tmem_slot, 512); else tc::tmem_alloc(&tmem_slot, 512);
    }
    tc::fence_before_sync();
    __syncthreads();
    if constexpr (kPair) tc::cluster_sync_all();      // the peer's barriers exist before anything arrives on them
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (warp == kG1ProducerWarp) {
        if (lane == 0) {
            uint32_t cc = 0;
            for (uint32_t unit = unit0; unit < unit1; ++unit) {
                const uint32_t pair = pair_of(unit);
                for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                    const uint32_t s = cc % kStages, ph = (cc / kStages) & 1u;
                    unsigned char* st = gsmem + (size_t)s * kStageBytes;
                    tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                    if ((p.debug & 3u) && cc >= (uint32_t)kStages) {   // timing experiments: the operands stay what the first fills left
                        tc::mbar_arrive_expect_tx(&full_bar[s], 0u);
                        continue;
                    }
                    tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                    // rows past the end of a partial tile are zero-filled by the tensor map and still count; an odd
                    // tile count leaves the last pair without a second tile: load the first again (its rows are
                    // masked out in the epilogue) rather than a box that lies entirely outside the matrix
                    const uint32_t tile_a = min(2u * pair, tiles_total - 1u), tile_b = min(2u * pair + 1u, tiles_total - 1u);
                    tc::tma_load_2d(st, &tmap_a, kc * 32, tile_a * kGmTile, &full_bar[s]);
                    tc::tma_load_2d(st + 16384, &tmap_a, kc * 32, tile_b * kGmTile, &full_bar[s]);
                    tma_bulk_g2s(st + 32768, q_blobs + ((size_t)qb * chunks + kc) * (kN * 128u) + crank * kBBytes, kBBytes, &full_bar[s]);
                }
            }
        }
    } else if (warp == kG1MmaWarp) {
        if (kPair && crank != 0u) {
            // ===== pair, peer CTA: no MMAs to issue — tell the leader when each of my stages has landed =====
            uint32_t cc = 0;
            for (uint32_t unit = unit0; unit < unit1; ++unit) {
                for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                    const uint32_t s = cc % kStages, ph = (cc / kStages) & 1u;
                    tc::mbar_wait(&full_bar[s], ph);
                    if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_shared(&peer_full[s], 0u));
                    __syncwarp();
                }
            }
        } else {
            const uint32_t idesc = tc::umma_idesc_tf32(kPair ? 2 * kGmTile : kGmTile, kN);
            uint32_t cc = 0, it = 0;
            for (uint32_t unit = unit0; unit < unit1; ++unit, ++it) {
                // pair: set it & 1 — while the epilogues filter the scores of unit it out of one set, the MMAs of
                // unit it + 1 fill the other
                const uint32_t set = it % kSets, free_par = ((it / kSets) & 1u) ^ 1u;
                const uint32_t acc0 = tbase + set * 2u * kN, acc1 = acc0 + kN;
                for (uint32_t kc = 0; kc < chunks; ++kc, ++cc) {
                    const uint32_t s = cc % kStages, ph = (cc / kStages) & 1u;
                    tc::mbar_wait(&full_bar[s], ph);
                    if constexpr (kPair) tc::mbar_wait(&peer_full[s], ph);
                    if (kc == 0) tc::mbar_wait(&d_free[set][0], free_par);   // accumulator 0's scores sit in registers (both CTAs of a pair)
                    tc::fence_after_sync();
                    const uint32_t st_addr = tc::smem_addr(gsmem + (size_t)s * kStageBytes);
                    const uint64_t a0 = tc::umma_smem_desc_sw128(st_addr);
                    const uint64_t a1 = tc::umma_smem_desc_sw128(st_addr + 16384);
                    const uint64_t b0 = tc::umma_smem_desc_sw128(st_addr + 32768);
                    if (tc::elect_one() && !(p.debug & 16u)) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) {
                            if constexpr (kPair) tc::umma_tf32_ss2(acc0, a0 + (uint64_t)(ks * 2u), b0 + (uint64_t)(ks * 2u), idesc, (kc | ks) != 0u);
                            else tc::umma_tf32_ss(acc0, a0 + (uint64_t)(ks * 2u), b0 + (uint64_t)(ks * 2u), idesc, (kc | ks) != 0u);
                        }
                    }
                    __syncwarp();
                    if (kc == 0) {
                        tc::mbar_wait(&d_free[set][1], free_par);
                        tc::fence_after_sync();
                    }
                    if (tc::elect_one()) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks)
                            if (!(p.debug & 16u)) {
                                if constexpr (kPair) tc::umma_tf32_ss2(acc1, a1 + (uint64_t)(ks * 2u), b0 + (uint64_t)(ks * 2u), idesc, (kc | ks) != 0u);
                                else tc::umma_tf32_ss(acc1, a1 + (uint64_t)(ks * 2u), b0 + (uint64_t)(ks * 2u), idesc, (kc | ks) != 0u);
                            }
                        if constexpr (kPair) {
                            tc::umma_commit2(&empty_bar[s], 3);
                            if (kc + 1 == chunks) tc::umma_commit2(&d_full[set], 3);
                        } else {
                            tc::umma_commit(&empty_bar[s]);
                            if (kc + 1 == chunks) tc::umma_commit(&d_full[set]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lanes of quarter (w & 3), query columns of part (w >> 2), both accumulators of the set =====
        const uint32_t quarter = warp & 3u, half = warp >> 2;   // `half`: this warp's column part (0 .. kG1ColParts - 1)
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        const size_t list_base = (size_t)blockIdx.x * kN;
        const float bias = p.metric == kCosine ? 1.0f : 0.0f;
        const float scale = p.rank_scale;
        uint32_t it = 0;
        for (uint32_t unit = unit0; unit < unit1; ++unit, ++it) {
            const uint32_t pair = pair_of(unit);
            const uint32_t set = it % kSets;
            tc::mbar_wait(&d_full[set], (it / kSets) & 1u);
            tc::fence_after_sync();
#pragma unroll 1
            for (uint32_t acc = 0; acc < 2; ++acc) {
                const uint32_t row = (2u * pair + acc) * kGmTile + quarter * 32u + lane;
                const bool valid = row < p.n;
                const uint32_t idr = valid ? (p.id_rank ? __ldg(p.id_rank + row) : row) : 0u;
                const float row_bias = valid ? (p.row_norm2 ? __ldg(p.row_norm2 + row) : bias) : __int_as_float(0x7fc00000);
                // One group of 32 query columns at a time: scores and the group's 32 bounds in registers (thresholds
                // only move in the cut phase below, behind the barrier, so reading them once per group is exact).
                // The per-score work is branch-free — FFMA, compare, a predicated bit and a predicated copy of the
                // passing rank — and the append code exists once, out of line. (The earlier form, a shared-memory
                // load + branch per score with 128 inlined copies of the append path, had 27 % of the kernel's samples
                // on the load wait and 11 % on instruction-cache misses; holding all 128 scores AND the bounds in
                // registers spilled.) The accumulator is released once its last group sits in registers.
#pragma unroll 1
                for (uint32_t g = 0; g < kGroups; ++g) {
                    const uint32_t cg = half * kGroups + g;
                    uint32_t r[32];
                    tc::tmem_ld32(lane_addr + set * 2u * kN + acc * kN + cg * 32u, r);
                    const float4* t4 = reinterpret_cast<const float4*>(&s_thr_rank[cg * 32u]);
                    tc::tmem_ld_wait();
                    if (g + 1u == kGroups) {
                        tc::fence_before_sync();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (kPair) tc::mbar_arrive_cluster(tc::mapa_shared(&d_free[set][acc], 0u)); else tc::mbar_arrive(&d_free[set][acc]);
                        }
                    }
                    if (p.debug & 8u) continue;
                    if (p.score_dump != nullptr) {   // pre-pass: ranks of this warp's 32 rows x 32 queries, 128 contiguous bytes per query
                        float* dst = p.score_dump + (size_t)(qb * kN + cg * 32u) * p.dump_stride + row;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (row < p.dump_stride) dst[(size_t)j * p.dump_stride] = valid ? fmaf(__uint_as_float(r[j]), scale, row_bias) : INFINITY;
                        continue;
                    }
                    // (padded query columns carry a bound nothing passes.) Dot family: rank = bias - dot, so the score
                    // is compared with a dot bound directly — compare + two predicated moves per score, no arithmetic.
                    // Non-finite scores cannot occur: the re-scoring kernel sends the batch to the exact path when
                    // |q| * max|row| could overflow (Cauchy-Schwarz bounds every partial sum).
                    uint32_t mask = 0u;
                    float one_dot = 0.0f;
                    if (dot_family) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {                   // the group's bounds, four at a time (registers are scarce at 576 threads)
                            const float4 t = t4[j4];
                            const float tb[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = 4 * j4 + e;
                                const float dot = __uint_as_float(r[j]);
                                if (dot >= tb[e]) { mask |= 1u << j; one_dot = dot; }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 t = t4[j4];
                            const float tb[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = 4 * j4 + e;
                                const float dot = __uint_as_float(r[j]);
                                if (fmaf(dot, scale, row_bias) <= tb[e]) { mask |= 1u << j; one_dot = dot; }
                            }
                        }
                    }
                    if (!valid) mask = 0u;                                 // rows past the end of the matrix
                    if (mask != 0u) {
                        if ((mask & (mask - 1u)) == 0u) {                  // the usual case: one score of the 32 passed
                            gemm1_append(s_thr, s_cnt, p.list_keys, p.list_pays, p.list_cap, (uint32_t)list_base,
                                         cg * 32u + (uint32_t)__ffs(mask) - 1u, fmaf(one_dot, scale, row_bias), idr, row);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (mask & (1u << j))
                                    gemm1_append(s_thr, s_cnt, p.list_keys, p.list_pays, p.list_cap, (uint32_t)list_base, cg * 32u + j,
                                                 fmaf(__uint_as_float(r[j]), scale, row_bias), idr, row);
                        }
                    }
                }
            }
            // lists that could overflow during the next pair of tiles are cut back to their best k
            asm volatile("bar.sync 2, %0;" ::"n"(kG1EpiWarps * 32) : "memory");
            for (uint32_t q = warp; q < kN; q += kG1EpiWarps) {
                const uint32_t cnt = min(s_cnt[q], p.list_cap);
                if (cnt + 2 * kGmTile > p.list_cap) {
                    u64* lkeys = p.list_keys + (list_base + q) * p.list_cap;
                    u64* lpays = p.list_pays + (list_base + q) * p.list_cap;
                    const u64 kth = warp_select_list<16>(lkeys, lpays, cnt, p.k, lane);
                    if (lane == 0) {
                        s_cnt[q] = min(cnt, p.k);
                        s_thr[q] = kth;
                        s_thr_rank[q] = filter_bound(kth == kKeyMax ? INFINITY : rank_from_key(kth), dot_family, dot_bias);
                    }
                }
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kG1EpiWarps * 32) : "memory");
        }
        for (uint32_t q = warp; q < kN; q += kG1EpiWarps) {
            const uint32_t cnt = min(s_cnt[q], p.list_cap);
            u64* lkeys = p.list_keys + (list_base + q) * p.list_cap;
            u64* lpays = p.list_pays + (list_base + q) * p.list_cap;
            if (cnt > 0) warp_select_list<16>(lkeys, lpays, cnt, p.k, lane);
            if (lane == 0) p.list_counts[list_base + q] = min(cnt, p.k);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if constexpr (kPair) tc::cluster_sync_all();      // neither CTA leaves (or frees TMEM) while the other may still signal it
    if (warp == kG1MmaWarp) {
        if constexpr (kPair) tc::tmem_dealloc2(tbase, 512); else tc::tmem_dealloc(tbase, 512);
    }
}

// queries [nq, dims] -> per (query block, 32-dim chunk) a 32 KB image [256 x 32] fp32 in the UMMA K-major
// SWIZZLE_128B shared-memory layout (zero rows beyond nq); the tensor core reads the words as TF32.
__global__ void pack_queries_kernel(const float* q, uint32_t nq, uint32_t qblocks, uint32_t nblk, uint32_t dims, unsigned char* blobs) {
    const uint32_t chunks = dims / 32;
    const size_t total = (size_t)qblocks * nblk * dims;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(i % dims);
        const size_t row = i / dims;
        const uint32_t qb = (uint32_t)(row / nblk), n = (uint32_t)(row % nblk);
        unsigned char* blob = blobs + ((size_t)qb * chunks + k / 32u) * (nblk * 128u);
        *reinterpret_cast<float*>(blob + tc::sw128_offset(n, k % 32u)) = row < nq ? q[row * dims + k] : 0.0f;
    }
}

// Pre-pass select (single-pass kernel): the k-th smallest rank among a query's `n` sample scores. One CTA per query.
// Phase 1 bisects the order keys of the first kSelHead scores in shared memory (bound t0); phase 2 streams the rest of
// the sample once from L2 / HBM and keeps what is <= t0 (about k * n / kSelHead keys) in a short list next to the
// phase-1 keys <= t0; the k-th smallest of that list is the k-th smallest of the whole sample. A list that overflows
// (ties, or a sample sorted best-last) leaves t0, which is still a valid bound. Fewer than k sample rows: +inf.
constexpr uint32_t kSelHead = 32768, kSelList = 8192, kSelThreads = 256;
constexpr size_t kSelSmem = (size_t)(kSelHead + kSelList) * sizeof(uint32_t);

__device__ __forceinline__ uint32_t select_kth_key(const uint32_t* keys, uint32_t n, uint32_t k, uint32_t hi, uint32_t* s_part) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t lo = 0u;
    while (lo < hi) {                          // smallest H with #(key <= H) >= k
        const uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t c = 0;
        for (uint32_t i = tid; i < n; i += kSelThreads) c += keys[i] <= mid ? 1u : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) s_part[warp] = c;
        __syncthreads();
        c = (s_part[0] + s_part[1] + s_part[2] + s_part[3]) + (s_part[4] + s_part[5] + s_part[6] + s_part[7]);
        __syncthreads();
        if (c >= k) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

__global__ void __launch_bounds__(kSelThreads) gemm1_sample_select_kernel(const float* dump, uint32_t stride, uint32_t n, uint32_t k,
                                                                          float* out_rank) {
    extern __shared__ __align__(1024) unsigned char gsmem[];
    __shared__ uint32_t s_part[8];
    __shared__ uint32_t s_cnt;
    uint32_t* keys = reinterpret_cast<uint32_t*>(gsmem);
    uint32_t* list = keys + kSelHead;
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const float* src = dump + (size_t)q * stride;
    const uint32_t n1 = min(n, kSelHead);
    for (uint32_t i = tid; i < n1; i += kSelThreads) keys[i] = order_key(src[i]);
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (n < k) { if (tid == 0) out_rank[q] = INFINITY; return; }
    uint32_t t = select_kth_key(keys, n1, k, 0xFFFFFFFFu, s_part);
    if (n > n1) {
        for (uint32_t i = tid; i < n1; i += kSelThreads) {
            const uint32_t key = keys[i];
            if (key <= t) { const uint32_t at = atomicAdd(&s_cnt, 1u); if (at < kSelList) list[at] = key; }
        }
        const uint4* src4 = reinterpret_cast<const uint4*>(src + n1);   // n1 and stride are multiples of 4
        const uint32_t n4 = (n - n1) / 4u;
        for (uint32_t i = tid; i < n4; i += kSelThreads) {
            const uint4 v = __ldcs(src4 + i);
            const uint32_t kk[4] = {order_key(__uint_as_float(v.x)), order_key(__uint_as_float(v.y)),
                                    order_key(__uint_as_float(v.z)), order_key(__uint_as_float(v.w))};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (kk[j] <= t) { const uint32_t at = atomicAdd(&s_cnt, 1u); if (at < kSelList) list[at] = kk[j]; }
        }
        for (uint32_t i = n1 + n4 * 4u + tid; i < n; i += kSelThreads) {
            const uint32_t key = order_key(src[i]);
            if (key <= t) { const uint32_t at = atomicAdd(&s_cnt, 1u); if (at < kSelList) list[at] = key; }
        }
        __syncthreads();
        const uint32_t cnt = s_cnt;
        if (cnt <= kSelList) t = select_kth_key(list, cnt, k, t, s_part);
    }
    if (tid == 0) {
        const uint32_t bits = (t & 0x80000000u) ? (t ^ 0x80000000u) : ~t;   // inverse of order_key
        out_rank[q] = __uint_as_float(bits);
    }
}

// Per query: merge the sorted lists of the CTAs that served its query block.
__global__ void __launch_bounds__(128)
flat_gemm_merge_kernel(const GemmParams p, uint32_t cap, u64* out_keys, u64* out_pays, uint32_t* out_counts) {
    extern __shared__ __align__(1024) unsigned char gsmem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    const uint32_t q = blockIdx.x, qb = q / p.nblk, ql = q % p.nblk;
    Collector col;
    col.init(gsmem, &s_thresh, &s_count, cap, p.k);
    __syncthreads();
    const GemmParams pp = p;
    // list l of this query: CTA l * qblocks + qb, or — CTA pairs — CTA 2 * ((l / 2) * qblocks + qb) + l % 2
    auto list_of = [pp, qb, ql](uint32_t l) {
        const size_t cta = pp.pair ? 2 * ((size_t)(l >> 1) * pp.qblocks + qb) + (l & 1u) : (size_t)l * pp.qblocks + qb;
        return cta * pp.nblk + ql;
    };
    collector_merge_lists(
        col, pp.pair ? 2 * p.ranges : p.ranges, p.k, [&](uint32_t l) { return pp.list_counts[list_of(l)]; },
        [&](uint32_t l, uint32_t i) { return pp.list_keys[list_of(l) * pp.list_cap + i]; },
        [&](uint32_t l, uint32_t i) { return pp.list_pays[list_of(l) * pp.list_cap + i]; });
    const uint32_t total = *col.count;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        out_keys[(size_t)q * p.k + i] = col.keys[i];
        out_pays[(size_t)q * p.k + i] = col.pays[i];
    }
    if (threadIdx.x == 0) out_counts[q] = total;
}

// The tensor-core scores are a FILTER: fp32 accumulation inside the MMA datapath truncates, so a
// score can be off by ~5e-8 * (3 * dims / 8) * |row| * |query| (measured 1.4e-5 at dims = 768 for an
// all-positive dot). Every query therefore keeps k' > k approximate candidates, which are re-scored
// here exactly like the single-query kernel does (same lane mapping and summation order as
// flat_scan_kernel, f64 recovery included) and re-ranked. The kept set provably contains the true
// top-k when the worst kept approximate rank, minus the error bound, is still beyond the exact
// k-th rank; otherwise the query is flagged and redone on the single-query path.
struct RescoreParams {
    const float* rows;
    size_t stride;
    uint32_t dims, n, k, kprime;
    const uint32_t* id_rank;
    const float* queries;        // [nq, dims]
    const u64* cand_keys;        // [nq][kprime] approximate keys, ascending
    const u64* cand_pays;        // [nq][kprime]
    const uint32_t* cand_counts; // [nq]
    float err_coeff;             // |approx dot - exact dot| <= err_coeff * |query| * max |row|
    float max_row_norm;
    int l2_family;               // approximate ranks are |x|^2 - 2 q.x (squared distance minus |q|^2)
    u64* out_keys;               // [nq][k] exact keys (optional)
    u64* out_pays;               // [nq][k]
    uint32_t* out_counts;        // [nq]
    uint32_t* flags;             // [nq]: 1 = redo on the single-query path, 2 = metric overflow
    uint32_t* bad;               // raised when |q| * max|row| could overflow fp32 inside the tensor-core pass
};


template <int M>
__global__ void __launch_bounds__(128) flat_gemm_rescore_kernel(const RescoreParams p) {
    __shared__ __align__(16) unsigned char col_mem[256 * 16];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ float s_qn2[4];
    const uint32_t q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Collector col;
    col.init(col_mem, &s_thresh, &s_count, 256, p.k);
    __syncthreads();
    const uint32_t cnt = min(p.cand_counts[q], p.kprime);
    const uint32_t nvec = p.dims >> 2;
    const float4* q4 = reinterpret_cast<const float4*>(p.queries + (size_t)q * p.dims);
    float qn2 = 0.0f;
    for (uint32_t idx = lane; idx < nvec; idx += 32) {
        const float4 v = __ldg(q4 + idx);
        qn2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    qn2 = warp_sum(qn2);
    bool fatal_any = false;
    for (uint32_t c = warp; c < cnt; c += 4) {
        const uint32_t row = (uint32_t)p.cand_pays[(size_t)q * p.kprime + c];
        const float4* rp = reinterpret_cast<const float4*>(p.rows + (size_t)row * p.stride);
        Scorer<M> sc;
        sc.init();
        for (uint32_t idx = lane; idx < nvec; idx += 32) sc.accum(__ldg(q4 + idx), ldg_stream(rp + idx));
        bool bad, fatal;
        float raw = sc.finish(0.0, bad, fatal);
        if (bad) {
            Recover<M> rc;
            rc.init();
            for (uint32_t idx = lane; idx < nvec; idx += 32) rc.accum(__ldg(q4 + idx), ldg_stream(rp + idx));
            raw = rc.finish(fatal);
        }
        fatal_any |= fatal;
        if (lane == 0) {
            const uint32_t idr = p.id_rank ? __ldg(p.id_rank + row) : row;
            col.push(((u64)order_key(rank_value(M, raw)) << 32) | idr, ((u64)__float_as_uint(raw) << 32) | row);
        }
    }
    if (fatal_any && lane == 0) p.flags[q] = 2u;
    // The single-pass filter does not test its scores for Inf / NaN: every partial sum of a dot product is bounded by
    // |q| |row| (and |x|^2 by max|row|^2), so below these magnitudes none can occur; above them the whole batch is
    // redone on the exact path (distances.rs:59-98 recovery semantics live there).
    if (threadIdx.x == 0 && p.bad != nullptr &&
        !(sqrtf(qn2) * p.max_row_norm < 1.0e37f && p.max_row_norm < 1.0e18f && qn2 < 1.0e36f))
        *p.bad = 1u;
    const uint32_t kept = col.compact();
    for (uint32_t i = threadIdx.x; i < kept; i += blockDim.x) {
        p.out_pays[(size_t)q * p.k + i] = col.pays[i];
        if (p.out_keys) p.out_keys[(size_t)q * p.k + i] = col.keys[i];
    }
    if (threadIdx.x == 0) {
        p.out_counts[q] = kept;
        // completeness check (see above); only needed when candidates were actually dropped
        if (cnt == p.kprime && p.kprime < p.n && kept == p.k) {
            float bound = p.err_coeff * sqrtf(qn2) * p.max_row_norm;
            float worst_kept_approx = rank_from_key(p.cand_keys[(size_t)q * p.kprime + cnt - 1]);
            float exact_kth = rank_from_key(col.keys[p.k - 1]);
            if (p.l2_family) {
                // compare squared distances: the approximate rank lacks |q|^2, carries twice the dot error and the
                // fp32 rounding of the row-norm mirror and of the subtraction
                worst_kept_approx += qn2;
                bound = 2.0f * bound + 1.0e-6f * (p.max_row_norm * p.max_row_norm + qn2);
                if (M == kL2) exact_kth = exact_kth * exact_kth;
                exact_kth *= 1.000001f;   // the exact value's own rounding (and the square of the rounded root)
            }
            if (!(worst_kept_approx - bound > exact_kth) && p.flags[q] == 0u) p.flags[q] = 1u;
        }
    }
    (void)s_qn2;
}


// queries [nq, dims] -> per (query block, 32-dim chunk) a 64 KB image: hi [256 x 32] | lo [256 x 32], each
// already in the UMMA K-major SWIZZLE_128B shared-memory layout (zero rows beyond nq).
__global__ void split_queries_kernel(const float* q, uint32_t nq, uint32_t qblocks, uint32_t dims, unsigned char* blobs) {
    const uint32_t chunks = dims / 32;
    const size_t total = (size_t)qblocks * kGmN * dims;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(i % dims);
        const size_t row = i / dims;                       // padded query index
        const uint32_t qb = (uint32_t)(row / kGmN), n = (uint32_t)(row % kGmN);
        const float x = row < nq ? q[row * dims + k] : 0.0f;
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        unsigned char* blob = blobs + ((size_t)qb * chunks + k / 32u) * 65536u;
        const uint32_t off = tc::sw128_offset(n, k % 32u);
        *reinterpret_cast<float*>(blob + off) = h;
        *reinterpret_cast<float*>(blob + 32768u + off) = x - h;
    }
}

bool flat_gemm_eligible(int metric, size_t dims, size_t stride, size_t nq, size_t k, size_t n) {
    if (std::getenv("VB_FLAT_NO_GEMM")) return false;
    if (metric != kCosine && metric != kInnerProduct && metric != kNegativeInnerProduct && metric != kL2 &&
        metric != kL2Squared)
        return false;
    if (dims % 32 != 0 || stride != dims) return false;
    const char* min_env = std::getenv("VB_FLAT_GEMM_MIN_BATCH");
    const size_t min_batch = min_env ? (size_t)std::atoi(min_env) : 16;
    if (nq < min_batch) return false;
    if (k == 0 || k > 100 || n < 1024) return false;   // k' <= 192 (one pass) / 128 (3xTF32) candidates kept per query
    return true;
}

// max over rows of |row| (f32 from an f64 sum), one warp per row; non-negative floats order like uints
__global__ void row_norm_max_kernel(const float* rows, size_t stride, uint32_t dims, uint32_t n, uint32_t* out_bits,
                                    float* norm2_out) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.0f;
    for (size_t r = warp; r < n; r += warps) {
        double s = 0.0;
        for (uint32_t c = lane; c < dims; c += 32) {
            const double v = rows[r * stride + c];
            s = fma(v, v, s);
        }
        s = warp_sum(s);
        best = fmaxf(best, (float)sqrt(s) * 1.0000002f);
        if (norm2_out && lane == 0) norm2_out[r] = (float)s;
    }
    if (lane == 0) atomicMax(out_bits, __float_as_uint(best));
}

Status flat_gemm_max_row_norm(SearchCtx& ctx, const float* d_rows, size_t stride, size_t n, size_t dims, float* out,
                              float* d_norm2_out) {
    VB_TRY(ctx.q_norms.reserve(16));
    uint32_t* d_bits = ctx.q_norms.as<uint32_t>() + 2;
    VB_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(uint32_t), ctx.stream));
    row_norm_max_kernel<<<148 * 8, 256, 0, ctx.stream>>>(d_rows, stride, (uint32_t)dims, (uint32_t)n, d_bits, d_norm2_out);
    uint32_t bits = 0;
    VB_CUDA(cudaMemcpyAsync(&bits, d_bits, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    VB_CUDA(cudaStreamSynchronize(ctx.stream));
    std::memcpy(out, &bits, 4);
    return Status::Ok();
}

typedef void (*RescoreKernel)(const RescoreParams);
static RescoreKernel rescore_lookup(int metric) {
    switch (metric) {
        case kCosine: return flat_gemm_rescore_kernel<kCosine>;
        case kInnerProduct: return flat_gemm_rescore_kernel<kInnerProduct>;
        case kNegativeInnerProduct: return flat_gemm_rescore_kernel<kNegativeInnerProduct>;
        case kL2: return flat_gemm_rescore_kernel<kL2>;
        case kL2Squared: return flat_gemm_rescore_kernel<kL2Squared>;
    }
    return nullptr;
}

// TF32 passes of this thread's most recent batched search (bench.py reports issued flops from it).
static thread_local int t_gemm_terms = 0;
extern "C" int vb_debug_gemm_terms() { return t_gemm_terms; }

// Device part: queries already in device memory ([nq, dims]); leaves, on `stream`, the exact top-k
// payloads [nq][k] (raw bits << 32 | row), optional exact keys, counts [nq] and flags [nq]
// (0 ok, 1 redo on the single-query path, 2 metric overflow); *d_bad != 0 when a tensor-core
// score was non-finite (redo the whole batch). No host synchronisation.
Status flat_gemm_search_device(SearchCtx& ctx, int metric, const float* d_rows, size_t stride,
                               const uint32_t* d_id_rank, size_t n, size_t dims, float max_row_norm,
                               const float* d_row_norm2, const float* d_queries, size_t nq, size_t k, u64* d_out_keys, u64* d_out_pays,
                               uint32_t* d_out_counts, uint32_t* d_out_flags, uint32_t* d_bad, cudaStream_t stream,
                               int force_terms, int* terms_used) {
    int dev = 0, sms = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // One TF32 pass with a wide candidate margin; 3xTF32 (narrow margin) only when forced — VB_GEMM_TERMS=3, or the
    // caller's second tier for the queries whose single-pass candidate set could not be proven complete.
    // Margins: the kept set must reach `bound` (2e-3 |q| |row| for one pass) below the exact k-th score. For unit
    // vectors of 768 dims the scores of a 12.5M-row shard put ~30 rows that close to the 100th best, so k' = 192
    // candidates for k = 100 (k' = k + 32 up to k = 32) leave a wide reserve; denser corpora flag and take tier two.
    // k' stays 64 below the point where a list is cut (256 of its 512 slots): a cut must buy room for many appends
    // (k' = 256 cut every list after every tile pair: 168 ms instead of 40 per batch on the C3 shard).
    int terms = 1;
    if (const char* e = std::getenv("VB_GEMM_TERMS")) {
        const int t = std::atoi(e);
        if (t == 3 || t == 1) terms = t;
    }
    if (force_terms == 3) terms = 3;
    if (terms_used) *terms_used = terms;
    t_gemm_terms = terms;
    // VB_GEMM_PAIR=1: the single-pass kernel as CTA pairs (clusters of 2, tcgen05 cta_group::2, two accumulator sets).
    // Measured slower than the single-CTA form (3.22 vs 2.99 ms per 1024 x 1M x 768 batch), so it is opt-in; see DESIGN.md.
    const bool pair = terms == 1 && sms >= 2 && std::getenv("VB_GEMM_PAIR") != nullptr;
    const uint32_t nblk = pair ? (uint32_t)kG2N : (uint32_t)kGmN;            // queries per block
    const uint32_t qblocks_total = (uint32_t)((nq + nblk - 1) / nblk);
    const size_t nq_pad = (size_t)qblocks_total * nblk;
    const uint32_t workers = pair ? (uint32_t)sms / 2u : (uint32_t)sms;      // CTAs, or CTA pairs, that can be resident
    const uint32_t group = std::min<uint32_t>(qblocks_total, workers);         // query blocks per launch
    const size_t margin = terms == 1 ? (k <= 32 ? 32 : std::max<size_t>(64, 2 * k)) : std::max<size_t>(8, k / 4);
    const size_t kprime = std::min<size_t>(std::min<size_t>(k + margin, terms == 1 ? 192 : 128), n);   // approximate candidates kept
    RescoreKernel rescore = rescore_lookup(metric);
    if (!rescore) return Status::Cuda("metric not served by the batched kernel");
    const bool l2_family = metric == kL2 || metric == kL2Squared;
    if (l2_family && d_row_norm2 == nullptr) return Status::Cuda("batched L2 search needs the row-norm mirror");

    VB_TRY(ctx.staging.reserve(2 * nq_pad * dims * sizeof(float)));
    unsigned char* q_blobs = ctx.staging.as<unsigned char>();
    if (terms == 1) pack_queries_kernel<<<148 * 4, 256, 0, stream>>>(d_queries, (uint32_t)nq, qblocks_total, nblk, (uint32_t)dims, q_blobs);
    else split_queries_kernel<<<148 * 4, 256, 0, stream>>>(d_queries, (uint32_t)nq, qblocks_total, (uint32_t)dims, q_blobs);
    VB_CUDA(cudaGetLastError());

    const size_t max_ctas = (size_t)sms;
    const uint32_t list_cap = (kprime <= 64 && terms != 1) ? kGmListSmall : kGmListLarge;   // single pass: 256 rows between cuts
    VB_TRY(ctx.cand_keys.reserve(max_ctas * nblk * list_cap * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve(max_ctas * nblk * list_cap * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve(max_ctas * nblk * sizeof(uint32_t) + 16));
    VB_TRY(ctx.dump_keys.reserve(nq_pad * kprime * sizeof(u64)));
    VB_TRY(ctx.dump_pays.reserve(nq_pad * kprime * sizeof(u64)));
    VB_TRY(ctx.staging_rank.reserve(nq_pad * sizeof(uint32_t)));

    const size_t smem_bytes = pair ? (size_t)kG2Stages * (2 * 16384 + kG2N * 128 / 2) + 1024
                            : terms == 1 ? (size_t)kG1Stages * kG1StageBytes + 1024 : (size_t)kGmStages * kGmStageBytes + 1024;
    if (pair) VB_TRY(ensure_dynamic_smem_for(flat_gemm1_topk_kernel<true>, smem_bytes));
    else if (terms == 1) VB_TRY(ensure_dynamic_smem_for(flat_gemm1_topk_kernel<false>, smem_bytes));
    else VB_TRY(ensure_dynamic_smem_for(flat_gemm_topk_kernel, smem_bytes));
    const size_t blob_bytes = terms == 1 ? (size_t)nblk * 128u : 65536u;

    CUtensorMap tmap_a;
    VB_TRY(make_tmap_rows_sw128(d_rows, n, stride, kGmTile, &tmap_a));
    uint32_t cap = 256;
    while (cap < 2 * kprime + 64) cap <<= 1;
    u64* apx_keys = ctx.dump_keys.as<u64>();
    u64* apx_pays = ctx.dump_pays.as<u64>();
    uint32_t* apx_counts = ctx.staging_rank.as<uint32_t>();
    VB_CUDA(cudaMemsetAsync(d_out_flags, 0, nq * sizeof(uint32_t), stream));
    VB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), stream));

    // Pre-pass over a sample of the rows: with 148 CTAs sharing ~4 query blocks, a CTA sees only 1/37 of the rows
    // of its queries, so a filter built from ITS rows lets ~k' ln(n) / (n / 37) of them through — with 32 rows per
    // warp that sent 20-40 % of all (warp, query) pairs down the append path, and the epilogue, not the tensor
    // pipe, set the pace (ncu: 20 % of the first-level branches taken). The k'-th best score of ANY subset of the
    // rows bounds the final k'-th from above: one small launch over the first rows + its merge gives every CTA of
    // the main launch a filter at quantile k' / sample from its first tile on.
    // Sample = 256 rows per CTA of the launch (one tile pair): while thresholds are open EVERY score is appended, and
    // that phase costs ~170 us per pair per CTA, so a longer sample buys a tighter filter at a steep price
    // (measured at 1M x 768, Q = 1024: 9.5k rows 313k queries/s, 19k 287k, 33k 282k, 66k 272k).
    const uint32_t ranges_full = std::max<uint32_t>(1, workers / group);
    size_t sample = std::min<size_t>(n / 8, (size_t)ranges_full * (pair ? 512 : 256));   // one work unit per CTA / CTA pair
    // Single-pass kernel with long lists (k' > 64): the dense pre-pass costs only sample / n of the main launch, and the
    // appends it saves grow with k' — n / 32 rows, at most 151 552 (measured, Q = 1024, k = 100: 1M rows 4.62 -> 3.61 ms
    // at 37 888; 12.5M rows 40.9 -> 37.9 / 35.2 / 34.1 / 34.4 ms at 37 888 / 75 776 / 151 552 / 303 104; k = 10 is
    // fastest with the short sample).
    // The dump is [queries x sample] floats: the long sample stays under 1 GiB of workspace whatever the batch size.
    if (terms == 1 && kprime > 64)
        sample = std::max(sample, std::min<size_t>({(n / 32) & ~(size_t)255, (size_t)151552, (((size_t)1 << 28) / nq_pad) & ~(size_t)255}));
    if (const char* e = std::getenv("VB_GEMM_SAMPLE")) sample = std::min<size_t>(n, std::max<size_t>(1024, (size_t)std::atol(e)));
    const bool prepass = n >= 65536 && sample >= 2048 && !std::getenv("VB_GEMM_NO_PREPASS");
    // Single-pass kernel: the pre-pass is a dense dump — ranks of [queries x sample rows] straight to HBM (coalesced, no
    // lists), then one small CTA per query bisects its k'-th smallest. (Through the list machinery the same pass cost
    // 290 us + a 70 us merge per 1024-query batch: with the bounds open every score took the append path.)
    const bool dense_prepass = prepass && terms == 1 && sample <= 524288 && !std::getenv("VB_GEMM_LIST_PREPASS");
    u64* pre_keys = nullptr;
    uint32_t* pre_counts = nullptr;
    float* pre_dump = nullptr;
    float* pre_rank = nullptr;
    const uint32_t dump_stride = (uint32_t)((sample + 3) & ~(size_t)3);
    if (dense_prepass) {
        VB_TRY(ctx.dump_keys2.reserve(nq_pad * (size_t)dump_stride * sizeof(float)));
        VB_TRY(ctx.hist.reserve(nq_pad * sizeof(float)));
        pre_dump = ctx.dump_keys2.as<float>();
        pre_rank = ctx.hist.as<float>();
        VB_TRY(ensure_dynamic_smem_for(gemm1_sample_select_kernel, kSelSmem));
    } else if (prepass) {
        VB_TRY(ctx.dump_keys2.reserve(nq_pad * kprime * sizeof(u64)));
        VB_TRY(ctx.dump_pays2.reserve(nq_pad * kprime * sizeof(u64)));
        VB_TRY(ctx.hist.reserve(nq_pad * sizeof(uint32_t)));
        pre_keys = ctx.dump_keys2.as<u64>();
        pre_counts = ctx.hist.as<uint32_t>();
    }
    for (uint32_t qb0 = 0; qb0 < qblocks_total; qb0 += group) {
        const uint32_t qblocks = std::min(group, qblocks_total - qb0);
        const size_t q0 = (size_t)qb0 * nblk;
        const uint32_t nq_here = (uint32_t)std::min<size_t>(nq - q0, (size_t)qblocks * nblk);
        GemmParams p{};
        p.dims = (uint32_t)dims;
        p.nq = nq_here;
        p.k = (uint32_t)kprime;
        p.metric = metric;
        p.rank_scale = l2_family ? -2.0f : -1.0f;
        p.row_norm2 = l2_family ? d_row_norm2 : nullptr;
        p.qblocks = qblocks;
        p.pair = pair ? 1u : 0u;
        p.nblk = nblk;
        p.ranges = std::max<uint32_t>(1, workers / qblocks);
        p.id_rank = d_id_rank;
        p.list_cap = list_cap;
        p.list_keys = ctx.cand_keys.as<u64>();
        p.list_pays = ctx.cand_pays.as<u64>();
        p.list_counts = ctx.cand_counts.as<uint32_t>();
        p.bad = d_bad;
        { const char* dbg = std::getenv("VB_GEMM_DEBUG"); p.debug = dbg ? (uint32_t)std::atoi(dbg) : 0u; }
        const uint32_t grid = p.qblocks * p.ranges * (pair ? 2u : 1u);
        const unsigned char* blobs = q_blobs + (size_t)qb0 * (dims / 32) * blob_bytes;
        for (int pass = prepass ? 0 : 1; pass < 2; ++pass) {
            p.n = pass == 0 ? (uint32_t)sample : (uint32_t)n;
            p.init_keys = pass == 0 || dense_prepass ? nullptr : (prepass ? pre_keys + q0 * kprime : nullptr);
            p.init_counts = pass == 0 || dense_prepass ? nullptr : (prepass ? pre_counts + q0 : nullptr);
            p.score_dump = pass == 0 && dense_prepass ? pre_dump + q0 * dump_stride : nullptr;
            p.dump_stride = dump_stride;
            p.init_rank = pass == 1 && dense_prepass ? pre_rank + q0 : nullptr;
            if (pair) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(kG1Threads);
                cfg.dynamicSmemBytes = smem_bytes;
                cfg.stream = stream;
                cudaLaunchAttribute attr{};
                attr.id = cudaLaunchAttributeClusterDimension;
                attr.val.clusterDim.x = 2;
                attr.val.clusterDim.y = 1;
                attr.val.clusterDim.z = 1;
                cfg.attrs = &attr;
                cfg.numAttrs = 1;
                VB_CUDA(cudaLaunchKernelEx(&cfg, flat_gemm1_topk_kernel<true>, tmap_a, blobs, p));
            } else if (terms == 1) {
                flat_gemm1_topk_kernel<false><<<grid, kG1Threads, smem_bytes, stream>>>(tmap_a, blobs, p);
            } else {
                flat_gemm_topk_kernel<<<grid, kGmThreads, smem_bytes, stream>>>(tmap_a, blobs, p);
            }
            VB_CUDA(cudaGetLastError());
            if (pass == 0 && dense_prepass)
                gemm1_sample_select_kernel<<<nq_here, kSelThreads, kSelSmem, stream>>>(
                    pre_dump + q0 * dump_stride, dump_stride, (uint32_t)sample, (uint32_t)kprime, pre_rank + q0);
            else if (pass == 0)
                flat_gemm_merge_kernel<<<nq_here, 128, (size_t)cap * 16, stream>>>(p, cap, pre_keys + q0 * kprime,
                                                                                  ctx.dump_pays2.as<u64>() + q0 * kprime, pre_counts + q0);
            else
                flat_gemm_merge_kernel<<<nq_here, 128, (size_t)cap * 16, stream>>>(p, cap, apx_keys + q0 * kprime,
                                                                                  apx_pays + q0 * kprime, apx_counts + q0);
            VB_CUDA(cudaGetLastError());
        }
    }
    RescoreParams rp{};
    rp.rows = d_rows;
    rp.stride = stride;
    rp.dims = (uint32_t)dims;
    rp.n = (uint32_t)n;
    rp.k = (uint32_t)k;
    rp.kprime = (uint32_t)kprime;
    rp.id_rank = d_id_rank;
    rp.queries = d_queries;
    rp.cand_keys = apx_keys;
    rp.cand_pays = apx_pays;
    rp.cand_counts = apx_counts;
    // |approx - exact| <= err_coeff * |q| * max|row|. One pass: both operands are cut to TF32 (< 2^-10 relative
    // each => 2^-9 on every product, Cauchy-Schwarz over the sum). 3xTF32: the dropped lo*lo term and the
    // truncation of the two lo operands (<= 2^-20 each, with head-room), plus fp32 accumulation over dims / 8
    // MMA steps per term in both cases.
    rp.err_coeff = (terms == 1 ? 1.96e-3f : 4.0e-6f) + 1.0e-7f * (float)(3 * dims / 8);
    rp.max_row_norm = max_row_norm;
    rp.l2_family = l2_family ? 1 : 0;
    rp.out_keys = d_out_keys;
    rp.out_pays = d_out_pays;
    rp.out_counts = d_out_counts;
    rp.flags = d_out_flags;
    rp.bad = d_bad;
    rescore<<<(unsigned)nq, 128, 0, stream>>>(rp);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

Status flat_gemm_search(SearchCtx& ctx, int metric, const float* d_rows, size_t stride, const uint32_t* d_id_rank,
                        size_t n, size_t dims, float max_row_norm, const float* d_row_norm2, const float* h_queries, size_t nq,
                        size_t k, GemmResult* out, int force_terms) {
    const size_t qbytes = nq * dims * sizeof(float);
    VB_TRY(ctx.h_queries.reserve(qbytes));
    VB_TRY(ctx.queries.reserve(qbytes));
    std::memcpy(ctx.h_queries.p, h_queries, qbytes);
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, ctx.h_queries.p, qbytes, cudaMemcpyHostToDevice, ctx.stream));
    // result block: pays [nq][k] | counts [nq] | flags [nq] | bad
    const size_t res_bytes = nq * k * sizeof(u64) + 2 * nq * sizeof(uint32_t) + 16;
    VB_TRY(ctx.result.reserve(res_bytes));
    u64* out_pays = ctx.result.as<u64>();
    uint32_t* out_counts = reinterpret_cast<uint32_t*>(out_pays + nq * k);
    uint32_t* out_flags = out_counts + nq;
    uint32_t* d_bad = out_flags + nq;
    Status s = flat_gemm_search_device(ctx, metric, d_rows, stride, d_id_rank, n, dims, max_row_norm, d_row_norm2,
                                       ctx.queries.as<float>(), nq, k, nullptr, out_pays, out_counts, out_flags, d_bad,
                                       ctx.stream, force_terms, &out->terms);
    if (!s.ok()) { ctx.poison(); return s; }
    VB_TRY(ctx.h_result.reserve(res_bytes));
    cudaError_t e = cudaMemcpyAsync(ctx.h_result.p, ctx.result.p, res_bytes, cudaMemcpyDeviceToHost, ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
    if (e != cudaSuccess) {
        ctx.poison();
        return Status::Cuda(cudaGetErrorString(e));
    }
    const u64* pays = ctx.h_result.as<u64>();
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(pays + nq * k);
    const uint32_t* flags = counts + nq;
    out->non_finite = flags[nq] != 0;
    out->k = k;
    out->counts.assign(counts, counts + nq);
    out->flags.resize(nq);
    out->rows.resize(nq * k);
    out->raws.resize(nq * k);
    for (size_t q = 0; q < nq; ++q) {
        out->flags[q] = (uint8_t)flags[q];
        for (size_t i = 0; i < k; ++i) {
            const u64 pay = pays[q * k + i];
            uint32_t bits = (uint32_t)(pay >> 32);
            std::memcpy(&out->raws[q * k + i], &bits, 4);
            out->rows[q * k + i] = (uint32_t)pay;
        }
    }
    return Status::Ok();
}

}  // namespace vb
