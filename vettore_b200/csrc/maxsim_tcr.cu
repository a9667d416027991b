// maxsim_tcr.cu — K5 on the tensor cores for RAGGED documents (reference multi_vector.rs:65-132, the
// inner-product family): documents of any length, packed back to back in the token matrix, up to 64 query
// tokens, any dimension up to 128. Same pipeline as maxsim_tc.cu (TMA ring -> hi/lo split into TMEM -> 3xTF32
// tcgen05.mma -> two epilogue groups on alternate tiles); what changes is the work split and the epilogue:
//
//  * A CTA owns a contiguous, DOCUMENT-ALIGNED token range (equal token counts, boundaries moved to the next
//    document start by a binary search over doc_off), cut into 128-token tiles that ignore document boundaries.
//    No document is shared between CTAs, so nothing is combined across the grid.
//  * An epilogue warp owns one 32-token chunk of a tile. `tok_doc[token]` names each lane's document; the lanes
//    where it changes cut the chunk into segments. A chunk inside one document (the common case) takes the same
//    31-shuffle transposing max butterfly as the uniform kernel; otherwise one masked butterfly per segment.
//  * A document that crosses chunk (tile, epilogue-group) boundaries is max-combined through a carry chain in
//    shared memory: chunk c publishes the running per-query maxima of its open last segment (release store of
//    c + 1 into the slot's flag), chunk c + 1 acquires them before it closes or extends the segment. Waits only
//    ever point at lower chunk numbers and a warp publishes before it can block on a CTA barrier, so the chain
//    cannot deadlock; 32 slots cover the 16 chunks two epilogue groups can be apart.
//  * The document whose last token lies in the chunk is finished there: per-query maxima go through a
//    shared-memory row, one lane adds them in query order (f32, multi_vector.rs:81-84) and pushes the score.
//  * Empty documents (score 0.0, multi_vector.rs:102-106) own no token: a closing pass over the CTA's document
//    range pushes them.
// A non-finite pair score (the reference recomputes those in f64, distances.rs:59-98) raises the error word;
// the host then repeats the query on the general kernel, which is the arbiter of those semantics.
#include "maxsim.h"

#include <cfloat>
#include <cstdlib>

#include "maxsim_tc.cuh"
#include "scan_driver.h"
#include "topk.cuh"

namespace vb {

constexpr uint32_t kTcrCarrySlots = 32;
constexpr uint32_t kTcrEmptyRound = 256;   // empty documents examined per checkpoint round (one per epilogue thread)

struct MaxSimTcrParams {
    uint32_t ndocs, ntok, dims, tq;
    int metric;                   // kInnerProduct, kNegativeInnerProduct or kCosineTrue
    const uint32_t* doc_off;      // [ndocs + 1]
    const uint32_t* doc_rank;     // [ndocs] or null (rank = document index)
    const uint32_t* tok_doc;      // [ntok] owning document of every token
    const float* inv_dnorm;       // [ntok] 1/|token| (0 for zero tokens), cosine only
    const float* query;           // [tq, dims] dense
    const float* inv_qnorm;       // [64], cosine only
    uint32_t cap, stages;
    uint32_t ckpt_tiles, slack;   // both epilogue groups meet every ckpt_tiles tiles (even); pushes in between <= slack
    uint32_t has_empty;           // some document has no token
    uint32_t no_scan;             // timing experiments (VB_MAXSIM_NO_SCAN): masked butterflies instead of the transposition tile
    u64* dump_keys;               // limit beyond the fused collector: every live document's key / payload goes to
    u64* dump_pays;               // [ndocs] arrays (pre-filled with kKeyMax) and the host radix-sorts them
    uint32_t* err;
    TopkWorkspace ws;
};

__device__ __forceinline__ uint32_t ld_acquire_smem(uint32_t smem_address) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_address) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_smem(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(tc::smem_addr(p)), "r"(v) : "memory");
}

// warp_transpose_max over the lanes of one segment only: lanes outside contribute -inf. The mask is folded
// into the first exchange level so only 16 extra registers are live next to v.
__device__ __forceinline__ float warp_transpose_max_masked(const float (&v)[32], bool inseg, int lane) {
    float w[16];
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = inseg ? v[i] : -INFINITY, b = inseg ? v[i + 16] : -INFINITY;
            const float send = hi ? a : b, keep = hi ? b : a;
            w[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
        }
    }
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool hi = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? w[i] : w[i + half];
            const float keep = hi ? w[i + half] : w[i];
            w[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, half));
        }
    }
    return w[0];
}

struct RingPos {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t stages) {
        if (++s == stages) { s = 0; ph ^= 1u; }
    }
};

template <int N>
__global__ void __launch_bounds__(kTcThreads, 1)
maxsim_tcr_kernel(const __grid_constant__ CUtensorMap tmap, const MaxSimTcrParams p) {
    constexpr uint32_t kChains = 128 / N;    // partial accumulators per buffer (4 x 32 or 2 x 64 columns)
    constexpr uint32_t kPasses = N / 32;     // the epilogue works on 32 query columns at a time
    constexpr uint32_t kBBlock = N * 128;    // bytes of one K block of the B operand
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages];
    __shared__ __align__(8) uint64_t a_ready[2], a_free[2], d_full[kTcAccBufs], d_free[kTcAccBufs];
    __shared__ uint32_t tmem_slot;
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;
    __shared__ uint32_t s_range[4];                              // T0, T1 (tokens), D0, D1 (documents)
    __shared__ float s_invq[N];
    __shared__ __align__(16) float s_carry[kTcrCarrySlots][N];
    __shared__ uint32_t s_flag[kTcrCarrySlots][kPasses];
    __shared__ __align__(16) float s_fin[kTcEpiWarps][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t KB = (p.dims + 31u) / 32u;                   // 128-byte K blocks per row (1..4)
    const uint32_t chunks_per_tile = (KB + 1) / 2;              // a ring chunk = up to 2 K blocks
    const uint32_t stages = p.stages;
    unsigned char* ring = smem;
    unsigned char* b_hi = ring + (size_t)stages * kTcChunkBytes;   // KB x [N rows x 128 B]
    unsigned char* b_lo = b_hi + (size_t)KB * kBBlock;
    unsigned char* col_mem = b_lo + (size_t)KB * kBBlock;
    // N = 32 only: one [32 queries][33] transposition tile per epilogue warp, behind the collector (multi-segment chunks)
    float* scan_tiles = reinterpret_cast<float*>(col_mem + (size_t)p.cap * 16);

    Collector col;
    col.init(col_mem, &s_thresh, &s_count, p.cap, p.ws.k, kTcEpiWarps * 32, 2);
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], kTcSplitWarps);
        }
        for (int u = 0; u < 2; ++u) {
            tc::mbar_init(&a_ready[u], kTcSplitWarps);
            tc::mbar_init(&a_free[u], 1);
        }
        for (int b = 0; b < kTcAccBufs; ++b) {
            tc::mbar_init(&d_full[b], 1);
            tc::mbar_init(&d_free[b], kTcEpiGroupWarps);
        }
        tc::mbar_fence_init();
    }
    if (warp == kTcMmaWarp) tc::tmem_alloc(&tmem_slot, 512);
    if (tid < N) s_invq[tid] = (p.inv_qnorm && (uint32_t)tid < p.tq) ? p.inv_qnorm[tid] : 0.0f;
    for (uint32_t i = tid; i < kTcrCarrySlots * kPasses; i += kTcThreads) (&s_flag[0][0])[i] = 0u;
    // This CTA's documents [D0, D1): boundary c = the first document that starts at or after token ntok * c / grid.
    if (tid == 64 || tid == 96) {
        const uint32_t c = blockIdx.x + (tid == 96 ? 1u : 0u);
        uint32_t d = p.ndocs;
        if (c < gridDim.x) {
            const uint32_t target = (uint32_t)(((uint64_t)p.ntok * c) / gridDim.x);
            uint32_t lo = 0, hi = p.ndocs;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(p.doc_off + mid) < target) lo = mid + 1; else hi = mid;
            }
            d = lo;
        }
        s_range[tid == 96 ? 3 : 2] = d;
        s_range[tid == 96 ? 1 : 0] = __ldg(p.doc_off + d);
    }
    // B operand: the query tokens, split hi/lo, UMMA K-major SWIZZLE_128B layout; rows >= tq and columns >= dims are zero.
    for (uint32_t idx = tid; idx < (uint32_t)N * KB * 32u; idx += kTcThreads) {
        const uint32_t n = idx / (KB * 32u), k = idx % (KB * 32u);
        const float x = (n < p.tq && k < p.dims) ? p.query[(size_t)n * p.dims + k] : 0.0f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const uint32_t off = (k / 32u) * kBBlock + tc::sw128_offset(n, k % 32u);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = x - hi;
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t T0 = s_range[0], T1 = s_range[1], D0 = s_range[2], D1 = s_range[3];
    const uint32_t my_tiles = (T1 - T0 + kTcTile - 1) / kTcTile;

    if (warp == kTcProducerWarp) {
        // ===== producer: one chunk (<= 2 K blocks of the tile) per ring stage =====
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap);
            RingPos r;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                for (uint32_t j = 0; j < chunks_per_tile; ++j, r.next(stages)) {
                    const uint32_t blocks = min(2u, KB - 2u * j);
                    tc::mbar_wait(&empty_bar[r.s], r.ph ^ 1u);
                    tc::mbar_arrive_expect_tx(&full_bar[r.s], blocks * 16384u);
                    for (uint32_t h = 0; h < blocks; ++h)
                        tc::tma_load_2d(ring + (size_t)r.s * kTcChunkBytes + (size_t)h * 16384, &tmap, (2u * j + h) * 32u,
                                        T0 + it * kTcTile, &full_bar[r.s]);
                }
            }
        }
    } else if (warp == kTcMmaWarp) {
        // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues =====
        const uint32_t idesc = tc::umma_idesc_tf32(kTcTile, N);
        const uint64_t bh0 = tc::umma_smem_desc_sw128(tc::smem_addr(b_hi));
        const uint64_t bl0 = tc::umma_smem_desc_sw128(tc::smem_addr(b_lo));
        uint32_t cc = 0;
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t b = it % kTcAccBufs;
            const uint32_t d_tmem = tbase + kTcAccCol + b * 128u;
            for (uint32_t j = 0; j < chunks_per_tile; ++j, ++cc) {
                const uint32_t u = cc & 1u;
                const uint32_t blocks = min(2u, KB - 2u * j);
                tc::mbar_wait(&a_ready[u], (cc >> 1) & 1u);
                if (j == 0) tc::mbar_wait(&d_free[b], ((it / kTcAccBufs) & 1u) ^ 1u);
                tc::fence_after_sync();
                // The k-steps of a block rotate over the partial accumulators, so back-to-back MMAs do not wait
                // on each other's result; the epilogue adds the partials. Descriptor start addresses count
                // 16-byte units: one K block of B = N * 8, one k-step = 2.
                const uint32_t a0 = tbase + u * 128u;
                const uint64_t kbo = (uint64_t)(2u * j) * (kBBlock / 16u);
                if (tc::elect_one()) {
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
                        if (h < blocks) {
                            const bool first_block = (j | h) == 0u;
#pragma unroll
                            for (uint32_t term = 0; term < 3; ++term) {
#pragma unroll
                                for (uint32_t ks = 0; ks < 4; ++ks) {
                                    const uint64_t bdesc = (term == 1 ? bl0 : bh0) + kbo + (uint64_t)(h * (kBBlock / 16u) + ks * 2u);
                                    const uint32_t a_addr = a0 + h * 32u + ks * 8u + (term == 2 ? 64u : 0u);
                                    const uint32_t accumulate = (first_block && term == 0 && ks < kChains) ? 0u : 1u;
                                    tc::umma_tf32_ts(d_tmem + (ks % kChains) * N, a_addr, bdesc, idesc, accumulate);
                                }
                            }
                        }
                    }
                    tc::umma_commit(&a_free[u]);                                 // this A buffer may be overwritten
                    if (j + 1 == chunks_per_tile) tc::umma_commit(&d_full[b]);   // the accumulator is complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= kTcEpiWarps) {
        // ===== split warps: smem chunk -> (hi, lo) -> TMEM. Two warps per TMEM lane quarter, one K block each =====
        const uint32_t sw = warp - kTcEpiWarps;
        const uint32_t quarter = warp & 3u, h = sw >> 2;
        const uint32_t row = quarter * 32u + lane;            // token row within the tile == TMEM lane
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        uint32_t cc = 0;
        RingPos r;
        for (uint32_t it = 0; it < my_tiles; ++it) {
            for (uint32_t j = 0; j < chunks_per_tile; ++j, ++cc, r.next(stages)) {
                const uint32_t u = cc & 1u;
                const uint32_t blocks = min(2u, KB - 2u * j);
                tc::mbar_wait(&full_bar[r.s], r.ph);
                tc::mbar_wait(&a_free[u], ((cc >> 1) & 1u) ^ 1u);   // MMAs that read this A buffer are done
                tc::fence_after_sync();
                if (h < blocks) {
                    const unsigned char* blk = ring + (size_t)r.s * kTcChunkBytes + (size_t)h * 16384 + row * 128u;
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
                    tc::tmem_st32(lane_addr + u * 128u + h * 32u, hi);
                    tc::tmem_st32(lane_addr + u * 128u + 64u + h * 32u, lo);
                    tc::tmem_st_wait();
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&a_ready[u]);
                    tc::mbar_arrive(&empty_bar[r.s]);   // this warp's share of the stage is consumed
                }
            }
        }
    } else {
        // ===== epilogue warps 0-7: group (warp >> 2) handles tiles it = group, group + 2, ... =====
        const uint32_t quarter = warp & 3u, group = warp >> 2;   // TMEM lanes [32 * quarter, +32)
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        const uint32_t b = group;                                // accumulator buffer of this group (kTcAccBufs == 2)
        const bool cosine = p.metric == kCosineTrue;
        const uint32_t C = p.ckpt_tiles;
        float* fin = s_fin[warp];
        float* tile = scan_tiles + (size_t)warp * (32 * 33);
        u64 g_prefetch = kKeyMax;
        uint32_t d_next = 0, before_next = 0xFFFFFFFFu, after_next = 0xFFFFFFFFu;
        {
            const uint32_t s0 = T0 + (group * 4u + quarter) * 32u;          // this warp's first chunk
            if (s0 < T1) {
                d_next = __ldg(p.tok_doc + min(s0 + (uint32_t)lane, T1 - 1u));
                before_next = s0 > T0 ? __ldg(p.tok_doc + s0 - 1u) : 0xFFFFFFFFu;
                after_next = s0 + 32u < T1 ? __ldg(p.tok_doc + s0 + 32u) : 0xFFFFFFFFu;
            }
        }
        for (uint32_t it = group; it < my_tiles; it += 2u) {
            const uint32_t chunk = it * 4u + quarter;            // CTA-local chunk number
            const uint32_t s_tok = T0 + chunk * 32u;             // first token of the chunk
            const uint32_t token = s_tok + (uint32_t)lane;
            const bool valid = token < T1;
            // Document of every lane (loaded one tile ahead, below); lanes past the CTA's range borrow the last
            // valid token's document (they are masked to -inf). The neighbours just outside the chunk tell whether
            // its first document began earlier / its last one goes on: no dependent doc_off gather on the chain.
            const uint32_t d = d_next, d_before = before_next, d_after = after_next;
            {
                const uint32_t nchunk = chunk + 8u, ns = T0 + nchunk * 32u;   // this warp's chunk of tile it + 2
                if (ns < T1) {
                    d_next = __ldg(p.tok_doc + min(ns + (uint32_t)lane, T1 - 1u));
                    before_next = __ldg(p.tok_doc + ns - 1u);                  // ns > T0: nchunk >= 8
                    after_next = ns + 32u < T1 ? __ldg(p.tok_doc + ns + 32u) : 0xFFFFFFFFu;
                }
            }
            const uint32_t rank = p.doc_rank ? __ldg(p.doc_rank + d) : d;
            const float inv_dn = cosine ? (valid ? __ldg(p.inv_dnorm + token) : 0.0f) : 1.0f;
            const uint32_t dprev = __shfl_up_sync(0xffffffffu, d, 1);
            const uint32_t heads = s_tok < T1 ? __ballot_sync(0xffffffffu, lane == 0 || d != dprev) : 0u;
            const uint32_t d_first = __shfl_sync(0xffffffffu, d, 0), d_last = __shfl_sync(0xffffffffu, d, 31);
            const bool first_continues = s_tok > T0 && d_before == d_first;   // same document as the token before the chunk
            const bool last_continues = d_after == d_last;                    // ... as the token after it (none: 0xFFFFFFFF)

            tc::mbar_wait(&d_full[b], (it / kTcAccBufs) & 1u);
            tc::fence_after_sync();
            float my_total = 0.0f;            // lane j: running score of the chunk's j-th segment
            uint32_t my_doc = 0, my_rank = 0xFFFFFFFFu;
            bool my_final = false;
#pragma unroll 1
            for (uint32_t h = 0; h < kPasses; ++h) {
                float v[32];
                {
                    const uint32_t acc = lane_addr + kTcAccCol + b * 128u + h * 32u;
#pragma unroll
                    for (int hcol = 0; hcol < 2; ++hcol) {
                        uint32_t r[kChains][16];
#pragma unroll
                        for (uint32_t c = 0; c < kChains; ++c) tc::tmem_ld16(acc + c * N + hcol * 16, r[c]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            if constexpr (kChains == 4)
                                v[hcol * 16 + q] = (__uint_as_float(r[0][q]) + __uint_as_float(r[1][q])) +
                                                   (__uint_as_float(r[2][q]) + __uint_as_float(r[3][q]));
                            else
                                v[hcol * 16 + q] = __uint_as_float(r[0][q]) + __uint_as_float(r[1][q]);
                        }
                    }
                }
                if (h + 1 == kPasses) {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&d_free[b]);
                }
                // similarity_value: inner product -> dot; negative inner product -> -(-dot) = dot; cosine ->
                // dot / (|q| |d|) clamped (distances.rs:170-172): 1/|q| >= 0 and the clamp are monotonic, so they
                // are applied once per query after the max. A non-finite pair sends the query to the general kernel.
                float chk = 0.0f;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    chk += v[q];
                    const float sim = cosine ? v[q] * inv_dn : v[q];
                    v[q] = valid ? sim : -INFINITY;
                }
                if (valid && !(fabsf(chk) <= FLT_MAX)) atomicMin(p.err, 0u);

                // Segments are taken LAST to FIRST: the chunk's open last segment is published before the first
                // segment waits for the previous chunk, so only chunks that lie wholly inside one document (wait ->
                // max -> publish) are links of the serial carry chain; every butterfly stays off it.
                // A chunk cut by document boundaries (N = 32): instead of one masked butterfly per segment, the warp
                // transposes its 32 x 32 scores through shared memory once; lane q then walks ITS query's column
                // segment by segment (32 loads in all, whatever the number of segments), parks the maxima of the
                // documents that end here in the columns it has already consumed, and after the loop lane j adds the
                // 32 maxima of segment j in query order — every finished document of the chunk in parallel.
                const bool scan = N == 32 && (heads & (heads - 1u)) != 0u && !p.no_scan;
                uint32_t final_mask = 0u;
                if (scan) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) tile[q * 33 + lane] = v[q];
                    __syncwarp();
                }
                uint32_t hm = heads, j = 0, l1 = 32u;
                while (hm) {                                                  // warp-uniform: one round per segment
                    const uint32_t l0 = 31u - (uint32_t)__clz(hm);
                    hm &= ~(1u << l0);
                    float m;
                    if (scan) {
                        m = -INFINITY;
                        for (uint32_t t = l0; t < l1; ++t) m = fmaxf(m, tile[lane * 33 + t]);
                    } else {
                        m = heads == 1u ? warp_transpose_max(v, lane)
                                        : warp_transpose_max_masked(v, (uint32_t)lane >= l0 && (uint32_t)lane < l1, lane);
                    }
                    const uint32_t sd = __shfl_sync(0xffffffffu, d, l0), srank = __shfl_sync(0xffffffffu, rank, l0);
                    const bool starts_here = l0 != 0u || !first_continues, ends_here = l1 != 32u || !last_continues;
                    l1 = l0;
                    if (!starts_here) {                                       // the document began in an earlier chunk
                        const uint32_t slot = (chunk - 1u) % kTcrCarrySlots;
                        const uint32_t flag = tc::smem_addr(&s_flag[slot][h]);
                        while (ld_acquire_smem(flag) != chunk) {}
                        m = fmaxf(m, s_carry[slot][h * 32u + lane]);
                    }
                    if (!ends_here) {                                         // ... and goes on in the next one
                        const uint32_t slot = chunk % kTcrCarrySlots;
                        s_carry[slot][h * 32u + lane] = m;
                        __syncwarp();
                        if (lane == 0) st_release_smem(&s_flag[slot][h], chunk + 1u);
                    } else if (scan) {
                        if (cosine) m = fminf(1.0f, fmaxf(-1.0f, m * s_invq[lane]));
                        tile[lane * 33 + (31u - j)] = m;                       // column 31 - j >= l0: already consumed by this lane
                        final_mask |= 1u << j;
                        if ((uint32_t)lane == j) { my_doc = sd; my_rank = srank; my_final = true; }
                    } else {
                        if (cosine) m = fminf(1.0f, fmaxf(-1.0f, m * s_invq[h * 32u + lane]));
                        fin[lane] = m;
                        __syncwarp();
                        if ((uint32_t)lane == j) {
                            // per-query maxima added in query order (multi_vector.rs:81-84)
                            const float4* mv = reinterpret_cast<const float4*>(fin);
                            float mq[32];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 t4 = mv[i];
                                mq[4 * i] = t4.x; mq[4 * i + 1] = t4.y; mq[4 * i + 2] = t4.z; mq[4 * i + 3] = t4.w;
                            }
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (h * 32u + (uint32_t)q < p.tq) my_total += mq[q];
                            my_doc = sd;
                            my_rank = srank;
                            my_final = true;
                        }
                        __syncwarp();
                    }
                    ++j;
                }
                if (scan) {
                    __syncwarp();
                    if ((final_mask >> lane) & 1u) {                         // lane j: the document that closed segment j
                        float total = 0.0f;
                        const float* colp = tile + (31 - lane);
                        for (uint32_t q = 0; q < p.tq; ++q) total += colp[q * 33];   // query order (multi_vector.rs:81-84)
                        my_total = total;
                    }
                    __syncwarp();
                }
            }
            if (my_final && my_rank != 0xFFFFFFFFu) {
                // a non-finite running sum can never become finite again, so one check suffices
                if (!isfinite(my_total)) { atomicMin(p.err, (my_doc << 1) | 1u); my_total = 0.0f; }
                const u64 key = ((u64)(~order_key(my_total)) << 32) | my_rank;
                const u64 pay = ((u64)__float_as_uint(my_total) << 32) | my_doc;
                if (p.dump_keys) { p.dump_keys[my_doc] = key; p.dump_pays[my_doc] = pay; }
                else if (key < col.threshold()) col.push(key, pay);
            }
            // Both groups meet (all 8 warps) each time the CTA has finished another C tiles; a window only
            // counts when it lies inside this CTA's tiles, so both groups pass the same number of checkpoints.
            if (!p.dump_keys && (it + 2u) / C != it / C && (it / C + 1u) * C <= my_tiles)
                collector_checkpoint(col, p.ws, 0, p.slack, g_prefetch);
        }
        if (p.has_empty) {
            for (uint32_t d0 = D0; d0 < D1; d0 += kTcrEmptyRound) {
                if (!p.dump_keys) collector_checkpoint(col, p.ws, 0, kTcrEmptyRound, g_prefetch);
                const uint32_t d = d0 + (uint32_t)tid;
                if (d < D1 && __ldg(p.doc_off + d) == __ldg(p.doc_off + d + 1u)) {
                    const uint32_t rank = p.doc_rank ? __ldg(p.doc_rank + d) : d;
                    const u64 key = ((u64)(~order_key(0.0f)) << 32) | rank;
                    if (rank == 0xFFFFFFFFu) continue;
                    if (p.dump_keys) { p.dump_keys[d] = key; p.dump_pays[d] = (u64)d; }
                    else if (key < col.threshold()) col.push(key, (u64)d);
                }
            }
        }
        if (!p.dump_keys) collector_publish_and_merge(col, p.ws, 0, &s_last);
    }
    // teardown: every role is done with TMEM before it is released
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kTcMmaWarp) tc::tmem_dealloc(tbase, 512);
}

bool maxsim_tcr_eligible(const MaxSimJob& job) {
    if (std::getenv("VB_MAXSIM_NO_TC") || std::getenv("VB_MAXSIM_NO_TCR")) return false;
    if (job.metric != kInnerProduct && job.metric != kNegativeInnerProduct && job.metric != kCosineTrue) return false;
    if (job.metric == kCosineTrue && !job.d_inv_dnorm) return false;
    if (!job.d_tok_doc || !job.d_doc_off) return false;
    if (job.dims == 0 || job.dims > 128 || job.stride % 4 != 0) return false;
    if (job.tq == 0 || job.tq > 64) return false;
    if (job.ntok == 0 || job.ntok >= (1ull << 31)) return false;   // TMA coordinates are signed 32-bit
    if (std::min<size_t>(job.k, job.ndocs) > (size_t)kMaxFusedK && job.d_keys_out) return false;
    return true;
}

Status maxsim_tcr_top_k(SearchCtx& ctx, const MaxSimJob& job, MaxSimResult* out) {
    out->rows.clear();
    out->scores.clear();
    out->err = kNoError;
    const uint32_t k_out = (uint32_t)std::min<size_t>(job.k, job.ndocs);
    const bool dump = k_out > (uint32_t)kMaxFusedK;   // every live document's score is written out and radix-sorted
    const uint32_t k = dump ? 1u : k_out;              // the collector is idle in dump mode
    const uint32_t KB = (job.dims + 31) / 32;
    const uint32_t N = job.tq <= 32 ? 32u : 64u;
    // query tokens + their inverse norms (f64 norm, reference distances.rs:165)
    const size_t qbytes = (size_t)job.tq * job.dims * sizeof(float);
    VB_TRY(ctx.h_queries.reserve(qbytes + 64 * sizeof(float)));
    VB_TRY(ctx.queries.reserve(qbytes + 64 * sizeof(float)));
    float* hq = ctx.h_queries.as<float>();
    std::memcpy(hq, job.h_query, qbytes);
    float* hinv = hq + (size_t)job.tq * job.dims;
    for (uint32_t q = 0; q < 64u; ++q) {
        double s = 0.0;
        if (q < job.tq)
            for (uint32_t i = 0; i < job.dims; ++i) {
                const double x = job.h_query[(size_t)q * job.dims + i];
                s += x * x;
            }
        hinv[q] = s > 0.0 ? (float)(1.0 / std::sqrt(s)) : 0.0f;
    }
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, hq, qbytes + 64 * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));

    CUtensorMap tmap;
    VB_TRY(make_tmap_rows_sw128(job.d_tokens, job.ntok, job.stride, kTcTile, &tmap));

    // Pushes between two checkpoints: every document that ENDS inside the window's tiles.
    const uint32_t min_td = std::max<uint32_t>(1u, std::min<uint32_t>(job.min_td ? job.min_td : 1u, (uint32_t)kTcTile));
    const uint32_t ends_per_tile = (uint32_t)kTcTile / min_td + 2u;
    uint32_t cap = 256;
    while (cap < 2 * k || cap < k + 64 || cap < k + 2 * ends_per_tile || (job.has_empty && cap < k + kTcrEmptyRound)) cap <<= 1;
    uint32_t ckpt = 2;
    while (ckpt < 16 && k + 2 * ckpt * ends_per_tile <= cap) ckpt *= 2;
    const uint32_t slack = ckpt * ends_per_tile;
    const size_t fixed = 2 * (size_t)KB * N * 128 + (size_t)cap * 16 + 1024 + (N == 32 ? (size_t)kTcEpiWarps * 32 * 33 * 4 : 0);
    uint32_t stages = kTcStages;
    while (stages > 2 && (size_t)stages * kTcChunkBytes + fixed > 208 * 1024) --stages;
    const size_t smem = (size_t)stages * kTcChunkBytes + fixed;
    if (smem > 208 * 1024) return Status::Cuda("ragged maxsim tensor-core kernel: shared memory budget exceeded");
    auto kernel = N == 32 ? maxsim_tcr_kernel<32> : maxsim_tcr_kernel<64>;
    VB_TRY(ensure_dynamic_smem_for(kernel, 208 * 1024));
    int dev = 0, sms = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const uint32_t tiles = (uint32_t)((job.ntok + kTcTile - 1) / kTcTile);
    const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(tiles, (uint32_t)sms));

    VB_TRY(ctx.arm_ctrl(1));
    VB_TRY(ctx.cand_keys.reserve((size_t)grid * k * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve((size_t)grid * k * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve((size_t)grid * sizeof(uint32_t)));
    VB_TRY(ctx.out_keys.reserve((size_t)k * sizeof(u64)));
    VB_TRY(ctx.result.reserve((size_t)k * sizeof(u64) + 8));
    if (dump) VB_TRY(maxsim_prepare_dump(ctx, job.ndocs));

    MaxSimTcrParams p{};
    p.dump_keys = dump ? ctx.dump_keys.as<u64>() : nullptr;
    p.dump_pays = dump ? ctx.dump_pays.as<u64>() : nullptr;
    p.ndocs = (uint32_t)job.ndocs;
    p.ntok = (uint32_t)job.ntok;
    p.dims = job.dims;
    p.tq = job.tq;
    p.metric = job.metric;
    p.doc_off = job.d_doc_off;
    p.doc_rank = job.d_doc_rank;
    p.tok_doc = job.d_tok_doc;
    p.inv_dnorm = job.d_inv_dnorm;
    p.query = ctx.queries.as<float>();
    p.inv_qnorm = ctx.queries.as<float>() + (size_t)job.tq * job.dims;
    p.cap = cap;
    p.stages = stages;
    p.ckpt_tiles = ckpt;
    p.slack = slack;
    p.has_empty = job.has_empty ? 1u : 0u;
    p.no_scan = std::getenv("VB_MAXSIM_NO_SCAN") ? 1u : 0u;
    p.err = ctx.err_row();
    p.ws.k = k;
    p.ws.cand_keys = ctx.cand_keys.as<u64>();
    p.ws.cand_pays = ctx.cand_pays.as<u64>();
    p.ws.cand_counts = ctx.cand_counts.as<uint32_t>();
    p.ws.done = ctx.done();
    p.ws.g_thresh = ctx.g_thresh();
    p.ws.out_keys = ctx.out_keys.as<u64>();
    p.ws.out_pays = ctx.result.as<u64>();
    p.ws.out_counts = reinterpret_cast<uint32_t*>(ctx.result.as<u64>() + k);
    p.ws.err_row = ctx.err_row();
    p.ws.out_err = p.ws.out_counts + 1;
    kernel<<<grid, kTcThreads, smem, ctx.stream>>>(tmap, p);
    if (dump) return maxsim_collect_dump(ctx, job.ndocs, k_out, cudaGetLastError(), out);
    return maxsim_collect_result(ctx, job, p.ws, k, cudaGetLastError(), out);
}

}  // namespace vb
