// maxsim_tc.cu — K5 on the 5th-gen tensor cores: ColBERT MaxSim for the inner-product
// family (inner product, negative inner product, renormalising cosine) over uniform-length
// documents, as a 3xTF32 tcgen05 contraction with the max-reduce + sum + top-k fused into
// the epilogue (reference multi_vector.rs:65-132).
//
// Shape of the work (C5: 1M docs x 128 tokens x 128 dims, 32 query tokens): the token matrix
// is streamed once (tokens * D * 4 bytes, HBM-bound at 16 flop/B), every 128-token tile is a
// [128 x D] . [D x 32] product. Per SM, one persistent CTA of 18 warps:
//   warp 16     producer: TMA 2D tiled loads (128-byte swizzle) of token chunks into a 4-stage ring
//   warps 8-15  split: each thread owns one token row (= one TMEM lane), reads it from the
//               swizzled tile (conflict-free 128-bit LDS), splits x = hi + lo and writes both
//               halves into TMEM with tcgen05.st (A operand from TMEM: no second smem round trip)
//   warp 17     MMA issuer: per tile 3 * D/8 tcgen05.mma kind::tf32 (A in TMEM, B = query
//               tokens resident in shared memory in the UMMA K-major SWIZZLE_128B layout),
//               fp32 accumulators in TMEM (2 buffers x 4 independent partial accumulators)
//   warps 0-7   epilogue, two groups of 4 warps that take alternate tiles (group g owns accumulator
//               buffer g): tcgen05.ld the [128 x 32] tile, similarity transform, per-query max over
//               each document's tokens (31-shuffle transpose butterfly + shared memory across the
//               group's warps), f32 sum in query order, collector push (topk.cuh). One tile's
//               epilogue is a ~2100-cycle dependent chain (tcgen05.ld, 5 shuffle levels, a barrier,
//               the 32-step ordered sum); a single group capped the kernel at one tile per chain.
#include "maxsim.h"

#include <cstdlib>
#include "scan_driver.h"
#include "maxsim_tc.cuh"
#include "topk.cuh"

namespace vb {

struct MaxSimTcParams {
    uint32_t ndocs, td, dims, tq;
    int metric;                   // kInnerProduct, kNegativeInnerProduct or kCosineTrue
    const uint32_t* doc_rank;     // [ndocs] or null
    const float* inv_dnorm;       // [ndocs * td] 1/|token| (0 for zero tokens), cosine only
    const float* query;           // [tq, dims]
    const float* inv_qnorm;       // [tq], cosine only
    uint32_t cap;
    uint32_t* err;
    uint32_t debug;               // timing experiments only (VB_MAXSIM_DEBUG): 1 skip MMAs, 2 skip the split work, 4 skip the TMA loads, 8 skip the epilogue math
    TopkWorkspace ws;
};

__global__ void __launch_bounds__(kTcThreads, 1)
maxsim_tc_kernel(const __grid_constant__ CUtensorMap tmap, const MaxSimTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages];
    __shared__ __align__(8) uint64_t a_ready[2], a_free[2], d_full[kTcAccBufs], d_free[kTcAccBufs];
    __shared__ uint32_t tmem_slot;
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;
    __shared__ __align__(16) float s_part[kTcAccBufs][4][kTcN];
    __shared__ float s_invq[kTcN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t KB = p.dims / 32;                        // 128-byte K blocks per row (1..4)
    const uint32_t chunks_per_tile = (KB + 1) / 2;          // a chunk = up to 2 K blocks
    unsigned char* ring = smem;
    unsigned char* b_hi = ring + (size_t)kTcStages * kTcChunkBytes;   // KB x [32 rows x 128 B]
    unsigned char* b_lo = b_hi + (size_t)KB * 4096;
    unsigned char* col_mem = b_lo + (size_t)KB * 4096;

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
    if (tid < kTcN) s_invq[tid] = (p.inv_qnorm && (uint32_t)tid < p.tq) ? p.inv_qnorm[tid] : 0.0f;
    // B operand: the query tokens, split hi/lo, UMMA K-major SWIZZLE_128B layout; rows >= tq are zero.
    for (uint32_t idx = tid; idx < (uint32_t)kTcN * p.dims; idx += kTcThreads) {
        const uint32_t n = idx / p.dims, k = idx % p.dims;
        const float x = n < p.tq ? p.query[(size_t)n * p.dims + k] : 0.0f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const uint32_t off = (k / 32u) * 4096u + tc::sw128_offset(n, k % 32u);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = x - hi;
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;

    const uint32_t ntok = p.ndocs * p.td;
    const uint32_t num_tiles = (ntok + kTcTile - 1) / kTcTile;
    const uint32_t docs_per_tile = kTcTile / p.td;          // td in {32, 64, 128}

    if (warp == kTcProducerWarp) {
        // ===== producer: one chunk (<= 2 K blocks of the tile) per ring stage =====
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap);
            uint32_t cc = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (uint32_t j = 0; j < chunks_per_tile; ++j, ++cc) {
                    const uint32_t s = cc % kTcStages, ph = (cc / kTcStages) & 1u;
                    const uint32_t blocks = min(2u, KB - 2u * j);
                    tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                    tc::mbar_arrive_expect_tx(&full_bar[s], (p.debug & 4u) ? 0u : blocks * 16384u);
                    for (uint32_t h = 0; h < blocks && !(p.debug & 4u); ++h)
                        tc::tma_load_2d(ring + (size_t)s * kTcChunkBytes + (size_t)h * 16384, &tmap, (2u * j + h) * 32u,
                                        tile * kTcTile, &full_bar[s]);
                }
            }
        }
    } else if (warp == kTcMmaWarp) {
        // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues =====
        const uint32_t idesc = tc::umma_idesc_tf32(kTcTile, kTcN);
        const uint64_t bh0 = tc::umma_smem_desc_sw128(tc::smem_addr(b_hi));
        const uint64_t bl0 = tc::umma_smem_desc_sw128(tc::smem_addr(b_lo));
        uint32_t cc = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it % kTcAccBufs;
            const uint32_t d_tmem = tbase + kTcAccCol + b * (kTcChains * kTcN);
            for (uint32_t j = 0; j < chunks_per_tile; ++j, ++cc) {
                const uint32_t u = cc & 1u;
                const uint32_t blocks = min(2u, KB - 2u * j);
                tc::mbar_wait(&a_ready[u], (cc >> 1) & 1u);
                if (j == 0) tc::mbar_wait(&d_free[b], ((it / kTcAccBufs) & 1u) ^ 1u);
                tc::fence_after_sync();
                // The 4 k-steps of a block feed 4 different partial accumulators, so back-to-back
                // MMAs never wait on each other's result; the epilogue adds the partials.
                // Descriptor start-address field counts 16-byte units: K block = 256, k-step = 2.
                const uint32_t a0 = tbase + u * 128u;
                const uint64_t kbo = (uint64_t)(2u * j) * 256u;
                if (tc::elect_one()) {
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
                        if (h < blocks && !(p.debug & 1u)) {
                            const uint32_t first = (j | h) == 0u ? 0u : 1u;
#pragma unroll
                            for (uint32_t term = 0; term < 3; ++term) {
#pragma unroll
                                for (uint32_t ks = 0; ks < 4; ++ks) {
                                    const uint64_t bdesc = (term == 1 ? bl0 : bh0) + kbo + (uint64_t)(h * 256u + ks * 2u);
                                    const uint32_t a_addr = a0 + h * 32u + ks * 8u + (term == 2 ? 64u : 0u);
                                    tc::umma_tf32_ts(d_tmem + ks * kTcN, a_addr, bdesc, idesc, term == 0 ? first : 1u);
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
        // ===== split warps: smem chunk -> (hi, lo) -> TMEM. Two warps per TMEM lane quarter, =====
        // ===== one K block of the chunk each.                                               =====
        const uint32_t sw = warp - kTcEpiWarps;
        const uint32_t quarter = warp & 3u, h = sw >> 2;      // warps 4..7 -> block 0, 8..11 -> block 1
        const uint32_t row = quarter * 32u + lane;            // token row within the tile == TMEM lane
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        uint32_t cc = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (uint32_t j = 0; j < chunks_per_tile; ++j, ++cc) {
                const uint32_t s = cc % kTcStages, ph = (cc / kTcStages) & 1u, u = cc & 1u;
                const uint32_t blocks = min(2u, KB - 2u * j);
                tc::mbar_wait(&full_bar[s], ph);
                tc::mbar_wait(&a_free[u], ((cc >> 1) & 1u) ^ 1u);   // MMAs that read this A buffer are done
                tc::fence_after_sync();
                if (h < blocks && !(p.debug & 2u)) {
                    const unsigned char* blk = ring + (size_t)s * kTcChunkBytes + (size_t)h * 16384 + row * 128u;
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
                    tc::mbar_arrive(&empty_bar[s]);   // this warp's share of the stage is consumed
                }
            }
        }
    } else {
        // ===== epilogue warps 0-7: group (warp >> 2) handles tiles it = group, group + 2, ... =====
        const uint32_t quarter = warp & 3u, group = warp >> 2;   // TMEM lanes [32 * quarter, +32)
        const uint32_t lane_addr = tbase + ((quarter * 32u) << 16);
        const uint32_t warps_per_doc = p.td / 32u;               // 1, 2 or 4
        const uint32_t my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
        const uint32_t b = group;                                // accumulator buffer of this group (kTcAccBufs == 2)
        u64 g_prefetch = kKeyMax;
        for (uint32_t it = group; it < my_tiles; it += 2u) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            const uint32_t token = tile * kTcTile + quarter * 32u + lane;
            float inv_dn = 1.0f;
            if (p.metric == kCosineTrue) inv_dn = token < ntok ? __ldg(p.inv_dnorm + token) : 0.0f;
            tc::mbar_wait(&d_full[b], (it / kTcAccBufs) & 1u);
            tc::fence_after_sync();
            float v[32];
            {
                const uint32_t acc = lane_addr + kTcAccCol + b * (kTcChains * kTcN);
#pragma unroll
                for (int hcol = 0; hcol < 2; ++hcol) {
                    uint32_t r0[16], r1[16], r2[16], r3[16];
                    tc::tmem_ld16(acc + hcol * 16, r0);
                    tc::tmem_ld16(acc + kTcN + hcol * 16, r1);
                    tc::tmem_ld16(acc + 2 * kTcN + hcol * 16, r2);
                    tc::tmem_ld16(acc + 3 * kTcN + hcol * 16, r3);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        v[hcol * 16 + q] = (__uint_as_float(r0[q]) + __uint_as_float(r1[q])) +
                                           (__uint_as_float(r2[q]) + __uint_as_float(r3[q]));
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d_free[b]);
            // similarity_value: inner product -> dot; negative inner product -> -(-dot) = dot; cosine ->
            // dot / (|q| |d|) clamped (distances.rs:170-172). Scaling by 1/|q| >= 0 and clamping are
            // monotonic, so they are applied once per query after the max instead of per pair.
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                float sim = v[q];
                if (p.metric == kCosineTrue) sim *= inv_dn;
                v[q] = token < ntok ? sim : -INFINITY;
            }
            float wmax = warp_transpose_max(v, lane);                      // lane q: max over this warp's tokens
            if (p.metric == kCosineTrue) wmax = fminf(1.0f, fmaxf(-1.0f, wmax * s_invq[lane]));
            s_part[b][quarter][lane] = wmax;
            asm volatile("bar.sync %0, 128;" ::"r"(3u + group) : "memory");   // the 4 warps of this group
            // document leaders: the first warp of each document combines its warps; one lane then adds
            // the per-query maxima in query order (multi_vector.rs:81-84) straight from shared memory
            // (32 pipelined loads + 32 dependent adds instead of a 32-deep shuffle chain).
            if ((quarter % warps_per_doc) == 0) {
                float m = wmax;
                for (uint32_t w = 1; w < warps_per_doc; ++w) m = fmaxf(m, s_part[b][quarter + w][lane]);
                s_part[b][quarter][lane] = m;
                __syncwarp();
                const uint32_t doc = tile * docs_per_tile + quarter / warps_per_doc;
                if (lane == 0 && doc < p.ndocs) {
                    const uint32_t rank = p.doc_rank ? __ldg(p.doc_rank + doc) : doc;
                    if (rank != 0xFFFFFFFFu) {
                        const float4* mv = reinterpret_cast<const float4*>(&s_part[b][quarter][0]);
                        float mq[kTcN];
#pragma unroll
                        for (int i = 0; i < kTcN / 4; ++i) {
                            const float4 t4 = mv[i];
                            mq[4 * i] = t4.x; mq[4 * i + 1] = t4.y; mq[4 * i + 2] = t4.z; mq[4 * i + 3] = t4.w;
                        }
                        float total = 0.0f;
#pragma unroll
                        for (int q = 0; q < kTcN; ++q)
                            if ((uint32_t)q < p.tq) total += mq[q];
                        // a non-finite running sum can never become finite again, so one check suffices
                        if (!isfinite(total)) { atomicMin(p.err, (doc << 1) | 1u); total = 0.0f; }
                        const u64 key = ((u64)(~order_key(total)) << 32) | rank;
                        if (key < col.threshold()) col.push(key, ((u64)__float_as_uint(total) << 32) | doc);
                    }
                }
            }
            // Both groups meet (all 8 warps) each time the CTA has finished another 16 tiles; a window only
            // counts when it lies inside this CTA's tiles, so both groups pass the same number of checkpoints.
            if ((it + 2u) / 16u != it / 16u && (it / 16u + 1u) * 16u <= my_tiles)
                collector_checkpoint(col, p.ws, 0, 16 * 4, g_prefetch);
        }
        collector_publish_and_merge(col, p.ws, 0, &s_last);
    }
    // teardown: every role is done with TMEM before it is released
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kTcMmaWarp) tc::tmem_dealloc(tbase, 512);
}

bool maxsim_tc_eligible(const MaxSimJob& job, uint32_t uniform_td) {
    if (std::getenv("VB_MAXSIM_NO_TC") || std::getenv("VB_MAXSIM_NO_TCU")) return false;
    if (job.metric != kInnerProduct && job.metric != kNegativeInnerProduct && job.metric != kCosineTrue) return false;
    if (job.dims % 32 != 0 || job.dims > 128 || job.stride != job.dims) return false;
    if (job.tq == 0 || job.tq > (uint32_t)kTcN) return false;
    if (uniform_td != 32 && uniform_td != 64 && uniform_td != 128) return false;
    if (std::min<size_t>(job.k, job.ndocs) > (size_t)kMaxFusedK) return false;
    return true;
}

Status maxsim_tc_top_k(SearchCtx& ctx, const MaxSimJob& job, uint32_t td, const float* d_inv_dnorm,
                       MaxSimResult* out) {
    out->rows.clear();
    out->scores.clear();
    out->err = kNoError;
    const uint32_t k = (uint32_t)std::min<size_t>(job.k, job.ndocs);
    const uint32_t KB = job.dims / 32;
    // query tokens + their inverse norms (f64 norm, reference distances.rs:165)
    const size_t qbytes = (size_t)job.tq * job.dims * sizeof(float);
    VB_TRY(ctx.h_queries.reserve(qbytes + kTcN * sizeof(float)));
    VB_TRY(ctx.queries.reserve(qbytes + kTcN * sizeof(float)));
    float* hq = ctx.h_queries.as<float>();
    std::memcpy(hq, job.h_query, qbytes);
    float* hinv = hq + (size_t)job.tq * job.dims;
    for (uint32_t q = 0; q < (uint32_t)kTcN; ++q) {
        double s = 0.0;
        if (q < job.tq)
            for (uint32_t i = 0; i < job.dims; ++i) {
                const double x = job.h_query[(size_t)q * job.dims + i];
                s += x * x;
            }
        hinv[q] = s > 0.0 ? (float)(1.0 / std::sqrt(s)) : 0.0f;
    }
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, hq, qbytes + kTcN * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));

    CUtensorMap tmap;
    VB_TRY(make_tmap_rows_sw128(job.d_tokens, (uint64_t)job.ndocs * td, job.stride, kTcTile, &tmap));

    uint32_t cap = 256;
    while (cap < 2 * k || cap < k + 64) cap <<= 1;
    const size_t smem = (size_t)kTcStages * kTcChunkBytes + 2 * (size_t)KB * 4096 + (size_t)cap * 16 + 1024;
    VB_TRY(ensure_dynamic_smem_for(maxsim_tc_kernel, 220 * 1024));
    if (smem > 220 * 1024) return Status::Cuda("maxsim tensor-core kernel: shared memory budget exceeded");
    int dev = 0, sms = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const uint32_t tiles = (uint32_t)(((uint64_t)job.ndocs * td + kTcTile - 1) / kTcTile);
    const uint32_t grid = std::min<uint32_t>(tiles, (uint32_t)sms);

    VB_TRY(ctx.arm_ctrl(1));
    VB_TRY(ctx.cand_keys.reserve((size_t)grid * k * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve((size_t)grid * k * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve((size_t)grid * sizeof(uint32_t)));
    VB_TRY(ctx.out_keys.reserve((size_t)k * sizeof(u64)));
    VB_TRY(ctx.result.reserve((size_t)k * sizeof(u64) + 8));

    MaxSimTcParams p{};
    p.ndocs = (uint32_t)job.ndocs;
    p.td = td;
    p.dims = job.dims;
    p.tq = job.tq;
    p.metric = job.metric;
    p.doc_rank = job.d_doc_rank;
    p.inv_dnorm = d_inv_dnorm;
    p.query = ctx.queries.as<float>();
    p.inv_qnorm = ctx.queries.as<float>() + (size_t)job.tq * job.dims;
    p.cap = cap;
    { const char* dbg = std::getenv("VB_MAXSIM_DEBUG"); p.debug = dbg ? (uint32_t)std::atoi(dbg) : 0u; }
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
    maxsim_tc_kernel<<<grid, kTcThreads, smem, ctx.stream>>>(tmap, p);
    return maxsim_collect_result(ctx, job, p.ws, k, cudaGetLastError(), out);
}

}  // namespace vb
