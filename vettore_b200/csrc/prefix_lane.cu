// prefix_lane.cu — kernel C of the K1/K4 family: true-cosine prefix scoring of EVERY row with ONE ROW PER LANE.
//
// The path: search::vector_top_k (reference search.rs:38-73) with metric cosine over the first `dims` <= 128
// columns — funnel stage 1 and the Matryoshka prefix scan (reference distances.rs:160-177: f64 dot, f64 row norm,
// dot / (|q| |row|), clamped, cast to f32).
//
// Why a third kernel: with a short prefix the warp-per-row kernels (flat_scan.cuh) spend ~85 warp instructions per
// row on what is NOT the dot product — two f64 butterflies, the f64 sqrt / divide of the tail, ring bookkeeping per
// 4-row tile — and run at 0.35-0.4 of the HBM rate with the issue slots half busy (profiles/
// r2_funnel_stage1_ncu_summary.txt). Here a lane owns a whole row: no cross-lane reduction exists, the tail is
// evaluated for 32 rows at once, and a tile costs ~15 warp instructions per row.
//
// Shape: persistent CTA per SM; a producer lane streams tiles of 32 rows through a shared-memory ring as
// ceil(dims / 32) TMA boxes of [32 rows x 32 columns] each, 128-byte swizzled, so lane L reads 16-byte chunk c of ITS
// row at chunk position c ^ (L & 7): the eight lanes of a 128-bit shared-memory wavefront hit eight different bank
// groups (an unswizzled [32][128 B] tile would put them all on one). The tensor map is `dims` columns wide: columns
// beyond the prefix are out of bounds and arrive as zeros, whatever the row stride holds there — the same kernel
// reads the dense prefix mirror or the main matrix. The query sits in shared memory as f64 (broadcast loads). Tile i
// of a CTA belongs to consumer warp i % W; the ring holds `depth` tiles per warp and every BOX has its own full /
// empty barrier pair, so a warp hands box j back as soon as it has read it (the next tile's box j loads while boxes
// j+1.. are scored), and a slot is only ever waited on by one warp — its barrier parity cannot alias. Scores go
// straight into the CTA's collector (topk.cuh).
#include "flat_scan.cuh"
#include "flat_scan.h"

namespace vb {

constexpr int kLaneMaxWarps = 12;       // consumer warps (runtime choice, blockDim.x / 32 - 1)
constexpr int kLaneMaxStages = 48;
constexpr int kLaneSyncRounds = 2;      // default rounds (one tile per consumer warp) between collector checkpoints
constexpr uint32_t kLaneBoxBytes = 32 * 128;

__global__ void __launch_bounds__(kLaneMaxWarps * 32 + 32, 1)
prefix_lane_kernel(const ScanParams p, const StreamGeom geom, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char lsmem[];
    __shared__ __align__(8) uint64_t full_bar[kLaneMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kLaneMaxStages];
    __shared__ __align__(16) double2 s_qa[32];   // (x, y) of query chunk c as f64
    __shared__ __align__(16) double2 s_qb[32];   // (z, w)
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t W = (blockDim.x >> 5) - 1u;
    const uint32_t qi = blockIdx.y;
    const uint32_t num_tiles = (p.n + 31u) / 32u;
    const uint32_t iters = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
    const uint32_t nb = (p.dims + 31u) / 32u;            // column boxes per tile
    unsigned char* ring = lsmem;
    unsigned char* col_mem = lsmem + (size_t)W * geom.stages * geom.tile_bytes;

    Collector col;
    col.init(col_mem, &s_thresh, &s_count, p.cap, p.ws.k, W * 32u, 1);
    collector_attach_pivots(col, p.ws);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < W * geom.stages * nb; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        const float* q = p.queries + (size_t)qi * p.q_stride;
        const uint32_t e = threadIdx.x * 4u;
        auto at = [&](uint32_t i) { return i < p.dims ? (double)q[i] : 0.0; };
        s_qa[threadIdx.x] = make_double2(at(e), at(e + 1u));
        s_qb[threadIdx.x] = make_double2(at(e + 2u), at(e + 3u));
    }
    __syncthreads();

    // geom.stages = ring depth in tiles per consumer warp; box slot of (tile i, box j) = ((i % (W * depth)) * nb + j)
    const uint32_t ring_tiles = W * geom.stages;
    if (warp == W) {
        // ===== producer: lane j streams box j of every tile of this CTA (one issuing lane per column box: a single
        // lane's wait / expect / issue round per 4 KB box was close to the rate the ring needs) =====
        if (lane < nb) {
            for (uint32_t i = 0; i < iters; ++i) {
                const uint32_t t = i % ring_tiles, ph = (i / ring_tiles) & 1u;
                const uint32_t row0 = (blockIdx.x + i * gridDim.x) * 32u;
                const uint32_t s = t * nb + lane;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], kLaneBoxBytes);
                tc::tma_load_2d(ring + (size_t)s * kLaneBoxBytes, &tmap, 32u * lane, row0, &full_bar[s]);
            }
        }
        return;
    }

    // ===== consumers: warp w takes tiles w, w + W, ... of this CTA; lane L owns row L of the tile =====
    const double q_norm = p.q_norms[qi];
    const uint32_t rounds = (iters + W - 1u) / W;
    const uint32_t sync_rounds = geom.tail_rem;          // kernel C: rounds between collector checkpoints (plan_lane)
    const uint32_t my_off = lane * 128u, my_x = lane & 7u;
    u64 g_prefetch = kKeyMax;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i = r * W + warp;
        if (i < iters) {
            const uint32_t t = i % ring_tiles, ph = (i / ring_tiles) & 1u;
            double d0a = 0.0, d0b = 0.0, d1a = 0.0, d1b = 0.0;
            for (uint32_t j = 0; j < nb; ++j) {
                const uint32_t s = t * nb + j;
                if (lane == 0) mbar_wait(&full_bar[s], ph);
                __syncwarp();
                const unsigned char* box = ring + (size_t)s * kLaneBoxBytes + my_off;
                if (!(p.debug & 2u))
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const float4 b = *reinterpret_cast<const float4*>(box + ((c ^ my_x) << 4));
                    const double2 qa = s_qa[j * 8u + c], qb = s_qb[j * 8u + c];
                    const double bx = (double)b.x, by = (double)b.y, bz = (double)b.z, bw = (double)b.w;
                    d0a = fma(qa.x, bx, d0a); d0b = fma(qa.y, by, d0b);
                    d1a = fma(bx, bx, d1a);   d1b = fma(by, by, d1b);
                    d0a = fma(qb.x, bz, d0a); d0b = fma(qb.y, bw, d0b);
                    d1a = fma(bz, bz, d1a);   d1b = fma(bw, bw, d1b);
                }
                __syncwarp();                              // every lane's reads of the box are done
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
            const uint32_t row = (blockIdx.x + i * gridDim.x) * 32u + lane;
            if (row < p.n && !(p.debug & 1u)) {
                const double v[2] = {d0a + d0b, d1a + d1b};
                bool bad, fatal;
                const float raw = PartialTraits<kCosineTrue>::finalize(v, q_norm, bad, fatal);
                if (fatal) atomicMin(p.err_row + qi, row);
                emit_row<kCosineTrue>(p, col, qi, false, col.threshold(), raw, row, row);
            }
        }
        if (!(p.debug & 4u) && (r + 1u) % sync_rounds == 0u)
            collector_checkpoint(col, p.ws, qi, sync_rounds * W * 32u, g_prefetch);
    }
    // every tile was consumed, so the ring is idle: the last CTA merges in ring + collector memory
    const uint32_t total_smem = W * geom.stages * geom.tile_bytes + p.cap * 16u;
    uint32_t big_cap = p.cap;
    while ((size_t)big_cap * 2u * 16u <= total_smem) big_cap *= 2u;
    collector_publish_and_merge(col, p.ws, qi, &s_last, lsmem, big_cap);
}

StreamKernel prefix_lane_kernel_entry() { return prefix_lane_kernel; }
int prefix_lane_max_warps() { return kLaneMaxWarps; }
int prefix_lane_max_stages() { return kLaneMaxStages; }
int prefix_lane_sync_rounds() { return kLaneSyncRounds; }

}  // namespace vb
