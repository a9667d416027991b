// hamming.cu — K3 Hamming scan and K6 sign packing (see hamming.h).
//
// K3 is a pure HBM stream: N * ceil(D/64) * 8 algorithmic bytes per query, one XOR+POPC
// per 8 bytes. Rows are split into 16-byte chunks (8-byte when the word count is odd); a
// group of g = pow2 >= chunks-per-row lanes owns one row, so a warp-level load covers
// 32/g whole consecutive rows (fully coalesced), U of them in flight per lane. Group
// popcounts are combined with log2(g) shuffles and the group leader feeds the collector.
#include "hamming.h"

#include <cub/device/device_radix_sort.cuh>

#include "select.h"
#include "topk.cuh"

namespace vb {

constexpr int kHamThreads = 256;
constexpr int kHamWarps = kHamThreads / 32;
constexpr int kHamUnroll = 8;
constexpr uint32_t kHamSlackRows = 1024;  // rows scanned per CTA between collector checks

struct HammingParams {
    const u64* codes;          // [n, nw]
    uint32_t n, nw, dims;
    const uint32_t* id_rank;   // optional
    const u64* queries;        // [nq, nw]
    uint32_t g;                // lanes per row (power of two)
    uint32_t sync_every;       // steps between collector checks (power of two)
    uint32_t cap;
    TopkWorkspace ws;
    u64* dump_keys;            // dump mode: [n] keys / pays, no collector
    u64* dump_pays;
};

template <bool WIDE>  // WIDE: 16-byte chunks (nw even); else 8-byte chunks
__global__ void __launch_bounds__(kHamThreads, 4) hamming_scan_kernel(const HammingParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    const bool dump = p.dump_keys != nullptr;
    Collector col;
    col.init(smem, &s_thresh, &s_count, p.cap, p.ws.k);
    __syncthreads();

    constexpr uint32_t WPC = WIDE ? 2 : 1;              // words per chunk
    const uint32_t cpr = p.nw / WPC;                    // chunks per row
    const uint32_t g = p.g;
    const uint32_t rpw = 32u / g;                       // rows per warp-level load
    const uint32_t sub = lane / g, c0 = lane % g;       // row within the load, first chunk
    const uint32_t rem = p.dims & 63u;
    const u64 last_mask = rem ? ((1ull << rem) - 1ull) : ~0ull;   // distances.rs:472-481
    const u64* q = p.queries + (size_t)qi * p.nw;

    // query chunk of this lane (first pass over the row); later passes re-read through L1
    u64 qa = 0, qb = 0;
    if (c0 < cpr) {
        qa = q[c0 * WPC];
        if (WIDE) qb = q[c0 * WPC + 1];
    }

    const uint32_t rows_per_step = kHamWarps * rpw * kHamUnroll;
    const uint32_t steps = (p.n + rows_per_step - 1u) / rows_per_step;
    u64 g_prefetch = kKeyMax;
    uint32_t it = 0;
    for (uint32_t step = blockIdx.x; step < steps; step += gridDim.x, ++it) {
        const uint32_t base = step * rows_per_step + warp * rpw * kHamUnroll + sub;
        uint32_t dist[kHamUnroll];
        // first chunk pass: all U loads issued back to back
        u64 wa[kHamUnroll], wb[kHamUnroll];
#pragma unroll
        for (int u = 0; u < kHamUnroll; ++u) {
            const uint32_t row = base + u * rpw;
            wa[u] = qa;
            wb[u] = qb;
            if (row < p.n && c0 < cpr) {
                const u64* src = p.codes + (size_t)row * p.nw + c0 * WPC;
                if (WIDE) {
                    uint4 v = ldg_stream(reinterpret_cast<const uint4*>(src));
                    wa[u] = ((u64)v.y << 32) | v.x;
                    wb[u] = ((u64)v.w << 32) | v.z;
                } else {
                    wa[u] = __ldg(src);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kHamUnroll; ++u) {
            u64 xa = wa[u] ^ qa, xb = wb[u] ^ qb;
            if (c0 * WPC == p.nw - 1u) xa &= last_mask;
            if (WIDE && c0 * WPC + 1u == p.nw - 1u) xb &= last_mask;
            dist[u] = __popcll(xa) + (WIDE ? __popcll(xb) : 0);
        }
        // rows longer than g chunks (only when cpr > 32): remaining passes
        for (uint32_t c = c0 + g; c < cpr; c += g) {
#pragma unroll
            for (int u = 0; u < kHamUnroll; ++u) {
                const uint32_t row = base + u * rpw;
                if (row >= p.n) continue;
                const u64* src = p.codes + (size_t)row * p.nw + c * WPC;
                u64 xa = __ldg(src) ^ __ldg(q + c * WPC), xb = 0;
                if (WIDE) xb = __ldg(src + 1) ^ __ldg(q + c * WPC + 1);
                if (c * WPC == p.nw - 1u) xa &= last_mask;
                if (WIDE && c * WPC + 1u == p.nw - 1u) xb &= last_mask;
                dist[u] += __popcll(xa) + (WIDE ? __popcll(xb) : 0);
            }
        }
#pragma unroll
        for (int u = 0; u < kHamUnroll; ++u)
            for (uint32_t o = g >> 1; o > 0; o >>= 1) dist[u] += __shfl_xor_sync(0xffffffffu, dist[u], o);

        if (c0 == 0) {
            const u64 T = dump ? kKeyMax : col.threshold();
#pragma unroll
            for (int u = 0; u < kHamUnroll; ++u) {
                const uint32_t row = base + u * rpw;
                if (row >= p.n) continue;
                const float raw = (float)dist[u];                       // distances.rs:436
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
        if (!dump && (it & (p.sync_every - 1)) == p.sync_every - 1)
            collector_checkpoint(col, p.ws, qi, p.sync_every * rows_per_step, g_prefetch);
    }
    if (dump) return;
    collector_publish_and_merge(col, p.ws, qi, &s_last);
}

static uint32_t pow2_at_least(uint32_t v, uint32_t lo) {
    uint32_t p = lo;
    while (p < v) p <<= 1;
    return p;
}

static Status launch_hamming(SearchCtx& ctx, HammingParams p, uint32_t nq, uint32_t k, bool dump,
                             cudaStream_t stream) {
    const bool wide = (p.nw % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.codes) & 15) == 0);
    const uint32_t cpr = wide ? p.nw / 2 : p.nw;
    p.g = std::min<uint32_t>(32, pow2_at_least(cpr, 1));
    const uint32_t rows_per_step = kHamWarps * (32 / p.g) * kHamUnroll;
    const uint32_t kk = dump ? 1 : k;
    p.sync_every = std::max<uint32_t>(1, kHamSlackRows / rows_per_step);  // both are powers of two
    p.cap = pow2_at_least(2 * kk + p.sync_every * rows_per_step, 256);
    const size_t smem = (size_t)p.cap * 16;
    auto kernel = wide ? hamming_scan_kernel<true> : hamming_scan_kernel<false>;
    if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kHamThreads, smem));
    if (per_sm < 1) return Status::Cuda("hamming kernel does not fit on an SM");
    cudaDeviceProp prop;
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    int sms = 0;
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    (void)prop;
    const uint32_t steps = (p.n + rows_per_step - 1) / rows_per_step;
    const uint32_t grid_x = std::min<uint32_t>(steps, (uint32_t)(sms * per_sm));
    if (!dump) {
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
        VB_TRY(ctx.dump_keys.reserve(n * sizeof(u64)));
        VB_TRY(ctx.dump_pays.reserve(n * sizeof(u64)));
        VB_TRY(ctx.dump_keys2.reserve(n * sizeof(u64)));
        VB_TRY(ctx.dump_pays2.reserve(n * sizeof(u64)));
        HammingParams p{};
        p.codes = d_codes;
        p.n = (uint32_t)n;
        p.nw = (uint32_t)nw;
        p.dims = (uint32_t)dims;
        p.id_rank = d_rank;
        p.queries = ctx.queries.as<u64>();
        p.dump_keys = ctx.dump_keys.as<u64>();
        p.dump_pays = ctx.dump_pays.as<u64>();
        VB_TRY(launch_hamming(ctx, p, 1, 1, true, ctx.stream));
        size_t tmp_bytes = 0;
        VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx.dump_keys.as<u64>(), ctx.dump_keys2.as<u64>(),
                                                ctx.dump_pays.as<u64>(), ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64,
                                                ctx.stream));
        VB_TRY(ctx.sort_tmp.reserve(tmp_bytes));
        VB_CUDA(cub::DeviceRadixSort::SortPairs(ctx.sort_tmp.p, tmp_bytes, ctx.dump_keys.as<u64>(),
                                                ctx.dump_keys2.as<u64>(), ctx.dump_pays.as<u64>(),
                                                ctx.dump_pays2.as<u64>(), (int64_t)n, 0, 64, ctx.stream));
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
