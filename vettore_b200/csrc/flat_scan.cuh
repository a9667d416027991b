// flat_scan.cuh — K1/K4: single-query exact scan + fused top-k over an HBM-resident
// row-major fp32 matrix. One kernel family serves
//   * FlatIndex::search            (reference flat.rs:96-124)      metric = index metric
//   * search::vector_top_k         (reference search.rs:38-73)     prefix dims, kCosineTrue
//   * exact rerank over a row list (reference collection.ex:821-851)
//
// Shape of the work: an N x D fp32 stream read exactly once per query (N*D*4 algorithmic
// bytes), 2 flops per element -> HBM-bound. Design:
//   - flat_stream_kernel (whole rows, the FlatIndex::search case): one persistent CTA per
//     SM; a producer warp streams contiguous row tiles HBM -> shared memory with TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx) through a multi-stage ring, so the bytes
//     in flight per SM are the ring size (~190 KB), independent of warp scheduling; eight
//     consumer warps read their rows back with conflict-free 128-bit LDS and score them
//     against the query held in registers;
//   - flat_scan_kernel (prefix scoring, row lists, dump mode): persistent grid, 8 warps per
//     CTA, each warp owns R whole rows per step, lane owns float4 slots {lane + 32 j}
//     (coalesced 512-byte segments, 128-bit read-only/no-L1-allocate loads), all R*NV loads
//     of a step issued before the first FMA;
//   - the score never goes back to HBM: rank key + id rank go straight into the CTA's
//     shared-memory collector (topk.cuh); the last CTA to finish merges the per-CTA
//     lists, so one launch per query returns the sorted top-k;
//   - non-finite f32 results are recomputed in f64 from the registers already loaded
//     (reference distances.rs:59-98); unrepresentable ones flag "metric overflow".
#pragma once
#include <type_traits>

#include "tc.cuh"
#include "topk.cuh"

namespace vb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kSyncEvery = 8;         // kernel A: steps between collector checks (power of two)
constexpr int kStreamSyncEvery = 16;  // kernel B: tiles between collector checks (power of two)

struct ScanParams {
    const float* rows;        // [*, row_stride] fp32, rows 16-byte aligned
    size_t row_stride;        // floats, multiple of 4
    const uint32_t* row_sel;  // optional: logical row -> device row
    const uint32_t* id_rank;  // optional: device row -> id tie-break rank (else the row)
    uint32_t n;               // logical rows
    uint32_t dims;            // elements scored per row (a prefix of the row)
    const float* queries;     // [nq, q_stride], zero beyond dims
    uint32_t q_stride;        // floats, multiple of 4
    const double* q_norms;    // [nq] f64 L2 norm of the query prefix (kCosineTrue only)
    uint32_t cap;             // collector capacity (entries, power of two)
    uint32_t debug;           // tuning experiments only (VB_SCAN_DEBUG): 1 no emit, 2 no scoring, 4 no checkpoint
    TopkWorkspace ws;         // per-query candidate lists / threshold / output (ws.k = results)
    uint32_t* err_row;        // [nq] smallest logical row with an unrecoverable overflow
    // dump mode (limit beyond the fused collector): every row's key/payload to HBM
    u64* dump_keys;           // [nq][n] or null
    u64* dump_pays;
};

// ---------------------------------------------------------------------------------------
// Per-row accumulation for one lane, by metric (reference distances.rs:197-347).
template <int M>
struct Scorer {
    float s0, s1, s2, s3;
    uint32_t c0, c1;
    double d0, d1;

    __device__ __forceinline__ void init() {
        s0 = s1 = s2 = s3 = 0.0f;
        c0 = c1 = 0u;
        d0 = d1 = 0.0;
    }

    __device__ __forceinline__ void accum(const float4& q, const float4& b) {
        if constexpr (M == kCosine || M == kInnerProduct || M == kNegativeInnerProduct) {
            s0 = fmaf(q.x, b.x, s0); s1 = fmaf(q.y, b.y, s1);
            s2 = fmaf(q.z, b.z, s2); s3 = fmaf(q.w, b.w, s3);
        } else if constexpr (M == kL2 || M == kL2Squared) {
            float dx = q.x - b.x, dy = q.y - b.y, dz = q.z - b.z, dw = q.w - b.w;
            s0 = fmaf(dx, dx, s0); s1 = fmaf(dy, dy, s1);
            s2 = fmaf(dz, dz, s2); s3 = fmaf(dw, dw, s3);
        } else if constexpr (M == kManhattan) {
            s0 += fabsf(q.x - b.x); s1 += fabsf(q.y - b.y);
            s2 += fabsf(q.z - b.z); s3 += fabsf(q.w - b.w);
        } else if constexpr (M == kChebyshev) {
            s0 = fmaxf(s0, fabsf(q.x - b.x)); s1 = fmaxf(s1, fabsf(q.y - b.y));
            s2 = fmaxf(s2, fabsf(q.z - b.z)); s3 = fmaxf(s3, fabsf(q.w - b.w));
        } else if constexpr (M == kHamming) {
            c0 += ((q.x != 0.0f) != (b.x != 0.0f)) + ((q.y != 0.0f) != (b.y != 0.0f)) +
                  ((q.z != 0.0f) != (b.z != 0.0f)) + ((q.w != 0.0f) != (b.w != 0.0f));
        } else if constexpr (M == kJaccard) {
            bool lx = q.x != 0.0f, ly = q.y != 0.0f, lz = q.z != 0.0f, lw = q.w != 0.0f;
            bool rx = b.x != 0.0f, ry = b.y != 0.0f, rz = b.z != 0.0f, rw = b.w != 0.0f;
            c0 += (lx && rx) + (ly && ry) + (lz && rz) + (lw && rw);
            c1 += (lx || rx) + (ly || ry) + (lz || rz) + (lw || rw);
        } else {  // kCosineTrue: f64 dot and f64 row norm (reference distances.rs:160-177)
            d0 = fma((double)q.x, (double)b.x, d0); d0 = fma((double)q.y, (double)b.y, d0);
            d0 = fma((double)q.z, (double)b.z, d0); d0 = fma((double)q.w, (double)b.w, d0);
            d1 = fma((double)b.x, (double)b.x, d1); d1 = fma((double)b.y, (double)b.y, d1);
            d1 = fma((double)b.z, (double)b.z, d1); d1 = fma((double)b.w, (double)b.w, d1);
        }
    }

    // Warp-wide: every lane returns the row's raw metric value. `bad` = the f32 result is
    // non-finite and must be recovered in f64; `fatal` = true-cosine overflow.
    __device__ __forceinline__ float finish(double q_norm, bool& bad, bool& fatal) {
        float v;
        bad = false;
        fatal = false;
        if constexpr (M == kCosine || M == kInnerProduct || M == kNegativeInnerProduct) {
            v = warp_sum((s0 + s1) + (s2 + s3));
            if constexpr (M == kNegativeInnerProduct) v = -v;
            bad = !isfinite(v);
        } else if constexpr (M == kL2 || M == kL2Squared) {
            v = warp_sum((s0 + s1) + (s2 + s3));
            bad = !isfinite(v);
            if constexpr (M == kL2) v = sqrtf(v);
        } else if constexpr (M == kManhattan) {
            v = warp_sum((s0 + s1) + (s2 + s3));
            bad = !isfinite(v);
        } else if constexpr (M == kChebyshev) {
            v = warp_max(fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
            bad = !isfinite(v);
        } else if constexpr (M == kHamming) {
            v = (float)warp_sum(c0);
        } else if constexpr (M == kJaccard) {
            uint32_t inter = warp_sum(c0), uni = warp_sum(c1);
            v = uni == 0u ? 0.0f : __fsub_rn(1.0f, __fdiv_rn((float)inter, (float)uni));
        } else {
            double dot = warp_sum(d0), nn = warp_sum(d1);
            double rn = sqrt(nn);
            if (q_norm == 0.0 || rn == 0.0) {
                v = 0.0f;
            } else {
                double s = dot / (q_norm * rn);
                if (!isfinite(s)) { fatal = true; s = 0.0; }
                s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
                v = (float)s;
            }
        }
        return v;
    }
};

// f64 recomputation of one row (reference recover_metric_overflow, distances.rs:70-98).
template <int M>
struct Recover {
    double d;
    __device__ __forceinline__ void init() { d = 0.0; }
    __device__ __forceinline__ void one(float a, float b) {
        if constexpr (M == kCosine || M == kInnerProduct || M == kNegativeInnerProduct) {
            d = fma((double)a, (double)b, d);
        } else if constexpr (M == kL2 || M == kL2Squared) {
            double t = (double)a - (double)b;
            d = fma(t, t, d);
        } else if constexpr (M == kManhattan) {
            d += fabs((double)a - (double)b);
        } else if constexpr (M == kChebyshev) {
            d = fmax(d, fabs((double)a - (double)b));
        }
    }
    __device__ __forceinline__ void accum(const float4& q, const float4& b) {
        one(q.x, b.x); one(q.y, b.y); one(q.z, b.z); one(q.w, b.w);
    }
    __device__ __forceinline__ float finish(bool& fatal) {
        double v;
        if constexpr (M == kChebyshev) {
            v = d;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        } else {
            v = warp_sum(d);
        }
        if constexpr (M == kL2) v = sqrt(v);
        if constexpr (M == kNegativeInnerProduct) v = -v;
        const double mx = 3.4028234663852886e38;
        fatal = !(isfinite(v) && v >= -mx && v <= mx);
        return fatal ? 0.0f : (float)v;
    }
};

template <int M>
constexpr bool kCanOverflow = (M != kHamming && M != kJaccard && M != kCosineTrue);

// Zeroes the components of the last float4 of a prefix that lie beyond `dims`.
__device__ __forceinline__ void mask_tail(float4& b, uint32_t rem) {
    if (rem < 2) b.y = 0.0f;
    if (rem < 3) b.z = 0.0f;
    if (rem < 4) b.w = 0.0f;
}

// ---------------------------------------------------------------------------------------
// Scores one row held in registers (b[NV]) against the query (q[NV]) and, when the f32
// result is non-finite, recomputes it in f64 from the same registers. Warp-wide.
template <int M, int NV>
__device__ __forceinline__ float score_row_regs(const float4 (&q)[NV], float4 (&b)[NV], double q_norm,
                                                bool& fatal) {
    Scorer<M> sc;
    sc.init();
#pragma unroll
    for (int j = 0; j < NV; ++j) sc.accum(q[j], b[j]);
    bool bad;
    float raw = sc.finish(q_norm, bad, fatal);
    if constexpr (kCanOverflow<M>) {
        if (bad) {  // warp-uniform cold path
            Recover<M> rc;
            rc.init();
#pragma unroll
            for (int j = 0; j < NV; ++j) rc.accum(q[j], b[j]);
            raw = rc.finish(fatal);
        }
    }
    return raw;
}

// Lane 0: turn a raw value into (key, payload) and hand it to the collector / dump arrays.
template <int M>
__device__ __forceinline__ void emit_row(const ScanParams& p, Collector& col, uint32_t qi, bool dump, u64 T,
                                         float raw, uint32_t row, uint32_t drow) {
    const uint32_t rk = order_key(rank_value(M, raw));
    if (rk > (uint32_t)(T >> 32)) return;
    const uint32_t idr = p.id_rank ? __ldg(p.id_rank + drow) : drow;
    const u64 key = ((u64)rk << 32) | idr;
    const u64 pay = ((u64)__float_as_uint(raw) << 32) | drow;
    if (dump) {
        p.dump_keys[(size_t)qi * p.n + row] = key;
        p.dump_pays[(size_t)qi * p.n + row] = pay;
    } else if (key < T) {
        col.push(key, pay);
    }
}

// Per-lane partial state of one row, reduced across the warp only once per group of rows.
template <int M>
struct PartialTraits {
    static constexpr bool kIsDouble = (M == kCosineTrue);
    static constexpr bool kIsCount = (M == kHamming || M == kJaccard);
    static constexpr bool kIsMax = (M == kChebyshev);
    static constexpr int kComps = (M == kJaccard || M == kCosineTrue) ? 2 : 1;
    using T = typename std::conditional<kIsDouble, double,
                                        typename std::conditional<kIsCount, uint32_t, float>::type>::type;
    __device__ static __forceinline__ T zero() { return T(0); }
    __device__ static __forceinline__ T op(T a, T b) {
        if constexpr (kIsMax) return fmaxf(a, b);
        else return a + b;
    }
    // lane-local partial of one row from the Scorer accumulators
    __device__ static __forceinline__ void from_scorer(const Scorer<M>& sc, T (&out)[kComps]) {
        if constexpr (kIsDouble) { out[0] = sc.d0; out[1] = sc.d1; }
        else if constexpr (M == kJaccard) { out[0] = sc.c0; out[1] = sc.c1; }
        else if constexpr (M == kHamming) { out[0] = sc.c0; }
        else if constexpr (kIsMax) { out[0] = fmaxf(fmaxf(sc.s0, sc.s1), fmaxf(sc.s2, sc.s3)); }
        else { out[0] = (sc.s0 + sc.s1) + (sc.s2 + sc.s3); }
    }
    // warp-reduced partial -> raw metric value (the tail of Scorer::finish)
    __device__ static __forceinline__ float finalize(const T (&v)[kComps], double q_norm, bool& bad, bool& fatal) {
        bad = false;
        fatal = false;
        float r;
        if constexpr (M == kCosine || M == kInnerProduct || M == kL2Squared || M == kManhattan || M == kChebyshev) {
            r = v[0];
            bad = !isfinite(r);
        } else if constexpr (M == kNegativeInnerProduct) {
            r = -v[0];
            bad = !isfinite(r);
        } else if constexpr (M == kL2) {
            bad = !isfinite(v[0]);
            r = sqrtf(v[0]);
        } else if constexpr (M == kHamming) {
            r = (float)v[0];
        } else if constexpr (M == kJaccard) {
            r = v[1] == 0u ? 0.0f : __fsub_rn(1.0f, __fdiv_rn((float)v[0], (float)v[1]));
        } else {
            const double rn = sqrt(v[1]);
            if (q_norm == 0.0 || rn == 0.0) {
                r = 0.0f;
            } else {
                double s = v[0] / (q_norm * rn);
                if (!isfinite(s)) { fatal = true; s = 0.0; }
                s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
                r = (float)s;
            }
        }
        return r;
    }
};

// Reduces R (1, 2, 4 or 8) per-lane partial rows across the warp, transposing as it goes: each of the first
// log2(R) stages halves the rows a lane still carries, the remaining stages are plain xor steps. The xor
// offsets run 16, 8, 4, 2, 1 for every row, i.e. the same summation tree as warp_sum. Afterwards v[0] of
// lane L holds the total of row row_slot_of_lane<R>(L).
template <typename Tr, typename T, int C, int R>
__device__ __forceinline__ void butterfly_rows(T (&v)[R][C], int lane) {
    int live = R;
#pragma unroll
    for (int bit = 16; bit >= 1; bit >>= 1) {
        if (live > 1) {
            const bool hi = (lane & bit) != 0;
            const int half = live / 2;
#pragma unroll
            for (int i = 0; i < R / 2; ++i)
                if (i < half) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const T send = hi ? v[i][c] : v[i + half][c], keep = hi ? v[i + half][c] : v[i][c];
                        v[i][c] = Tr::op(keep, __shfl_xor_sync(0xffffffffu, send, bit));
                    }
                }
            live = half;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) v[0][c] = Tr::op(v[0][c], __shfl_xor_sync(0xffffffffu, v[0][c], bit));
        }
    }
}
// First lane of the group that ends up with row slot r (inverse of row_slot_of_lane).
template <int R>
__device__ __forceinline__ constexpr uint32_t lead_lane_of_slot(int r) {
    return R == 8 ? (uint32_t)((((r >> 2) & 1) << 4) | (((r >> 1) & 1) << 3) | ((r & 1) << 2))
         : R == 4 ? (uint32_t)((((r >> 1) & 1) << 4) | ((r & 1) << 3))
         : R == 2 ? (uint32_t)((r & 1) << 4) : 0u;
}
template <int R>
__device__ __forceinline__ int row_slot_of_lane(int lane) {
    if constexpr (R == 8) return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    else if constexpr (R == 4) return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    else if constexpr (R == 2) return (lane >> 4) & 1;
    else return 0;
}

// ---------------------------------------------------------------------------------------
// Kernel A. NV > 0: query and R rows in registers (dims <= 128 * NV). NV == 0: generic loop.
template <int M, int NV, int R>
__global__ void __launch_bounds__(kScanThreads, (NV * R <= 12) ? 2 : 1)
flat_scan_kernel(const ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    const bool dump = p.dump_keys != nullptr;

    Collector col;
    col.init(smem, &s_thresh, &s_count, p.cap, p.ws.k);
    collector_attach_pivots(col, p.ws);
    __syncthreads();

    const uint32_t nvec = (p.dims + 3u) >> 2;           // float4 slots in the prefix
    const uint32_t tail_idx = nvec - 1u;                // slot holding the prefix tail
    const uint32_t tail_rem = p.dims - 4u * tail_idx;   // valid components in it (1..4)
    const bool need_mask = tail_rem != 4u;              // uniform: rows carry data beyond dims
    const float4* q4 = reinterpret_cast<const float4*>(p.queries + (size_t)qi * p.q_stride);
    const double q_norm = (M == kCosineTrue) ? p.q_norms[qi] : 0.0;

    constexpr int NQ = NV > 0 ? NV : 1;
    float4 q[NQ];
    if constexpr (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint32_t idx = lane + 32u * j;
            q[j] = idx < nvec ? q4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    constexpr uint32_t kTileRows = kScanWarps * R;
    const uint32_t num_tiles = (p.n + kTileRows - 1u) / kTileRows;
    const uint32_t slack = kSyncEvery * kTileRows;
    uint32_t step = 0;
    u64 g_prefetch = kKeyMax;

    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++step) {
        const uint32_t r0 = tile * kTileRows + warp * R;
        uint32_t drow[R];
        const float4* rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t row = r0 + r;
            bool valid = row < p.n;
            drow[r] = valid ? (p.row_sel ? p.row_sel[row] : row) : 0u;
            rp[r] = reinterpret_cast<const float4*>(p.rows + (size_t)drow[r] * p.row_stride);
        }

        float raw[R];
        if constexpr (NV > 0) {
            float4 b[R][NV];
            // issue every load of the step before anything consumes one
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool valid = (r0 + r) < p.n;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const uint32_t idx = lane + 32u * j;
                    b[r][j] = (valid && idx < nvec) ? ldg_stream(rp[r] + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // Lane-local partials of the R rows, ONE transposing butterfly for all of them, and the metric's
            // tail (for the true cosine an f64 sqrt and divide) evaluated once for R rows, one row per lane group.
            using Tr = PartialTraits<M>;
            using PT = typename Tr::T;
            constexpr int C = Tr::kComps;
            PT part[R][C];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (need_mask) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        if (lane + 32u * j == tail_idx) mask_tail(b[r][j], tail_rem);
                }
                Scorer<M> sc;
                sc.init();
#pragma unroll
                for (int j = 0; j < NV; ++j) sc.accum(q[j], b[r][j]);
                Tr::from_scorer(sc, part[r]);
            }
            butterfly_rows<Tr, PT, C, R>(part, lane);
            const int slot = row_slot_of_lane<R>(lane);
            const bool owner = (lane & (32 / R - 1)) == 0 && (r0 + slot) < p.n;
            bool bad, fatal;
            float rawv = Tr::finalize(part[0], q_norm, bad, fatal);
            if constexpr (kCanOverflow<M>) {
                uint32_t todo = __ballot_sync(0xffffffffu, owner && bad);
                if (todo) {   // cold path: redo the overflowed rows in f64 from the registers, warp-wide, one row at a time
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (todo & (1u << lead_lane_of_slot<R>(r))) {
                            Recover<M> rc;
                            rc.init();
#pragma unroll
                            for (int j = 0; j < NV; ++j) rc.accum(q[j], b[r][j]);
                            bool f2;
                            const float rec = rc.finish(f2);
                            if (slot == r) { rawv = rec; fatal = f2; }
                        }
                    }
                }
            }
            uint32_t drow_s = drow[0];
#pragma unroll
            for (int r = 1; r < R; ++r)
                if (slot == r) drow_s = drow[r];
            if (owner) {
                if (fatal) atomicMin(p.err_row + qi, r0 + slot);
                emit_row<M>(p, col, qi, dump, dump ? kKeyMax : col.threshold(), rawv, r0 + slot, drow_s);
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool valid = (r0 + r) < p.n;
                Scorer<M> sc;
                sc.init();
                if (valid) {
                    for (uint32_t idx = lane; idx < nvec; idx += 32u) {
                        float4 b = ldg_stream(rp[r] + idx);
                        if (idx == tail_idx) mask_tail(b, tail_rem);
                        sc.accum(__ldg(q4 + idx), b);
                    }
                }
                bool bad, fatal;
                raw[r] = sc.finish(q_norm, bad, fatal);
                if constexpr (kCanOverflow<M>) {
                    if (bad) {
                        Recover<M> rc;
                        rc.init();
                        for (uint32_t idx = lane; idx < nvec; idx += 32u) {
                            float4 b = ldg_stream(rp[r] + idx);
                            if (idx == tail_idx) mask_tail(b, tail_rem);
                            rc.accum(__ldg(q4 + idx), b);
                        }
                        raw[r] = rc.finish(fatal);
                    }
                }
                if (fatal && valid && lane == 0) atomicMin(p.err_row + qi, r0 + r);
            }
        }

        if (NV == 0 && lane == 0) {
            const u64 T = dump ? kKeyMax : col.threshold();
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (r0 + r < p.n) emit_row<M>(p, col, qi, dump, T, raw[r], r0 + r, drow[r]);
        }

        if (!dump && (step & (kSyncEvery - 1)) == kSyncEvery - 1)
            collector_checkpoint(col, p.ws, qi, slack, g_prefetch);
    }
    if (dump) return;

    collector_publish_and_merge(col, p.ws, qi, &s_last);
}

// ---------------------------------------------------------------------------------------
// Kernel B: TMA-staged whole-row stream. 8 consumer warps + 1 producer warp.
constexpr int kStreamMaxStages = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct StreamGeom {
    uint32_t stages;      // ring depth (kernel C, prefix_lane.cu: tiles per consumer warp)
    uint32_t tile_bytes;  // bytes per stage (tile_rows * row_floats * 4), multiple of 128
    uint32_t row_floats;  // floats per row IN SHARED MEMORY: the row stride (whole rows) or the padded prefix
    uint32_t tail_rem;    // valid components of the last float4 of a row (4 = nothing to mask); kernel C masks no
                          // tail (its tensor map is `dims` wide) and reads this word as its checkpoint cadence
    uint32_t use_tmap;    // 1: prefix scan, tiles arrive through the 2D tensor map (only the scored columns)
};

constexpr int kGroupRows = 8;  // rows whose partials one warp reduces with a single butterfly

// Reduces 8 per-lane partial rows across the warp with 4+2+1+1+1 = 9 shuffles per component
// (instead of 8 x 5). Afterwards v[0] of lane L holds the total of row slot_of_lane(L).
template <typename Tr, typename T, int C>
__device__ __forceinline__ void butterfly8(T (&v)[kGroupRows][C], int lane) {
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            T send = hi ? v[i][c] : v[i + 4][c], keep = hi ? v[i + 4][c] : v[i][c];
            v[i][c] = Tr::op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
        }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            T send = hi ? v[i][c] : v[i + 2][c], keep = hi ? v[i + 2][c] : v[i][c];
            v[i][c] = Tr::op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
        }
    hi = (lane & 4) != 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        T send = hi ? v[0][c] : v[1][c], keep = hi ? v[1][c] : v[0][c];
        v[0][c] = Tr::op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
        v[0][c] = Tr::op(v[0][c], __shfl_xor_sync(0xffffffffu, v[0][c], 2));
        v[0][c] = Tr::op(v[0][c], __shfl_xor_sync(0xffffffffu, v[0][c], 1));
    }
}
__device__ __forceinline__ int slot_of_lane(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// Rows must be whole (dims covers the stride's float4 slots) and contiguous (no row_sel).
// W consumer warps + 1 producer warp; every consumer warp owns RPW rows of each tile and
// defers the cross-lane reduction until it has kGroupRows row partials.
template <int M, int NV, int RPW, int W>
__global__ void __launch_bounds__(W * 32 + 32, 1)
flat_stream_kernel(const ScanParams p, const StreamGeom geom, const __grid_constant__ CUtensorMap tmap) {
    using Tr = PartialTraits<M>;
    using T = typename Tr::T;
    constexpr int C = Tr::kComps;
    static_assert(kGroupRows % RPW == 0, "RPW must divide the group size");
    constexpr int kTilesPerGroup = kGroupRows / RPW;

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kStreamMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kStreamMaxStages];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    constexpr uint32_t kTileRows = W * RPW;
    const uint32_t num_tiles = (p.n + kTileRows - 1u) / kTileRows;
    const uint32_t row_bytes = geom.row_floats * 4u;
    unsigned char* ring = smem;
    unsigned char* col_mem = smem + (size_t)geom.stages * geom.tile_bytes;

    Collector col;
    col.init(col_mem, &s_thresh, &s_count, p.cap, p.ws.k, W * 32, 1);
    collector_attach_pivots(col, p.ws);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < geom.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == W) {
        // ===== producer: one elected lane streams this CTA's tiles through the ring =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % geom.stages, ph = (it / geom.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const uint32_t row0 = tile * kTileRows;
                if (geom.use_tmap) {
                    // box = [kTileRows rows x row_floats columns]; rows past the matrix are zero-filled and still count
                    mbar_arrive_expect_tx(&full_bar[s], kTileRows * row_bytes);
                    tc::tma_load_2d(ring + (size_t)s * geom.tile_bytes, &tmap, 0u, row0, &full_bar[s]);
                } else {
                    const uint32_t rows = min(kTileRows, p.n - row0);
                    const uint32_t bytes = rows * row_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], bytes);
                    tma_bulk_g2s(ring + (size_t)s * geom.tile_bytes, p.rows + (size_t)row0 * p.row_stride, bytes,
                                 &full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const uint32_t nvec = geom.row_floats >> 2;
    const uint32_t tail_idx = nvec - 1u;
    const bool need_mask = geom.tail_rem != 4u;         // prefix not a multiple of 4: the box carries columns beyond it
    const float4* q4 = reinterpret_cast<const float4*>(p.queries + (size_t)qi * p.q_stride);
    const double q_norm = (M == kCosineTrue) ? p.q_norms[qi] : 0.0;
    float4 q[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        uint32_t idx = lane + 32u * j;
        q[j] = idx < nvec ? q4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // Reduce + emit the current group: slot s = tile_in_group * RPW + r.
    T part[kGroupRows][C];

    auto flush_group = [&](uint32_t tile0) {
        butterfly8<Tr, T, C>(part, lane);
        const int slot = slot_of_lane(lane);
        const uint32_t tile = tile0 + (uint32_t)(slot / RPW) * gridDim.x;
        const uint32_t row = tile * kTileRows + (uint32_t)warp * RPW + (uint32_t)(slot % RPW);
        const bool owner = (lane & 3) == 0 && tile < num_tiles && row < p.n;
        bool bad = false, fatal = false;
        float raw = Tr::finalize(part[0], q_norm, bad, fatal);
        if constexpr (kCanOverflow<M>) {
            // cold path: recompute overflowed rows in f64 straight from HBM, one row at a time
            uint32_t todo = __ballot_sync(0xffffffffu, owner && bad);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t rrow = __shfl_sync(0xffffffffu, row, src);
                const float4* rp = reinterpret_cast<const float4*>(p.rows + (size_t)rrow * p.row_stride);
                Recover<M> rc;
                rc.init();
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const uint32_t idx = lane + 32u * j;
                    if (idx < nvec) {
                        float4 bv = ldg_stream(rp + idx);
                        if (need_mask && idx == tail_idx) mask_tail(bv, geom.tail_rem);
                        rc.accum(q[j], bv);
                    }
                }
                bool f2;
                const float rec = rc.finish(f2);
                if (lane == src) { raw = rec; fatal = f2; }
            }
        }
        if (owner) {
            if (fatal) atomicMin(p.err_row + qi, row);
            emit_row<M>(p, col, qi, false, col.threshold(), raw, row, row);
        }
    };

    constexpr uint32_t kGroupsPerSync = kStreamSyncEvery / kTilesPerGroup > 0 ? kStreamSyncEvery / kTilesPerGroup : 1;
    uint32_t it = 0, grp = 0;
    u64 g_prefetch = kKeyMax;
    for (uint32_t tile0 = blockIdx.x; tile0 < num_tiles; tile0 += gridDim.x * kTilesPerGroup, ++grp) {
#pragma unroll
        for (int g = 0; g < kGroupRows; ++g)
#pragma unroll
            for (int c = 0; c < C; ++c) part[g][c] = Tr::zero();
#pragma unroll
        for (int t = 0; t < kTilesPerGroup; ++t) {
            const uint32_t tile = tile0 + (uint32_t)t * gridDim.x;
            if (tile >= num_tiles) break;  // CTA-uniform
            const uint32_t s = it % geom.stages, ph = (it / geom.stages) & 1u;
            ++it;
            if (lane == 0) mbar_wait(&full_bar[s], ph);
            __syncwarp();
            const float4* tile4 = reinterpret_cast<const float4*>(ring + (size_t)s * geom.tile_bytes);
            const uint32_t r0 = tile * kTileRows + warp * RPW;
            float4 b[RPW][NV];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const bool valid = (r0 + r) < p.n;
                const float4* rp = tile4 + (size_t)(warp * RPW + r) * nvec;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const uint32_t idx = lane + 32u * j;
                    b[r][j] = (valid && idx < nvec) ? rp[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (need_mask) {   // CTA-uniform and AFTER every load was issued: a select between loads serialises them
#pragma unroll
                for (int r = 0; r < RPW; ++r)
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        if (lane + 32u * j == tail_idx) mask_tail(b[r][j], geom.tail_rem);
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                Scorer<M> sc;
                sc.init();
                if (!(p.debug & 2u)) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) sc.accum(q[j], b[r][j]);
                }
                Tr::from_scorer(sc, part[t * RPW + r]);
            }
            // every lane's shared-memory reads have been consumed: hand the slot back
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        if (!(p.debug & 1u)) flush_group(tile0);
        if (!(p.debug & 4u) && (grp & (kGroupsPerSync - 1)) == kGroupsPerSync - 1)
            collector_checkpoint(col, p.ws, qi, kGroupsPerSync * kGroupRows * W, g_prefetch);
    }
    // every tile was consumed, so the ring is idle: the last CTA merges in ring + collector memory
    const uint32_t total_smem = geom.stages * geom.tile_bytes + p.cap * 16u;
    uint32_t big_cap = p.cap;
    while ((size_t)big_cap * 2u * 16u <= total_smem) big_cap *= 2u;
    collector_publish_and_merge(col, p.ws, qi, &s_last, smem, big_cap);
}

}  // namespace vb
