// flat_scan.cuh — K1/K4: single-query exact scan + fused top-k over an HBM-resident
// row-major fp32 matrix. One kernel family serves
//   * FlatIndex::search            (reference flat.rs:96-124)      metric = index metric
//   * search::vector_top_k         (reference search.rs:38-73)     prefix dims, kCosineTrue
//   * exact rerank over a row list (reference collection.ex:821-851)
//
// Shape of the work: an N x D fp32 stream read exactly once per query (N*D*4 algorithmic
// bytes), 2 flops per element -> HBM-bound. Design:
//   - persistent grid (SM count x resident CTAs), 8 warps per CTA, each warp owns R whole
//     rows per step; a lane owns float4 slots {lane + 32 j}: every warp-level load is one
//     fully coalesced 512-byte segment, 128-bit per lane, read-only/no-L1-allocate;
//   - the query lives in registers (NV float4 per lane), all R*NV loads of a step are
//     issued before the first FMA so each lane keeps R*NV 16-byte requests in flight;
//   - the score never goes back to HBM: rank key + id rank go straight into the CTA's
//     shared-memory collector (topk.cuh); the last CTA to finish merges the per-CTA
//     lists, so one launch per query returns the sorted top-k;
//   - non-finite f32 results are recomputed in f64 from the registers already loaded
//     (reference distances.rs:59-98); unrepresentable ones flag "metric overflow".
#pragma once
#include "topk.cuh"

namespace vb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kSyncEvery = 4;  // steps between collector checks (power of two)

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
// NV > 0: query and R rows in registers (dims <= 128 * NV). NV == 0: generic loop (any dims).
template <int M, int NV, int R>
__global__ void __launch_bounds__(kScanThreads, (NV * R <= 12) ? 2 : 1)
flat_scan_kernel(const ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.y;
    const bool dump = p.dump_keys != nullptr;

    Collector col;
    col.init(smem, &s_thresh, &s_count, p.cap, p.ws.k);
    __syncthreads();

    const uint32_t nvec = (p.dims + 3u) >> 2;           // float4 slots in the prefix
    const uint32_t tail_idx = nvec - 1u;                // slot holding the prefix tail
    const uint32_t tail_rem = p.dims - 4u * tail_idx;   // valid components in it (1..4)
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
        bool fatal_any = false;

        if constexpr (NV > 0) {
            float4 b[R][NV];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool valid = (r0 + r) < p.n;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    uint32_t idx = lane + 32u * j;
                    if (valid && idx < nvec) {
                        b[r][j] = ldg_stream(rp[r] + idx);
                        if (idx == tail_idx) mask_tail(b[r][j], tail_rem);
                    } else {
                        b[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                Scorer<M> sc;
                sc.init();
#pragma unroll
                for (int j = 0; j < NV; ++j) sc.accum(q[j], b[r][j]);
                bool bad, fatal;
                raw[r] = sc.finish(q_norm, bad, fatal);
                if constexpr (kCanOverflow<M>) {
                    if (bad) {  // warp-uniform cold path
                        Recover<M> rc;
                        rc.init();
#pragma unroll
                        for (int j = 0; j < NV; ++j) rc.accum(q[j], b[r][j]);
                        raw[r] = rc.finish(fatal);
                    }
                }
                fatal_any |= fatal && ((r0 + r) < p.n);
                if (fatal && (r0 + r) < p.n && lane == 0) atomicMin(p.err_row + qi, r0 + r);
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
        (void)fatal_any;

        if (lane == 0) {
            const u64 T = dump ? kKeyMax : col.threshold();
            const uint32_t t_hi = (uint32_t)(T >> 32);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t row = r0 + r;
                if (row >= p.n) continue;
                const uint32_t rk = order_key(rank_value(M, raw[r]));
                if (rk > t_hi) continue;
                const uint32_t idr = p.id_rank ? __ldg(p.id_rank + drow[r]) : drow[r];
                const u64 key = ((u64)rk << 32) | idr;
                const u64 pay = ((u64)__float_as_uint(raw[r]) << 32) | drow[r];
                if (dump) {
                    p.dump_keys[(size_t)qi * p.n + row] = key;
                    p.dump_pays[(size_t)qi * p.n + row] = pay;
                } else if (key < T) {
                    col.push(key, pay);
                }
            }
        }

        if (!dump && (step & (kSyncEvery - 1)) == kSyncEvery - 1) collector_checkpoint(col, p.ws, qi, slack);
    }
    if (dump) return;

    collector_publish_and_merge(col, p.ws, qi, &s_last);
}

}  // namespace vb
