// maxsim.cu — K5 (general form): ColBERT MaxSim scoring + fused top-k over a device-resident
// token matrix (reference multi_vector.rs:65-132). All nine metrics plus the f64
// renormalising cosine; the tensor-core kernel in maxsim_tc.cu takes the inner-product
// family when shapes allow, this one is the general path and the arbiter of semantics.
//
// Per document: S[q][d] = metric(query token q, doc token d) for every pair, each query
// token keeps max_d similarity_value(S), the document score is the f32 sum of those
// maxima in query-token order (multi_vector.rs:70-86), "score overflow" when the running
// sum turns non-finite. One CTA owns one document at a time; S is produced in 32 x 128
// register-tiled blocks (each thread a 4 x 4 patch) from shared-memory chunks of 32 dims.
#include "maxsim.h"
#include "scan_driver.h"

#include <cstdlib>

#include "topk.cuh"

namespace vb {

constexpr int kMsThreads = 256;
constexpr int kMsQB = 32;    // query tokens per block
constexpr int kMsDT = 128;   // doc tokens per tile
constexpr int kMsKC = 32;    // dims per shared-memory chunk
constexpr int kMsLdQ = kMsQB + 4;
constexpr int kMsLdD = kMsDT + 4;

struct MaxSimParams {
    const float* tokens;       // [ntok, stride]
    size_t stride;             // floats, multiple of 4
    const uint32_t* doc_off;   // [ndocs + 1] token offsets
    const uint32_t* doc_rank;  // [ndocs] id rank, 0xFFFFFFFF = deleted
    uint32_t ndocs, dims;
    const float* query;        // [tq, stride]
    const double* q_norms;     // [tq] (kCosineTrue)
    uint32_t tq;
    uint32_t cap;
    uint32_t* err;             // atomicMin of (doc << 1 | kind): kind 0 metric overflow, 1 score overflow
    u64* dump_keys;            // limit beyond the fused collector: every live document's key / payload goes to
    u64* dump_pays;            // [ndocs] arrays (pre-filled with kKeyMax) and the host radix-sorts them
    TopkWorkspace ws;
};

template <int M>
struct PairAcc {
    static constexpr bool kDouble = (M == kCosineTrue);
    static constexpr bool kCount = (M == kHamming || M == kJaccard);
    using T = typename std::conditional<kDouble, double, typename std::conditional<kCount, uint32_t, float>::type>::type;
    T a;   // main accumulator
    T b;   // jaccard: union count
    __device__ __forceinline__ void init() { a = T(0); b = T(0); }
    __device__ __forceinline__ void step(float q, float d) {
        if constexpr (M == kCosine || M == kInnerProduct || M == kNegativeInnerProduct) a = fmaf(q, d, a);
        else if constexpr (M == kL2 || M == kL2Squared) { float t = q - d; a = fmaf(t, t, a); }
        else if constexpr (M == kManhattan) a += fabsf(q - d);
        else if constexpr (M == kChebyshev) a = fmaxf(a, fabsf(q - d));
        else if constexpr (M == kHamming) a += ((q != 0.0f) != (d != 0.0f));
        else if constexpr (M == kJaccard) { bool l = q != 0.0f, r = d != 0.0f; a += (l && r); b += (l || r); }
        else a = fma((double)q, (double)d, a);
    }
};

// f64 recomputation of one pair straight from global memory (distances.rs:70-98).
template <int M>
__device__ float recover_pair(const float* q, const float* d, uint32_t dims, bool& fatal) {
    double v = 0.0;
    for (uint32_t i = 0; i < dims; ++i) {
        const double a = q[i], b = d[i];
        if constexpr (M == kCosine || M == kInnerProduct || M == kNegativeInnerProduct) v = fma(a, b, v);
        else if constexpr (M == kL2 || M == kL2Squared) v = fma(a - b, a - b, v);
        else if constexpr (M == kManhattan) v += fabs(a - b);
        else if constexpr (M == kChebyshev) v = fmax(v, fabs(a - b));
    }
    if constexpr (M == kL2) v = sqrt(v);
    if constexpr (M == kNegativeInnerProduct) v = -v;
    const double mx = 3.4028234663852886e38;
    fatal = !(isfinite(v) && v >= -mx && v <= mx);
    return fatal ? 0.0f : (float)v;
}

template <int M>
__global__ void __launch_bounds__(kMsThreads) maxsim_kernel(const MaxSimParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_thresh;
    __shared__ uint32_t s_count;
    __shared__ int s_last;
    __shared__ __align__(16) float Qs[kMsKC][kMsLdQ];
    __shared__ __align__(16) float Ds[kMsKC][kMsLdD];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Collector col;
    col.init(smem, &s_thresh, &s_count, p.cap, p.ws.k);
    float* s_best = reinterpret_cast<float*>(smem + (size_t)p.cap * 16);   // [tq]
    __syncthreads();

    using Acc = PairAcc<M>;
    const uint32_t kchunks = (p.dims + kMsKC - 1) / kMsKC;
    u64 g_prefetch = kKeyMax;
    uint32_t iter = 0;

    for (uint32_t doc = blockIdx.x; doc < p.ndocs; doc += gridDim.x, ++iter) {
        const uint32_t rank = p.doc_rank ? p.doc_rank[doc] : doc;
        const uint32_t t0 = p.doc_off[doc], td = p.doc_off[doc + 1] - t0;
        const bool live = rank != 0xFFFFFFFFu;
        if (live && td > 0) {
            for (uint32_t qb = 0; qb < p.tq; qb += kMsQB) {
                float rowbest[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                for (uint32_t dt = 0; dt < td; dt += kMsDT) {
                    Acc acc[4][4];
                    double dn[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j].init();
                    for (uint32_t kc = 0; kc < kchunks; ++kc) {
                        const uint32_t k0 = kc * kMsKC;
                        __syncthreads();
                        {   // query chunk: 32 tokens x 32 dims, transposed into Qs[k][q]
                            const uint32_t qt = tid >> 3, kq = tid & 7;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (qb + qt < p.tq && k0 + kq * 4 < p.stride)
                                v = *reinterpret_cast<const float4*>(p.query + (size_t)(qb + qt) * p.stride + k0 + kq * 4);
                            Qs[kq * 4 + 0][qt] = v.x; Qs[kq * 4 + 1][qt] = v.y;
                            Qs[kq * 4 + 2][qt] = v.z; Qs[kq * 4 + 3][qt] = v.w;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {   // doc chunk: 128 tokens x 32 dims -> Ds[k][d]
                            const uint32_t f = tid + kMsThreads * i, tok = f >> 3, kq = f & 7;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (dt + tok < td && k0 + kq * 4 < p.stride)
                                v = ldg_stream(reinterpret_cast<const float4*>(p.tokens + (size_t)(t0 + dt + tok) * p.stride + k0 + kq * 4));
                            Ds[kq * 4 + 0][tok] = v.x; Ds[kq * 4 + 1][tok] = v.y;
                            Ds[kq * 4 + 2][tok] = v.z; Ds[kq * 4 + 3][tok] = v.w;
                        }
                        __syncthreads();
#pragma unroll 8
                        for (int k = 0; k < kMsKC; ++k) {
                            const float4 a = *reinterpret_cast<const float4*>(&Qs[k][warp * 4]);
                            const float4 b = *reinterpret_cast<const float4*>(&Ds[k][lane * 4]);
                            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[i][j].step(av[i], bv[j]);
                            if constexpr (M == kCosineTrue) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) dn[j] = fma((double)bv[j], (double)bv[j], dn[j]);
                            }
                        }
                    }
                    // finalise the 4 x 4 patch: raw -> similarity -> running max over doc tokens
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t q = qb + warp * 4 + i;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t d = dt + lane * 4 + j;
                            if (q >= p.tq || d >= td) continue;
                            float raw;
                            bool bad = false, fatal = false;
                            if constexpr (M == kCosineTrue) {
                                const double qn = p.q_norms[q], rn = sqrt(dn[j]);
                                if (qn == 0.0 || rn == 0.0) raw = 0.0f;
                                else {
                                    double s = acc[i][j].a / (qn * rn);
                                    if (!isfinite(s)) { fatal = true; s = 0.0; }
                                    s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
                                    raw = (float)s;
                                }
                            } else if constexpr (M == kHamming) {
                                raw = (float)acc[i][j].a;
                            } else if constexpr (M == kJaccard) {
                                raw = acc[i][j].b == 0u ? 0.0f
                                                        : __fsub_rn(1.0f, __fdiv_rn((float)acc[i][j].a, (float)acc[i][j].b));
                            } else {
                                raw = acc[i][j].a;
                                bad = !isfinite(raw);
                                if constexpr (M == kL2) raw = sqrtf(raw);
                                if constexpr (M == kNegativeInnerProduct) raw = -raw;
                                if (bad)
                                    raw = recover_pair<M>(p.query + (size_t)q * p.stride,
                                                          p.tokens + (size_t)(t0 + d) * p.stride, p.dims, fatal);
                            }
                            if (fatal) atomicMin(p.err, doc << 1);
                            rowbest[i] = fmaxf(rowbest[i], similarity_value(M, raw));   // multi_vector.rs:79
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float m = warp_max(rowbest[i]);
                    if (lane == 0 && qb + warp * 4 + i < p.tq) s_best[qb + warp * 4 + i] = m;
                }
            }
            __syncthreads();
        }
        if (tid == 0 && live) {
            float total = 0.0f;
            if (td > 0) {
                for (uint32_t q = 0; q < p.tq; ++q) {   // multi_vector.rs:81-84: f32, query order
                    total += s_best[q];
                    if (!isfinite(total)) { atomicMin(p.err, (doc << 1) | 1u); total = 0.0f; break; }
                }
            }
            // descending score (total order), ascending id: multi_vector.rs:22-31
            const u64 key = ((u64)(~order_key(total)) << 32) | rank;
            const u64 pay = ((u64)__float_as_uint(total) << 32) | doc;
            if (p.dump_keys) { p.dump_keys[doc] = key; p.dump_pays[doc] = pay; }
            else if (key < col.threshold()) col.push(key, pay);
        }
        if (!p.dump_keys && (iter & 63u) == 63u) collector_checkpoint(col, p.ws, 0, 64, g_prefetch);
    }
    if (!p.dump_keys) collector_publish_and_merge(col, p.ws, 0, &s_last);
}

typedef void (*MaxSimKernel)(const MaxSimParams);
static MaxSimKernel maxsim_lookup(int metric) {
    switch (metric) {
        case kL2: return maxsim_kernel<kL2>;
        case kL2Squared: return maxsim_kernel<kL2Squared>;
        case kCosine: return maxsim_kernel<kCosine>;
        case kInnerProduct: return maxsim_kernel<kInnerProduct>;
        case kNegativeInnerProduct: return maxsim_kernel<kNegativeInnerProduct>;
        case kManhattan: return maxsim_kernel<kManhattan>;
        case kChebyshev: return maxsim_kernel<kChebyshev>;
        case kHamming: return maxsim_kernel<kHamming>;
        case kJaccard: return maxsim_kernel<kJaccard>;
        case kCosineTrue: return maxsim_kernel<kCosineTrue>;
    }
    return nullptr;
}

// Dump mode (limit > 1024): [ndocs] key / payload arrays pre-filled with kKeyMax (deleted documents write nothing).
Status maxsim_prepare_dump(SearchCtx& ctx, size_t ndocs) {
    VB_TRY(ctx.dump_keys.reserve(ndocs * sizeof(u64)));
    VB_TRY(ctx.dump_pays.reserve(ndocs * sizeof(u64)));
    VB_TRY(ctx.dump_keys2.reserve(ndocs * sizeof(u64)));
    VB_TRY(ctx.dump_pays2.reserve(ndocs * sizeof(u64)));
    VB_CUDA(cudaMemsetAsync(ctx.dump_keys.p, 0xFF, ndocs * sizeof(u64), ctx.stream));
    return Status::Ok();
}

static void maxsim_decode(SearchCtx& ctx, uint32_t k, MaxSimResult* out) {
    const u64* pays = ctx.h_result.as<u64>();
    const uint32_t* tail = reinterpret_cast<const uint32_t*>(pays + k);
    const uint32_t count = tail[0];
    out->err = tail[1];
    out->rows.resize(count);
    out->scores.resize(count);
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t bits = (uint32_t)(pays[i] >> 32);
        std::memcpy(&out->scores[i], &bits, 4);
        out->rows[i] = (uint32_t)pays[i];
    }
}

Status maxsim_collect_dump(SearchCtx& ctx, size_t ndocs, uint32_t k, cudaError_t e, MaxSimResult* out) {
    if (e != cudaSuccess) { ctx.poison(); return Status::Cuda(cudaGetErrorString(e)); }
    Status s = sort_dump_and_fetch(ctx, ndocs, k);
    if (!s.ok()) { ctx.poison(); return s; }
    maxsim_decode(ctx, k, out);
    return Status::Ok();
}

Status maxsim_collect_result(SearchCtx& ctx, const MaxSimJob& job, const TopkWorkspace& ws, uint32_t k, cudaError_t e,
                             MaxSimResult* out) {
    if (e == cudaSuccess && job.d_keys_out) {
        Status u = unpack_device_results(ws.out_keys, ws.out_pays, ws.out_counts, 1, k, job.d_keys_out, job.d_values_out,
                                         job.d_rows_out, job.d_counts_out, ctx.stream);
        if (!u.ok()) { ctx.poison(); return u; }
    }
    const size_t bytes = (size_t)k * sizeof(u64) + 8;
    if (e == cudaSuccess) e = ctx.h_result.reserve(bytes).ok() ? cudaSuccess : cudaErrorMemoryAllocation;
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx.h_result.p, ctx.result.p, bytes, cudaMemcpyDeviceToHost, ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
    if (e != cudaSuccess) {
        ctx.poison();
        return Status::Cuda(cudaGetErrorString(e));
    }
    maxsim_decode(ctx, k, out);
    return Status::Ok();
}

// Which kernel answered this thread's last maxsim_top_k (tests): 0 general, 1 tensor-core uniform, 2 tensor-core
// ragged, 3 / 4 = 1 / 2 flagged a non-finite score and the general kernel repeated the query.
static thread_local int t_last_path = 0;
extern "C" int vb_debug_maxsim_path() { return t_last_path; }

Status maxsim_top_k(SearchCtx& ctx, const MaxSimJob& job, MaxSimResult* out) {
    t_last_path = 0;
    out->rows.clear();
    out->scores.clear();
    out->err = kNoError;
    if (job.ndocs == 0 || job.k == 0 || job.tq == 0) return Status::Cuda("empty maxsim job");
    const uint32_t k = (uint32_t)std::min<size_t>(job.k, job.ndocs);
    // A limit beyond the fused collector (1024): every live document's score is written out and radix-sorted.
    const bool dump = k > (uint32_t)kMaxFusedK;
    if (dump && job.d_keys_out) return Status::Cuda("sharded multi-vector search is limited to 1024 hits per shard");
    // Tensor-core kernels first. They raise the error word for a non-finite pair or sum; which of the reference's
    // two overflow errors that is (or whether f64 recovery rescues the pair, distances.rs:59-98) is decided by
    // repeating the query on the general kernel below.
    if (maxsim_tc_eligible(job, job.uniform_td) && (job.metric != kCosineTrue || job.d_inv_dnorm)) {
        VB_TRY(maxsim_tc_top_k(ctx, job, job.uniform_td, job.d_inv_dnorm, out));
        t_last_path = 1;
        if (out->err == kNoError) return Status::Ok();
    } else if (maxsim_tcr_eligible(job)) {
        VB_TRY(maxsim_tcr_top_k(ctx, job, out));
        t_last_path = 2;
        if (out->err == kNoError) return Status::Ok();
    }
    t_last_path = t_last_path ? t_last_path + 2 : 0;   // 3 / 4: a tensor-core kernel flagged the query, redone here
    out->rows.clear();
    out->scores.clear();
    out->err = kNoError;
    MaxSimKernel kernel = maxsim_lookup(job.metric);
    if (!kernel) return Status::Ref("unknown metric");

    // stage the query tokens (zero padded to the token stride) and their f64 norms
    const size_t qbytes = (size_t)job.tq * job.stride * sizeof(float);
    VB_TRY(ctx.h_queries.reserve(qbytes + job.tq * sizeof(double)));
    VB_TRY(ctx.queries.reserve(qbytes));
    VB_TRY(ctx.q_norms.reserve(job.tq * sizeof(double)));
    float* hq = ctx.h_queries.as<float>();
    double* hn = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(hq) + qbytes);
    for (uint32_t q = 0; q < job.tq; ++q) {
        const float* src = job.h_query + (size_t)q * job.dims;
        float* dst = hq + (size_t)q * job.stride;
        std::memcpy(dst, src, job.dims * sizeof(float));
        for (size_t c = job.dims; c < job.stride; ++c) dst[c] = 0.0f;
        double s = 0.0;
        for (uint32_t i = 0; i < job.dims; ++i) s += (double)src[i] * (double)src[i];
        hn[q] = std::sqrt(s);
    }
    VB_CUDA(cudaMemcpyAsync(ctx.queries.p, hq, qbytes, cudaMemcpyHostToDevice, ctx.stream));
    VB_CUDA(cudaMemcpyAsync(ctx.q_norms.p, hn, job.tq * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));

    const uint32_t kc = dump ? 1u : k;   // the collector is idle in dump mode
    uint32_t cap = 256;
    while (cap < 2 * kc || cap < kc + 64) cap <<= 1;
    const size_t smem = (size_t)cap * 16 + (size_t)job.tq * sizeof(float);
    if (smem > 160 * 1024) return Status::Cuda("too many query tokens for the multi-vector kernel");
    VB_TRY(ensure_dynamic_smem_for(kernel, smem));
    int per_sm = 0, dev = 0, sms = 0;
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMsThreads, smem));
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm < 1) return Status::Cuda("multi-vector kernel does not fit on an SM");
    const uint32_t grid = (uint32_t)std::min<size_t>(job.ndocs, (size_t)sms * per_sm);

    VB_TRY(ctx.arm_ctrl(1));
    VB_TRY(ctx.cand_keys.reserve((size_t)grid * kc * sizeof(u64)));
    VB_TRY(ctx.cand_pays.reserve((size_t)grid * kc * sizeof(u64)));
    VB_TRY(ctx.cand_counts.reserve((size_t)grid * sizeof(uint32_t)));
    VB_TRY(ctx.out_keys.reserve((size_t)kc * sizeof(u64)));
    VB_TRY(ctx.result.reserve((size_t)kc * sizeof(u64) + 8));
    if (dump) VB_TRY(maxsim_prepare_dump(ctx, job.ndocs));

    MaxSimParams p{};
    p.dump_keys = dump ? ctx.dump_keys.as<u64>() : nullptr;
    p.dump_pays = dump ? ctx.dump_pays.as<u64>() : nullptr;
    p.tokens = job.d_tokens;
    p.stride = job.stride;
    p.doc_off = job.d_doc_off;
    p.doc_rank = job.d_doc_rank;
    p.ndocs = (uint32_t)job.ndocs;
    p.dims = job.dims;
    p.query = ctx.queries.as<float>();
    p.q_norms = ctx.q_norms.as<double>();
    p.tq = job.tq;
    p.cap = cap;
    p.err = ctx.err_row();
    p.ws.k = kc;
    p.ws.cand_keys = ctx.cand_keys.as<u64>();
    p.ws.cand_pays = ctx.cand_pays.as<u64>();
    p.ws.cand_counts = ctx.cand_counts.as<uint32_t>();
    p.ws.done = ctx.done();
    p.ws.g_thresh = ctx.g_thresh();
    p.ws.out_keys = ctx.out_keys.as<u64>();
    p.ws.out_pays = ctx.result.as<u64>();
    p.ws.out_counts = reinterpret_cast<uint32_t*>(ctx.result.as<u64>() + kc);
    p.ws.err_row = ctx.err_row();
    p.ws.out_err = p.ws.out_counts + 1;
    kernel<<<grid, kMsThreads, smem, ctx.stream>>>(p);
    if (dump) return maxsim_collect_dump(ctx, job.ndocs, k, cudaGetLastError(), out);
    return maxsim_collect_result(ctx, job, p.ws, k, cudaGetLastError(), out);
}

}  // namespace vb
