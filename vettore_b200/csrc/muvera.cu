// muvera.cu — MUVERA fixed-dimensional encoding on the device, batched over documents
// (reference native/vettore/src/muvera.rs:26-74 behind nifs.rs:430-476).
//
// The reference encodes ONE multi-vector per NIF call with three nested scalar loops; every weight and
// sign is re-derived from a 64-bit hash for every (vector, projection, coordinate). A corpus encode
// (Vettore.Encoding.Muvera.encode_document per document before the flat inner-product index is built) is
// embarrassingly parallel over documents, repetitions and projections — but the per-slot arithmetic is
// ORDER-DEPENDENT: a document slot is a running mean rounded to f32 after every vector
// (muvera.rs:163-176), a projected value is an f64 sum in coordinate order (:149-160). The device version
// keeps exactly those orders (one thread owns a slot column and walks the document's vectors in order;
// f64 products and sums are separate IEEE operations, never contracted), so its output is bit-identical
// to the reference arithmetic — the tests compare with the oracle for equality, not within a tolerance.
//
// Launches per batch: hash tables (weights, signs: config only) -> partition of every (vector, repetition)
// (SimHash, :111-131) -> arrival number of every vector inside its partition -> accumulation -> optional
// count sketch (:179-196; slot / sign per input index are config-only and built on the host as a CSR list).
#include "muvera.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace vb {

namespace {

constexpr size_t kMaxOutputDimensions = 16777216;   // muvera.rs:23

__host__ __device__ inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// muvera.rs:215-221
__host__ __device__ inline uint64_t hash4(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t x = a ^ rotl64(b, 17) ^ rotl64(c, 31) ^ rotl64(d, 47);
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// weights[rep][proj][d] (muvera.rs:199-203) and signs[rep][proj][d] (:206-212, seed + 17)
__global__ void muvera_tables_kernel(uint64_t seed, uint32_t reps, uint32_t ks, uint32_t pdim, uint32_t dim, float* weights,
                                     signed char* signs) {
    const size_t nw = (size_t)reps * ks * dim, ns = signs ? (size_t)reps * pdim * dim : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw + ns; i += (size_t)gridDim.x * blockDim.x) {
        if (i < nw) {
            const uint64_t d = i % dim, proj = (i / dim) % ks, rep = i / ((size_t)dim * ks);
            const uint64_t h = hash4(seed, rep, proj, d);
            const float unit = (float)(__ddiv_rn((double)h, 18446744073709551615.0));   // u64 -> f64 rounds to nearest, like `as f64`
            weights[i] = __fsub_rn(__fmul_rn(unit, 2.0f), 1.0f);
        } else {
            const size_t j = i - nw;
            const uint64_t d = j % dim, proj = (j / dim) % pdim, rep = j / ((size_t)dim * pdim);
            signs[j] = (hash4(seed + 17ull, rep, proj, d) & 1ull) == 0ull ? 1 : -1;
        }
    }
}

// SimHash partition of every (repetition, vector): muvera.rs:111-131. part[rep][v].
__global__ void muvera_partition_kernel(const float* vecs, size_t nvec, uint32_t dim, uint32_t reps, uint32_t ks,
                                        const float* weights, uint32_t* part) {
    const size_t total = nvec * reps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i % nvec, rep = i / nvec;
        const float* x = vecs + v * dim;
        uint32_t p = 0;
        for (uint32_t proj = 0; proj < ks; ++proj) {
            const float* w = weights + ((size_t)rep * ks + proj) * dim;
            double dot = 0.0;
            for (uint32_t d = 0; d < dim; ++d) dot = __dadd_rn(dot, __dmul_rn((double)x[d], (double)w[d]));
            p = (p << 1) + (dot >= 0.0 ? 1u : 0u);
        }
        part[i] = p;
    }
}

// arrival[rep][v] = how many vectors of v's document, up to and including v, fell into v's partition:
// the `count` the reference passes to accumulate (muvera.rs:49-51).
__global__ void muvera_arrival_kernel(const uint32_t* part, const uint32_t* doc_of, const uint32_t* doc_first, size_t nvec,
                                      uint32_t reps, uint32_t* arrival) {
    const size_t total = nvec * reps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i % nvec, rep = i / nvec;
        const uint32_t* pr = part + rep * nvec;
        const uint32_t mine = pr[v];
        uint32_t c = 1;
        for (size_t u = doc_first[doc_of[v]]; u < v; ++u) c += pr[u] == mine ? 1u : 0u;
        arrival[i] = c;
    }
}

// muvera.rs:163-176; returns false on "encoding overflow"
__device__ __forceinline__ bool accumulate(float* slot, double value, int mode, uint32_t count) {
    const double current = (double)*slot;
    const double next = mode == 0 ? __dadd_rn(current, value)
                                  : __dadd_rn(current, __ddiv_rn(__dsub_rn(value, current), (double)count));
    if (isfinite(next) && next >= -3.4028234663852886e+38 && next <= 3.4028234663852886e+38) {
        *slot = (float)next;
        return true;
    }
    return false;
}

// One thread owns the slots [doc][rep][*][j] and walks the document's vectors in order (muvera.rs:45-63).
__global__ void muvera_accumulate_kernel(const float* vecs, const uint32_t* doc_first, uint32_t ndocs, size_t nvec,
                                         uint32_t dim, uint32_t reps, uint32_t partitions, uint32_t pdim,
                                         const signed char* signs, const uint32_t* part, const uint32_t* arrival, int mode,
                                         float* full, uint32_t* overflow) {
    const size_t total = (size_t)ndocs * reps * pdim;
    const size_t rep_size = (size_t)partitions * pdim, out_size = rep_size * reps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i % pdim), rep = (uint32_t)((i / pdim) % reps), doc = (uint32_t)(i / ((size_t)pdim * reps));
        float* base = full + (size_t)doc * out_size + (size_t)rep * rep_size + j;
        const signed char* sg = signs ? signs + ((size_t)rep * pdim + j) * dim : nullptr;
        for (size_t v = doc_first[doc]; v < doc_first[doc + 1]; ++v) {
            const float* x = vecs + v * dim;
            double value;
            if (!sg) {
                value = (double)x[j];                                           // identity projection (:142-147)
            } else {
                value = 0.0;
                for (uint32_t d = 0; d < dim; ++d) value = __dadd_rn(value, __dmul_rn((double)x[d], (double)sg[d]));
            }
            const uint32_t p = part[(size_t)rep * nvec + v];
            if (!accumulate(base + (size_t)p * pdim, value, mode, arrival[(size_t)rep * nvec + v])) *overflow = 1u;
        }
    }
}

// Count sketch (muvera.rs:179-196): final slot s sums, in input-index order, sign * full[index] over the indices
// that hash to it (CSR: idx[beg[s] .. beg[s+1]), sign in the top bit).
__global__ void muvera_sketch_kernel(const float* full, size_t out_size, const uint32_t* beg, const uint32_t* idx,
                                     uint32_t final_dim, uint32_t ndocs, float* out, uint32_t* overflow) {
    const size_t total = (size_t)ndocs * final_dim;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(i % final_dim), doc = (uint32_t)(i / final_dim);
        const float* in = full + (size_t)doc * out_size;
        float acc = 0.0f;
        for (uint32_t e = beg[s]; e < beg[s + 1]; ++e) {
            const uint32_t w = idx[e];
            const float sign = (w & 0x80000000u) ? -1.0f : 1.0f;
            const double next = __dadd_rn((double)acc, (double)__fmul_rn(sign, in[w & 0x7fffffffu]));
            if (!isfinite(next) || next < -3.4028234663852886e+38 || next > 3.4028234663852886e+38) { *overflow = 1u; break; }
            acc = (float)next;
        }
        out[i] = acc;
    }
}

bool mul_overflow(size_t a, size_t b, size_t* out) { return __builtin_mul_overflow(a, b, out); }

}  // namespace

Status muvera_output_dimension(const MuveraConfig& c, size_t* full_size, size_t* final_size) {
    // muvera.rs:77-108 (configuration part), :29-44
    if (c.dimension == 0) return Status::Ref("dimension must be positive");
    if (c.num_repetitions == 0) return Status::Ref("num_repetitions must be positive");
    if (c.num_simhash_projections >= 31) return Status::Ref("num_simhash_projections must be < 31");
    if (c.projection_dimension == 0) return Status::Ref("projection_dimension must be positive");
    if (c.has_final && c.final_projection_dimension == 0) return Status::Ref("final_projection_dimension must be positive");
    const size_t partitions = (size_t)1 << c.num_simhash_projections;
    size_t rep_size, out_size, counts;
    if (mul_overflow(partitions, c.projection_dimension, &rep_size)) return Status::Ref("fde dimension overflow");
    if (mul_overflow(c.num_repetitions, rep_size, &out_size)) return Status::Ref("fde dimension overflow");
    const size_t fin = c.has_final ? c.final_projection_dimension : out_size;
    if (out_size > kMaxOutputDimensions || fin > kMaxOutputDimensions) return Status::Ref("fde dimension exceeds safety limit");
    if (mul_overflow(c.num_repetitions, partitions, &counts)) return Status::Ref("fde dimension overflow");
    *full_size = out_size;
    *final_size = fin;
    return Status::Ok();
}

Status muvera_encode_batch(SearchCtx& ctx, const MuveraConfig& c, size_t ndocs, const float* vals, const uint64_t* vec_off,
                           const uint64_t* doc_vec, int mode, float* out, size_t out_capacity) {
    // muvera.rs:77-108 in the reference's order, document by document (each is one encode call there)
    if (ndocs == 0) return Status::Ok();
    for (size_t d = 0; d < ndocs; ++d) {
        if (doc_vec[d + 1] == doc_vec[d]) return Status::Ref("empty vectors");
        if (d == 0) {
            size_t a, b;
            Status s = muvera_output_dimension(c, &a, &b);
            if (!s.ok() && (s.msg == "dimension must be positive" || s.msg == "num_repetitions must be positive" ||
                            s.msg == "num_simhash_projections must be < 31" || s.msg == "projection_dimension must be positive" ||
                            s.msg == "final_projection_dimension must be positive"))
                return s;
        }
        for (size_t v = doc_vec[d]; v < doc_vec[d + 1]; ++v)
            if (vec_off[v + 1] - vec_off[v] != c.dimension) return Status::Ref("dimension mismatch");
        for (size_t v = doc_vec[d]; v < doc_vec[d + 1]; ++v)
            for (size_t i = vec_off[v]; i < vec_off[v + 1]; ++i)
                if (!std::isfinite(vals[i])) return Status::Ref("vector contains a non-finite value");
    }
    size_t out_size = 0, fin = 0;
    VB_TRY(muvera_output_dimension(c, &out_size, &fin));
    if (ndocs * fin > out_capacity) return Status::Cuda("muvera: output buffer too small");
    const size_t v0 = doc_vec[0], nvec = doc_vec[ndocs] - v0;
    const uint32_t dim = (uint32_t)c.dimension, reps = (uint32_t)c.num_repetitions, ks = (uint32_t)c.num_simhash_projections;
    const uint32_t pdim = (uint32_t)c.projection_dimension, partitions = 1u << ks;
    if (nvec >= 0x7fffffffull || ndocs >= 0x7fffffffull) return Status::Cuda("muvera: batch too large");
    const bool identity = c.projection_dimension == c.dimension;
    cudaStream_t st = ctx.stream;

    // ---- config-only tables
    const size_t nw = (size_t)reps * ks * dim, ns = identity ? 0 : (size_t)reps * pdim * dim;
    VB_TRY(ctx.staging_rank.reserve(std::max<size_t>(nw, 1) * sizeof(float) + ns + 64));
    float* d_weights = ctx.staging_rank.as<float>();
    signed char* d_signs = identity ? nullptr : reinterpret_cast<signed char*>(d_weights + std::max<size_t>(nw, 1));
    if (nw + ns > 0) {
        muvera_tables_kernel<<<(unsigned)std::min<size_t>((nw + ns + 255) / 256, 148 * 8), 256, 0, st>>>(c.seed, reps, ks, pdim, dim,
                                                                                                      d_weights, d_signs);
        VB_CUDA(cudaGetLastError());
    }
    // count sketch CSR (host: two hashes per input index)
    std::vector<uint32_t> h_beg, h_idx;
    if (c.has_final) {
        const size_t fd = c.final_projection_dimension;
        std::vector<uint32_t> slot(out_size);
        h_beg.assign(fd + 1, 0);
        for (size_t i = 0; i < out_size; ++i) {
            slot[i] = (uint32_t)(hash4(c.seed, 0x9E3779B97F4A7C15ull, i, 0) % fd);
            ++h_beg[slot[i] + 1];
        }
        for (size_t s = 0; s < fd; ++s) h_beg[s + 1] += h_beg[s];
        h_idx.resize(out_size);
        std::vector<uint32_t> cur(h_beg.begin(), h_beg.end() - 1);
        for (size_t i = 0; i < out_size; ++i) {   // ascending i inside every slot: the reference's accumulation order
            const uint32_t neg = (hash4(c.seed, 0xD1B54A32D192ED03ull, i, slot[i]) & 1ull) ? 0x80000000u : 0u;
            h_idx[cur[slot[i]]++] = (uint32_t)i | neg;
        }
    }

    // ---- documents in chunks that keep the full (pre-sketch) encodings within ~256 MB of HBM
    const size_t chunk_docs = std::max<size_t>(1, std::min<size_t>(ndocs, (64u << 20) / std::max<size_t>(out_size, 1)));
    PinnedBuf hb;
    for (size_t d0 = 0; d0 < ndocs; d0 += chunk_docs) {
        const size_t nd = std::min(chunk_docs, ndocs - d0);
        const size_t cv0 = doc_vec[d0], cn = doc_vec[d0 + nd] - cv0;
        // staging: vectors [cn][dim] | doc_first [nd + 1] | doc_of [cn]
        VB_TRY(hb.reserve(cn * dim * sizeof(float) + (nd + 1 + cn) * sizeof(uint32_t)));
        float* hv = hb.as<float>();
        uint32_t* hfirst = reinterpret_cast<uint32_t*>(hv + cn * dim);
        uint32_t* hdoc = hfirst + nd + 1;
        for (size_t v = 0; v < cn; ++v) std::memcpy(hv + v * dim, vals + vec_off[cv0 + v], dim * sizeof(float));
        for (size_t d = 0; d <= nd; ++d) hfirst[d] = (uint32_t)(doc_vec[d0 + d] - cv0);
        for (size_t d = 0; d < nd; ++d)
            for (uint32_t v = hfirst[d]; v < hfirst[d + 1]; ++v) hdoc[v] = (uint32_t)d;
        const size_t meta_bytes = (nd + 1 + cn) * sizeof(uint32_t);
        VB_TRY(ctx.staging.reserve(cn * dim * sizeof(float) + meta_bytes));
        float* d_vecs = ctx.staging.as<float>();
        uint32_t* d_first = reinterpret_cast<uint32_t*>(d_vecs + cn * dim);
        uint32_t* d_doc = d_first + nd + 1;
        VB_CUDA(cudaMemcpyAsync(d_vecs, hv, cn * dim * sizeof(float) + meta_bytes, cudaMemcpyHostToDevice, st));
        VB_TRY(ctx.dump_keys.reserve(2 * cn * reps * sizeof(uint32_t) + 64));
        uint32_t* d_part = ctx.dump_keys.as<uint32_t>();
        uint32_t* d_arr = d_part + cn * reps;
        VB_TRY(ctx.dump_pays.reserve(nd * out_size * sizeof(float)));
        float* d_full = ctx.dump_pays.as<float>();
        VB_TRY(ctx.misc.reserve(64));
        uint32_t* d_over = ctx.misc.as<uint32_t>() + 8;
        VB_CUDA(cudaMemsetAsync(d_full, 0, nd * out_size * sizeof(float), st));
        VB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(uint32_t), st));
        const auto blocks = [](size_t items) { return (unsigned)std::max<size_t>(1, std::min<size_t>((items + 127) / 128, 148 * 16)); };
        muvera_partition_kernel<<<blocks(cn * reps), 128, 0, st>>>(d_vecs, cn, dim, reps, ks, d_weights, d_part);
        VB_CUDA(cudaGetLastError());
        muvera_arrival_kernel<<<blocks(cn * reps), 128, 0, st>>>(d_part, d_doc, d_first, cn, reps, d_arr);
        VB_CUDA(cudaGetLastError());
        muvera_accumulate_kernel<<<blocks(nd * reps * pdim), 128, 0, st>>>(d_vecs, d_first, (uint32_t)nd, cn, dim, reps, partitions,
                                                                         pdim, d_signs, d_part, d_arr, mode, d_full, d_over);
        VB_CUDA(cudaGetLastError());
        const float* d_result = d_full;
        if (c.has_final) {
            const size_t fd = c.final_projection_dimension;
            VB_TRY(ctx.dump_keys2.reserve((fd + 1 + out_size) * sizeof(uint32_t)));
            uint32_t* d_beg = ctx.dump_keys2.as<uint32_t>();
            uint32_t* d_idx = d_beg + fd + 1;
            VB_CUDA(cudaMemcpyAsync(d_beg, h_beg.data(), (fd + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            VB_CUDA(cudaMemcpyAsync(d_idx, h_idx.data(), out_size * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            VB_TRY(ctx.dump_pays2.reserve(nd * fd * sizeof(float)));
            muvera_sketch_kernel<<<blocks(nd * fd), 128, 0, st>>>(d_full, out_size, d_beg, d_idx, (uint32_t)fd, (uint32_t)nd,
                                                                 ctx.dump_pays2.as<float>(), d_over);
            VB_CUDA(cudaGetLastError());
            d_result = ctx.dump_pays2.as<float>();
        }
        uint32_t h_over = 0;
        VB_CUDA(cudaMemcpyAsync(out + d0 * fin, d_result, nd * fin * sizeof(float), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(&h_over, d_over, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        if (h_over) { hb.release(); return Status::Ref("encoding overflow"); }   // muvera.rs:171-175, :190-192
    }
    hb.release();
    return Status::Ok();
}

}  // namespace vb
