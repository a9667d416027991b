// vettore_oracle.cpp — CPU restatement of Vettore's scan path. TEST INFRASTRUCTURE ONLY.
//
// This file is the parity checker for the CUDA path in vettore_b200/csrc and the
// timed CPU baseline of bench.py. Nothing under vettore_b200/ may import, link or
// call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do.
//
// Pinning status: the reference is Rust behind Rustler NIFs (no cargo/rustc/erl in
// this image), so it cannot be compiled or run here. The oracle is pinned against
// every known-answer test the reference holds for this path (tests/test_oracle_golden.py
// restates native/vettore/src/{distances,flat,search,multi_vector}.rs #[cfg(test)]
// and test/vector_algorithms_hardening_test.exs:90-121 etc.).
//
// Third-party arithmetic restated here: wide 1.5.0 `f32x8::reduce_add`
// (Cargo.lock:139-146; source not vendored). Restated as the AVX tree
// ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7)) — from the crate's published AVX path; the
// non-AVX targets use two f32x4 halves. The reference's own tests pin these kernels
// only to 2e-6 relative (distances.rs:570-609), so float parity is 1e-5, not bitwise.
//
// Each function cites the reference lines it follows (paths relative to
// /root/reference/native/vettore/src/).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#if defined(__AVX__)
#include <immintrin.h>
#endif

namespace {

thread_local std::string g_err;

int fail(const char* msg) {
    g_err = msg;
    return 1;
}

enum Metric : uint8_t {  // distances.rs:10-39
    L2 = 0, L2Squared = 1, Cosine = 2, InnerProduct = 3, NegativeInnerProduct = 4,
    Manhattan = 5, Chebyshev = 6, Hamming = 7, Jaccard = 8
};

// wide::f32x8::reduce_add, AVX association order (see header).
inline float reduce_add8(const float* l) {
    float s04 = l[0] + l[4], s15 = l[1] + l[5], s26 = l[2] + l[6], s37 = l[3] + l[7];
    return (s04 + s26) + (s15 + s37);
}

#if defined(__AVX__)
inline float reduce_add8(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v);
    __m128 hi = _mm256_extractf128_ps(v, 1);
    __m128 q = _mm_add_ps(lo, hi);              // (l0+l4, l1+l5, l2+l6, l3+l7)
    __m128 d = _mm_add_ps(q, _mm_movehl_ps(q, q));  // (s04+s26, s15+s37, ..)
    __m128 s = _mm_add_ss(d, _mm_shuffle_ps(d, d, 0x1));
    return _mm_cvtss_f32(s);
}
#endif

// distances.rs:236-270
float simd_dot(const float* a, const float* b, size_t n) {
    float acc = 0.0f;
    size_t i = 0;
#if defined(__AVX__)
    for (; i + 8 <= n; i += 8)
        acc += reduce_add8(_mm256_mul_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i)));
#else
    for (; i + 8 <= n; i += 8) {
        float p[8];
        for (int j = 0; j < 8; ++j) p[j] = a[i + j] * b[i + j];
        acc += reduce_add8(p);
    }
#endif
    for (; i < n; ++i) acc += a[i] * b[i];
    return acc;
}

// distances.rs:197-233
float simd_l2_squared(const float* a, const float* b, size_t n) {
    float acc = 0.0f;
    size_t i = 0;
#if defined(__AVX__)
    for (; i + 8 <= n; i += 8) {
        __m256 d = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        acc += reduce_add8(_mm256_mul_ps(d, d));
    }
#else
    for (; i + 8 <= n; i += 8) {
        float p[8];
        for (int j = 0; j < 8; ++j) { float d = a[i + j] - b[i + j]; p[j] = d * d; }
        acc += reduce_add8(p);
    }
#endif
    for (; i < n; ++i) { float d = a[i] - b[i]; acc += d * d; }
    return acc;
}

// distances.rs:273-308
float manhattan(const float* a, const float* b, size_t n) {
    float acc = 0.0f;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        float p[8];
        for (int j = 0; j < 8; ++j) p[j] = std::fabs(a[i + j] - b[i + j]);
        acc += reduce_add8(p);
    }
    for (; i < n; ++i) acc += std::fabs(a[i] - b[i]);
    return acc;
}

// Rust f32::max: returns the non-NaN operand when one is NaN.
inline float rust_max(float x, float y) { return std::fmax(x, y); }

// distances.rs:311-316
float chebyshev(const float* a, const float* b, size_t n) {
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) m = rust_max(m, std::fabs(a[i] - b[i]));
    return m;
}

// distances.rs:319-324
float hamming(const float* a, const float* b, size_t n) {
    size_t c = 0;
    for (size_t i = 0; i < n; ++i) c += ((a[i] != 0.0f) != (b[i] != 0.0f));
    return static_cast<float>(c);
}

// distances.rs:327-347
float jaccard(const float* a, const float* b, size_t n) {
    size_t inter = 0, uni = 0;
    for (size_t i = 0; i < n; ++i) {
        bool l = a[i] != 0.0f, r = b[i] != 0.0f;
        uni += (l || r);
        inter += (l && r);
    }
    if (uni == 0) return 0.0f;
    return 1.0f - static_cast<float>(inter) / static_cast<float>(uni);
}

// distances.rs:179-194
double f64_dot(const float* a, const float* b, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += static_cast<double>(a[i]) * static_cast<double>(b[i]);
    return s;
}
double f64_l2_squared(const float* a, const float* b, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double d = static_cast<double>(a[i]) - static_cast<double>(b[i]);
        s += d * d;
    }
    return s;
}

// distances.rs:140-147
float l2(const float* a, const float* b, size_t n) {
    float sq = simd_l2_squared(a, b, n);
    if (std::isfinite(sq)) return std::sqrt(sq);
    return static_cast<float>(std::sqrt(f64_l2_squared(a, b, n)));
}

// distances.rs:92-98
bool f64_to_f32(double v, float* out) {
    const double mx = static_cast<double>(std::numeric_limits<float>::max());
    if (std::isfinite(v) && v >= -mx && v <= mx) { *out = static_cast<float>(v); return true; }
    return false;
}

// distances.rs:70-90
bool recover_metric_overflow(uint8_t metric, const float* a, const float* b, size_t n, float* out) {
    double r;
    switch (metric) {
        case L2: r = std::sqrt(f64_l2_squared(a, b, n)); break;
        case L2Squared: r = f64_l2_squared(a, b, n); break;
        case Cosine: case InnerProduct: r = f64_dot(a, b, n); break;
        case NegativeInnerProduct: r = -f64_dot(a, b, n); break;
        case Manhattan: {
            r = 0.0;
            for (size_t i = 0; i < n; ++i) r += std::fabs(static_cast<double>(a[i]) - static_cast<double>(b[i]));
            break;
        }
        case Chebyshev: {
            r = 0.0;
            for (size_t i = 0; i < n; ++i) r = std::fmax(r, std::fabs(static_cast<double>(a[i]) - static_cast<double>(b[i])));
            break;
        }
        default: return false;
    }
    return f64_to_f32(r, out);
}

// distances.rs:42-68 (lengths already equal here; the length check is at the call sites)
int compute(uint8_t metric, const float* a, const float* b, size_t n, float* out) {
    float v;
    switch (metric) {
        case L2: v = l2(a, b, n); break;
        case L2Squared: v = simd_l2_squared(a, b, n); break;
        case Cosine: case InnerProduct: v = simd_dot(a, b, n); break;
        case NegativeInnerProduct: v = -simd_dot(a, b, n); break;
        case Manhattan: v = manhattan(a, b, n); break;
        case Chebyshev: v = chebyshev(a, b, n); break;
        case Hamming: v = hamming(a, b, n); break;
        case Jaccard: v = jaccard(a, b, n); break;
        default: return fail("unknown metric");
    }
    if (std::isfinite(v)) { *out = v; return 0; }
    if (recover_metric_overflow(metric, a, b, n, out)) return 0;
    return fail("metric overflow");
}

// distances.rs:160-177
int cosine(const float* a, const float* b, size_t n, float* out) {
    double ln = std::sqrt(f64_dot(a, a, n));
    double rn = std::sqrt(f64_dot(b, b, n));
    if (ln == 0.0 || rn == 0.0) { *out = 0.0f; return 0; }
    double s = f64_dot(a, b, n) / (ln * rn);
    if (!std::isfinite(s)) return fail("metric overflow");
    s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
    *out = static_cast<float>(s);
    return 0;
}

// distances.rs:113-128
inline float rank_value(uint8_t metric, float raw) {
    if (metric == Cosine) return 1.0f - raw;
    if (metric == InnerProduct) return -raw;
    return raw;
}
inline float similarity_value(uint8_t metric, float raw) {
    if (metric == Cosine || metric == InnerProduct) return raw;
    if (metric == NegativeInnerProduct) return -raw;
    return 1.0f / (1.0f + raw);
}

// distances.rs:131-137
bool all_finite(const float* v, size_t n) {
    for (size_t i = 0; i < n; ++i) if (!std::isfinite(v[i])) return false;
    return true;
}

// f32::total_cmp as an unsigned key (ascending).
inline uint32_t total_order_key(float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b ^ 0x80000000u);
}

struct IdView { const char* p; size_t n; };
inline int id_cmp(const IdView& a, const IdView& b) {  // Rust String::cmp = byte-lexicographic
    int c = std::memcmp(a.p, b.p, std::min(a.n, b.n));
    if (c != 0) return c;
    return a.n < b.n ? -1 : (a.n > b.n ? 1 : 0);
}

// flat.rs:19-46 / search.rs:8-35: (rank.total_cmp, id.cmp)
struct Hit {
    uint32_t key;   // total-order key of rank
    uint64_t idx;   // position in the caller's arrays
    float raw;
};

struct HitLess {
    const char* ids; const uint64_t* off;
    bool operator()(const Hit& a, const Hit& b) const {
        if (a.key != b.key) return a.key < b.key;
        IdView ia{ids + off[a.idx], static_cast<size_t>(off[a.idx + 1] - off[a.idx])};
        IdView ib{ids + off[b.idx], static_cast<size_t>(off[b.idx + 1] - off[b.idx])};
        return id_cmp(ia, ib) < 0;
    }
};

// Bounded max-heap of the best `limit` hits (flat.rs:103-118, search.rs:94-104).
class TopK {
  public:
    TopK(size_t limit, HitLess less) : limit_(limit), less_(less) {}
    void push(const Hit& h) {
        if (limit_ == 0) return;
        if (heap_.size() < limit_) {
            heap_.push_back(h);
            std::push_heap(heap_.begin(), heap_.end(), less_);
        } else if (less_(h, heap_.front())) {
            std::pop_heap(heap_.begin(), heap_.end(), less_);
            heap_.back() = h;
            std::push_heap(heap_.begin(), heap_.end(), less_);
        }
    }
    size_t finish(uint64_t* out_idx, float* out_raw) {  // flat.rs:120-123
        std::sort(heap_.begin(), heap_.end(), less_);
        for (size_t i = 0; i < heap_.size(); ++i) { out_idx[i] = heap_[i].idx; out_raw[i] = heap_[i].raw; }
        return heap_.size();
    }
  private:
    size_t limit_;
    HitLess less_;
    std::vector<Hit> heap_;
};

// distances.rs:472-481
inline uint64_t word_mask(size_t index, size_t dims) {
    size_t words = (dims + 63) / 64, rem = dims % 64;
    if (index + 1 == words && rem != 0) return (1ull << rem) - 1;
    return ~0ull;
}

// distances.rs:459-470
int validate_packed_pair(size_t ln, size_t rn, size_t dims) {
    size_t words = (dims + 63) / 64;
    if (dims == 0) return fail("dimensions must be positive");
    if (ln != words || rn != words) return fail("dimension mismatch");
    return 0;
}

// multi_vector.rs:144-152 over a ragged token list
int validate_vectors(const float* vals, const uint64_t* tok_off, size_t t0, size_t t1, size_t dim) {
    for (size_t t = t0; t < t1; ++t) {
        size_t len = tok_off[t + 1] - tok_off[t];
        if (len != dim) return fail("dimension mismatch");
        if (!all_finite(vals + tok_off[t], len)) return fail("vector contains a non-finite value");
    }
    return 0;
}
// multi_vector.rs:134-142
int validate_standalone(const float* vals, const uint64_t* tok_off, size_t t0, size_t t1) {
    if (t0 == t1) return 0;
    size_t first = tok_off[t0 + 1] - tok_off[t0];
    if (first == 0) return fail("vectors must not be empty");
    return validate_vectors(vals, tok_off, t0, t1, first);
}

// multi_vector.rs:65-87
int score_validated(const float* qv, const uint64_t* qoff, size_t nq,
                    const float* dv, const uint64_t* doff, size_t t0, size_t t1,
                    size_t dim, uint8_t metric, float* out) {
    float total = 0.0f;
    for (size_t q = 0; q < nq; ++q) {
        float best = -std::numeric_limits<float>::infinity();
        for (size_t t = t0; t < t1; ++t) {
            float raw;
            int rc = (metric == Cosine) ? cosine(qv + qoff[q], dv + doff[t], dim, &raw)
                                        : compute(metric, qv + qoff[q], dv + doff[t], dim, &raw);
            if (rc) return rc;
            best = rust_max(best, similarity_value(metric, raw));
        }
        total += best;
        if (!std::isfinite(total)) return fail("score overflow");
    }
    *out = total;
    return 0;
}

}  // namespace

extern "C" {

const char* vo_last_error() { return g_err.c_str(); }

// distances.rs:42-68 with the length check of :43-45; `checked` adds :101-105.
int vo_compute(uint8_t metric, const float* a, size_t na, const float* b, size_t nb, int checked, float* out) {
    if (metric > 8) return fail("unknown metric");
    if (checked) {
        if (!all_finite(a, na) || !all_finite(b, nb)) return fail("vector contains a non-finite value");
    }
    if (na != nb) return fail("dimension mismatch");
    return compute(metric, a, b, na, out);
}

// distances.rs:160-177
int vo_cosine(const float* a, size_t na, const float* b, size_t nb, float* out) {
    if (na != nb) return fail("dimension mismatch");
    return cosine(a, b, na, out);
}

float vo_rank_value(uint8_t metric, float raw) { return rank_value(metric, raw); }
float vo_similarity_value(uint8_t metric, float raw) { return similarity_value(metric, raw); }

// distances.rs:350-361
int vo_normalize_l2(const float* v, size_t n, float* out) {
    if (!all_finite(v, n)) return fail("vector contains a non-finite value");
    double norm = std::sqrt(f64_dot(v, v, n));
    for (size_t i = 0; i < n; ++i)
        out[i] = norm == 0.0 ? 0.0f : static_cast<float>(static_cast<double>(v[i]) / norm);
    return 0;
}

// distances.rs:413-423
void vo_compress_sign_bits(const float* v, size_t n, uint64_t* words) {
    size_t nw = (n + 63) / 64;
    for (size_t w = 0; w < nw; ++w) words[w] = 0;
    for (size_t i = 0; i < n; ++i)
        if (v[i] >= 0.0f) words[i / 64] |= 1ull << (i % 64);
}

// distances.rs:426-437
int vo_packed_hamming(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, size_t dims, float* out) {
    if (validate_packed_pair(na, nb, dims)) return 1;
    uint64_t d = 0;
    for (size_t i = 0; i < na; ++i) d += __builtin_popcountll((a[i] ^ b[i]) & word_mask(i, dims));
    *out = static_cast<float>(d);
    return 0;
}

// distances.rs:440-457
int vo_packed_jaccard(const uint64_t* a, size_t na, const uint64_t* b, size_t nb, size_t dims, float* out) {
    if (validate_packed_pair(na, nb, dims)) return 1;
    uint64_t inter = 0, uni = 0;
    for (size_t i = 0; i < na; ++i) {
        uint64_t m = word_mask(i, dims);
        inter += __builtin_popcountll((a[i] & b[i]) & m);
        uni += __builtin_popcountll((a[i] | b[i]) & m);
    }
    *out = uni == 0 ? 0.0f : 1.0f - static_cast<float>(inter) / static_cast<float>(uni);
    return 0;
}

// FlatIndex::search, flat.rs:96-124, over a dense snapshot of the index (rows are
// the HashMap's values in arbitrary order; order cannot matter, flat.rs:34-40).
// `index_dim` < 0 encodes dimension == None (empty index).
int vo_flat_search(uint8_t metric, const float* rows, size_t n, size_t d, long long index_dim,
                   const char* ids, const uint64_t* id_off,
                   const float* q, size_t qlen, size_t limit,
                   uint64_t* out_idx, float* out_raw, size_t* out_n) {
    *out_n = 0;
    if (metric > 8) return fail("unknown metric");
    if (limit == 0) return 0;                                             // flat.rs:97-99
    if (qlen == 0) return fail("vector must not be empty");              // flat.rs:136-144
    if (index_dim >= 0 && static_cast<size_t>(index_dim) != qlen) return fail("dimension mismatch");
    if (!all_finite(q, qlen)) return fail("vector contains a non-finite value");
    TopK top(limit, HitLess{ids, id_off});
    for (size_t r = 0; r < n; ++r) {
        float raw;
        if (compute(metric, q, rows + r * d, d, &raw)) return 1;          // flat.rs:105
        top.push(Hit{total_order_key(rank_value(metric, raw)), r, raw});
    }
    *out_n = top.finish(out_idx, out_raw);
    return 0;
}

// search::vector_top_k, search.rs:38-73. Vectors are ragged: row r = vals[voff[r]..voff[r+1]).
int vo_vector_top_k(const float* vals, const uint64_t* voff, size_t n,
                    const char* ids, const uint64_t* id_off,
                    const float* q, size_t qlen, int metric_code, size_t dims, size_t limit,
                    uint64_t* out_idx, float* out_raw, size_t* out_n) {
    *out_n = 0;
    if (metric_code < 0 || metric_code > 8) return fail("unknown metric");   // nifs.rs:160
    uint8_t metric = static_cast<uint8_t>(metric_code);
    if (dims == 0 || dims > qlen) return fail("invalid prefix dimensions");
    if (!all_finite(q, dims)) return fail("vector contains a non-finite value");
    TopK top(limit, HitLess{ids, id_off});
    for (size_t r = 0; r < n; ++r) {
        size_t len = voff[r + 1] - voff[r];
        const float* v = vals + voff[r];
        if (dims > len) return fail("dimension mismatch");
        if (!all_finite(v, dims)) return fail("vector contains a non-finite value");
        float raw;
        int rc = (metric == Cosine) ? cosine(q, v, dims, &raw) : compute(metric, q, v, dims, &raw);
        if (rc) return rc;
        top.push(Hit{total_order_key(rank_value(metric, raw)), r, raw});
    }
    *out_n = top.finish(out_idx, out_raw);
    return 0;
}

// search::binary_top_k, search.rs:76-92. Codes are ragged u64 word lists.
int vo_binary_top_k(const uint64_t* words, const uint64_t* woff, size_t n,
                    const char* ids, const uint64_t* id_off,
                    const uint64_t* q, size_t qwords, size_t dims, size_t limit,
                    uint64_t* out_idx, float* out_raw, size_t* out_n) {
    *out_n = 0;
    float self;
    if (vo_packed_hamming(q, qwords, q, qwords, dims, &self)) return 1;      // search.rs:84
    TopK top(limit, HitLess{ids, id_off});
    for (size_t r = 0; r < n; ++r) {
        float raw;
        if (vo_packed_hamming(q, qwords, words + woff[r], woff[r + 1] - woff[r], dims, &raw)) return 1;
        top.push(Hit{total_order_key(raw), r, raw});
    }
    *out_n = top.finish(out_idx, out_raw);
    return 0;
}

// multi_vector::score, multi_vector.rs:40-63. Tokens are ragged float lists.
int vo_multi_vector_score(const float* qv, const uint64_t* qoff, size_t nq,
                          const float* dv, const uint64_t* doff, size_t nd,
                          int metric_code, float* out) {
    if (metric_code < 0 || metric_code > 8) return fail("unknown metric");
    uint8_t metric = static_cast<uint8_t>(metric_code);
    if (nq == 0) {
        if (validate_standalone(dv, doff, 0, nd)) return 1;
        *out = 0.0f;
        return 0;
    }
    size_t dim = qoff[1] - qoff[0];
    if (dim == 0) return fail("vectors must not be empty");
    if (validate_vectors(qv, qoff, 0, nq, dim)) return 1;
    if (nd == 0) { *out = 0.0f; return 0; }
    if (validate_vectors(dv, doff, 0, nd, dim)) return 1;
    return score_validated(qv, qoff, nq, dv, doff, 0, nd, dim, metric, out);
}

// multi_vector::top_k, multi_vector.rs:90-132. Document i owns tokens
// [doc_tok[i], doc_tok[i+1]) of the ragged token list (dv, doff).
int vo_multi_vector_top_k(const float* dv, const uint64_t* doff, const uint64_t* doc_tok, size_t ndocs,
                          const char* ids, const uint64_t* id_off,
                          const float* qv, const uint64_t* qoff, size_t nq,
                          int metric_code, size_t limit,
                          uint64_t* out_idx, float* out_score, size_t* out_n) {
    *out_n = 0;
    if (metric_code < 0 || metric_code > 8) return fail("unknown metric");
    uint8_t metric = static_cast<uint8_t>(metric_code);
    if (validate_standalone(qv, qoff, 0, nq)) return 1;                     // :96
    bool has_dim = nq > 0;
    size_t dim = has_dim ? static_cast<size_t>(qoff[1] - qoff[0]) : 0;
    // Reverse ordering (multi_vector.rs:22-31): higher score first == ascending on the
    // complemented total-order key, then id ascending.
    TopK top(limit, HitLess{ids, id_off});
    for (size_t i = 0; i < ndocs; ++i) {
        size_t t0 = doc_tok[i], t1 = doc_tok[i + 1];
        float score = 0.0f;
        if (!has_dim) {
            if (validate_standalone(dv, doff, t0, t1)) return 1;
        } else if (t0 != t1) {
            if (validate_vectors(dv, doff, t0, t1, dim)) return 1;
            if (score_validated(qv, qoff, nq, dv, doff, t0, t1, dim, metric, &score)) return 1;
        }
        top.push(Hit{~total_order_key(score), i, score});
    }
    *out_n = top.finish(out_idx, out_score);
    return 0;
}

// Timed CPU baseline: `nq` flat searches over a dense [n, d] snapshot, `threads` host
// threads each running whole queries sequentially (one query = one sequential scan,
// exactly the unit of work of one dirty-scheduler flat_search call, nifs.rs:297-309;
// several BEAM dirty schedulers may run such calls concurrently). Ids are implicit
// zero-padded row numbers, so id order == row order and the id comparison is on idx.
// Returns wall seconds; writes every query's top-`limit` ([nq][min(limit, n)]) for parity checks.
double vo_flat_scan_timed(uint8_t metric, const float* rows, size_t n, size_t d,
                          const float* queries, size_t nq, size_t limit, int threads,
                          uint64_t* out_idx, float* out_raw) {
    if (threads < 1) threads = 1;
    std::atomic<size_t> next{0};
    auto worker = [&](int tid) {
        struct H { uint32_t key; uint64_t idx; float raw; };
        auto less = [](const H& a, const H& b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; };
        std::vector<H> heap;
        for (;;) {
            size_t qi = next.fetch_add(1);
            if (qi >= nq) break;
            const float* q = queries + qi * d;
            heap.clear();
            for (size_t r = 0; r < n; ++r) {
                float raw;
                if (compute(metric, q, rows + r * d, d, &raw)) return;
                H h{total_order_key(rank_value(metric, raw)), r, raw};
                if (heap.size() < limit) {
                    heap.push_back(h);
                    std::push_heap(heap.begin(), heap.end(), less);
                } else if (less(h, heap.front())) {
                    std::pop_heap(heap.begin(), heap.end(), less);
                    heap.back() = h;
                    std::push_heap(heap.begin(), heap.end(), less);
                }
            }
            std::sort(heap.begin(), heap.end(), less);
            if (out_idx && out_raw) {   // every query's hits: [nq][min(limit, n)] (the caller checks parity with them)
                const size_t cap = limit < n ? limit : n;
                for (size_t i = 0; i < heap.size(); ++i) {
                    out_idx[qi * cap + i] = heap[i].idx;
                    out_raw[qi * cap + i] = heap[i].raw;
                }
            }
            (void)tid;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Timed CPU baseline of the Hamming candidate pass (search::binary_top_k, search.rs:76-92) over a dense
// [n, nw] u64 code matrix: `nq` queries, `threads` host threads each running whole queries (one query =
// one sequential scan, as in the reference). Ids are implicit zero-padded row numbers (id order == row
// order). Writes every query's top-`limit` ([nq][min(limit, n)]). Returns wall seconds.
double vo_binary_scan_timed(const uint64_t* codes, size_t n, size_t nw, size_t dims, const uint64_t* queries,
                            size_t nq, size_t limit, int threads, uint64_t* out_idx, float* out_raw) {
    if (threads < 1) threads = 1;
    std::atomic<size_t> next{0};
    const size_t cap = limit < n ? limit : n;
    auto worker = [&]() {
        struct H { uint32_t key; uint64_t idx; };
        auto less = [](const H& a, const H& b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; };
        std::vector<H> heap;
        for (;;) {
            size_t qi = next.fetch_add(1);
            if (qi >= nq) break;
            const uint64_t* q = queries + qi * nw;
            heap.clear();
            for (size_t r = 0; r < n; ++r) {
                const uint64_t* c = codes + r * nw;
                uint64_t d = 0;
                for (size_t w = 0; w < nw; ++w) d += __builtin_popcountll((q[w] ^ c[w]) & word_mask(w, dims));
                H h{static_cast<uint32_t>(d), r};   // distances are small non-negative integers: key order == f32 order
                if (heap.size() < limit) {
                    heap.push_back(h);
                    std::push_heap(heap.begin(), heap.end(), less);
                } else if (less(h, heap.front())) {
                    std::pop_heap(heap.begin(), heap.end(), less);
                    heap.back() = h;
                    std::push_heap(heap.begin(), heap.end(), less);
                }
            }
            std::sort(heap.begin(), heap.end(), less);
            if (out_idx && out_raw)
                for (size_t i = 0; i < heap.size(); ++i) {
                    out_idx[qi * cap + i] = heap[i].idx;
                    out_raw[qi * cap + i] = static_cast<float>(heap[i].key);
                }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Timed CPU baseline of MaxSim top-k (multi_vector::top_k, multi_vector.rs:90-132) over uniform documents
// stored densely as [ndocs, td, dim]: `nq` queries of `tq` tokens each, `threads` host threads each running
// whole queries. Same scoring as score_validated (f64 true cosine when metric == Cosine). Writes every
// query's top-`limit` (score desc, doc index asc). Returns wall seconds, or -1 on a scoring error.
double vo_maxsim_scan_timed(const float* tokens, size_t ndocs, size_t td, size_t dim, const float* queries, size_t nq,
                            size_t tq, int metric_code, size_t limit, int threads, uint64_t* out_idx, float* out_score) {
    if (threads < 1) threads = 1;
    const uint8_t metric = static_cast<uint8_t>(metric_code);
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    const size_t cap = limit < ndocs ? limit : ndocs;
    std::vector<uint64_t> qoff(tq + 1), doff(td + 1);
    for (size_t i = 0; i <= tq; ++i) qoff[i] = i * dim;
    for (size_t i = 0; i <= td; ++i) doff[i] = i * dim;
    auto worker = [&]() {
        struct H { uint32_t key; uint64_t idx; float score; };
        auto less = [](const H& a, const H& b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; };
        std::vector<H> heap;
        for (;;) {
            size_t qi = next.fetch_add(1);
            if (qi >= nq) break;
            const float* q = queries + qi * tq * dim;
            heap.clear();
            for (size_t d = 0; d < ndocs; ++d) {
                float score = 0.0f;
                if (score_validated(q, qoff.data(), tq, tokens + d * td * dim, doff.data(), 0, td, dim, metric, &score)) {
                    failed = 1;
                    return;
                }
                H h{~total_order_key(score), d, score};
                if (heap.size() < limit) {
                    heap.push_back(h);
                    std::push_heap(heap.begin(), heap.end(), less);
                } else if (less(h, heap.front())) {
                    std::pop_heap(heap.begin(), heap.end(), less);
                    heap.back() = h;
                    std::push_heap(heap.begin(), heap.end(), less);
                }
            }
            std::sort(heap.begin(), heap.end(), less);
            if (out_idx && out_score)
                for (size_t i = 0; i < heap.size(); ++i) {
                    out_idx[qi * cap + i] = heap[i].idx;
                    out_score[qi * cap + i] = heap[i].score;
                }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    return failed ? -1.0 : std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

// ---- MUVERA fixed-dimensional encoding: muvera.rs:26-74 and its helpers -------------
namespace {
// muvera.rs:215-221
inline uint64_t mv_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t mv_hash4(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t x = a ^ mv_rotl(b, 17) ^ mv_rotl(c, 31) ^ mv_rotl(d, 47);
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// muvera.rs:199-203: (hash as f64 / u64::MAX as f64) as f32, then * 2 - 1 in f32
inline float mv_random_weight(uint64_t seed, uint64_t rep, uint64_t proj, uint64_t dim) {
    const uint64_t h = mv_hash4(seed, rep, proj, dim);
    const float unit = static_cast<float>(static_cast<double>(h) / static_cast<double>(UINT64_MAX));
    return unit * 2.0f - 1.0f;
}
// muvera.rs:206-212
inline float mv_random_sign(uint64_t seed, uint64_t rep, uint64_t proj, uint64_t dim) {
    return (mv_hash4(seed, rep, proj, dim) & 1) == 0 ? 1.0f : -1.0f;
}
// muvera.rs:163-176
inline int mv_accumulate(float* slot, double value, int mode, size_t count) {
    const double current = static_cast<double>(*slot);
    const double next = mode == 0 ? current + value : current + (value - current) / static_cast<double>(count);
    if (std::isfinite(next) && next >= -static_cast<double>(std::numeric_limits<float>::max()) &&
        next <= static_cast<double>(std::numeric_limits<float>::max())) {
        *slot = static_cast<float>(next);
        return 0;
    }
    return fail("encoding overflow");
}
}  // namespace

extern "C" {

// muvera.rs:26-74 (encode) incl. validate (:77-108) and count_sketch (:179-196). `vectors` is a ragged list
// (vals, off[nvec + 1]); has_final = final_projection_dimension is Some(final_dim); mode 0 = Query (sum),
// 1 = Document (running average). Writes *out_len floats to out (capacity out_cap).
int vo_muvera_encode(const float* vals, const uint64_t* off, size_t nvec, size_t dimension, size_t num_repetitions,
                     size_t num_simhash, uint64_t seed, size_t projection_dimension, int has_final, size_t final_dim,
                     int mode, float* out, size_t out_cap, size_t* out_len) {
    *out_len = 0;
    const size_t kMaxOut = 16777216;
    if (nvec == 0) return fail("empty vectors");
    if (dimension == 0) return fail("dimension must be positive");
    if (num_repetitions == 0) return fail("num_repetitions must be positive");
    if (num_simhash >= 31) return fail("num_simhash_projections must be < 31");
    if (projection_dimension == 0) return fail("projection_dimension must be positive");
    if (has_final && final_dim == 0) return fail("final_projection_dimension must be positive");
    for (size_t v = 0; v < nvec; ++v)
        if (off[v + 1] - off[v] != dimension) return fail("dimension mismatch");
    for (size_t v = 0; v < nvec; ++v)
        if (!all_finite(vals + off[v], dimension)) return fail("vector contains a non-finite value");
    const size_t partitions = static_cast<size_t>(1) << num_simhash;
    size_t repetition_size, output_size, counts_size;
    if (__builtin_mul_overflow(partitions, projection_dimension, &repetition_size)) return fail("fde dimension overflow");
    if (__builtin_mul_overflow(num_repetitions, repetition_size, &output_size)) return fail("fde dimension overflow");
    const size_t final_size = has_final ? final_dim : output_size;
    if (output_size > kMaxOut || final_size > kMaxOut) return fail("fde dimension exceeds safety limit");
    if (__builtin_mul_overflow(num_repetitions, partitions, &counts_size)) return fail("fde dimension overflow");
    if (final_size > out_cap) return fail("output buffer too small");
    std::vector<float> full(output_size, 0.0f);
    std::vector<size_t> counts(counts_size, 0);
    for (size_t rep = 0; rep < num_repetitions; ++rep) {
        for (size_t v = 0; v < nvec; ++v) {
            const float* vec = vals + off[v];
            size_t partition = 0;                                             // muvera.rs:111-131
            for (size_t proj = 0; proj < num_simhash; ++proj) {
                double dot = 0.0;
                for (size_t d = 0; d < dimension; ++d)
                    dot += static_cast<double>(vec[d]) * static_cast<double>(mv_random_weight(seed, rep, proj, d));
                partition = (partition << 1) + (dot >= 0.0 ? 1 : 0);
            }
            const size_t ci = rep * partitions + partition;
            counts[ci] += 1;
            const size_t base = rep * repetition_size + partition * projection_dimension;
            if (projection_dimension == dimension) {                           // muvera.rs:142-147
                for (size_t d = 0; d < dimension; ++d)
                    if (mv_accumulate(&full[base + d], static_cast<double>(vec[d]), mode, counts[ci])) return 1;
            } else {
                for (size_t proj = 0; proj < projection_dimension; ++proj) {   // :149-160
                    double value = 0.0;
                    for (size_t d = 0; d < dimension; ++d)
                        value += static_cast<double>(vec[d]) * static_cast<double>(mv_random_sign(seed + 17, rep, proj, d));
                    if (mv_accumulate(&full[base + proj], value, mode, counts[ci])) return 1;
                }
            }
        }
    }
    if (!has_final) {
        std::memcpy(out, full.data(), output_size * sizeof(float));
        *out_len = output_size;
        return 0;
    }
    std::vector<float> fin(final_dim, 0.0f);                                    // muvera.rs:179-196
    for (size_t i = 0; i < output_size; ++i) {
        const size_t slot = static_cast<size_t>(mv_hash4(seed, 0x9E3779B97F4A7C15ull, i, 0)) % final_dim;
        const float sign = (mv_hash4(seed, 0xD1B54A32D192ED03ull, i, slot) & 1) == 0 ? 1.0f : -1.0f;
        const double next = static_cast<double>(fin[slot]) + static_cast<double>(sign * full[i]);
        if (!std::isfinite(next) || next < -static_cast<double>(std::numeric_limits<float>::max()) ||
            next > static_cast<double>(std::numeric_limits<float>::max()))
            return fail("encoding overflow");
        fin[slot] = static_cast<float>(next);
    }
    std::memcpy(out, fin.data(), final_dim * sizeof(float));
    *out_len = final_dim;
    return 0;
}

uint64_t vo_muvera_hash4(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { return mv_hash4(a, b, c, d); }

}  // extern "C"
