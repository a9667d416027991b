/* vettore_b200_nif.c — erl_nif shim that maps the scan-path functions of `Vettore.Nifs`
 * (reference lib/vettore_nifs.ex:70-172; Rust side native/vettore/src/nifs.rs:125-309) onto the C ABI of
 * include/vettore_b200.h. Module name: Elixir.Vettore.B200.Nifs (lib/vettore/b200/nifs.ex).
 *
 * Build next to the Rust crate (needs the OTP headers, which this repository's image does not have):
 *   cc -O2 -fPIC -shared -I"$ERL_INCLUDE" -Iinclude nif/vettore_b200_nif.c \
 *      -Lvettore_b200 -lvettore_b200 -o priv/native/libvettore_b200_nif.so
 * In this repository the file is compiled against tests/mock_erl/erl_nif.h (the documented OTP signatures)
 * and driven term-in / term-out by tests/test_nif_shim.py (CPU: the (name, arity) table, argument decoding,
 * badarg and validation errors; GPU: every function end to end against the oracle).
 *
 * Conventions reproduced from Rustler (nifs.rs): every function is a dirty CPU-bound NIF
 * (schedule = "DirtyCpu"); Result<T, String> -> {:ok, T} | {:error, binary}; Result<(), String> -> {:ok, {}};
 * flat_new_* return the bare resource; floats arrive as f64 (integers are accepted like Rustler's f32
 * decoder accepts them) and are narrowed to f32; u64 words may be bignums; mistyped terms -> badarg.
 * HNSW and the pairwise helpers stay in the Rust library behind `Vettore.Nifs`.
 */
#include <erl_nif.h>
#include <stdlib.h>
#include <string.h>

#include "vettore_b200.h"

static ErlNifResourceType* FLAT_TYPE;
static ErlNifResourceType* MV_TYPE;

typedef struct { vb_flat* index; } flat_res;
typedef struct { vb_mv* index; } mv_res;

static void flat_dtor(ErlNifEnv* env, void* obj) { (void)env; vb_flat_free(((flat_res*)obj)->index); }
static void mv_dtor(ErlNifEnv* env, void* obj) { (void)env; vb_mv_free(((mv_res*)obj)->index); }

/* ---- term construction ---------------------------------------------------------------------------- */
static ERL_NIF_TERM mk_binary(ErlNifEnv* env, const char* p, size_t n) {
    ERL_NIF_TERM bin;
    unsigned char* dst = enif_make_new_binary(env, n, &bin);
    if (n) memcpy(dst, p, n);
    return bin;
}
static ERL_NIF_TERM mk_error(ErlNifEnv* env) {           /* {:error, "message"} */
    const char* msg = vb_last_error();
    return enif_make_tuple2(env, enif_make_atom(env, "error"), mk_binary(env, msg, strlen(msg)));
}
static ERL_NIF_TERM mk_ok(ErlNifEnv* env, ERL_NIF_TERM value) { return enif_make_tuple2(env, enif_make_atom(env, "ok"), value); }
static ERL_NIF_TERM mk_ok_unit(ErlNifEnv* env) { return mk_ok(env, enif_make_tuple(env, 0)); }   /* Result<(), String> */
static ERL_NIF_TERM status_unit(ErlNifEnv* env, int rc) { return rc == VB_OK ? mk_ok_unit(env) : mk_error(env); }

/* vb_hits -> [{id_binary, float}] (consumes the hits) */
static ERL_NIF_TERM hits_to_term(ErlNifEnv* env, vb_hits* h) {
    const char* blob; const uint64_t* off; const float* val; const uint64_t* idx;
    size_t n = vb_hits_export(h, &blob, &off, &val, &idx);
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (size_t i = n; i-- > 0;) {
        ERL_NIF_TERM id = mk_binary(env, blob + off[i], (size_t)(off[i + 1] - off[i]));
        list = enif_make_list_cell(env, enif_make_tuple2(env, id, enif_make_double(env, val[i])), list);
    }
    vb_hits_free(h);
    return list;
}
static ERL_NIF_TERM hits_result(ErlNifEnv* env, int rc, vb_hits* hits) {
    return rc == VB_OK ? mk_ok(env, hits_to_term(env, hits)) : mk_error(env);
}

/* ---- growable arrays for the decoders ---------------------------------------------------------------- */
typedef struct { float* v; size_t n, cap; } fbuf;
typedef struct { uint64_t* v; size_t n, cap; } ubuf;
typedef struct { char* v; size_t n, cap; } cbuf;

static int fbuf_room(fbuf* b, size_t more) {
    if (b->n + more <= b->cap) return 1;
    size_t cap = b->cap ? b->cap : 256;
    while (cap < b->n + more) cap *= 2;
    float* p = (float*)realloc(b->v, cap * sizeof(float));
    if (!p) return 0;
    b->v = p; b->cap = cap;
    return 1;
}
static int ubuf_push(ubuf* b, uint64_t x) {
    if (b->n == b->cap) {
        size_t cap = b->cap ? 2 * b->cap : 64;
        uint64_t* p = (uint64_t*)realloc(b->v, cap * sizeof(uint64_t));
        if (!p) return 0;
        b->v = p; b->cap = cap;
    }
    b->v[b->n++] = x;
    return 1;
}
static int cbuf_append(cbuf* b, const unsigned char* p, size_t n) {
    if (b->n + n > b->cap) {
        size_t cap = b->cap ? b->cap : 256;
        while (cap < b->n + n) cap *= 2;
        char* q = (char*)realloc(b->v, cap);
        if (!q) return 0;
        b->v = q; b->cap = cap;
    }
    if (n) memcpy(b->v + b->n, p, n);
    b->n += n;
    return 1;
}

/* [number] appended to `out` (f64 -> f32 like Rustler's Vec<f32> decoder; integers accepted). */
static int append_floats(ErlNifEnv* env, ERL_NIF_TERM list, fbuf* out) {
    unsigned len;
    if (!enif_get_list_length(env, list, &len) || !fbuf_room(out, len)) return 0;
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; i < len; ++i) {
        double d; long l;
        enif_get_list_cell(env, tail, &head, &tail);
        if (enif_get_double(env, head, &d)) out->v[out->n++] = (float)d;
        else if (enif_get_long(env, head, &l)) out->v[out->n++] = (float)l;
        else return 0;
    }
    return 1;
}
/* [non_neg_integer] appended to `out` (u64 words; values >= 2^60 are bignums on the BEAM). */
static int append_words(ErlNifEnv* env, ERL_NIF_TERM list, ubuf* out) {
    unsigned len;
    if (!enif_get_list_length(env, list, &len)) return 0;
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; i < len; ++i) {
        ErlNifUInt64 w;
        enif_get_list_cell(env, tail, &head, &tail);
        if (!enif_get_uint64(env, head, &w) || !ubuf_push(out, (uint64_t)w)) return 0;
    }
    return 1;
}

/* A decoded by-value batch: ids as blob + n+1 offsets, one ragged payload per id. */
typedef struct {
    size_t n;
    cbuf ids; ubuf id_off;
    fbuf vals; ubuf val_off;      /* float payloads: row i = vals[val_off[i] .. val_off[i+1]) */
    ubuf words; ubuf word_off;    /* u64 payloads */
    ubuf doc_tok;                 /* documents: doc i owns token vectors [doc_tok[i], doc_tok[i+1]) of (vals, val_off) */
} batch;

static void batch_free(batch* b) {
    free(b->ids.v); free(b->id_off.v); free(b->vals.v); free(b->val_off.v);
    free(b->words.v); free(b->word_off.v); free(b->doc_tok.v);
    memset(b, 0, sizeof(*b));
}
/* Non-NULL pointers even for empty batches (the C ABI takes plain pointers). */
static const char* b_ids(const batch* b) { return b->ids.v ? b->ids.v : ""; }
static const float* b_vals(const batch* b) { static const float z = 0.0f; return b->vals.v ? b->vals.v : &z; }
static const uint64_t* b_words(const batch* b) { static const uint64_t z = 0; return b->words.v ? b->words.v : &z; }

enum { PAYLOAD_FLOATS, PAYLOAD_WORDS, PAYLOAD_VECTORS };

/* [{id_binary, payload}] -> batch. payload: [float] | [u64] | [[float]] */
static int decode_batch(ErlNifEnv* env, ERL_NIF_TERM list, int kind, batch* b) {
    unsigned n;
    memset(b, 0, sizeof(*b));
    if (!enif_get_list_length(env, list, &n)) return 0;
    b->n = n;
    int ok = ubuf_push(&b->id_off, 0) && ubuf_push(&b->val_off, 0) && ubuf_push(&b->word_off, 0) && ubuf_push(&b->doc_tok, 0);
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; ok && i < n; ++i) {
        const ERL_NIF_TERM* tup; int arity; ErlNifBinary id;
        enif_get_list_cell(env, tail, &head, &tail);
        ok = enif_get_tuple(env, head, &arity, &tup) && arity == 2 && enif_inspect_binary(env, tup[0], &id) &&
             cbuf_append(&b->ids, id.data, id.size) && ubuf_push(&b->id_off, b->ids.n);
        if (!ok) break;
        if (kind == PAYLOAD_FLOATS) {
            ok = append_floats(env, tup[1], &b->vals) && ubuf_push(&b->val_off, b->vals.n);
        } else if (kind == PAYLOAD_WORDS) {
            ok = append_words(env, tup[1], &b->words) && ubuf_push(&b->word_off, b->words.n);
        } else {
            unsigned nv;
            ERL_NIF_TERM vh, vt = tup[1];
            ok = enif_get_list_length(env, tup[1], &nv);
            for (unsigned j = 0; ok && j < nv; ++j) {
                enif_get_list_cell(env, vt, &vh, &vt);
                ok = append_floats(env, vh, &b->vals) && ubuf_push(&b->val_off, b->vals.n);
            }
            ok = ok && ubuf_push(&b->doc_tok, b->val_off.n - 1);
        }
    }
    if (!ok) batch_free(b);
    return ok;
}

/* [[float]] -> ragged vectors (vals, off) */
typedef struct { fbuf vals; ubuf off; size_t n; } vecs;
static void vecs_free(vecs* v) { free(v->vals.v); free(v->off.v); memset(v, 0, sizeof(*v)); }
static const float* v_vals(const vecs* v) { static const float z = 0.0f; return v->vals.v ? v->vals.v : &z; }
static int decode_vectors(ErlNifEnv* env, ERL_NIF_TERM list, vecs* out) {
    unsigned n;
    memset(out, 0, sizeof(*out));
    if (!enif_get_list_length(env, list, &n)) return 0;
    out->n = n;
    int ok = ubuf_push(&out->off, 0);
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; ok && i < n; ++i) {
        enif_get_list_cell(env, tail, &head, &tail);
        ok = append_floats(env, head, &out->vals) && ubuf_push(&out->off, out->vals.n);
    }
    if (!ok) vecs_free(out);
    return ok;
}

static int get_size(ErlNifEnv* env, ERL_NIF_TERM t, size_t* out) {   /* usize arguments */
    ErlNifUInt64 v;
    if (!enif_get_uint64(env, t, &v)) return 0;
    *out = (size_t)v;
    return 1;
}
static int get_metric(ErlNifEnv* env, ERL_NIF_TERM t, int* out) {    /* u8 metric codes: 256.. is badarg, 9..255 "unknown metric" */
    ErlNifUInt64 v;
    if (!enif_get_uint64(env, t, &v) || v > 255) return 0;
    *out = (int)v;
    return 1;
}

/* ---- by-value NIFs: lib/vettore_nifs.ex:70-119 -------------------------------------------------------- */
/* compress_sign_bits/1 (nifs.rs:125-129): bare list of u64 words */
static ERL_NIF_TERM compress_sign_bits(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    fbuf v = {0};
    (void)argc;
    if (!append_floats(env, argv[0], &v)) { free(v.v); return enif_make_badarg(env); }
    size_t nw = (v.n + 63) / 64;
    uint64_t* words = (uint64_t*)calloc(nw ? nw : 1, sizeof(uint64_t));
    static const float z = 0.0f;
    vb_compress_sign_bits(v.v ? v.v : &z, v.n, words);
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (size_t i = nw; i-- > 0;) list = enif_make_list_cell(env, enif_make_uint64(env, words[i]), list);
    free(words); free(v.v);
    return list;
}

/* vector_top_k/5 (nifs.rs:151-162): vectors, query, metric_code, dimensions, limit */
static ERL_NIF_TERM vector_top_k(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    batch b; fbuf q = {0}; int metric; size_t dims, limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_metric(env, argv[2], &metric) || !get_size(env, argv[3], &dims) || !get_size(env, argv[4], &limit) ||
        !append_floats(env, argv[1], &q)) { free(q.v); return enif_make_badarg(env); }
    if (!decode_batch(env, argv[0], PAYLOAD_FLOATS, &b)) { free(q.v); return enif_make_badarg(env); }
    static const float z = 0.0f;
    int rc = vb_vector_top_k(b.n, b_ids(&b), b.id_off.v, b_vals(&b), b.val_off.v, q.v ? q.v : &z, q.n, metric, dims, limit, &hits);
    batch_free(&b); free(q.v);
    return hits_result(env, rc, hits);
}

/* binary_top_k/4 (nifs.rs:164-175): vectors, query words, dimensions, limit */
static ERL_NIF_TERM binary_top_k(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    batch b; ubuf q = {0}; size_t dims, limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_size(env, argv[2], &dims) || !get_size(env, argv[3], &limit) || !append_words(env, argv[1], &q)) {
        free(q.v); return enif_make_badarg(env);
    }
    if (!decode_batch(env, argv[0], PAYLOAD_WORDS, &b)) { free(q.v); return enif_make_badarg(env); }
    static const uint64_t z = 0;
    int rc = vb_binary_top_k(b.n, b_ids(&b), b.id_off.v, b_words(&b), b.word_off.v, q.v ? q.v : &z, q.n, dims, limit, &hits);
    batch_free(&b); free(q.v);
    return hits_result(env, rc, hits);
}

/* multi_vector_score/3 (nifs.rs:177-186): query vectors, document vectors, metric_code -> {:ok, float} */
static ERL_NIF_TERM multi_vector_score(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    vecs q, d; int metric; float score = 0.0f;
    (void)argc;
    if (!get_metric(env, argv[2], &metric) || !decode_vectors(env, argv[0], &q)) return enif_make_badarg(env);
    if (!decode_vectors(env, argv[1], &d)) { vecs_free(&q); return enif_make_badarg(env); }
    int rc = vb_multi_vector_score(v_vals(&q), q.off.v, q.n, v_vals(&d), d.off.v, d.n, metric, &score);
    vecs_free(&q); vecs_free(&d);
    return rc == VB_OK ? mk_ok(env, enif_make_double(env, score)) : mk_error(env);
}

/* multi_vector_top_k/4 (nifs.rs:188-198): documents, query vectors, metric_code, limit */
static ERL_NIF_TERM multi_vector_top_k(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    batch b; vecs q; int metric; size_t limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_metric(env, argv[2], &metric) || !get_size(env, argv[3], &limit) || !decode_vectors(env, argv[1], &q))
        return enif_make_badarg(env);
    if (!decode_batch(env, argv[0], PAYLOAD_VECTORS, &b)) { vecs_free(&q); return enif_make_badarg(env); }
    int rc = vb_multi_vector_top_k(b.n, b_ids(&b), b.id_off.v, b_vals(&b), b.val_off.v, b.doc_tok.v, v_vals(&q), q.off.v, q.n,
                                   metric, limit, &hits);
    batch_free(&b); vecs_free(&q);
    return hits_result(env, rc, hits);
}

/* ---- resident flat index: lib/vettore_nifs.ex:122-172 -------------------------------------------------- */
/* flat_new_<metric>/0: returns the bare resource (nifs.rs:200-257). A missing device raises: the reference
 * constructors cannot fail, and an index that silently does nothing would be worse. */
static ERL_NIF_TERM flat_new(ErlNifEnv* env, int metric) {
    vb_flat* index = NULL;
    /* VETTORE_B200_GPUS=N (N > 1): the same resource spread over N GPUs of the box (vb_flat_new_sharded);
     * the callers cannot tell the difference. */
    const char* gpus = getenv("VETTORE_B200_GPUS");
    const int n_gpus = gpus ? atoi(gpus) : 1;
    const int rc = n_gpus > 1 ? vb_flat_new_sharded(metric, n_gpus, NULL, &index) : vb_flat_new(metric, &index);
    if (rc != VB_OK) return enif_raise_exception(env, mk_error(env));
    flat_res* r = (flat_res*)enif_alloc_resource(FLAT_TYPE, sizeof(flat_res));
    r->index = index;
    ERL_NIF_TERM t = enif_make_resource(env, r);
    enif_release_resource(r);
    return t;
}
#define FLAT_NEW(name, code) \
    static ERL_NIF_TERM name(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) { (void)argc; (void)argv; return flat_new(env, code); }
FLAT_NEW(flat_new_l2, VB_METRIC_L2)
FLAT_NEW(flat_new_l2_squared, VB_METRIC_L2_SQUARED)
FLAT_NEW(flat_new_cosine, VB_METRIC_COSINE)
FLAT_NEW(flat_new_inner_product, VB_METRIC_INNER_PRODUCT)
FLAT_NEW(flat_new_negative_inner_product, VB_METRIC_NEGATIVE_INNER_PRODUCT)
FLAT_NEW(flat_new_manhattan, VB_METRIC_MANHATTAN)
FLAT_NEW(flat_new_chebyshev, VB_METRIC_CHEBYSHEV)
FLAT_NEW(flat_new_hamming, VB_METRIC_HAMMING)
FLAT_NEW(flat_new_jaccard, VB_METRIC_JACCARD)

static int get_flat(ErlNifEnv* env, ERL_NIF_TERM t, flat_res** r) { return enif_get_resource(env, t, FLAT_TYPE, (void**)r); }

/* flat_insert/3 (nifs.rs:259-271) */
static ERL_NIF_TERM flat_insert(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; ErlNifBinary id; fbuf v = {0};
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &id) || !append_floats(env, argv[2], &v)) {
        free(v.v); return enif_make_badarg(env);
    }
    static const float z = 0.0f;
    int rc = vb_flat_insert(r->index, (const char*)id.data, id.size, v.v ? v.v : &z, v.n);
    free(v.v);
    return status_unit(env, rc);
}

/* flat_insert_many/2 (nifs.rs:273-284): [{id, [float]}] -> blob + offsets, ONE C call, one bulk H2D */
static ERL_NIF_TERM flat_insert_many(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; batch b;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !decode_batch(env, argv[1], PAYLOAD_FLOATS, &b)) return enif_make_badarg(env);
    int rc = vb_flat_insert_many(r->index, b.n, b_ids(&b), b.id_off.v, b_vals(&b), b.val_off.v);
    batch_free(&b);
    return status_unit(env, rc);
}

/* flat_delete/2 (nifs.rs:286-295) */
static ERL_NIF_TERM flat_delete(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; ErlNifBinary id;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &id)) return enif_make_badarg(env);
    return status_unit(env, vb_flat_delete(r->index, (const char*)id.data, id.size));
}

/* flat_search/3 (nifs.rs:297-309) -> {:ok, [{id, raw}]} */
static ERL_NIF_TERM flat_search(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; fbuf q = {0}; size_t limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !get_size(env, argv[2], &limit) || !append_floats(env, argv[1], &q)) {
        free(q.v); return enif_make_badarg(env);
    }
    static const float z = 0.0f;
    int rc = vb_flat_search(r->index, q.v ? q.v : &z, q.n, limit, &hits);
    free(q.v);
    return hits_result(env, rc, hits);
}

/* ---- additive NIFs (no counterpart in the reference: resident pipelines) ---- */
/* flat_reserve/2: capacity hint before rebuild_index streams a snapshot in (collection.ex:426-433) */
static ERL_NIF_TERM flat_reserve(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; size_t rows;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !get_size(env, argv[1], &rows)) return enif_make_badarg(env);
    return status_unit(env, vb_flat_reserve(r->index, rows));
}

/* flat_search_shaped/5: flat_search + Distance.result_values (vettore_distance.ex:525-543) for the whole hit
 * list in one call: index, query, limit, metric_code, score_mode (0 raw, 1 similarity) ->
 * {:ok, [{id, raw, score, distance}]}; only the ETS lookups are left to the caller (index/flat.ex:72-91). */
static ERL_NIF_TERM flat_search_shaped(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; fbuf q = {0}; size_t limit, mode; int metric; vb_hits* hits = NULL;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !get_size(env, argv[2], &limit) || !get_metric(env, argv[3], &metric) ||
        !get_size(env, argv[4], &mode) || mode > 1 || !append_floats(env, argv[1], &q)) {
        free(q.v); return enif_make_badarg(env);
    }
    static const float z = 0.0f;
    int rc = vb_flat_search(r->index, q.v ? q.v : &z, q.n, limit, &hits);
    free(q.v);
    if (rc != VB_OK) return mk_error(env);
    const char* blob; const uint64_t* off; const float* val; const uint64_t* idx;
    size_t n = vb_hits_export(hits, &blob, &off, &val, &idx);
    double* score = (double*)malloc((n ? n : 1) * sizeof(double));
    double* dist = (double*)malloc((n ? n : 1) * sizeof(double));
    if (vb_result_values(metric, (int)mode, val, n, score, dist) != VB_OK) {
        free(score); free(dist); vb_hits_free(hits);
        return mk_error(env);
    }
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (size_t i = n; i-- > 0;) {
        ERL_NIF_TERM id = mk_binary(env, blob + off[i], (size_t)(off[i + 1] - off[i]));
        list = enif_make_list_cell(env, enif_make_tuple4(env, id, enif_make_double(env, val[i]), enif_make_double(env, score[i]),
                                                         enif_make_double(env, dist[i])), list);
    }
    free(score); free(dist); vb_hits_free(hits);
    return mk_ok(env, list);
}

/* flat_funnel_search/6 (collection.ex:244-260 on the resident matrix): index, query, metric_code, stages, candidates, limit */
static ERL_NIF_TERM flat_funnel_search(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; fbuf q = {0}; ubuf st = {0}; int metric; size_t candidates, limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !get_metric(env, argv[2], &metric) || !get_size(env, argv[4], &candidates) ||
        !get_size(env, argv[5], &limit) || !append_floats(env, argv[1], &q) || !append_words(env, argv[3], &st)) {
        free(q.v); free(st.v); return enif_make_badarg(env);
    }
    size_t* stages = (size_t*)malloc((st.n ? st.n : 1) * sizeof(size_t));
    for (size_t i = 0; i < st.n; ++i) stages[i] = (size_t)st.v[i];
    static const float z = 0.0f;
    int rc = vb_flat_funnel_search(r->index, q.v ? q.v : &z, q.n, metric, stages, st.n, candidates, limit, &hits);
    free(stages); free(q.v); free(st.v);
    return hits_result(env, rc, hits);
}

/* flat_quantized_search/5 (collection.ex:266-295 on the resident sign codes): index, query, metric_code, candidates, limit */
static ERL_NIF_TERM flat_quantized_search(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; fbuf q = {0}; int metric; size_t candidates, limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_flat(env, argv[0], &r) || !get_metric(env, argv[2], &metric) || !get_size(env, argv[3], &candidates) ||
        !get_size(env, argv[4], &limit) || !append_floats(env, argv[1], &q)) {
        free(q.v); return enif_make_badarg(env);
    }
    static const float z = 0.0f;
    int rc = vb_flat_quantized_search(r->index, q.v ? q.v : &z, q.n, metric, candidates, limit, &hits);
    free(q.v);
    return hits_result(env, rc, hits);
}

/* mv_new/1 (metric_code) -> resource; mv_insert_many/2; mv_delete/2; mv_search/3: the HBM-resident
 * multi-vector collection behind multi_vector_search (collection.ex:313-323) */
static ERL_NIF_TERM mv_new(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    int metric; vb_mv* index = NULL;
    (void)argc;
    if (!get_metric(env, argv[0], &metric)) return enif_make_badarg(env);
    {   /* VETTORE_B200_GPUS=N (N > 1): the collection spread over N GPUs, like flat_new */
        const char* gpus = getenv("VETTORE_B200_GPUS");
        const int n_gpus = gpus ? atoi(gpus) : 1;
        const int rc = n_gpus > 1 ? vb_mv_new_sharded(metric, n_gpus, NULL, &index) : vb_mv_new(metric, &index);
        if (rc != VB_OK) return mk_error(env);
    }
    mv_res* r = (mv_res*)enif_alloc_resource(MV_TYPE, sizeof(mv_res));
    r->index = index;
    ERL_NIF_TERM t = enif_make_resource(env, r);
    enif_release_resource(r);
    return mk_ok(env, t);
}
static int get_mv(ErlNifEnv* env, ERL_NIF_TERM t, mv_res** r) { return enif_get_resource(env, t, MV_TYPE, (void**)r); }
static ERL_NIF_TERM mv_insert_many(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    mv_res* r; batch b;
    (void)argc;
    if (!get_mv(env, argv[0], &r) || !decode_batch(env, argv[1], PAYLOAD_VECTORS, &b)) return enif_make_badarg(env);
    int rc = vb_mv_insert_many(r->index, b.n, b_ids(&b), b.id_off.v, b_vals(&b), b.val_off.v, b.doc_tok.v);
    batch_free(&b);
    return status_unit(env, rc);
}
static ERL_NIF_TERM mv_delete(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    mv_res* r; ErlNifBinary id;
    (void)argc;
    if (!get_mv(env, argv[0], &r) || !enif_inspect_binary(env, argv[1], &id)) return enif_make_badarg(env);
    return status_unit(env, vb_mv_delete(r->index, (const char*)id.data, id.size));
}
static ERL_NIF_TERM mv_search(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    mv_res* r; vecs q; size_t limit; vb_hits* hits = NULL;
    (void)argc;
    if (!get_mv(env, argv[0], &r) || !get_size(env, argv[2], &limit) || !decode_vectors(env, argv[1], &q)) return enif_make_badarg(env);
    int rc = vb_mv_search(r->index, v_vals(&q), q.off.v, q.n, limit, &hits);
    vecs_free(&q);
    return hits_result(env, rc, hits);
}

/* muvera_encode_query/7 and muvera_encode_document/7 (lib/vettore_nifs.ex:219-258; nifs.rs:430-476): vectors,
 * dimension, num_repetitions, num_simhash_projections, seed, projection_dimension, final_projection_dimension
 * (an integer or nil = Option<usize>) -> {:ok, [float]}. One multi-vector per call like the reference; the C
 * entry itself is batched over documents (vb_muvera_encode). */
static ERL_NIF_TERM muvera_encode(ErlNifEnv* env, const ERL_NIF_TERM argv[], int mode) {
    vecs v; size_t dim, reps, ks, pdim, fin = 0; ErlNifUInt64 seed; int has_final = 1; char atom[8];
    if (!get_size(env, argv[1], &dim) || !get_size(env, argv[2], &reps) || !get_size(env, argv[3], &ks) ||
        !enif_get_uint64(env, argv[4], &seed) || !get_size(env, argv[5], &pdim)) return enif_make_badarg(env);
    if (!get_size(env, argv[6], &fin)) {
        if (!enif_get_atom(env, argv[6], atom, sizeof(atom), ERL_NIF_LATIN1) || strcmp(atom, "nil") != 0) return enif_make_badarg(env);
        has_final = 0;
    }
    if (!decode_vectors(env, argv[0], &v)) return enif_make_badarg(env);
    /* the output size is known only after validation: out == NULL validates and sizes without computing */
    size_t fde = 0;
    const uint64_t doc_vec[2] = {0, v.n};
    int rc = vb_muvera_encode(1, v_vals(&v), v.off.v, doc_vec, dim, reps, ks, (uint64_t)seed, pdim, has_final, fin, mode, NULL, 0, &fde);
    float* out = NULL;
    if (rc == VB_OK) {
        out = (float*)malloc((fde ? fde : 1) * sizeof(float));
        rc = vb_muvera_encode(1, v_vals(&v), v.off.v, doc_vec, dim, reps, ks, (uint64_t)seed, pdim, has_final, fin, mode, out, fde, &fde);
    }
    vecs_free(&v);
    if (rc != VB_OK) { free(out); return mk_error(env); }
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (size_t i = fde; i-- > 0;) list = enif_make_list_cell(env, enif_make_double(env, out[i]), list);
    free(out);
    return mk_ok(env, list);
}
static ERL_NIF_TERM muvera_encode_query(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) { (void)argc; return muvera_encode(env, argv, 0); }
static ERL_NIF_TERM muvera_encode_document(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) { (void)argc; return muvera_encode(env, argv, 1); }

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
    (void)priv; (void)info;
    FLAT_TYPE = enif_open_resource_type(env, NULL, "vettore_b200_flat", flat_dtor, ERL_NIF_RT_CREATE, NULL);
    MV_TYPE = enif_open_resource_type(env, NULL, "vettore_b200_mv", mv_dtor, ERL_NIF_RT_CREATE, NULL);
    return (FLAT_TYPE && MV_TYPE) ? 0 : 1;
}

#define DIRTY ERL_NIF_DIRTY_JOB_CPU_BOUND
static ErlNifFunc nif_funcs[] = {
    /* the scan-path subset of Vettore.Nifs, same names and arities (lib/vettore_nifs.ex) */
    {"compress_sign_bits", 1, compress_sign_bits, DIRTY},
    {"vector_top_k", 5, vector_top_k, DIRTY},
    {"binary_top_k", 4, binary_top_k, DIRTY},
    {"multi_vector_score", 3, multi_vector_score, DIRTY},
    {"multi_vector_top_k", 4, multi_vector_top_k, DIRTY},
    {"flat_new_l2", 0, flat_new_l2, DIRTY},
    {"flat_new_l2_squared", 0, flat_new_l2_squared, DIRTY},
    {"flat_new_cosine", 0, flat_new_cosine, DIRTY},
    {"flat_new_inner_product", 0, flat_new_inner_product, DIRTY},
    {"flat_new_negative_inner_product", 0, flat_new_negative_inner_product, DIRTY},
    {"flat_new_manhattan", 0, flat_new_manhattan, DIRTY},
    {"flat_new_chebyshev", 0, flat_new_chebyshev, DIRTY},
    {"flat_new_hamming", 0, flat_new_hamming, DIRTY},
    {"flat_new_jaccard", 0, flat_new_jaccard, DIRTY},
    {"flat_insert", 3, flat_insert, DIRTY},
    {"flat_insert_many", 2, flat_insert_many, DIRTY},
    {"flat_delete", 2, flat_delete, DIRTY},
    {"flat_search", 3, flat_search, DIRTY},
    {"muvera_encode_query", 7, muvera_encode_query, DIRTY},
    {"muvera_encode_document", 7, muvera_encode_document, DIRTY},
    /* additive */
    {"flat_reserve", 2, flat_reserve, DIRTY},
    {"flat_search_shaped", 5, flat_search_shaped, DIRTY},
    {"flat_funnel_search", 6, flat_funnel_search, DIRTY},
    {"flat_quantized_search", 5, flat_quantized_search, DIRTY},
    {"mv_new", 1, mv_new, DIRTY},
    {"mv_insert_many", 2, mv_insert_many, DIRTY},
    {"mv_delete", 2, mv_delete, DIRTY},
    {"mv_search", 3, mv_search, DIRTY},
};
ERL_NIF_INIT(Elixir.Vettore.B200.Nifs, nif_funcs, load, NULL, NULL, NULL)
