/* vettore_b200_nif.c — erl_nif shim that maps the scan-path functions of `Vettore.Nifs`
 * (reference lib/vettore_nifs.ex) onto the C ABI of include/vettore_b200.h.
 *
 * NOT compiled in this repository's image (no erl_nif.h / OTP here); it is the binding a
 * maintainer adds on the reference side. Build next to the Rust crate:
 *   cc -O2 -fPIC -shared -I$ERL_INCLUDE -Iinclude nif/vettore_b200_nif.c \
 *      -Lvettore_b200 -lvettore_b200 -o priv/native/libvettore_b200_nif.so
 * and load it from `Vettore.B200.Nifs`. HNSW / MUVERA / pairwise
 * helpers keep coming from the Rust NIF library (`Vettore.Nifs`).
 *
 * Every function is a dirty CPU-bound NIF like the reference (nifs.rs: schedule = "DirtyCpu").
 */
#if defined(__has_include)
#if __has_include(<erl_nif.h>)
#define VB_HAVE_ERL_NIF 1
#endif
#endif

#ifdef VB_HAVE_ERL_NIF
#include <erl_nif.h>
#include <stdlib.h>
#include <string.h>

#include "vettore_b200.h"

static ErlNifResourceType* FLAT_TYPE;

typedef struct { vb_flat* index; } flat_res;

static void flat_dtor(ErlNifEnv* env, void* obj) { (void)env; vb_flat_free(((flat_res*)obj)->index); }

static ERL_NIF_TERM mk_error(ErlNifEnv* env) {           /* {:error, "message"} */
    const char* msg = vb_last_error();
    ERL_NIF_TERM bin;
    unsigned char* p = enif_make_new_binary(env, strlen(msg), &bin);
    memcpy(p, msg, strlen(msg));
    return enif_make_tuple2(env, enif_make_atom(env, "error"), bin);
}
static ERL_NIF_TERM mk_ok_unit(ErlNifEnv* env) {         /* {:ok, {}} like Result<(), String> */
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), enif_make_tuple(env, 0));
}

/* Erlang list of floats -> malloc'd float array (Rustler narrows f64 -> f32 the same way). */
static int get_floats(ErlNifEnv* env, ERL_NIF_TERM list, float** out, size_t* n) {
    unsigned len;
    if (!enif_get_list_length(env, list, &len)) return 0;
    float* v = (float*)malloc((len ? len : 1) * sizeof(float));
    ERL_NIF_TERM head, tail = list;
    for (unsigned i = 0; i < len; ++i) {
        double d; long l;
        enif_get_list_cell(env, tail, &head, &tail);
        if (enif_get_double(env, head, &d)) v[i] = (float)d;
        else if (enif_get_long(env, head, &l)) v[i] = (float)l;
        else { free(v); return 0; }
    }
    *out = v; *n = len;
    return 1;
}

/* vb_hits -> [{id_binary, float}] */
static ERL_NIF_TERM hits_to_term(ErlNifEnv* env, vb_hits* h) {
    const char* blob; const uint64_t* off; const float* val; const uint64_t* idx;
    size_t n = vb_hits_export(h, &blob, &off, &val, &idx);
    ERL_NIF_TERM list = enif_make_list(env, 0);
    for (size_t i = n; i-- > 0;) {
        ERL_NIF_TERM bin;
        size_t len = (size_t)(off[i + 1] - off[i]);
        memcpy(enif_make_new_binary(env, len, &bin), blob + off[i], len);
        list = enif_make_list_cell(env, enif_make_tuple2(env, bin, enif_make_double(env, val[i])), list);
    }
    vb_hits_free(h);
    return list;
}

/* flat_new_<metric>/0: returns the bare resource (nifs.rs:200-257) */
static ERL_NIF_TERM flat_new(ErlNifEnv* env, int metric) {
    flat_res* r = (flat_res*)enif_alloc_resource(FLAT_TYPE, sizeof(flat_res));
    if (vb_flat_new(metric, &r->index) != VB_OK) { enif_release_resource(r); return enif_raise_exception(env, mk_error(env)); }
    ERL_NIF_TERM t = enif_make_resource(env, r);
    enif_release_resource(r);
    return t;
}
#define FLAT_NEW(name, code) \
    static ERL_NIF_TERM name(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) { (void)argc; (void)argv; return flat_new(env, code); }
FLAT_NEW(flat_new_l2, VB_METRIC_L2) FLAT_NEW(flat_new_l2_squared, VB_METRIC_L2_SQUARED)
FLAT_NEW(flat_new_cosine, VB_METRIC_COSINE) FLAT_NEW(flat_new_inner_product, VB_METRIC_INNER_PRODUCT)
FLAT_NEW(flat_new_negative_inner_product, VB_METRIC_NEGATIVE_INNER_PRODUCT)
FLAT_NEW(flat_new_manhattan, VB_METRIC_MANHATTAN) FLAT_NEW(flat_new_chebyshev, VB_METRIC_CHEBYSHEV)
FLAT_NEW(flat_new_hamming, VB_METRIC_HAMMING) FLAT_NEW(flat_new_jaccard, VB_METRIC_JACCARD)

/* flat_insert/3 (nifs.rs:259-271) */
static ERL_NIF_TERM flat_insert(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; ErlNifBinary id; float* v; size_t n;
    (void)argc;
    if (!enif_get_resource(env, argv[0], FLAT_TYPE, (void**)&r) || !enif_inspect_binary(env, argv[1], &id) ||
        !get_floats(env, argv[2], &v, &n)) return enif_make_badarg(env);
    int rc = vb_flat_insert(r->index, (const char*)id.data, id.size, v, n);
    free(v);
    return rc == VB_OK ? mk_ok_unit(env) : mk_error(env);
}

/* flat_insert_many/2 (nifs.rs:273-284): [{id, [float]}] -> blob + offsets, one C call */
static ERL_NIF_TERM flat_insert_many(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; unsigned n;
    (void)argc;
    if (!enif_get_resource(env, argv[0], FLAT_TYPE, (void**)&r) || !enif_get_list_length(env, argv[1], &n))
        return enif_make_badarg(env);
    uint64_t* id_off = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    uint64_t* val_off = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    size_t id_cap = 64 * (size_t)n + 1, val_cap = 1024, id_len = 0, val_len = 0;
    char* ids = (char*)malloc(id_cap);
    float* vals = (float*)malloc(val_cap * sizeof(float));
    ERL_NIF_TERM head, tail = argv[1];
    int ok = 1;
    for (unsigned i = 0; ok && i < n; ++i) {
        const ERL_NIF_TERM* tup; int arity; ErlNifBinary id; float* v; size_t vn;
        enif_get_list_cell(env, tail, &head, &tail);
        ok = enif_get_tuple(env, head, &arity, &tup) && arity == 2 && enif_inspect_binary(env, tup[0], &id) &&
             get_floats(env, tup[1], &v, &vn);
        if (!ok) break;
        if (id_len + id.size > id_cap) { id_cap = 2 * (id_len + id.size); ids = (char*)realloc(ids, id_cap); }
        memcpy(ids + id_len, id.data, id.size); id_len += id.size; id_off[i + 1] = id_len;
        if (val_len + vn > val_cap) { val_cap = 2 * (val_len + vn); vals = (float*)realloc(vals, val_cap * sizeof(float)); }
        memcpy(vals + val_len, v, vn * sizeof(float)); val_len += vn; val_off[i + 1] = val_len;
        free(v);
    }
    ERL_NIF_TERM res = enif_make_badarg(env);
    if (ok) res = vb_flat_insert_many(r->index, n, ids, id_off, vals, val_off) == VB_OK ? mk_ok_unit(env) : mk_error(env);
    free(ids); free(vals); free(id_off); free(val_off);
    return res;
}

/* flat_delete/2 (nifs.rs:286-295) */
static ERL_NIF_TERM flat_delete(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; ErlNifBinary id;
    (void)argc;
    if (!enif_get_resource(env, argv[0], FLAT_TYPE, (void**)&r) || !enif_inspect_binary(env, argv[1], &id))
        return enif_make_badarg(env);
    return vb_flat_delete(r->index, (const char*)id.data, id.size) == VB_OK ? mk_ok_unit(env) : mk_error(env);
}

/* flat_search/3 (nifs.rs:297-309) -> {:ok, [{id, raw}]} */
static ERL_NIF_TERM flat_search(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]) {
    flat_res* r; float* q; size_t n; ErlNifUInt64 limit; vb_hits* hits;
    (void)argc;
    if (!enif_get_resource(env, argv[0], FLAT_TYPE, (void**)&r) || !get_floats(env, argv[1], &q, &n) ||
        !enif_get_uint64(env, argv[2], &limit)) return enif_make_badarg(env);
    int rc = vb_flat_search(r->index, q, n, (size_t)limit, &hits);
    free(q);
    if (rc != VB_OK) return mk_error(env);
    return enif_make_tuple2(env, enif_make_atom(env, "ok"), hits_to_term(env, hits));
}

/* vector_top_k/5, binary_top_k/4, multi_vector_top_k/4, multi_vector_score/3 and the additive
 * flat_funnel_search / flat_quantized_search / mv_* follow the same pattern: decode lists into
 * blob+offset arrays, one vb_* call, hits_to_term. Omitted here for brevity; their C signatures
 * are in include/vettore_b200.h next to the NIF each one replaces. */

static int load(ErlNifEnv* env, void** priv, ERL_NIF_TERM info) {
    (void)priv; (void)info;
    FLAT_TYPE = enif_open_resource_type(env, NULL, "vettore_b200_flat", flat_dtor, ERL_NIF_RT_CREATE, NULL);
    return FLAT_TYPE ? 0 : 1;
}

#define DIRTY ERL_NIF_DIRTY_JOB_CPU_BOUND
static ErlNifFunc nif_funcs[] = {
    {"flat_new_l2", 0, flat_new_l2, DIRTY}, {"flat_new_l2_squared", 0, flat_new_l2_squared, DIRTY},
    {"flat_new_cosine", 0, flat_new_cosine, DIRTY}, {"flat_new_inner_product", 0, flat_new_inner_product, DIRTY},
    {"flat_new_negative_inner_product", 0, flat_new_negative_inner_product, DIRTY},
    {"flat_new_manhattan", 0, flat_new_manhattan, DIRTY}, {"flat_new_chebyshev", 0, flat_new_chebyshev, DIRTY},
    {"flat_new_hamming", 0, flat_new_hamming, DIRTY}, {"flat_new_jaccard", 0, flat_new_jaccard, DIRTY},
    {"flat_insert", 3, flat_insert, DIRTY}, {"flat_insert_many", 2, flat_insert_many, DIRTY},
    {"flat_delete", 2, flat_delete, DIRTY}, {"flat_search", 3, flat_search, DIRTY},
};
ERL_NIF_INIT(Elixir.Vettore.B200.Nifs, nif_funcs, load, NULL, NULL, NULL)
#endif /* VB_HAVE_ERL_NIF */
