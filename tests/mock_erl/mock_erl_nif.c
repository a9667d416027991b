/* mock_erl_nif.c — a tiny term heap implementing the enif_* subset of tests/mock_erl/erl_nif.h, plus the
 * mock_* entry points pytest uses to build argument terms, call a NIF by (name, arity) through the
 * ErlNifEntry table that ERL_NIF_INIT produced, and read the result term back. Test infrastructure only;
 * terms are never freed (a test process is short-lived), resources run their destructor on mock_release. */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "erl_nif.h"

enum { T_INT = 1, T_DOUBLE, T_ATOM, T_BINARY, T_NIL, T_CONS, T_TUPLE, T_RESOURCE, T_BADARG, T_EXCEPTION };
typedef struct term {
    int type;
    uint64_t u; int neg;            /* T_INT: magnitude + sign */
    double d;
    char* s; size_t len;            /* atom name / binary bytes */
    struct term *head, *tail;       /* T_CONS; T_EXCEPTION keeps the reason in head */
    ERL_NIF_TERM* elems; int arity; /* T_TUPLE */
    void* obj;                      /* T_RESOURCE */
} term;
struct enif_resource_type_t { ErlNifResourceDtor* dtor; };
typedef struct { ErlNifResourceType* type; int refs; } res_hdr;
struct enif_environment_t { int unused; };
static ErlNifEnv g_env;

static term* T(ERL_NIF_TERM t) { return (term*)t; }
static ERL_NIF_TERM mk(int type) { term* t = (term*)calloc(1, sizeof(term)); t->type = type; return (ERL_NIF_TERM)t; }

int enif_get_list_length(ErlNifEnv* e, ERL_NIF_TERM l, unsigned* len) {
    (void)e; unsigned n = 0; term* t = T(l);
    while (t->type == T_CONS) { ++n; t = t->tail; }
    if (t->type != T_NIL) return 0;
    *len = n; return 1;
}
int enif_get_list_cell(ErlNifEnv* e, ERL_NIF_TERM l, ERL_NIF_TERM* h, ERL_NIF_TERM* tl) {
    (void)e; if (T(l)->type != T_CONS) return 0;
    *h = (ERL_NIF_TERM)T(l)->head; *tl = (ERL_NIF_TERM)T(l)->tail; return 1;
}
int enif_get_double(ErlNifEnv* e, ERL_NIF_TERM t, double* d) { (void)e; if (T(t)->type != T_DOUBLE) return 0; *d = T(t)->d; return 1; }
int enif_get_long(ErlNifEnv* e, ERL_NIF_TERM t, long* ip) {
    (void)e; if (T(t)->type != T_INT || T(t)->u > (uint64_t)INT64_MAX) return 0;
    *ip = T(t)->neg ? -(long)T(t)->u : (long)T(t)->u; return 1;
}
int enif_get_int(ErlNifEnv* e, ERL_NIF_TERM t, int* ip) {
    long l; if (!enif_get_long(e, t, &l) || l > 2147483647L || l < -2147483647L - 1) return 0; *ip = (int)l; return 1;
}
int enif_get_uint64(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifUInt64* ip) {
    (void)e; if (T(t)->type != T_INT || T(t)->neg) return 0; *ip = T(t)->u; return 1;
}
int enif_get_atom(ErlNifEnv* e, ERL_NIF_TERM t, char* buf, unsigned len, ErlNifCharEncoding enc) {
    (void)e; (void)enc; if (T(t)->type != T_ATOM || T(t)->len + 1 > len) return 0;
    memcpy(buf, T(t)->s, T(t)->len + 1); return (int)T(t)->len + 1;
}
int enif_get_tuple(ErlNifEnv* e, ERL_NIF_TERM t, int* arity, const ERL_NIF_TERM** arr) {
    (void)e; if (T(t)->type != T_TUPLE) return 0; *arity = T(t)->arity; *arr = T(t)->elems; return 1;
}
int enif_inspect_binary(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifBinary* b) {
    (void)e; if (T(t)->type != T_BINARY) return 0; b->size = T(t)->len; b->data = (unsigned char*)T(t)->s; return 1;
}
int enif_get_resource(ErlNifEnv* e, ERL_NIF_TERM t, ErlNifResourceType* type, void** objp) {
    (void)e; if (T(t)->type != T_RESOURCE) return 0;
    res_hdr* h = (res_hdr*)T(t)->obj - 1;
    if (h->type != type) return 0;
    *objp = T(t)->obj; return 1;
}
ERL_NIF_TERM enif_make_atom(ErlNifEnv* e, const char* name) { (void)e; ERL_NIF_TERM t = mk(T_ATOM); T(t)->len = strlen(name); T(t)->s = (char*)malloc(T(t)->len + 1); memcpy(T(t)->s, name, T(t)->len + 1); return t; }
ERL_NIF_TERM enif_make_double(ErlNifEnv* e, double d) { (void)e; ERL_NIF_TERM t = mk(T_DOUBLE); T(t)->d = d; return t; }
ERL_NIF_TERM enif_make_uint64(ErlNifEnv* e, ErlNifUInt64 i) { (void)e; ERL_NIF_TERM t = mk(T_INT); T(t)->u = i; return t; }
static ERL_NIF_TERM make_tuple_v(unsigned cnt, va_list ap) {
    ERL_NIF_TERM t = mk(T_TUPLE); T(t)->arity = (int)cnt;
    T(t)->elems = (ERL_NIF_TERM*)calloc(cnt ? cnt : 1, sizeof(ERL_NIF_TERM));
    for (unsigned i = 0; i < cnt; ++i) T(t)->elems[i] = va_arg(ap, ERL_NIF_TERM);
    return t;
}
ERL_NIF_TERM enif_make_tuple(ErlNifEnv* e, unsigned cnt, ...) { (void)e; va_list ap; va_start(ap, cnt); ERL_NIF_TERM t = make_tuple_v(cnt, ap); va_end(ap); return t; }
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv* e, ERL_NIF_TERM a, ERL_NIF_TERM b) { return enif_make_tuple(e, 2, a, b); }
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv* e, ERL_NIF_TERM a, ERL_NIF_TERM b, ERL_NIF_TERM c, ERL_NIF_TERM d) { return enif_make_tuple(e, 4, a, b, c, d); }
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv* e, ERL_NIF_TERM car, ERL_NIF_TERM cdr) { (void)e; ERL_NIF_TERM t = mk(T_CONS); T(t)->head = T(car); T(t)->tail = T(cdr); return t; }
ERL_NIF_TERM enif_make_list(ErlNifEnv* e, unsigned cnt, ...) {
    ERL_NIF_TERM tmp[16]; va_list ap; va_start(ap, cnt);
    for (unsigned i = 0; i < cnt && i < 16; ++i) tmp[i] = va_arg(ap, ERL_NIF_TERM);
    va_end(ap);
    ERL_NIF_TERM l = mk(T_NIL);
    for (unsigned i = cnt; i-- > 0;) l = enif_make_list_cell(e, tmp[i], l);
    return l;
}
unsigned char* enif_make_new_binary(ErlNifEnv* e, size_t size, ERL_NIF_TERM* termp) {
    (void)e; ERL_NIF_TERM t = mk(T_BINARY); T(t)->s = (char*)calloc(size ? size : 1, 1); T(t)->len = size; *termp = t; return (unsigned char*)T(t)->s;
}
ERL_NIF_TERM enif_make_badarg(ErlNifEnv* e) { (void)e; return mk(T_BADARG); }
ERL_NIF_TERM enif_raise_exception(ErlNifEnv* e, ERL_NIF_TERM reason) { (void)e; ERL_NIF_TERM t = mk(T_EXCEPTION); T(t)->head = T(reason); return t; }
void* enif_alloc_resource(ErlNifResourceType* type, size_t size) {
    res_hdr* h = (res_hdr*)calloc(1, sizeof(res_hdr) + size); h->type = type; h->refs = 1; return h + 1;
}
ERL_NIF_TERM enif_make_resource(ErlNifEnv* e, void* obj) { (void)e; ERL_NIF_TERM t = mk(T_RESOURCE); T(t)->obj = obj; ((res_hdr*)obj - 1)->refs++; return t; }
void enif_release_resource(void* obj) {
    res_hdr* h = (res_hdr*)obj - 1;
    if (--h->refs == 0) { if (h->type->dtor) h->type->dtor(&g_env, obj); free(h); }
}
ErlNifResourceType* enif_open_resource_type(ErlNifEnv* e, const char* m, const char* n, ErlNifResourceDtor* dtor,
                                            ErlNifResourceFlags flags, ErlNifResourceFlags* tried) {
    (void)e; (void)m; (void)n; (void)flags; (void)tried;
    ErlNifResourceType* t = (ErlNifResourceType*)calloc(1, sizeof(*t)); t->dtor = dtor; return t;
}

/* ---- driver side (pytest via ctypes) ---------------------------------------------------------------- */
ErlNifEntry* nif_init(void);
static ErlNifEntry* g_entry;
int mock_load(void) {
    g_entry = nif_init();
    void* priv = NULL;
    return g_entry->load ? g_entry->load(&g_env, &priv, mk(T_NIL)) : 0;
}
const char* mock_module_name(void) { return g_entry->name; }
int mock_num_funcs(void) { return g_entry->num_of_funcs; }
const char* mock_func_name(int i) { return g_entry->funcs[i].name; }
unsigned mock_func_arity(int i) { return g_entry->funcs[i].arity; }
unsigned mock_func_flags(int i) { return g_entry->funcs[i].flags; }
ERL_NIF_TERM mock_call(const char* name, int argc, const ERL_NIF_TERM* argv) {
    for (int i = 0; i < g_entry->num_of_funcs; ++i)
        if (!strcmp(g_entry->funcs[i].name, name) && (int)g_entry->funcs[i].arity == argc)
            return g_entry->funcs[i].fptr(&g_env, argc, argv);
    return 0;
}
ERL_NIF_TERM mock_int(uint64_t mag, int neg) { ERL_NIF_TERM t = mk(T_INT); T(t)->u = mag; T(t)->neg = neg; return t; }
ERL_NIF_TERM mock_double(double d) { return enif_make_double(&g_env, d); }
ERL_NIF_TERM mock_atom(const char* s) { return enif_make_atom(&g_env, s); }
ERL_NIF_TERM mock_binary(const char* p, size_t n) { ERL_NIF_TERM t; memcpy(enif_make_new_binary(&g_env, n, &t), p, n); return t; }
ERL_NIF_TERM mock_nil(void) { return mk(T_NIL); }
ERL_NIF_TERM mock_cons(ERL_NIF_TERM h, ERL_NIF_TERM t) { return enif_make_list_cell(&g_env, h, t); }
ERL_NIF_TERM mock_tuple(int n, const ERL_NIF_TERM* elems) {
    ERL_NIF_TERM t = mk(T_TUPLE); T(t)->arity = n; T(t)->elems = (ERL_NIF_TERM*)calloc(n ? n : 1, sizeof(ERL_NIF_TERM));
    memcpy(T(t)->elems, elems, n * sizeof(ERL_NIF_TERM)); return t;
}
/* A dense list of floats / list of u64 without one ctypes call per element. */
ERL_NIF_TERM mock_float_list(const double* v, size_t n) { ERL_NIF_TERM l = mk(T_NIL); for (size_t i = n; i-- > 0;) l = mock_cons(mock_double(v[i]), l); return l; }
ERL_NIF_TERM mock_u64_list(const uint64_t* v, size_t n) { ERL_NIF_TERM l = mk(T_NIL); for (size_t i = n; i-- > 0;) l = mock_cons(mock_int(v[i], 0), l); return l; }
int mock_type(ERL_NIF_TERM t) { return T(t)->type; }
uint64_t mock_get_int(ERL_NIF_TERM t, int* neg) { *neg = T(t)->neg; return T(t)->u; }
double mock_get_double(ERL_NIF_TERM t) { return T(t)->d; }
const char* mock_get_bytes(ERL_NIF_TERM t, size_t* n) { *n = T(t)->len; return T(t)->s; }
ERL_NIF_TERM mock_head(ERL_NIF_TERM t) { return (ERL_NIF_TERM)T(t)->head; }
ERL_NIF_TERM mock_tail(ERL_NIF_TERM t) { return (ERL_NIF_TERM)T(t)->tail; }
int mock_arity(ERL_NIF_TERM t) { return T(t)->arity; }
ERL_NIF_TERM mock_elem(ERL_NIF_TERM t, int i) { return T(t)->elems[i]; }
void mock_release(ERL_NIF_TERM t) { if (T(t)->type == T_RESOURCE) enif_release_resource(T(t)->obj); }
