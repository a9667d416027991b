/* erl_nif.h — MOCK of the Erlang/OTP NIF API, test infrastructure only.
 *
 * The image has no Erlang/OTP, so the real <erl_nif.h> is absent. This header declares the subset of the
 * NIF API that nif/vettore_b200_nif.c uses, with the documented OTP signatures (erts/emulator/beam/erl_nif.h,
 * erl_nif_api_funcs.h), so that the shim goes through a C compiler (tests/test_nif_shim.py: gcc -Wall
 * -Werror) and, linked against mock_erl_nif.c (a tiny term heap), can be DRIVEN from pytest: Erlang terms in,
 * Erlang terms out, through the very functions a BEAM dirty scheduler would call. It is not a BEAM. */
#ifndef MOCK_ERL_NIF_H
#define MOCK_ERL_NIF_H
#include <stddef.h>
#include <stdint.h>

typedef uintptr_t ERL_NIF_TERM;
typedef struct enif_environment_t ErlNifEnv;
typedef uint64_t ErlNifUInt64;
typedef int64_t ErlNifSInt64;
typedef struct { size_t size; unsigned char* data; void* ref_bin; void* spare[2]; } ErlNifBinary;
typedef struct enif_resource_type_t ErlNifResourceType;
typedef void ErlNifResourceDtor(ErlNifEnv*, void*);
typedef enum { ERL_NIF_RT_CREATE = 1, ERL_NIF_RT_TAKEOVER = 2 } ErlNifResourceFlags;
typedef struct {
    const char* name;
    unsigned arity;
    ERL_NIF_TERM (*fptr)(ErlNifEnv* env, int argc, const ERL_NIF_TERM argv[]);
    unsigned flags;
} ErlNifFunc;
typedef struct {
    int major, minor;
    const char* name;
    int num_of_funcs;
    ErlNifFunc* funcs;
    int (*load)(ErlNifEnv*, void** priv_data, ERL_NIF_TERM load_info);
    int (*reload)(ErlNifEnv*, void** priv_data, ERL_NIF_TERM load_info);
    int (*upgrade)(ErlNifEnv*, void** priv_data, void** old_priv_data, ERL_NIF_TERM load_info);
    void (*unload)(ErlNifEnv*, void* priv_data);
    const char* vm_variant;
    unsigned options;
    size_t sizeof_ErlNifResourceTypeInit;
    const char* min_erts;
} ErlNifEntry;

#define ERL_NIF_DIRTY_JOB_CPU_BOUND 1
#define ERL_NIF_DIRTY_JOB_IO_BOUND 2

int enif_get_list_length(ErlNifEnv*, ERL_NIF_TERM list, unsigned* len);
int enif_get_list_cell(ErlNifEnv*, ERL_NIF_TERM list, ERL_NIF_TERM* head, ERL_NIF_TERM* tail);
int enif_get_double(ErlNifEnv*, ERL_NIF_TERM term, double* dp);
int enif_get_long(ErlNifEnv*, ERL_NIF_TERM term, long* ip);
int enif_get_int(ErlNifEnv*, ERL_NIF_TERM term, int* ip);
int enif_get_uint64(ErlNifEnv*, ERL_NIF_TERM term, ErlNifUInt64* ip);
typedef enum { ERL_NIF_LATIN1 = 1, ERL_NIF_UTF8 = 2 } ErlNifCharEncoding;
int enif_get_atom(ErlNifEnv*, ERL_NIF_TERM atom, char* buf, unsigned len, ErlNifCharEncoding);
int enif_get_tuple(ErlNifEnv*, ERL_NIF_TERM tpl, int* arity, const ERL_NIF_TERM** array);
int enif_inspect_binary(ErlNifEnv*, ERL_NIF_TERM bin_term, ErlNifBinary* bin);
int enif_get_resource(ErlNifEnv*, ERL_NIF_TERM term, ErlNifResourceType* type, void** objp);
ERL_NIF_TERM enif_make_atom(ErlNifEnv*, const char* name);
ERL_NIF_TERM enif_make_double(ErlNifEnv*, double d);
ERL_NIF_TERM enif_make_uint64(ErlNifEnv*, ErlNifUInt64 i);
ERL_NIF_TERM enif_make_tuple(ErlNifEnv*, unsigned cnt, ...);
ERL_NIF_TERM enif_make_tuple2(ErlNifEnv*, ERL_NIF_TERM e1, ERL_NIF_TERM e2);
ERL_NIF_TERM enif_make_tuple4(ErlNifEnv*, ERL_NIF_TERM e1, ERL_NIF_TERM e2, ERL_NIF_TERM e3, ERL_NIF_TERM e4);
ERL_NIF_TERM enif_make_list(ErlNifEnv*, unsigned cnt, ...);
ERL_NIF_TERM enif_make_list_cell(ErlNifEnv*, ERL_NIF_TERM car, ERL_NIF_TERM cdr);
unsigned char* enif_make_new_binary(ErlNifEnv*, size_t size, ERL_NIF_TERM* termp);
ERL_NIF_TERM enif_make_badarg(ErlNifEnv*);
ERL_NIF_TERM enif_raise_exception(ErlNifEnv*, ERL_NIF_TERM reason);
ERL_NIF_TERM enif_make_resource(ErlNifEnv*, void* obj);
void* enif_alloc_resource(ErlNifResourceType* type, size_t size);
void enif_release_resource(void* obj);
ErlNifResourceType* enif_open_resource_type(ErlNifEnv*, const char* module_str, const char* name_str,
                                            ErlNifResourceDtor* dtor, ErlNifResourceFlags flags, ErlNifResourceFlags* tried);

#define ERL_NIF_INIT(NAME, FUNCS, LOAD, RELOAD, UPGRADE, UNLOAD)                                            \
    ErlNifEntry* nif_init(void) {                                                                           \
        static ErlNifEntry entry = {2, 16, #NAME, (int)(sizeof(FUNCS) / sizeof(*FUNCS)), FUNCS, LOAD, RELOAD, \
                                    UPGRADE, UNLOAD, "mock", 0, 0, "mock"};                                   \
        return &entry;                                                                                      \
    }
#endif
