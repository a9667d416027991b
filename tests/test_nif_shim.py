"""The erl_nif shim (nif/vettore_b200_nif.c) through a C compiler and, term in / term out, through the very
functions a BEAM dirty scheduler would call — against a mock of the OTP NIF API (tests/mock_erl/: the image
has no Erlang/OTP). CPU part: it compiles with -Wall -Wextra -Werror, its ErlNifFunc table carries the
scan-path subset of the reference's `Vettore.Nifs` with identical names and arities
(tests/golden/nifs_surface.json, parsed from lib/vettore_nifs.ex), argument decoding / badarg / the
validation errors that need no device. GPU part: the reference's NIF-level known answers
(test/vector_algorithms_hardening_test.exs:20-36, 90-121) and oracle parity for every function."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
from helpers import assert_hits_match

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "nif_mock")
SO = os.path.join(OUT, "libvettore_b200_nif_mock.so")

T_INT, T_DOUBLE, T_ATOM, T_BINARY, T_NIL, T_CONS, T_TUPLE, T_RESOURCE, T_BADARG, T_EXCEPTION = range(1, 11)


class Atom(str):
    pass


class Resource:
    def __init__(self, term):
        self.term = term


class BadArg(Exception):
    pass


class NifException(Exception):
    pass


@pytest.fixture(scope="module")
def nif():
    from vettore_b200 import _lib
    _lib.lib()   # the product library must exist (the shim links against it)
    os.makedirs(OUT, exist_ok=True)
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "tests", "mock_erl"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "nif", "vettore_b200_nif.c"), os.path.join(ROOT, "tests", "mock_erl", "mock_erl_nif.c"),
           "-L", os.path.join(ROOT, "vettore_b200"), "-lvettore_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "vettore_b200"), "-o", SO]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    L = C.CDLL(SO)
    u = C.c_size_t   # ERL_NIF_TERM = uintptr_t
    for name, restype, argtypes in [
        ("mock_load", C.c_int, []), ("mock_module_name", C.c_char_p, []), ("mock_num_funcs", C.c_int, []),
        ("mock_func_name", C.c_char_p, [C.c_int]), ("mock_func_arity", C.c_uint, [C.c_int]), ("mock_func_flags", C.c_uint, [C.c_int]),
        ("mock_call", u, [C.c_char_p, C.c_int, C.POINTER(u)]), ("mock_int", u, [C.c_uint64, C.c_int]),
        ("mock_double", u, [C.c_double]), ("mock_atom", u, [C.c_char_p]), ("mock_binary", u, [C.c_char_p, C.c_size_t]),
        ("mock_nil", u, []), ("mock_cons", u, [u, u]), ("mock_tuple", u, [C.c_int, C.POINTER(u)]),
        ("mock_float_list", u, [C.POINTER(C.c_double), C.c_size_t]), ("mock_u64_list", u, [C.POINTER(C.c_uint64), C.c_size_t]),
        ("mock_type", C.c_int, [u]), ("mock_get_int", C.c_uint64, [u, C.POINTER(C.c_int)]), ("mock_get_double", C.c_double, [u]),
        ("mock_get_bytes", C.POINTER(C.c_char), [u, C.POINTER(C.c_size_t)]), ("mock_head", u, [u]), ("mock_tail", u, [u]),
        ("mock_arity", C.c_int, [u]), ("mock_elem", u, [u, C.c_int]), ("mock_release", None, [u]),
    ]:
        fn = getattr(L, name)
        fn.restype, fn.argtypes = restype, argtypes
    assert L.mock_load() == 0
    return Shim(L)


class Shim:
    def __init__(self, L):
        self.L = L

    # ---- python -> term
    def term(self, x):
        L = self.L
        if isinstance(x, Resource):
            return x.term
        if isinstance(x, Atom):
            return L.mock_atom(x.encode())
        if isinstance(x, bool):
            return L.mock_atom(b"true" if x else b"false")
        if isinstance(x, (int, np.integer)):
            return L.mock_int(abs(int(x)), 1 if int(x) < 0 else 0)
        if isinstance(x, (float, np.floating)):
            return L.mock_double(float(x))
        if isinstance(x, str):
            b = x.encode()
            return L.mock_binary(b, len(b))
        if isinstance(x, bytes):
            return L.mock_binary(x, len(x))
        if isinstance(x, tuple):
            arr = (C.c_size_t * max(1, len(x)))(*[self.term(e) for e in x])
            return L.mock_tuple(len(x), arr)
        if isinstance(x, np.ndarray) and x.ndim == 1 and x.dtype.kind == "f":
            a = np.ascontiguousarray(x, dtype=np.float64)
            return L.mock_float_list(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
        if isinstance(x, np.ndarray) and x.ndim == 1 and x.dtype.kind == "u":
            a = np.ascontiguousarray(x, dtype=np.uint64)
            return L.mock_u64_list(a.ctypes.data_as(C.POINTER(C.c_uint64)), a.size)
        if isinstance(x, (list, np.ndarray)):
            t = L.mock_nil()
            for e in reversed(list(x)):
                t = L.mock_cons(self.term(e), t)
            return t
        raise TypeError(type(x))

    # ---- term -> python
    def py(self, t):
        L = self.L
        ty = L.mock_type(t)
        if ty == T_INT:
            neg = C.c_int()
            v = L.mock_get_int(t, C.byref(neg))
            return -v if neg.value else v
        if ty == T_DOUBLE:
            return L.mock_get_double(t)
        if ty in (T_ATOM, T_BINARY):
            n = C.c_size_t()
            p = L.mock_get_bytes(t, C.byref(n))
            s = C.string_at(p, n.value).decode()
            return Atom(s) if ty == T_ATOM else s
        if ty == T_NIL:
            return []
        if ty == T_CONS:
            out = []
            while L.mock_type(t) == T_CONS:
                out.append(self.py(L.mock_head(t)))
                t = L.mock_tail(t)
            return out
        if ty == T_TUPLE:
            return tuple(self.py(L.mock_elem(t, i)) for i in range(L.mock_arity(t)))
        if ty == T_RESOURCE:
            return Resource(t)
        if ty == T_BADARG:
            raise BadArg()
        if ty == T_EXCEPTION:
            raise NifException(self.py(L.mock_head(t)))
        raise AssertionError(ty)

    def call(self, name, *args):
        argv = (C.c_size_t * max(1, len(args)))(*[self.term(a) for a in args])
        t = self.L.mock_call(name.encode(), len(args), argv)
        assert t != 0, f"{name}/{len(args)} is not in the ErlNifFunc table"
        return self.py(t)

    def table(self):
        L = self.L
        return [(L.mock_func_name(i).decode(), L.mock_func_arity(i), L.mock_func_flags(i)) for i in range(L.mock_num_funcs())]


OK, ERROR = Atom("ok"), Atom("error")


# ------------------------------------------------------------------------------------------- CPU
def test_shim_compiles_and_exports_the_scan_path_surface_of_vettore_nifs(nif):
    surface = json.load(open(os.path.join(ROOT, "tests", "golden", "nifs_surface.json")))
    table = nif.table()
    have = {(n, a) for n, a, _ in table}
    assert len(have) == len(table), "duplicate (name, arity)"
    missing = [tuple(e) for e in surface["scan_path"] if tuple(e) not in have]
    assert not missing, missing
    assert all(flags == 1 for _, _, flags in table), "every NIF is dirty CPU-bound (nifs.rs: schedule = DirtyCpu)"
    assert nif.L.mock_module_name() == b"Elixir.Vettore.B200.Nifs"
    # nothing outside the scan path leaks into the table under a reference name
    ref_all = {tuple(e) for e in surface["all"]}
    assert {e for e in have if e in ref_all} == {tuple(e) for e in surface["scan_path"]}


def test_elixir_stub_module_matches_the_shim_table(nif):
    import re
    src = open(os.path.join(ROOT, "lib", "vettore", "b200", "nifs.ex")).read()
    stubs = set()
    for m in re.finditer(r"def\s+([a-z_0-9]+)(?:\(([^)]*)\))?\s*,\s*do:\s*:erlang\.nif_error", src):
        args = (m.group(2) or "").strip()
        stubs.add((m.group(1), 0 if not args else len(args.split(","))))
    assert stubs == {(n, a) for n, a, _ in nif.table()}
    assert "defmodule Vettore.B200.Nifs" in src


def test_compress_sign_bits_known_answers(nif):
    assert nif.call("compress_sign_bits", [1.0, -2.0, 0.0]) == [5]          # vettore_distance.ex doctest
    assert nif.call("compress_sign_bits", [-0.0]) == [1]                      # -0.0 >= 0.0 (distances.rs:416-420)
    assert nif.call("compress_sign_bits", []) == []
    v = np.linspace(-1, 1, 130)
    assert nif.call("compress_sign_bits", v) == oracle.compress_sign_bits(v.astype(np.float32))
    assert nif.call("compress_sign_bits", [1, -2, 0]) == [5]                  # integers decode like floats


def test_badarg_on_mistyped_terms(nif):
    for args in [("compress_sign_bits", Atom("x")), ("compress_sign_bits", ["a"]),
                 ("vector_top_k", [("a", [1.0])], [1.0], -1, 1, 1), ("vector_top_k", [("a", [1.0])], [1.0], 0, 1.5, 1),
                 ("vector_top_k", [["a", [1.0]]], [1.0], 0, 1, 1), ("vector_top_k", [(1, [1.0])], [1.0], 0, 1, 1),
                 ("vector_top_k", [("a", [1.0])], [1.0], 256, 1, 1),
                 ("binary_top_k", [("a", [-1])], [1], 1, 1), ("binary_top_k", [("a", [1.5])], [1], 1, 1),
                 ("multi_vector_score", [[1.0]], [1.0], 0), ("flat_search", Atom("not_a_resource"), [1.0], 1),
                 ("flat_insert_many", 7, []), ("mv_search", 7, [[1.0]], 1)]:
        with pytest.raises(BadArg):
            nif.call(*args)


def test_validation_errors_that_need_no_device(nif):
    rows = [("b", [1.0, 0.0]), ("a", [1.0, 0.0]), ("c", [0.0, 1.0])]
    assert nif.call("vector_top_k", rows, [1.0, 0.0], 9, 2, 2) == (ERROR, "unknown metric")          # hardening_test:98-101
    assert nif.call("vector_top_k", rows, [1.0, 0.0], 0, 0, 2) == (ERROR, "invalid prefix dimensions")  # :119
    assert nif.call("vector_top_k", rows, [1.0, 0.0], 0, 3, 2) == (ERROR, "invalid prefix dimensions")
    assert nif.call("vector_top_k", rows, [float("nan"), 0.0], 0, 1, 2) == (ERROR, "vector contains a non-finite value")
    assert nif.call("vector_top_k", [], [1.0, 0.0], 0, 2, 2) == (OK, [])
    assert nif.call("binary_top_k", [], [], 0, 1) == (ERROR, "dimensions must be positive")          # search.rs:246-260
    assert nif.call("binary_top_k", [], [0], 1, 1) == (OK, [])
    assert nif.call("binary_top_k", [], [2 ** 64 - 1, 1], 65, 1) == (OK, [])                             # bignum words decode
    assert nif.call("binary_top_k", [("a", [1, 2])], [1], 64, 1) == (ERROR, "dimension mismatch")
    assert nif.call("multi_vector_score", [], [[1.0], [2.0]], 0) == (OK, 0.0)                            # multi_vector.rs:45-48
    assert nif.call("multi_vector_score", [[1.0]], [], 0) == (OK, 0.0)
    assert nif.call("multi_vector_score", [[1.0]], [[1.0, 2.0]], 0) == (ERROR, "dimension mismatch")
    assert nif.call("multi_vector_score", [[1.0]], [[1.0]], 9) == (ERROR, "unknown metric")
    assert nif.call("multi_vector_top_k", [("a", [[1.0]])], [[]], 0, 1) == (ERROR, "vectors must not be empty")
    assert nif.call("multi_vector_top_k", [("b", [[1.0]]), ("a", [[2.0]])], [], 0, 10) == (OK, [("a", 0.0), ("b", 0.0)])  # :298-306


def test_muvera_argument_decoding_and_validation(nif):
    assert nif.call("muvera_encode_query", [], 2, 2, 1, 42, 2, Atom("nil")) == (ERROR, "empty vectors")
    assert nif.call("muvera_encode_query", [], 0, 2, 1, 42, 2, Atom("nil")) == (ERROR, "empty vectors")   # before the config checks
    assert nif.call("muvera_encode_query", [[1.0]], 2, 2, 1, 42, 2, Atom("nil")) == (ERROR, "dimension mismatch")
    assert nif.call("muvera_encode_document", [[1.0, 0.0]], 2, 2, 31, 42, 2, Atom("nil")) == (ERROR, "num_simhash_projections must be < 31")
    assert nif.call("muvera_encode_document", [[1.0, 0.0]], 2, 2, 1, 42, 2, 0) == (ERROR, "final_projection_dimension must be positive")
    assert nif.call("muvera_encode_query", [[1.0, 0.0]], 2, 2, 30, 2 ** 64 - 1, 2, Atom("nil")) == (ERROR, "fde dimension exceeds safety limit")
    with pytest.raises(BadArg):
        nif.call("muvera_encode_query", [[1.0, 0.0]], 2, 2, 1, 42, 2, Atom("none"))
    with pytest.raises(BadArg):
        nif.call("muvera_encode_query", [[1.0, 0.0]], 2, 2, 1, -1, 2, Atom("nil"))


def test_flat_new_without_a_device_raises_instead_of_falling_back(nif):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(NifException) as e:
        nif.call("flat_new_cosine")
    assert e.value.args[0][0] == ERROR and e.value.args[0][1].startswith("cuda: ")
    assert nif.call("mv_new", 2)[0] == ERROR


# ------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
def test_nif_level_known_answers_of_the_reference(nif):
    rows = [("b", [1.0, 0.0]), ("a", [1.0, 0.0]), ("c", [0.0, 1.0])]
    for code in range(9):                                                      # hardening_test:90-97
        st, hits = nif.call("vector_top_k", rows, [1.0, 0.0], code, 2, 2)
        assert st == OK and [h[0] for h in hits] == ["a", "b"], code
    assert nif.call("binary_top_k", [("b", [1]), ("a", [3])], [3], 2, 2) == (OK, [("a", 0.0), ("b", 1.0)])   # :103-110
    assert nif.call("binary_top_k", [("same", [2 ** 64 - 1, 1]), ("far", [0, 0])], [2 ** 64 - 1, 1], 65, 2) == \
        (OK, [("same", 0.0), ("far", 65.0)])                                   # search.rs:246-260
    for metric in ["l2", "l2_squared", "cosine", "inner_product", "negative_inner_product", "manhattan", "chebyshev",
                   "hamming", "jaccard"]:                                       # hardening_test:20-36
        idx = nif.call(f"flat_new_{metric}")
        assert isinstance(idx, Resource)
        assert nif.call("flat_insert_many", idx, [("b", [0.0, 1.0]), ("a", [1.0, 0.0]), ("c", [1.0, 0.0])]) == (OK, ())
        st, hits = nif.call("flat_search", idx, [1.0, 0.0], 2)
        assert st == OK and [h[0] for h in hits] == ["a", "c"], metric
        nif.L.mock_release(idx.term)
    assert nif.call("multi_vector_score", [[1.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]], 3) == (OK, 2.0)   # multi_vector.rs:193-206
    assert nif.call("multi_vector_score", [[1e19]] * 4, [[1e19]], 3) == (ERROR, "score overflow")                 # :251-258
    docs = [("b", [[1.0, 0.0]]), ("a", [[1.0, 0.0]]), ("c", [[-1.0, 0.0]])]
    assert nif.call("multi_vector_top_k", docs, [[1.0, 0.0]], 3, 2) == (OK, [("a", 1.0), ("b", 1.0)])           # :209-222


@gpu
def test_muvera_through_the_shim_equals_the_oracle(nif):
    vectors = [[1.0, 2.0], [3.0, 4.0], [-2.0, 0.0]]                              # muvera.rs:334-355
    assert nif.call("muvera_encode_query", vectors, 2, 1, 0, 0, 2, Atom("nil")) == (OK, [2.0, 6.0])
    assert nif.call("muvera_encode_document", vectors, 2, 1, 0, 0, 2, Atom("nil")) == (OK, [float(np.float32(2.0 / 3.0)), 2.0])
    rng = np.random.default_rng(1)
    doc = rng.standard_normal((9, 16)).astype(np.float32)
    for final in (Atom("nil"), 40):
        got = nif.call("muvera_encode_document", [t for t in doc], 16, 4, 3, 99, 8, final)
        assert got == oracle.muvera_encode(doc, 16, 4, 3, 99, 8, None if isinstance(final, Atom) else final, "document")


@gpu
def test_flat_lifecycle_and_pipelines_through_the_shim_match_the_oracle(nif):
    rng = np.random.default_rng(5)
    n, d = 3000, 96
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ids = [f"{(i * 7919) % n:05d}" for i in range(n)]
    q = rows[17] + 0.05 * rng.standard_normal(d).astype(np.float32)
    idx = nif.call("flat_new_cosine")
    assert nif.call("flat_reserve", idx, n) == (OK, ())
    assert nif.call("flat_insert_many", idx, [(ids[i], rows[i]) for i in range(n)]) == (OK, ())
    assert nif.call("flat_insert", idx, "zzz", np.zeros(d + 1)) == (ERROR, "dimension mismatch")
    st, hits = nif.call("flat_search", idx, q, 10)
    assert st == OK
    assert_hits_match(hits, oracle.flat_search_dense("cosine", rows, ids, q, 10)[1])
    # shaped: flat_search + Distance.result_values in one call
    st, shaped = nif.call("flat_search_shaped", idx, q, 10, 2, 0)
    assert st == OK and [(h[0], h[1]) for h in shaped] == hits
    for _id, raw, score, dist in shaped:
        assert score == raw and dist == 1.0 - raw                              # vettore_distance.ex:531-532 (f64 on the BEAM)
    st, shaped = nif.call("flat_search_shaped", idx, q, 3, 2, 1)
    assert all(s == (r + 1.0) / 2.0 for _i, r, s, _d in shaped)
    # delete, then the resident pipelines against the oracle composition
    assert nif.call("flat_delete", idx, ids[17]) == (OK, ())
    keep = [i for i in range(n) if i != 17]
    vectors = [(ids[i], rows[i]) for i in keep]
    from test_pipelines_gpu import ref_funnel, ref_quantized
    st, got = nif.call("flat_funnel_search", idx, q, 2, [32, 64], 200, 10)
    assert st == OK
    assert_hits_match(got, ref_funnel(vectors, q, 2, [32, 64], 200, 10))
    st, got = nif.call("flat_quantized_search", idx, q, 2, 1500, 10)           # beyond the 1024 collector
    assert st == OK
    assert_hits_match(got, ref_quantized(vectors, q, 2, 1500, 10))
    # by-value forms over the same data
    st, got = nif.call("vector_top_k", vectors[:500], q, 2, 48, 7)
    assert st == OK
    assert_hits_match(got, oracle.vector_top_k(vectors[:500], q, 2, 48, 7)[1])
    codes = [(i, np.array(oracle.compress_sign_bits(v), dtype=np.uint64)) for i, v in vectors[:800]]
    qb = np.array(oracle.compress_sign_bits(q), dtype=np.uint64)
    assert nif.call("binary_top_k", codes, qb, d, 25) == (OK, oracle.binary_top_k(codes, qb, d, 25)[1])
    nif.L.mock_release(idx.term)


@gpu
def test_multi_vector_resident_index_through_the_shim(nif):
    rng = np.random.default_rng(9)
    docs = [(f"doc-{i:03d}", rng.standard_normal((int(rng.integers(1, 9)), 24)).astype(np.float32)) for i in range(120)]
    qv = rng.standard_normal((5, 24)).astype(np.float32)
    st, mv = nif.call("mv_new", 3)
    assert st == OK
    assert nif.call("mv_insert_many", mv, [(i, [t for t in toks]) for i, toks in docs]) == (OK, ())
    st, hits = nif.call("mv_search", mv, [t for t in qv], 8)
    assert st == OK
    assert_hits_match(hits, oracle.multi_vector_top_k(docs, qv, 3, 8)[1])
    st, by_value = nif.call("multi_vector_top_k", [(i, [t for t in toks]) for i, toks in docs], [t for t in qv], 3, 8)
    assert st == OK and by_value == hits
    assert nif.call("mv_delete", mv, hits[0][0]) == (OK, ())
    st, hits2 = nif.call("mv_search", mv, [t for t in qv], 8)
    assert st == OK and hits2[0][0] != hits[0][0]
    nif.L.mock_release(mv.term)


@gpu
def test_sharded_resource_behind_the_same_nifs(nif, monkeypatch):
    """VETTORE_B200_GPUS=N: flat_new_* returns ONE resource spread over N shards (one per GPU; they share the
    device of a single-GPU box); insert / search / delete answer exactly like the single-GPU resource."""
    rng = np.random.default_rng(13)
    n, d = 2000, 48
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"{(i * 7919) % n:05d}" for i in range(n)]
    q = rng.standard_normal(d).astype(np.float32)
    one = nif.call("flat_new_l2")
    monkeypatch.setenv("VETTORE_B200_GPUS", "4")
    many = nif.call("flat_new_l2")
    for idx in (one, many):
        assert nif.call("flat_insert_many", idx, [(ids[i], rows[i]) for i in range(n)]) == (OK, ())
        assert nif.call("flat_delete", idx, ids[5]) == (OK, ())
    a, b = nif.call("flat_search", one, q, 12), nif.call("flat_search", many, q, 12)
    assert a == b and a[0] == OK
    keep = [i for i in range(n) if i != 5]
    assert_hits_match(a[1], oracle.flat_search_dense("l2", rows[keep], [ids[i] for i in keep], q, 12)[1])
    # the resident pipelines answer through the sharded resource too (every stage on all shards, merged per stage)
    for call in (("flat_funnel_search", q, 0, [16, 32], 100, 8), ("flat_quantized_search", q, 2, 300, 8)):
        x, y = nif.call(call[0], one, *call[1:]), nif.call(call[0], many, *call[1:])
        assert x == y and x[0] == OK and len(x[1]) == 8
    nif.L.mock_release(one.term)
    nif.L.mock_release(many.term)
