"""MUVERA fixed-dimensional encoding: the oracle restatement against the reference's own
known answers (native/vettore/src/muvera.rs:225-418), the C ABI's validation strings (no device needed), and — on
the GPU — the batched device encode against the oracle for EQUALITY (the device keeps the reference's accumulation
orders, so there is no tolerance to state)."""
import numpy as np
import pytest

import oracle
from vettore_b200 import nifs

F32_MAX = float(np.finfo(np.float32).max)
CFG = dict(dimension=2, num_repetitions=2, num_simhash_projections=1, seed=42, projection_dimension=2,
           final_projection_dimension=None)   # muvera.rs:228-237


def enc(impl, vectors, mode, **over):
    c = {**CFG, **over}
    fn = impl.muvera_encode if impl is oracle else None
    if impl is oracle:
        return fn(vectors, c["dimension"], c["num_repetitions"], c["num_simhash_projections"], c["seed"],
                  c["projection_dimension"], c["final_projection_dimension"], mode)
    f = nifs.muvera_encode_query if mode == "query" else nifs.muvera_encode_document
    return f(vectors, c["dimension"], c["num_repetitions"], c["num_simhash_projections"], c["seed"],
             c["projection_dimension"], c["final_projection_dimension"])


# ------------------------------------------------------------------------------------------ oracle, CPU
def test_oracle_reproduces_the_reference_known_answers():
    vectors = [[1.0, 2.0], [3.0, 4.0], [-2.0, 0.0]]                       # muvera.rs:334-355
    ident = dict(num_repetitions=1, num_simhash_projections=0, seed=0)
    assert enc(oracle, vectors, "query", **ident) == ("ok", [2.0, 6.0])
    assert enc(oracle, vectors, "document", **ident) == ("ok", [float(np.float32(2.0 / 3.0)), 2.0])
    one = dict(dimension=1, projection_dimension=1, num_repetitions=1, num_simhash_projections=0)   # :277-293
    assert enc(oracle, [[F32_MAX], [F32_MAX]], "query", **one) == ("error", "encoding overflow")
    assert enc(oracle, [[F32_MAX], [F32_MAX]], "document", **one) == ("ok", [F32_MAX])
    v2 = [[1.0, 0.0], [0.0, 1.0]]                                          # :240-249
    q, d = enc(oracle, v2, "query"), enc(oracle, v2, "document")
    assert q[0] == "ok" and d[0] == "ok" and q == enc(oracle, v2, "query") and q != d and len(q[1]) == 8
    assert len(enc(oracle, [[1.0, 2.0]], "query", projection_dimension=3, final_projection_dimension=5)[1]) == 5   # :252-259
    for ks in range(5):                                                    # :381-389
        r = enc(oracle, [[1.0, -2.0]], "query", num_repetitions=3, num_simhash_projections=ks, projection_dimension=5)
        assert r[0] == "ok" and len(r[1]) == 3 * (1 << ks) * 5
    v3 = [[1.0, 0.0], [0.0, 1.0], [-1.0, 0.5]]                             # :358-377
    assert enc(oracle, v3, "query") == enc(oracle, v3[::-1], "query")
    a, b = enc(oracle, v3, "document")[1], enc(oracle, v3[::-1], "document")[1]
    assert all(abs(x - y) <= 1e-6 for x, y in zip(a, b))
    assert enc(oracle, v3, "query") != enc(oracle, v3, "query", seed=43)


ERRORS = [  # muvera.rs:262-276, 296-331
    (([], "query", {}), "empty vectors"),
    (([[1.0]], "query", {}), "dimension mismatch"),
    (([[float("nan"), 0.0]], "query", {}), "vector contains a non-finite value"),
    (([[1.0, 0.0]], "query", dict(num_simhash_projections=30)), "fde dimension exceeds safety limit"),
    (([[1.0, 0.0]], "query", dict(num_simhash_projections=31)), "num_simhash_projections must be < 31"),
    (([[1.0, 0.0]], "query", dict(final_projection_dimension=0)), "final_projection_dimension must be positive"),
    (([[1.0, 0.0]], "query", dict(dimension=0)), "dimension must be positive"),
    (([[1.0, 0.0]], "query", dict(num_repetitions=0)), "num_repetitions must be positive"),
    (([[1.0, 0.0]], "query", dict(projection_dimension=0)), "projection_dimension must be positive"),
    (([[1.0, 0.0]], "query", dict(num_repetitions=16_777_217, num_simhash_projections=0, projection_dimension=1)),
     "fde dimension exceeds safety limit"),
    (([[1.0, 0.0]], "query", dict(final_projection_dimension=16_777_217)), "fde dimension exceeds safety limit"),
]


@pytest.mark.parametrize("args,msg", ERRORS, ids=[m[1][:28] + str(i) for i, m in enumerate(ERRORS)])
def test_validation_errors_match_the_reference_in_oracle_and_c_abi(args, msg):
    vectors, mode, over = args
    if msg == "final_projection_dimension must be positive":
        # Option<usize>: Some(0) must be spelled explicitly (None would be "no count sketch")
        c = {**CFG, **over}
        assert oracle.lib().vo_muvera_encode  # restated path: has_final = 1, final_dim = 0
        import ctypes as C
        vals, off = oracle._ragged_f32(vectors)
        out, n = np.zeros(8, np.float32), C.c_size_t()
        rc = oracle.lib().vo_muvera_encode(oracle._p(vals, C.c_float), oracle._p(off, C.c_uint64), C.c_size_t(1), C.c_size_t(2),
                                           C.c_size_t(2), C.c_size_t(1), C.c_uint64(42), C.c_size_t(2), C.c_int(1), C.c_size_t(0),
                                           C.c_int(0), oracle._p(out, C.c_float), C.c_size_t(8), C.byref(n))
        assert rc != 0 and oracle._err() == ("error", msg)
        from vettore_b200 import _lib
        dv = np.array([0, 1], dtype=np.uint64)
        rc = _lib.lib().vb_muvera_encode(1, vals.ctypes.data_as(C.POINTER(C.c_float)), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         dv.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 2, 1, 42, 2, 1, 0, 0,
                                         out.ctypes.data_as(C.POINTER(C.c_float)), 8, C.byref(n))
        assert rc == 1 and _lib.last_error() == msg
        return
    assert enc(oracle, vectors, mode, **over) == ("error", msg)
    assert enc(nifs, vectors, mode, **over) == ("error", msg)          # comes back before any device is touched


# ------------------------------------------------------------------------------------------ device, GPU
@pytest.mark.gpu
def test_device_encode_reproduces_the_reference_known_answers():
    vectors = [[1.0, 2.0], [3.0, 4.0], [-2.0, 0.0]]
    ident = dict(num_repetitions=1, num_simhash_projections=0, seed=0)
    assert enc(nifs, vectors, "query", **ident) == ("ok", [2.0, 6.0])
    assert enc(nifs, vectors, "document", **ident) == ("ok", [float(np.float32(2.0 / 3.0)), 2.0])
    one = dict(dimension=1, projection_dimension=1, num_repetitions=1, num_simhash_projections=0)
    assert enc(nifs, [[F32_MAX], [F32_MAX]], "query", **one) == ("error", "encoding overflow")
    assert enc(nifs, [[F32_MAX], [F32_MAX]], "document", **one) == ("ok", [F32_MAX])
    v2 = [[1.0, 0.0], [0.0, 1.0]]
    for mode in ("query", "document"):
        assert enc(nifs, v2, mode) == enc(oracle, v2, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["query", "document"])
@pytest.mark.parametrize("dim,reps,ks,pdim,final", [(16, 3, 2, 16, None), (24, 4, 3, 8, None), (128, 20, 5, 16, None),
                                                    (32, 5, 4, 8, 100), (7, 2, 0, 3, 5), (64, 10, 6, 64, 2048)])
def test_device_batch_encode_is_bit_identical_to_the_oracle(mode, dim, reps, ks, pdim, final):
    rng = np.random.default_rng(dim * 1000 + reps)
    docs = [rng.standard_normal((int(rng.integers(1, 40)), dim)).astype(np.float32) for _ in range(23)]
    st, got = nifs.muvera_encode_batch(docs, dim, reps, ks, 2026, pdim, final, mode)
    assert st == "ok", got
    for i, d in enumerate(docs):
        ref = oracle.muvera_encode(d, dim, reps, ks, 2026, pdim, final, mode)
        assert ref[0] == "ok"
        assert got[i].tolist() == ref[1], (i, mode)                          # equality, not tolerance
    # one document through the reference-shaped single call == its row of the batch
    single = (nifs.muvera_encode_query if mode == "query" else nifs.muvera_encode_document)(docs[3], dim, reps, ks, 2026, pdim, final)
    assert single == ("ok", got[3].tolist())


@pytest.mark.gpu
def test_device_encode_feeds_the_flat_inner_product_index():
    """Query FDE . document FDE approximates MaxSim (the reason the encoding exists): the document that IS the
    query's tokens must rank first among random documents under the flat inner-product scan."""
    rng = np.random.default_rng(3)
    dim, reps, ks, pdim = 32, 20, 3, 8
    docs = [rng.standard_normal((12, dim)).astype(np.float32) for _ in range(200)]
    docs = [d / np.linalg.norm(d, axis=1, keepdims=True) for d in docs]
    st, fde = nifs.muvera_encode_batch(docs, dim, reps, ks, 7, pdim, None, "document")
    assert st == "ok"
    idx = nifs.flat_new_inner_product()
    assert nifs.flat_insert_matrix(idx, [f"d{i:03d}" for i in range(200)], fde) == ("ok", ())
    st, q = nifs.muvera_encode_query(docs[77], dim, reps, ks, 7, pdim, None)
    assert st == "ok"
    st, hits = nifs.flat_search(idx, q, 3)
    assert st == "ok" and hits[0][0] == "d077"
