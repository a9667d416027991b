"""GPU parity: the CUDA path behind the C ABI (vettore_b200.nifs) against the CPU oracle,
on the reference's own fixtures and on seeded random inputs. Integer / id results must be
equal; float values within 1e-5 (|a-b| <= 1e-5 * max(1, |a|, |b|), distances.rs:487-493)
with id order equal except for ties inside that tolerance."""
import math

import numpy as np
import pytest

import oracle
from helpers import METRICS, assert_hits_match, close
from test_oracle_golden import flat_fixture, search_fixture
from vettore_b200 import nifs

pytestmark = pytest.mark.gpu


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def err(x):
    assert x[0] == "error", x
    return x[1]


def new_index(metric):
    return getattr(nifs, f"flat_new_{metric}")()


# ------------------------------------------------------------------ reference unit tests
def test_inserts_replaces_deletes_and_returns_stable_top_k():  # flat.rs:164-180
    idx = nifs.flat_new_l2()
    ok(nifs.flat_insert(idx, "b", [2.0])); ok(nifs.flat_insert(idx, "a", [0.0])); ok(nifs.flat_insert(idx, "c", [2.0]))
    assert ok(nifs.flat_search(idx, [1.0], 2)) == [("a", 1.0), ("b", 1.0)]
    ok(nifs.flat_insert(idx, "a", [10.0]))
    assert ok(nifs.flat_search(idx, [2.0], 1))[0][0] == "b"
    ok(nifs.flat_delete(idx, "b"))
    assert ok(nifs.flat_search(idx, [2.0], 1))[0][0] == "c"


def test_batch_validation_is_atomic():  # flat.rs:182-196
    idx = nifs.flat_new_inner_product()
    ok(nifs.flat_insert(idx, "existing", [1.0, 0.0]))
    assert err(nifs.flat_insert_many(idx, [("valid", [0.0, 1.0]), ("invalid", [1.0])])) == "dimension mismatch"
    assert nifs.flat_info(idx) == (1, 2)
    assert [h[0] for h in ok(nifs.flat_search(idx, [0.0, 1.0], 10))] == ["existing"]
    assert err(nifs.flat_insert(idx, "nan", [float("nan"), 0.0])) == "vector contains a non-finite value"


def test_rejects_invalid_queries_and_handles_empty_limits():  # flat.rs:198-206
    idx = nifs.flat_new_cosine()
    assert err(nifs.flat_insert(idx, "empty", [])) == "vector must not be empty"
    ok(nifs.flat_insert(idx, "a", [1.0, 0.0]))
    assert err(nifs.flat_search(idx, [1.0], 1)) == "dimension mismatch"
    assert err(nifs.flat_search(idx, [float("inf"), 0.0], 1)) == "vector contains a non-finite value"
    assert ok(nifs.flat_search(idx, [1.0, 0.0], 0)) == []
    assert ok(nifs.flat_search(idx, [1.0], 0)) == []  # limit 0 short-circuits validation (flat.rs:97-99)


@pytest.mark.parametrize("metric", METRICS)
def test_exact_heap_matches_a_full_sort_for_all_metrics(metric):  # flat.rs:208-249
    vectors, query = flat_fixture()
    idx = new_index(metric)
    ok(nifs.flat_insert_many(idx, vectors))
    ref = oracle.FlatIndex(metric)
    ok(ref.insert_many(vectors))
    for limit in (1, 7, 51, 100):
        assert_hits_match(ok(nifs.flat_search(idx, query, limit)), ok(ref.search(query, limit)))


def test_empty_batches_unknown_deletes_and_dimension_resets_are_total():  # flat.rs:251-267
    idx = nifs.flat_new_l2()
    assert nifs.flat_insert_many(idx, []) == ("ok", ())
    assert nifs.flat_search(idx, [1.0], 10) == ("ok", [])
    ok(nifs.flat_delete(idx, "missing"))
    ok(nifs.flat_insert(idx, "one", [1.0]))
    ok(nifs.flat_delete(idx, "missing"))
    assert nifs.flat_info(idx) == (1, 1)
    ok(nifs.flat_delete(idx, "one"))
    assert nifs.flat_info(idx) == (0, None)
    ok(nifs.flat_insert(idx, "two", [1.0, 2.0]))
    assert nifs.flat_info(idx) == (1, 2)
    assert len(ok(nifs.flat_search(idx, [1.0, 2.0], 2 ** 64 - 1))) == 1


def test_duplicate_batch_ids_replace_deterministically_and_large_l2_stays_finite():  # flat.rs:269-281
    idx = nifs.flat_new_l2()
    ok(nifs.flat_insert_many(idx, [("same", [0.0]), ("same", [1.0e20])]))
    assert nifs.flat_info(idx)[0] == 1
    hit = ok(nifs.flat_search(idx, [0.0], 1))[0]
    assert hit[0] == "same" and math.isfinite(hit[1]) and close(hit[1], 1.0e20, 1e-6)


@pytest.mark.parametrize("metric", METRICS)
def test_every_flat_metric_gives_stable_ties(metric):  # vector_algorithms_hardening_test.exs:20-36
    idx = new_index(metric)
    ok(nifs.flat_insert_many(idx, [("b", [0.0, 1.0]), ("a", [1.0, 0.0]), ("c", [1.0, 0.0])]))
    assert [h[0] for h in ok(nifs.flat_search(idx, [1.0, 0.0], 2))] == ["a", "c"]


def test_overflow_recovery_and_metric_overflow():  # distances.rs:611-635 through the index
    F = float(np.finfo(np.float32).max)
    idx = nifs.flat_new_inner_product()
    ok(nifs.flat_insert(idx, "x", [F, F]))
    assert ok(nifs.flat_search(idx, [2.0, -2.0], 1)) == [("x", 0.0)]
    idx = nifs.flat_new_l2_squared()
    ok(nifs.flat_insert(idx, "x", [1.0e20]))
    assert err(nifs.flat_search(idx, [0.0], 1)) == "metric overflow"
    idx = nifs.flat_new_manhattan()
    ok(nifs.flat_insert(idx, "x", [F, F]))
    assert err(nifs.flat_search(idx, [0.0, 0.0], 1)) == "metric overflow"
    idx = nifs.flat_new_chebyshev()
    ok(nifs.flat_insert(idx, "x", [F]))
    assert err(nifs.flat_search(idx, [-F], 1)) == "metric overflow"


# ------------------------------------------------------------------ by-value helpers
def test_vector_top_k_handles_prefixes_similarity_and_ties():  # search.rs:158-173
    vectors = [("b", [1.0, 10.0]), ("a", [1.0, -10.0]), ("c", [-1.0, 0.0])]
    assert ok(nifs.vector_top_k(vectors, [1.0, 0.0], 0, 1, 2)) == [("a", 0.0), ("b", 0.0)]
    assert ok(nifs.vector_top_k(vectors, [1.0, 1.0], 3, 2, 1))[0][0] == "b"


def test_vector_top_k_rejects_bad_dimensions_and_values():  # search.rs:175-184
    assert err(nifs.vector_top_k([], [1.0], 0, 0, 1)) == "invalid prefix dimensions"
    assert err(nifs.vector_top_k([("a", [1.0])], [1.0, 2.0], 0, 2, 1)) == "dimension mismatch"
    assert err(nifs.vector_top_k([("a", [float("nan")])], [1.0], 0, 1, 1)) == "vector contains a non-finite value"


@pytest.mark.parametrize("code", range(9))
def test_vector_top_k_matches_full_sort_for_every_metric_and_limit(code):  # search.rs:205-232
    vectors, query = search_fixture()
    for dims in (1, 3, 4):
        for limit in (0, 1, 5, 37, 100):
            assert_hits_match(ok(nifs.vector_top_k(vectors, query, code, dims, limit)),
                              ok(oracle.vector_top_k(vectors, query, code, dims, limit)))


def test_vector_top_k_validates_queries_and_only_reads_the_requested_prefix():  # search.rs:234-244
    nan = float("nan")
    assert nifs.vector_top_k([], [nan], 0, 1, 1)[0] == "error"
    assert nifs.vector_top_k([], [1.0], 0, 2, 1)[0] == "error"
    assert nifs.vector_top_k([("a", [1.0, nan])], [1.0, nan], 0, 1, 1) == ("ok", [("a", 0.0)])


def test_stable_ties_do_not_depend_on_candidate_order():  # search.rs:262-281
    fwd = [("c", [1.0]), ("a", [1.0]), ("b", [1.0])]
    exp = [("a", 0.0), ("b", 0.0)]
    assert ok(nifs.vector_top_k(fwd, [1.0], 0, 1, 2)) == exp
    assert ok(nifs.vector_top_k(fwd[::-1], [1.0], 0, 1, 2)) == exp


def test_nif_level_known_answers():  # vector_algorithms_hardening_test.exs:90-121
    vectors = [("b", [1.0, 0.0]), ("a", [1.0, 0.0]), ("c", [0.0, 1.0])]
    for code in range(9):
        assert [h[0] for h in ok(nifs.vector_top_k(vectors, [1.0, 0.0], code, 2, 2))] == ["a", "b"]
    assert ok(nifs.binary_top_k([("b", [1]), ("a", [3])], [3], 2, 2)) == [("a", 0.0), ("b", 1.0)]
    assert err(nifs.vector_top_k(vectors, [1.0, 0.0], 9, 2, 2)) == "unknown metric"
    assert err(nifs.vector_top_k(vectors, [1.0, 0.0], 0, 0, 2)) == "invalid prefix dimensions"


def test_vector_top_k_error_precedence_follows_row_order():
    """search.rs:51-60: the first failing row decides the error (overflow vs validation)."""
    rows = [("a", [1.0e20]), ("b", [float("nan")])]
    assert err(nifs.vector_top_k(rows, [0.0], 1, 1, 1)) == err(oracle.vector_top_k(rows, [0.0], 1, 1, 1)) == "metric overflow"
    rows = [("b", [float("nan")]), ("a", [1.0e20])]
    assert err(nifs.vector_top_k(rows, [0.0], 1, 1, 1)) == "vector contains a non-finite value"


# ------------------------------------------------------------------ Hamming (integer: exact)
def test_binary_top_k_masks_padding_and_orders_ids():  # search.rs:186-203
    q = nifs.compress_sign_bits([1.0, -1.0, 1.0])
    vectors = [("b", nifs.compress_sign_bits([1.0, 1.0, 1.0])), ("a", nifs.compress_sign_bits([1.0, -1.0, 1.0]))]
    assert ok(nifs.binary_top_k(vectors, q, 3, 2)) == [("a", 0.0), ("b", 1.0)]


def test_binary_top_k_validates_empty_batches_limits_and_word_boundaries():  # search.rs:246-260
    M = (1 << 64) - 1
    q = [M, 1]
    vectors = [("same", q), ("far", [0, 0])]
    assert nifs.binary_top_k(vectors, q, 65, 0) == ("ok", [])
    assert ok(nifs.binary_top_k(vectors, q, 65, 10)) == [("same", 0.0), ("far", 65.0)]
    assert err(nifs.binary_top_k([("bad", [0])], q, 65, 1)) == "dimension mismatch"


@pytest.mark.parametrize("dims", [1, 63, 64, 65, 127, 128, 129, 256, 300, 384, 576, 768, 1024, 2048, 2100, 4096, 4500, 8192])
@pytest.mark.parametrize("no_stream", [False, True])
def test_binary_top_k_random_codes_bit_exact(dims, no_stream, monkeypatch):  # distances.rs:675-707 + search.rs:76-92
    if no_stream:   # K3's register-staged kernel also for the shapes the TMA ring would take
        monkeypatch.setenv("VB_HAMMING_NO_STREAM", "1")
    rng = np.random.default_rng(dims)
    nw = (dims + 63) // 64
    n = 777
    codes = rng.integers(0, 2 ** 64, size=(n, nw), dtype=np.uint64)  # padding bits are garbage on purpose
    codes[5] = codes[9]  # force distance ties broken by id
    q = rng.integers(0, 2 ** 64, size=nw, dtype=np.uint64)
    ids = [f"id-{(i * 7919) % n:04d}" for i in range(n)]
    vectors = [(ids[i], [int(w) for w in codes[i]]) for i in range(n)]
    for limit in (1, 10, 100, 777, 1500):
        got = ok(nifs.binary_top_k(vectors, [int(w) for w in q], dims, limit))
        assert got == ok(oracle.binary_top_k(vectors, [int(w) for w in q], dims, limit))


# ------------------------------------------------------------------ seeded random parity
def _random_rows(n, d, seed, normalise=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if normalise:
        x = (x.astype(np.float64) / np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    return x


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d", [(1000, 384), (3001, 70), (513, 1030), (2000, 1600)])
def test_flat_search_random_parity(metric, n, d):
    rows = _random_rows(n, d, seed=n + d, normalise=metric in ("cosine", "inner_product"))
    if metric in ("hamming", "jaccard"):
        rows[rows < 0.3] = 0.0  # make truthiness informative
    ids = [f"{(i * 7919) % n:09d}" for i in range(n)]
    q = _random_rows(1, d, seed=99, normalise=True)[0]
    idx = new_index(metric)
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    for limit in (1, 10, 100):
        assert_hits_match(ok(nifs.flat_search(idx, q, limit)), ok(oracle.flat_search_dense(metric, rows, ids, q, limit)))


def test_config1_flat_cosine_10k_384_k10():
    """BASELINE.json configs[0]: 10k x 384 fp32 L2-normalised, single query, k=10."""
    rows = _random_rows(10_000, 384, seed=20_260_721)
    q = _random_rows(1, 384, seed=20_260_722)[0]
    ids = [f"{i:09d}" for i in range(10_000)]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    assert_hits_match(ok(nifs.flat_search(idx, q, 10)), ok(oracle.flat_search_dense("cosine", rows, ids, q, 10)))
    got = ok(nifs.flat_search_batch(idx, np.stack([q, rows[17], -q]), 10))
    for g, qq in zip(got, [q, rows[17], -q]):
        assert_hits_match(g, ok(oracle.flat_search_dense("cosine", rows, ids, qq, 10)))
    assert got[1][0][0] == ids[17]


def test_large_limit_uses_the_sort_path_and_limit_beyond_n():
    n, d = 5000, 96
    rows = _random_rows(n, d, seed=5)
    ids = [f"{i:09d}" for i in range(n)]
    q = _random_rows(1, d, seed=6)[0]
    idx = nifs.flat_new_l2()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    for limit in (1024, 1025, 3000, 5000, 10 ** 9):
        assert_hits_match(ok(nifs.flat_search(idx, q, limit)), ok(oracle.flat_search_dense("l2", rows, ids, q, limit)))


def test_mutations_keep_the_id_order_labels_valid():
    """Out-of-order inserts, upserts and deletes: ties must still resolve by id bytes."""
    rng = np.random.default_rng(3)
    idx = nifs.flat_new_l2()
    ref = oracle.FlatIndex("l2")
    names = [f"k{rng.integers(0, 10 ** 6):06d}" for _ in range(400)]
    for step, name in enumerate(names):
        v = [float(rng.integers(0, 3)), float(rng.integers(0, 3))]  # many exact distance ties
        ok(nifs.flat_insert(idx, name, v)); ok(ref.insert(name, v))
        if step % 7 == 3:
            victim = names[rng.integers(0, step + 1)]
            ok(nifs.flat_delete(idx, victim)); ok(ref.delete(victim))
        if step % 25 == 0:
            assert ok(nifs.flat_search(idx, [1.0, 1.0], 15)) == ok(ref.search([1.0, 1.0], 15))
    assert ok(nifs.flat_search(idx, [1.0, 1.0], 1000)) == ok(ref.search([1.0, 1.0], 1000))


def test_resident_prefix_top_k_matches_by_value_semantics():
    """Funnel stage over the resident matrix == vector_top_k over store.all (collection.ex:674-691)."""
    n, d = 2000, 128
    rows = _random_rows(n, d, seed=11, normalise=False)
    ids = [f"{(i * 31) % n:05d}" for i in range(n)]
    q = _random_rows(1, d, seed=12)[0]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    vectors = [(ids[i], rows[i]) for i in range(n)]
    for code in (0, 2, 3, 5):
        for dims in (5, 32, 127, 128):
            exp = ok(oracle.vector_top_k(vectors, q, code, dims, 50))
            assert_hits_match(ok(nifs.flat_prefix_top_k(idx, None, q, code, dims, 50)), exp)
    subset = ids[::3] + ["no-such-id"]
    sub_vectors = [(ids[i], rows[i]) for i in range(0, n, 3)]
    assert_hits_match(ok(nifs.flat_prefix_top_k(idx, subset, q, 2, 128, 10)), ok(oracle.vector_top_k(sub_vectors, q, 2, 128, 10)))
    assert err(nifs.flat_prefix_top_k(idx, None, q, 2, 0, 10)) == "invalid prefix dimensions"
    assert err(nifs.flat_prefix_top_k(idx, None, q, 9, 8, 10)) == "unknown metric"


@pytest.mark.parametrize("dim", [96, 127, 768])
def test_device_bulk_ingest_equals_host_ingest(dim):
    """vb_flat_insert_many_device (rows already in HBM) must leave the index exactly as insert_many would:
    same validation, upserts in place, duplicate ids in a batch -> last wins, sign codes follow."""
    import torch
    rng = np.random.default_rng(dim)
    n = 3000
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    ids = [f"r{i:05d}" for i in range(n)]
    ids[17] = ids[5]                       # duplicate id inside the batch: row 17 wins
    ids[2999] = ids[2998]
    host, devi = nifs.flat_new_l2(), nifs.flat_new_l2()
    assert nifs.flat_insert_matrix(host, ids, rows) == ("ok", ())
    d_rows = torch.from_numpy(rows).cuda()
    assert nifs.flat_insert_device(devi, ids, d_rows.data_ptr(), dim) == ("ok", ())
    assert nifs.flat_info(host) == nifs.flat_info(devi)
    # second batch: upserts of existing ids mixed with new ones
    rows2 = rng.standard_normal((200, dim)).astype(np.float32)
    ids2 = [ids[i * 3] if i % 2 else f"n{i:04d}" for i in range(200)]
    assert nifs.flat_insert_matrix(host, ids2, rows2) == ("ok", ())
    d_rows2 = torch.from_numpy(rows2).cuda()
    assert nifs.flat_insert_device(devi, ids2, d_rows2.data_ptr(), dim) == ("ok", ())
    for q in (rows[5], rows[17], rows2[1], rng.standard_normal(dim).astype(np.float32)):
        assert nifs.flat_search(devi, q, 20) == nifs.flat_search(host, q, 20)
        assert nifs.flat_quantized_search(devi, q, nifs.METRIC_CODE["l2"], 50, 10) == \
            nifs.flat_quantized_search(host, q, nifs.METRIC_CODE["l2"], 50, 10)
    # all-or-nothing validation on the device copy
    bad = rows2.copy()
    bad[150, dim // 2] = np.inf
    before = nifs.flat_info(devi)
    assert nifs.flat_insert_device(devi, [f"x{i}" for i in range(200)], torch.from_numpy(bad).cuda().data_ptr(), dim) == \
        ("error", "vector contains a non-finite value")
    assert nifs.flat_insert_device(devi, ["y"], d_rows2.data_ptr(), dim + 1) == ("error", "dimension mismatch")
    assert nifs.flat_insert_device(nifs.flat_new_l2(), ["y"], d_rows2.data_ptr(), 0) == ("error", "vector must not be empty")
    assert nifs.flat_info(devi) == before


@pytest.mark.parametrize("ints", [False, True])
@pytest.mark.parametrize("metric,n,d,k", [("cosine", 400_000, 64, 100), ("l2", 300_000, 96, 1000), ("inner_product", 250_000, 128, 333)])
def test_large_k_flat_search_many_rows(metric, n, d, k, ints):
    """k >= 32 takes the launch-wide pivot ladder (topk.cuh): every CTA prunes with a bound proven over the
    whole launch. Small-integer rows make the bound fall into huge groups of tied ranks."""
    rng = np.random.default_rng(n + k)
    if ints:
        rows = rng.integers(-2, 3, size=(n, d)).astype(np.float32)
        q = rng.integers(-2, 3, size=d).astype(np.float32)
    else:
        rows = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal(d).astype(np.float32)
    ids = [f"{(i * 7919) % n:07d}" for i in range(n)]
    idx = getattr(nifs, f"flat_new_{metric}")()
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    got = ok(nifs.flat_search(idx, q, k))
    exp = ok(oracle.flat_search_dense(metric, rows, ids, q, k))
    assert_hits_match(got, exp, exact_ids=ints)
    # prefix scoring over all rows (K4, register-staged kernel) with the same ladder; for metrics other than
    # cosine vector_top_k on a prefix is the flat search over the truncated rows (search.rs:56-60)
    h = d // 2
    got = ok(nifs.flat_prefix_top_k(idx, None, q, nifs.METRIC_CODE[metric], h, k))
    if metric != "cosine":
        exp = ok(oracle.flat_search_dense(metric, np.ascontiguousarray(rows[:, :h]), ids, q[:h].copy(), k))
        assert_hits_match(got, exp, exact_ids=ints)
    else:
        r64, q64 = rows[:, :h].astype(np.float64), q[:h].astype(np.float64)
        den = np.linalg.norm(r64, axis=1) * np.linalg.norm(q64)
        cos = np.where(den > 0, (r64 @ q64) / np.where(den > 0, den, 1.0), 0.0).clip(-1.0, 1.0).astype(np.float32)
        order = sorted(range(n), key=lambda i: (float(np.float32(1.0) - cos[i]), ids[i]))[:k]
        assert_hits_match(got, [(ids[i], float(cos[i])) for i in order], exact_ids=False)


def test_small_batch_large_k_uses_one_ladder_per_query(monkeypatch):
    """A batch below the K2 threshold runs K1 with one grid row per query: every query slot owns its pivot
    ladder, counters and bound."""
    rng = np.random.default_rng(11)
    n, d, nq, k = 200_000, 64, 5, 64
    rows = rng.standard_normal((n, d)).astype(np.float32)
    queries = rng.standard_normal((nq, d)).astype(np.float32)
    queries[2] = -queries[1]          # opposite rankings in neighbouring slots
    ids = [f"{i:06d}" for i in range(n)]
    idx = nifs.flat_new_inner_product()
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in range(nq):
        assert_hits_match(got[qi], ok(oracle.flat_search_dense("inner_product", rows, ids, queries[qi], k)))
