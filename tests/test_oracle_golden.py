"""Pins the CPU oracle to the reference's own known-answer tests.

Each test names the reference test it restates (paths under /root/reference/). Expected
values are the literals asserted there; where the reference compares against a full sort
or an f64 scalar oracle, the same comparison is rebuilt here independently in numpy.
"""
import math

import numpy as np
import pytest

import oracle
from helpers import METRICS, close, total_order_key

F32_MAX = float(np.finfo(np.float32).max)


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def err(x):
    assert x[0] == "error", x
    return x[1]


# ---------------------------------------------------------------- distances.rs tests
def test_computes_every_metric_and_rank_semantics():  # distances.rs:495-515
    l, r = [1.0, 0.0, 1.0], [0.0, 1.0, 1.0]
    assert ok(oracle.compute("l2_squared", l, r)) == 2.0
    assert abs(ok(oracle.compute("l2", l, r)) - math.sqrt(2.0)) < 1e-6
    assert ok(oracle.compute("cosine", l, r)) == 1.0
    assert ok(oracle.compute("inner_product", l, r)) == 1.0
    assert ok(oracle.compute("negative_inner_product", l, r)) == -1.0
    assert ok(oracle.compute("manhattan", l, r)) == 2.0
    assert ok(oracle.compute("chebyshev", l, r)) == 1.0
    assert ok(oracle.compute("hamming", l, r)) == 2.0
    assert abs(ok(oracle.compute("jaccard", l, r)) - 2.0 / 3.0) < 1e-6
    assert oracle.rank_value("inner_product", 2.0) == -2.0
    assert oracle.rank_value("cosine", 0.25) == 0.75
    assert oracle.similarity_value("negative_inner_product", -3.0) == 3.0


def test_validates_dimensions_normalization_and_finite_values():  # distances.rs:517-537
    assert err(oracle.compute("l2", [1.0], [1.0, 2.0])) == "dimension mismatch"
    assert list(ok(oracle.normalize_l2([3.0, 4.0]))) == [np.float32(0.6), np.float32(0.8)]
    assert list(ok(oracle.normalize_l2([0.0, 0.0]))) == [0.0, 0.0]
    assert ok(oracle.cosine([2.0, 0.0], [4.0, 0.0])) == 1.0
    assert ok(oracle.cosine([0.0, 0.0], [4.0, 0.0])) == 0.0
    n = ok(oracle.normalize_l2([F32_MAX, F32_MAX]))
    assert abs(n[0] - math.sqrt(0.5)) < 1e-6
    assert oracle.compute("inner_product", [F32_MAX], [F32_MAX], checked=True)[0] == "error"
    assert oracle.compute("hamming", [float("nan")], [0.0], checked=True)[0] == "error"


def test_packs_bits_and_masks_unused_coordinates():  # distances.rs:539-548
    left = oracle.compress_sign_bits([1.0, -1.0, 0.0])
    right = oracle.compress_sign_bits([-1.0, -1.0, 0.0])
    assert left == [5]
    assert ok(oracle.packed_hamming(left, right, 3)) == 1.0
    assert ok(oracle.packed_jaccard(left, right, 3)) == 0.5
    assert oracle.packed_hamming(left, right, 0)[0] == "error"
    assert oracle.packed_hamming(left, [], 3)[0] == "error"
    assert oracle.compress_sign_bits([-0.0]) == [1]  # `>= 0.0` is true for -0.0


def test_decodes_metric_codes():  # distances.rs:550-568
    for code in range(9):
        assert oracle.compute(code, [1.0], [1.0])[0] == "ok"
    assert err(oracle.compute(9, [1.0], [1.0])) == "unknown metric"
    assert err(oracle.compute(255, [1.0], [1.0])) == "unknown metric"


def test_simd_and_tail_kernels_match_scalar_oracles():  # distances.rs:570-609
    for n in range(0, 41):
        left = np.array([((i * 37 % 23) - 11.0) / 3.0 for i in range(n)], dtype=np.float32)
        right = np.array([((i * 19 % 29) - 14.0) / 5.0 for i in range(n)], dtype=np.float32)
        l64, r64 = left.astype(np.float64), right.astype(np.float64)
        e_dot = float(np.float32(sum(l64 * r64)))
        e_l2s = float(np.float32(sum((l64 - r64) ** 2)))
        e_l1 = float(np.float32(sum(abs(l64 - r64))))
        e_linf = float(max([np.float32(0.0)] + [abs(a - b) for a, b in zip(left, right)]))
        if n == 0:
            # reference kernels accept empty slices; compute() on equal empty lengths -> 0
            assert ok(oracle.compute("inner_product", left, right)) == 0.0
            continue
        assert close(ok(oracle.compute("inner_product", left, right)), e_dot, 2e-6)
        assert close(ok(oracle.compute("l2_squared", left, right)), e_l2s, 2e-6)
        assert close(ok(oracle.compute("manhattan", left, right)), e_l1, 2e-6)
        assert ok(oracle.compute("chebyshev", left, right)) == e_linf


def test_recovers_representable_results_after_f32_intermediate_overflow():  # distances.rs:611-635
    large = 1.0e20
    assert close(ok(oracle.compute("l2", [large], [0.0])), float(np.float32(large)), 1e-6)
    assert ok(oracle.compute("inner_product", [F32_MAX, F32_MAX], [2.0, -2.0])) == 0.0
    assert ok(oracle.compute("negative_inner_product", [F32_MAX, F32_MAX], [2.0, -2.0])) == 0.0
    assert err(oracle.compute("l2_squared", [large], [0.0])) == "metric overflow"
    assert oracle.compute("l2", [F32_MAX, F32_MAX], [0.0, 0.0])[0] == "error"
    assert oracle.compute("manhattan", [F32_MAX, F32_MAX], [0.0, 0.0])[0] == "error"
    assert oracle.compute("chebyshev", [F32_MAX], [-F32_MAX])[0] == "error"
    assert ok(oracle.compute("jaccard", [0.0, 0.0], [0.0, 0.0])) == 0.0


def test_cosine_obeys_numerical_invariants():  # distances.rs:637-673 (cosine part)
    assert ok(oracle.cosine([], [])) == 0.0
    assert err(oracle.cosine([1.0], [1.0, 2.0])) == "dimension mismatch"
    assert close(ok(oracle.cosine([2.0, 0.0], [-5.0, 0.0])), -1.0, 1e-6)
    assert close(ok(oracle.cosine([3.0, 4.0], [6.0, 8.0])), 1.0, 1e-6)
    assert list(ok(oracle.normalize_l2([]))) == []
    n = ok(oracle.normalize_l2([3.0, -4.0, 12.0])).astype(np.float64)
    assert close(float(np.float32(n @ n)), 1.0, 1e-6)
    for bad in (float("nan"), float("inf"), float("-inf")):
        assert oracle.normalize_l2([bad])[0] == "error"


def test_packed_distances_cover_word_boundaries_and_ignore_padding():  # distances.rs:675-707
    M = (1 << 64) - 1
    for dims in (1, 63, 64, 65, 127, 128, 129):
        words = (dims + 63) // 64
        left = [M] * words
        right = list(left)
        flipped = [0] + ([dims - 1] if dims > 1 else [])
        for c in flipped:
            right[c // 64] ^= 1 << (c % 64)
        if dims % 64:
            used = (1 << (dims % 64)) - 1
            right[words - 1] ^= (~used) & M
        assert ok(oracle.packed_hamming(left, right, dims)) == float(len(flipped))
        assert close(ok(oracle.packed_jaccard(left, right, dims)), len(flipped) / dims, 1e-6)
    assert ok(oracle.packed_jaccard([0], [0], 64)) == 0.0
    assert oracle.packed_jaccard([], [], 1)[0] == "error"


# ---------------------------------------------------------------- flat.rs tests
def test_flat_inserts_replaces_deletes_and_returns_stable_top_k():  # flat.rs:164-180
    idx = oracle.FlatIndex("l2")
    ok(idx.insert("b", [2.0])); ok(idx.insert("a", [0.0])); ok(idx.insert("c", [2.0]))
    assert ok(idx.search([1.0], 2)) == [("a", 1.0), ("b", 1.0)]
    ok(idx.insert("a", [10.0]))
    assert ok(idx.search([2.0], 1))[0][0] == "b"
    idx.delete("b")
    assert ok(idx.search([2.0], 1))[0][0] == "c"


def test_flat_batch_validation_is_atomic():  # flat.rs:182-196
    idx = oracle.FlatIndex("inner_product")
    ok(idx.insert("existing", [1.0, 0.0]))
    assert idx.insert_many([("valid", [0.0, 1.0]), ("invalid", [1.0])])[0] == "error"
    assert len(idx.vectors) == 1 and "valid" not in idx.vectors
    assert idx.insert("nan", [float("nan"), 0.0])[0] == "error"


def test_flat_rejects_invalid_queries_and_handles_empty_limits():  # flat.rs:198-206
    idx = oracle.FlatIndex("cosine")
    assert err(idx.insert("empty", [])) == "vector must not be empty"
    ok(idx.insert("a", [1.0, 0.0]))
    assert err(idx.search([1.0], 1)) == "dimension mismatch"
    assert err(idx.search([float("inf"), 0.0], 1)) == "vector contains a non-finite value"
    assert ok(idx.search([1.0, 0.0], 0)) == []


def _full_sort(vectors, query, metric, scorer):
    scored = [(i, scorer(v)) for i, v in vectors]
    scored.sort(key=lambda h: (total_order_key(oracle.rank_value(metric, h[1])), h[0].encode()))
    return scored


def flat_fixture():  # flat.rs:210-222
    vectors = [(f"v-{i:02d}", [np.float32(i - 25.0) / np.float32(9.0),
                               np.float32((i * 13 % 31) - 15.0) / np.float32(7.0),
                               0.0 if i % 2 == 0 else 1.0]) for i in range(51)]
    return vectors, [0.5, -1.25, 1.0]


@pytest.mark.parametrize("metric", METRICS)
def test_flat_exact_heap_matches_a_full_sort_for_all_metrics(metric):  # flat.rs:208-249
    vectors, query = flat_fixture()
    idx = oracle.FlatIndex(metric)
    ok(idx.insert_many(vectors))
    expected = _full_sort(vectors, query, metric, lambda v: ok(oracle.compute(metric, query, v)))
    for limit in (1, 7, 51, 100):
        assert ok(idx.search(query, limit)) == expected[:limit]


def test_flat_empty_batches_unknown_deletes_and_dimension_resets_are_total():  # flat.rs:251-267
    idx = oracle.FlatIndex("l2")
    assert idx.insert_many([]) == ("ok", ())
    assert idx.search([1.0], 10) == ("ok", [])
    idx.delete("missing")
    ok(idx.insert("one", [1.0]))
    idx.delete("missing")
    assert idx.dimension == 1
    idx.delete("one")
    assert idx.dimension is None
    ok(idx.insert("two", [1.0, 2.0]))
    assert idx.dimension == 2
    assert len(ok(idx.search([1.0, 2.0], 2 ** 64 - 1))) == 1


def test_flat_duplicate_batch_ids_replace_and_large_l2_stays_finite():  # flat.rs:269-281
    idx = oracle.FlatIndex("l2")
    ok(idx.insert_many([("same", [0.0]), ("same", [1.0e20])]))
    assert len(idx.vectors) == 1
    hit = ok(idx.search([0.0], 1))[0]
    assert hit[0] == "same" and math.isfinite(hit[1])


# ---------------------------------------------------------------- search.rs tests
def test_vector_top_k_handles_prefixes_similarity_and_ties():  # search.rs:158-173
    vectors = [("b", [1.0, 10.0]), ("a", [1.0, -10.0]), ("c", [-1.0, 0.0])]
    assert ok(oracle.vector_top_k(vectors, [1.0, 0.0], 0, 1, 2)) == [("a", 0.0), ("b", 0.0)]
    assert ok(oracle.vector_top_k(vectors, [1.0, 1.0], 3, 2, 1))[0][0] == "b"


def test_vector_top_k_rejects_bad_dimensions_and_values():  # search.rs:175-184
    assert err(oracle.vector_top_k([], [1.0], 0, 0, 1)) == "invalid prefix dimensions"
    assert err(oracle.vector_top_k([("a", [1.0])], [1.0, 2.0], 0, 2, 1)) == "dimension mismatch"
    assert err(oracle.vector_top_k([("a", [float("nan")])], [1.0], 0, 1, 1)) == "vector contains a non-finite value"


def test_binary_top_k_masks_padding_and_orders_ids():  # search.rs:186-203
    q = oracle.compress_sign_bits([1.0, -1.0, 1.0])
    vectors = [("b", oracle.compress_sign_bits([1.0, 1.0, 1.0])), ("a", oracle.compress_sign_bits([1.0, -1.0, 1.0]))]
    assert ok(oracle.binary_top_k(vectors, q, 3, 2)) == [("a", 0.0), ("b", 1.0)]


def search_fixture():  # search.rs:207-220
    f = np.float32
    vectors = [(f"id-{i:02d}", [f(i - 18.0) / f(7.0), f((i * 11 % 17) - 8.0) / f(5.0),
                                f((i * 7 % 13) - 6.0) / f(3.0), 0.0 if i % 3 == 0 else 1.0]) for i in range(37)]
    return vectors, [0.25, -0.75, 1.5, 0.0]


@pytest.mark.parametrize("code", range(9))
def test_vector_top_k_matches_full_sort_for_every_metric_and_limit(code):  # search.rs:205-232
    vectors, query = search_fixture()
    metric = METRICS[code]
    for dims in (1, 3, 4):
        def scorer(v):
            if metric == "cosine":
                return ok(oracle.cosine(query[:dims], v[:dims]))
            return ok(oracle.compute(metric, query[:dims], v[:dims]))
        expected = _full_sort(vectors, query, metric, scorer)
        for limit in (0, 1, 5, 37, 100):
            assert ok(oracle.vector_top_k(vectors, query, code, dims, limit)) == expected[:limit]


def test_vector_top_k_true_cosine_is_independent_f64():
    """The cosine branch (search.rs:56-60) must equal a numpy f64 cosine, not the flat dot."""
    vectors, query = search_fixture()
    q = np.asarray(query, dtype=np.float32).astype(np.float64)
    got = dict(ok(oracle.vector_top_k(vectors, query, 2, 4, 100)))
    for i, v in vectors:
        v = np.asarray(v, dtype=np.float32).astype(np.float64)
        n = math.sqrt(q @ q) * math.sqrt(v @ v)
        e = 0.0 if n == 0 else min(1.0, max(-1.0, (q @ v) / n))
        assert close(got[i], e, 1e-6)


def test_vector_top_k_validates_queries_and_only_reads_the_requested_prefix():  # search.rs:234-244
    nan = float("nan")
    assert oracle.vector_top_k([], [nan], 0, 1, 1)[0] == "error"
    assert oracle.vector_top_k([], [1.0], 0, 2, 1)[0] == "error"
    assert oracle.vector_top_k([("a", [1.0, nan])], [1.0, nan], 0, 1, 1) == ("ok", [("a", 0.0)])


def test_binary_top_k_validates_empty_batches_limits_and_word_boundaries():  # search.rs:246-260
    M = (1 << 64) - 1
    assert err(oracle.binary_top_k([], [], 0, 1)) == "dimensions must be positive"
    assert err(oracle.binary_top_k([], [], 1, 1)) == "dimension mismatch"
    assert oracle.binary_top_k([], [0], 1, 1) == ("ok", [])
    q = [M, 1]
    vectors = [("same", q), ("far", [0, 0])]
    assert oracle.binary_top_k(vectors, q, 65, 0) == ("ok", [])
    assert ok(oracle.binary_top_k(vectors, q, 65, 10)) == [("same", 0.0), ("far", 65.0)]
    assert oracle.binary_top_k([("bad", [0])], q, 65, 1)[0] == "error"


def test_stable_ties_do_not_depend_on_candidate_order():  # search.rs:262-281
    fwd = [("c", [1.0]), ("a", [1.0]), ("b", [1.0])]
    exp = [("a", 0.0), ("b", 0.0)]
    assert ok(oracle.vector_top_k(fwd, [1.0], 0, 1, 2)) == exp
    assert ok(oracle.vector_top_k(fwd[::-1], [1.0], 0, 1, 2)) == exp


# ---------------------------------------------------------------- multi_vector.rs tests
def test_mv_scores_similarity_and_distance_metrics():  # multi_vector.rs:193-206
    q = [[1.0, 0.0], [0.0, 1.0]]
    d = [[1.0, 0.0], [0.0, 1.0]]
    for code in (3, 4, 2, 0):
        assert ok(oracle.multi_vector_score(q, d, code)) == 2.0
    assert ok(oracle.multi_vector_score([], d, 0)) == 0.0
    assert ok(oracle.multi_vector_score(q, [], 0)) == 0.0


def test_mv_top_k_is_stable_and_rejects_bad_shapes():  # multi_vector.rs:208-222
    q = [[1.0, 0.0]]
    docs = [("b", [[1.0, 0.0]]), ("a", [[1.0, 0.0]]), ("c", [[-1.0, 0.0]])]
    assert ok(oracle.multi_vector_top_k(docs, q, 3, 2)) == [("a", 1.0), ("b", 1.0)]
    assert oracle.multi_vector_score(q, [[1.0]], 3)[0] == "error"
    assert oracle.multi_vector_score([[float("nan"), 0.0]], q, 3)[0] == "error"


def _mv_score_indep(q, d, metric):  # multi_vector.rs:172-191 (independent composition)
    total = np.float32(0.0)
    for qv in q:
        sims = []
        for dv in d:
            raw = ok(oracle.cosine(qv, dv)) if metric == "cosine" else ok(oracle.compute(metric, qv, dv))
            sims.append(oracle.similarity_value(metric, raw))
        total = np.float32(total + np.float32(max(sims, key=total_order_key)))
    return float(total)


@pytest.mark.parametrize("code", range(9))
def test_mv_every_metric_matches_an_independent_maxsim_oracle(code):  # multi_vector.rs:224-238
    q = [[1.0, -0.5, 0.0], [0.0, 1.0, 1.0]]
    d = [[1.0, 0.0, 0.0], [0.0, 1.0, -1.0], [-1.0, 0.5, 1.0]]
    assert abs(ok(oracle.multi_vector_score(q, d, code)) - _mv_score_indep(q, d, METRICS[code])) <= 1e-6


def test_mv_validates_the_nonempty_side_even_when_the_other_side_is_empty():  # multi_vector.rs:240-248
    nan, inf = float("nan"), float("inf")
    assert err(oracle.multi_vector_score([], [[]], 0)) == "vectors must not be empty"
    assert oracle.multi_vector_score([], [[nan]], 0)[0] == "error"
    assert oracle.multi_vector_score([[]], [], 0)[0] == "error"
    assert oracle.multi_vector_score([[inf]], [], 0)[0] == "error"
    assert oracle.multi_vector_top_k([], [[]], 0, 1)[0] == "error"
    assert oracle.multi_vector_top_k([], [[nan]], 0, 1)[0] == "error"


def test_mv_detects_total_score_overflow_after_finite_pair_scores():  # multi_vector.rs:250-258
    assert err(oracle.multi_vector_score([[1.0e19]] * 4, [[1.0e19]], 3)) == "score overflow"


def mv_fixture():  # multi_vector.rs:262-273
    f = np.float32
    docs = [(f"doc-{i:02d}", [[f(i - 12.0) / f(5.0), 1.0], [0.0, f((i * 7 % 11) - 5.0) / f(3.0)]]) for i in range(25)]
    return docs, [[1.0, 0.0], [0.0, 1.0]]


@pytest.mark.parametrize("code", range(9))
def test_mv_batched_top_k_matches_full_sort_for_all_metrics_and_limits(code):  # multi_vector.rs:260-296
    docs, q = mv_fixture()
    scored = [(i, ok(oracle.multi_vector_score(q, vs, code))) for i, vs in docs]
    scored.sort(key=lambda h: (-total_order_key(h[1]), h[0].encode()))
    for limit in (0, 1, 7, 25, 100):
        assert ok(oracle.multi_vector_top_k(docs, q, code, limit)) == scored[:limit]


def test_mv_empty_queries_still_validate_documents_and_order_zero_score_ties():  # multi_vector.rs:298-306
    docs = [("b", [[1.0]]), ("a", [[2.0]])]
    assert oracle.multi_vector_top_k(docs, [], 0, 10) == ("ok", [("a", 0.0), ("b", 0.0)])


# ---------------------------------------------------------------- Elixir-level NIF answers
def test_nif_level_known_answers():  # test/vector_algorithms_hardening_test.exs:90-121
    vectors = [("b", [1.0, 0.0]), ("a", [1.0, 0.0]), ("c", [0.0, 1.0])]
    for code in range(9):
        assert [h[0] for h in ok(oracle.vector_top_k(vectors, [1.0, 0.0], code, 2, 2))] == ["a", "b"]
    assert ok(oracle.binary_top_k([("b", [1]), ("a", [3])], [3], 2, 2)) == [("a", 0.0), ("b", 1.0)]
    assert err(oracle.vector_top_k(vectors, [1.0, 0.0], 9, 2, 2)) == "unknown metric"
    assert err(oracle.vector_top_k(vectors, [1.0, 0.0], 0, 0, 2)) == "invalid prefix dimensions"


@pytest.mark.parametrize("metric", METRICS)
def test_every_flat_metric_gives_stable_ties(metric):  # test/vector_algorithms_hardening_test.exs:20-36
    idx = oracle.FlatIndex(metric)
    ok(idx.insert_many([("b", [0.0, 1.0]), ("a", [1.0, 0.0]), ("c", [1.0, 0.0])]))
    assert [h[0] for h in ok(idx.search([1.0, 0.0], 2))] == ["a", "c"]


def test_mode_equivalence_fixture_exact_flat():  # test/vector_adversarial_test.exs:376-421 (exact leg)
    f = np.float32
    rows = [(f"id-{i:02d}", [f(i) / f(10.0), f(7 * i % 17) / f(5.0), f(11 * i % 19) / f(7.0), float(i % 3)]) for i in range(64)]
    q = [2.25, 1.5, 0.75, 1.0]
    idx = oracle.FlatIndex("l2")
    ok(idx.insert_many(rows))
    exact = [h[0] for h in ok(idx.search(q, 10))]
    # full-candidate funnel (stages [2, 4], candidates 64) and full-candidate quantized
    # must reproduce the exact id list (collection.ex:244-295 composition of the NIFs).
    stage = ok(oracle.vector_top_k(rows, q, 0, 2, 64))
    keep = {h[0] for h in stage}
    survivors = [r for r in rows if r[0] in keep]
    funnel = [h[0] for h in ok(oracle.vector_top_k(survivors, q, 0, 4, 10))]
    assert funnel == exact
    codes = [(i, oracle.compress_sign_bits(v)) for i, v in rows]
    cands = ok(oracle.binary_top_k(codes, oracle.compress_sign_bits(q), 4, 64))
    keep = {h[0] for h in cands}
    quant = [h[0] for h in ok(oracle.vector_top_k([r for r in rows if r[0] in keep], q, 0, 4, 10))]
    assert quant == exact
