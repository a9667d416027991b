"""GPU parity of the resident funnel / quantized pipelines and the Collection mirror against
the reference composition of by-value NIFs (collection.ex:244-323), restated with the oracle."""
import numpy as np
import pytest

import oracle
from helpers import assert_hits_match
from vettore_b200 import nifs
from vettore_b200.collection import Collection, Embedding, normalize_l2

pytestmark = pytest.mark.gpu


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def ref_funnel(vectors, q, code, stages, candidates, limit):
    """collection.ex:674-691 + exact_rerank (:821-851) with the by-value NIFs."""
    cur = vectors
    for d in stages:
        keep = {h[0] for h in ok(oracle.vector_top_k(cur, q, code, d, candidates))}
        cur = [v for v in cur if v[0] in keep]
    return ok(oracle.vector_top_k(cur, q, code, len(q), limit))


def ref_quantized(vectors, q, code, candidates, limit):
    """collection.ex:699-713 + exact_rerank."""
    dims = len(q)
    codes = [(i, oracle.compress_sign_bits(v)) for i, v in vectors]
    keep = {h[0] for h in ok(oracle.binary_top_k(codes, oracle.compress_sign_bits(q), dims, candidates))}
    return ok(oracle.vector_top_k([v for v in vectors if v[0] in keep], q, code, dims, limit))


def adversarial_fixture():  # test/vector_adversarial_test.exs:376-421
    f = np.float32
    rows = [(f"id-{i:02d}", [f(i) / f(10.0), f(7 * i % 17) / f(5.0), f(11 * i % 19) / f(7.0), float(i % 3)]) for i in range(64)]
    return rows, [2.25, 1.5, 0.75, 1.0]


def test_full_candidate_modes_equal_exact_flat_ids():  # vector_adversarial_test.exs:376-421
    rows, q = adversarial_fixture()
    idx = nifs.flat_new_l2()
    ok(nifs.flat_insert_many(idx, rows))
    exact = [h[0] for h in ok(nifs.flat_search(idx, q, 10))]
    assert exact == [h[0] for h in ok(oracle.flat_search_dense("l2", np.array([r[1] for r in rows], np.float32), [r[0] for r in rows], q, 10))]
    assert [h[0] for h in ok(nifs.flat_funnel_search(idx, q, 0, [2, 4], 64, 10))] == exact
    assert [h[0] for h in ok(nifs.flat_quantized_search(idx, q, 0, 64, 10))] == exact


def test_db_level_known_answers():  # test/vector_db_test.exs:135-174
    c = Collection("l2", normalize="none")
    ok(c.put_many([Embedding("exact", [1.0, 0.0, 0.0]), Embedding("prefix", [1.0, 5.0, 0.0]), Embedding("far", [-1.0, 0.0, 0.0])]))
    res = ok(c.funnel_search([1.0, 0.0, 0.0], limit=1, candidates=2, stages=[1]))
    assert [r.id for r in res] == ["exact"]
    c = Collection("l2", normalize="none")
    ok(c.put_many([Embedding("exact", [1.0, 1.0]), Embedding("same_bits_far", [100.0, 100.0]), Embedding("opposite", [-1.0, -1.0])]))
    res = ok(c.quantized_search([1.0, 1.0], limit=1, candidates=2))
    assert [r.id for r in res] == ["exact"] and res[0].distance == 0.0
    assert nifs.compress_sign_bits([1.0, 1.0]) == [3]
    c = Collection("inner_product")
    ok(c.put_many([Embedding("both_axes", [0.5, 0.5], vectors=[[1.0, 0.0], [0.0, 1.0]]), Embedding("one_axis", [1.0, 0.0], vectors=[[1.0, 0.0]])]))
    res = ok(c.multi_vector_search([[1.0, 0.0], [0.0, 1.0]], limit=10))
    assert [(r.id, r.score, r.distance) for r in res] == [("both_axes", 2.0, None), ("one_axis", 1.0, None)]


def _rows(n, d, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d)).astype(np.float32)


@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product", "manhattan"])
@pytest.mark.parametrize("n,d,stages,cand", [(3000, 256, [64, 128], 200), (1500, 768, [192, 384, 768], 100), (900, 100, [33], 1024)])
def test_funnel_pipeline_matches_reference_composition(metric, n, d, stages, cand):
    rows = _rows(n, d, n + d)
    if metric == "cosine":
        rows = np.stack([normalize_l2(r) for r in rows])
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    q = normalize_l2(_rows(1, d, 3)[0]) if metric == "cosine" else _rows(1, d, 3)[0]
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    code = nifs.METRIC_CODE[metric]
    vectors = [(ids[i], rows[i]) for i in range(n)]
    cand = min(cand, n)
    got = ok(nifs.flat_funnel_search(idx, q, code, stages, cand, 10))
    assert_hits_match(got, ref_funnel(vectors, q, code, stages, cand, 10))


@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product"])
@pytest.mark.parametrize("n,d,cand", [(4000, 128, 100), (2500, 1024, 1000), (700, 200, 64)])
def test_quantized_pipeline_matches_reference_composition(metric, n, d, cand):
    rows = _rows(n, d, n * 3 + d)
    if metric == "cosine":
        rows = np.stack([normalize_l2(r) for r in rows])
    ids = [f"{(i * 104729) % n:06d}" for i in range(n)]
    q = normalize_l2(_rows(1, d, 5)[0])
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    code = nifs.METRIC_CODE[metric]
    vectors = [(ids[i], rows[i]) for i in range(n)]
    assert_hits_match(ok(nifs.flat_quantized_search(idx, q, code, cand, 10)), ref_quantized(vectors, q, code, cand, 10))
    # the code mirror must follow mutations: upsert a row to the query itself, delete another
    ok(nifs.flat_insert(idx, ids[7], q))
    ok(nifs.flat_delete(idx, ids[11]))
    vectors2 = [(i, (q if i == ids[7] else v)) for i, v in vectors if i != ids[11]]
    got = ok(nifs.flat_quantized_search(idx, q, code, cand, 10))
    assert_hits_match(got, ref_quantized(vectors2, q, code, cand, 10))
    if metric != "inner_product":  # an unnormalised row can out-score the query itself under raw dot
        assert got[0][0] == ids[7]


def test_collection_mirror_shapes_results_like_the_reference():
    c = Collection("cosine")
    rows = _rows(500, 64, 1)
    ok(c.put_many([Embedding(f"e{i:03d}", rows[i], value=f"v{i}", metadata={"i": i}) for i in range(500)]))
    q = rows[42] + 0.01
    res = ok(c.search(q, limit=5))
    assert res[0].id == "e042" and res[0].value == "v42" and res[0].metadata == {"i": 42}
    assert abs(res[0].score - (1.0 - res[0].distance)) < 1e-6 and res[0].metric == "cosine"
    assert [r.id for r in ok(c.funnel_search(q, limit=5, candidates=500, stages=[16, 64]))] == [r.id for r in res]
    assert [r.id for r in ok(c.quantized_search(q, limit=5, candidates=500))] == [r.id for r in res]
    c.delete("e042")
    assert ok(c.search(q, limit=5))[0].id != "e042"
    assert c.search(q, limit=0) == ("error", "invalid_limit")


@pytest.mark.parametrize("n,dims,k,narrow", [(300_000, 256, 1000, False), (250_000, 192, 500, False),
                                             (200_000, 64, 300, True), (150_000, 100, 1000, False)])
@pytest.mark.parametrize("no_stream", [False, True])
def test_hamming_large_k_over_many_rows(n, dims, k, narrow, no_stream, monkeypatch):
    """K3 with k > 64 (histogram thresholds) over enough rows that every CTA prunes many times; the expected order is
    (distance, id bytes) as binary_top_k's (search.rs:186-203). `narrow` packs the distances into few values so the
    k-th distance is a large tie. The grid is capped so each CTA meets many collector checkpoints."""
    monkeypatch.setenv("VB_HAMMING_MAX_GRID", "3")
    if no_stream:   # the register-staged kernel instead of the TMA ring
        monkeypatch.setenv("VB_HAMMING_NO_STREAM", "1")
    rng = np.random.default_rng(n + dims)
    nw = (dims + 63) // 64
    codes = rng.integers(0, 2 ** 63, size=(n, nw), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, nw), dtype=np.uint64)
    if narrow:
        codes &= np.uint64(0xFF)
    q = codes[17].copy()
    ids = [f"{(i * 7919) % n:07d}" for i in range(n)]
    mask = np.full(nw, np.uint64(0xFFFFFFFFFFFFFFFF))
    if dims % 64:
        mask[-1] = np.uint64((1 << (dims % 64)) - 1)
    dist = np.bitwise_count((codes ^ q) & mask).sum(axis=1).astype(np.int64)
    order = sorted(range(n), key=lambda i: (dist[i], ids[i]))[:k]
    got = ok(nifs.binary_top_k([(ids[i], codes[i]) for i in range(n)], q, dims, k))
    assert [h[0] for h in got] == [ids[i] for i in order]
    assert [h[1] for h in got] == [float(dist[i]) for i in order]


# ---- candidates / limit beyond the fused collector (1024): the reference has no such bound
# (collection.ex:509-510: candidates default to 10 x limit; limits up to 2^32 - 1) ------------------
@pytest.mark.parametrize("metric", ["cosine", "l2"])
@pytest.mark.parametrize("n,d,cand,limit", [(6000, 96, 1025, 10), (6000, 96, 5000, 200), (3000, 64, 3000, 1500)])
def test_quantized_pipeline_beyond_1024_candidates(metric, n, d, cand, limit):
    rows = _rows(n, d, n + 11 * d)
    if metric == "cosine":
        rows = np.stack([normalize_l2(r) for r in rows])
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    q = normalize_l2(_rows(1, d, 9)[0])
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    code = nifs.METRIC_CODE[metric]
    vectors = [(ids[i], rows[i]) for i in range(n)]
    assert_hits_match(ok(nifs.flat_quantized_search(idx, q, code, cand, limit)), ref_quantized(vectors, q, code, cand, limit))


@pytest.mark.parametrize("metric", ["cosine", "inner_product"])
@pytest.mark.parametrize("n,d,stages,cand,limit", [(6000, 128, [32, 64], 1025, 10), (6000, 128, [64], 5000, 200),
                                                   (2500, 64, [16, 32], 2500, 1100)])
def test_funnel_pipeline_beyond_1024_candidates(metric, n, d, stages, cand, limit):
    rows = _rows(n, d, n + 13 * d)
    if metric == "cosine":
        rows = np.stack([normalize_l2(r) for r in rows])
    ids = [f"{(i * 104729) % n:06d}" for i in range(n)]
    q = normalize_l2(_rows(1, d, 4)[0])
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    code = nifs.METRIC_CODE[metric]
    vectors = [(ids[i], rows[i]) for i in range(n)]
    assert_hits_match(ok(nifs.flat_funnel_search(idx, q, code, stages, cand, limit)),
                      ref_funnel(vectors, q, code, stages, cand, limit))


def test_collection_default_candidates_with_limit_200():
    """`limit: 200` with the default candidates (10 x limit = 2000) used to hit the 1024 wall."""
    n, d = 5000, 64
    rows = _rows(n, d, 77)
    c = Collection("cosine")
    ok(c.put_many([Embedding(f"e{i:05d}", rows[i]) for i in range(n)]))
    q = rows[5] + 0.05
    vectors = [(f"e{i:05d}", normalize_l2(rows[i])) for i in range(n)]
    qn = normalize_l2(q)
    got = ok(c.quantized_search(q, limit=200))
    assert_hits_match([(r.id, r.score) for r in got], ref_quantized(vectors, qn, 2, 2000, 200))
    got = ok(c.funnel_search(q, limit=200, stages=[32]))
    assert_hits_match([(r.id, r.score) for r in got], ref_funnel(vectors, qn, 2, [32], 2000, 200))


@pytest.mark.parametrize("metric,stages", [("cosine", [64, 128]), ("l2", [50]), ("inner_product", [128])])
def test_funnel_stage_one_over_the_dense_prefix_mirror(metric, stages):
    """Enough rows for the dense mirror of the first-stage columns (flat_index.h: prefix_wanted): same answers as
    the reference composition, before and after inserts, an in-place upsert and deletes touch the mirror."""
    n, d, cand, limit = 40_000, 256, 300, 10
    rows = _rows(n, d, 1234)
    if metric == "cosine":
        rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ids = [f"{(i * 104729) % n:06d}" for i in range(n)]
    q = normalize_l2(_rows(1, d, 8)[0])
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    code = nifs.METRIC_CODE[metric]
    vectors = {ids[i]: rows[i] for i in range(n)}
    got = ok(nifs.flat_funnel_search(idx, q, code, stages, cand, limit))
    assert_hits_match(got, ref_funnel(list(vectors.items()), q, code, stages, cand, limit))
    # the mirror now exists: mutate through every path that must keep it in sync
    best = got[0][0]
    ok(nifs.flat_delete(idx, best))                                   # hole filled by the last row
    del vectors[best]
    up = got[1][0]
    newv = normalize_l2(_rows(1, d, 99)[0])
    ok(nifs.flat_insert(idx, up, newv))                               # in-place upsert
    vectors[up] = newv
    extra = _rows(50, d, 77)
    extra[0, :stages[0]] = q[:stages[0]] * 3.0                        # a strong first-stage candidate among the appended rows
    if metric == "cosine":
        extra = (extra / np.linalg.norm(extra.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ok(nifs.flat_insert_many(idx, [(f"new-{i:02d}", extra[i]) for i in range(50)]))
    for i in range(50):
        vectors[f"new-{i:02d}"] = extra[i]
    got = ok(nifs.flat_funnel_search(idx, q, code, stages, cand, limit))
    assert_hits_match(got, ref_funnel(list(vectors.items()), q, code, stages, cand, limit))
    # a different first stage rebuilds the mirror with the new width
    other = [stages[0] // 2] + stages
    got = ok(nifs.flat_funnel_search(idx, q, code, other, cand, limit))
    assert_hits_match(got, ref_funnel(list(vectors.items()), q, code, other, cand, limit))
