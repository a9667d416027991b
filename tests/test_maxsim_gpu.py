"""GPU parity of ColBERT MaxSim (multi_vector_score / multi_vector_top_k and the resident
multi-vector index) against the oracle: the reference's own unit tests plus random shapes."""
import numpy as np
import pytest

import oracle
from helpers import METRICS, assert_hits_match, close
from test_oracle_golden import mv_fixture
from vettore_b200 import nifs

pytestmark = pytest.mark.gpu


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def err(x):
    assert x[0] == "error", x
    return x[1]


def test_scores_similarity_and_distance_metrics():  # multi_vector.rs:193-206
    q = [[1.0, 0.0], [0.0, 1.0]]
    d = [[1.0, 0.0], [0.0, 1.0]]
    for code in (3, 4, 2, 0):
        assert ok(nifs.multi_vector_score(q, d, code)) == 2.0
    assert ok(nifs.multi_vector_score([], d, 0)) == 0.0
    assert ok(nifs.multi_vector_score(q, [], 0)) == 0.0


def test_top_k_is_stable_and_rejects_bad_shapes():  # multi_vector.rs:208-222
    q = [[1.0, 0.0]]
    docs = [("b", [[1.0, 0.0]]), ("a", [[1.0, 0.0]]), ("c", [[-1.0, 0.0]])]
    assert ok(nifs.multi_vector_top_k(docs, q, 3, 2)) == [("a", 1.0), ("b", 1.0)]
    assert err(nifs.multi_vector_score(q, [[1.0]], 3)) == "dimension mismatch"
    assert err(nifs.multi_vector_score([[float("nan"), 0.0]], q, 3)) == "vector contains a non-finite value"


@pytest.mark.parametrize("code", range(9))
def test_every_metric_matches_the_oracle_score(code):  # multi_vector.rs:224-238
    q = [[1.0, -0.5, 0.0], [0.0, 1.0, 1.0]]
    d = [[1.0, 0.0, 0.0], [0.0, 1.0, -1.0], [-1.0, 0.5, 1.0]]
    assert close(ok(nifs.multi_vector_score(q, d, code)), ok(oracle.multi_vector_score(q, d, code)), 1e-6)


def test_validates_the_nonempty_side_even_when_the_other_side_is_empty():  # multi_vector.rs:240-248
    nan, inf = float("nan"), float("inf")
    assert err(nifs.multi_vector_score([], [[]], 0)) == "vectors must not be empty"
    assert nifs.multi_vector_score([], [[nan]], 0)[0] == "error"
    assert nifs.multi_vector_score([[]], [], 0)[0] == "error"
    assert nifs.multi_vector_score([[inf]], [], 0)[0] == "error"
    assert nifs.multi_vector_top_k([], [[]], 0, 1)[0] == "error"
    assert nifs.multi_vector_top_k([], [[nan]], 0, 1)[0] == "error"
    assert err(nifs.multi_vector_score([[1.0]], [[1.0]], 9)) == "unknown metric"


def test_detects_total_score_overflow_after_finite_pair_scores():  # multi_vector.rs:250-258
    assert err(nifs.multi_vector_score([[1.0e19]] * 4, [[1.0e19]], 3)) == "score overflow"
    assert err(nifs.multi_vector_score([[3.0e38]], [[3.0e38]], 3)) == "metric overflow"


@pytest.mark.parametrize("code", range(9))
def test_batched_top_k_matches_full_sort_for_all_metrics_and_limits(code):  # multi_vector.rs:260-296
    docs, q = mv_fixture()
    for limit in (0, 1, 7, 25, 100):
        assert_hits_match(ok(nifs.multi_vector_top_k(docs, q, code, limit)), ok(oracle.multi_vector_top_k(docs, q, code, limit)))


def test_empty_queries_still_validate_documents_and_order_zero_score_ties():  # multi_vector.rs:298-306
    docs = [("b", [[1.0]]), ("a", [[2.0]])]
    assert nifs.multi_vector_top_k(docs, [], 0, 10) == ("ok", [("a", 0.0), ("b", 0.0)])
    assert nifs.multi_vector_top_k([("b", [[1.0]]), ("a", [[]])], [], 0, 10)[0] == "error"


def test_db_level_maxsim_known_answers():  # test/vector_db_test.exs:176-218
    docs = [("both_axes", [[1.0, 0.0], [0.0, 1.0]]), ("one_axis", [[1.0, 0.0]])]
    hits = ok(nifs.multi_vector_top_k(docs, [[1.0, 0.0], [0.0, 1.0]], 3, 10))
    assert hits == [("both_axes", 2.0), ("one_axis", 1.0)]


def _random_docs(ndocs, tmin, tmax, dim, seed):
    rng = np.random.default_rng(seed)
    docs = []
    for i in range(ndocs):
        t = int(rng.integers(tmin, tmax + 1))
        v = rng.standard_normal((t, dim)).astype(np.float32)
        if t:
            v = (v / np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
        docs.append((f"doc-{(i * 7919) % ndocs:05d}", v))
    return docs


@pytest.mark.parametrize("code", [0, 2, 3, 4, 5, 6, 7, 8, 1])
@pytest.mark.parametrize("ndocs,tmin,tmax,dim,tq", [(300, 1, 40, 64, 8), (200, 0, 200, 128, 32), (64, 120, 140, 70, 37)])
def test_random_ragged_parity(code, ndocs, tmin, tmax, dim, tq):
    docs = _random_docs(ndocs, tmin, tmax, dim, seed=ndocs + dim)
    rng = np.random.default_rng(5)
    q = rng.standard_normal((tq, dim)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    if code in (7, 8):
        docs = [(i, np.where(v < 0.05, 0.0, v).astype(np.float32)) for i, v in docs]
        q = np.where(q < 0.05, 0.0, q).astype(np.float32)
    for limit in (1, 10, 1000):
        assert_hits_match(ok(nifs.multi_vector_top_k(docs, q, code, limit)), ok(oracle.multi_vector_top_k(docs, q, code, limit)))


def test_resident_index_matches_by_value_and_handles_mutation():
    dim, tq = 128, 32
    docs = _random_docs(500, 100, 128, dim, seed=1)
    rng = np.random.default_rng(9)
    q = rng.standard_normal((tq, dim)).astype(np.float32)
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    assert nifs.mv_info(idx)[0] == 500
    exp = ok(oracle.multi_vector_top_k(docs, q, 3, 10))
    assert_hits_match(ok(nifs.mv_search(idx, q, 10)), exp)
    # upsert the best document with a useless one, delete the runner-up
    best, second = exp[0][0], exp[1][0]
    new_doc = (best, np.zeros((3, dim), dtype=np.float32))
    assert nifs.mv_insert_many(idx, [new_doc]) == ("ok", ())
    assert nifs.mv_delete(idx, second) == ("ok", ())
    docs2 = [d for d in docs if d[0] not in (best, second)] + [new_doc]
    assert_hits_match(ok(nifs.mv_search(idx, q, 10)), ok(oracle.multi_vector_top_k(docs2, q, 3, 10)))
    assert err(nifs.mv_search(idx, q[:, :5], 10)) == "dimension mismatch"
    assert ok(nifs.mv_search(idx, [], 3)) == ok(oracle.multi_vector_top_k(docs2, [], 3, 3))


# ------------------------------------------------------------------ tensor-core path (uniform documents)
def _uniform_docs(ndocs, td, dim, seed, normalise=True):
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((ndocs, td, dim)).astype(np.float32)
    if normalise:
        t = (t / np.linalg.norm(t.astype(np.float64), axis=2, keepdims=True)).astype(np.float32)
    ids = [f"doc-{(i * 7919) % ndocs:06d}" for i in range(ndocs)]
    return ids, t


@pytest.mark.parametrize("metric", ["inner_product", "negative_inner_product", "cosine"])
@pytest.mark.parametrize("ndocs,td,dim,tq", [(700, 128, 128, 32), (1000, 64, 64, 7), (900, 32, 96, 1), (333, 128, 32, 20)])
def test_tensor_core_maxsim_matches_oracle(metric, ndocs, td, dim, tq):
    ids, toks = _uniform_docs(ndocs, td, dim, seed=ndocs + td, normalise=(metric != "cosine"))
    rng = np.random.default_rng(11)
    q = rng.standard_normal((tq, dim)).astype(np.float32)
    if metric != "cosine":
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    idx = nifs.mv_new(metric)
    assert nifs.mv_insert_tensor(idx, ids, toks) == ("ok", ())
    docs = [(ids[i], toks[i]) for i in range(ndocs)]
    code = nifs.METRIC_CODE[metric]
    for limit in (1, 10, 100):
        assert_hits_match(ok(nifs.mv_search(idx, q, limit)), ok(oracle.multi_vector_top_k(docs, q, code, limit)))
    # deleting and upserting keeps the uniform layout (tombstones are skipped by the epilogue)
    best = ok(nifs.mv_search(idx, q, 1))[0][0]
    assert nifs.mv_delete(idx, best) == ("ok", ())
    docs2 = [d for d in docs if d[0] != best]
    assert_hits_match(ok(nifs.mv_search(idx, q, 10)), ok(oracle.multi_vector_top_k(docs2, q, code, 10)))


def test_tensor_core_and_general_kernels_agree(monkeypatch):
    ids, toks = _uniform_docs(512, 128, 128, seed=3)
    q = _uniform_docs(1, 32, 128, seed=4)[1][0]
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_tensor(idx, ids, toks) == ("ok", ())
    tc = ok(nifs.mv_search(idx, q, 50))
    monkeypatch.setenv("VB_MAXSIM_NO_TC", "1")
    general = ok(nifs.mv_search(idx, q, 50))
    assert_hits_match(tc, general)
