"""GPU parity of the batched flat search (K2: tcgen05 3xTF32 GEMM + fused per-query top-k)
against the oracle, and its agreement with the single-query kernel."""
import numpy as np
import pytest

import oracle
from helpers import assert_hits_match
from vettore_b200 import nifs

pytestmark = pytest.mark.gpu


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def _rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)


@pytest.fixture(params=["auto", "1", "3"])
def gemm_terms(request, monkeypatch):
    """K2 precision mode: one TF32 pass with a wide candidate margin (k <= 32 by default), 3xTF32 otherwise; both
    are filters in front of the exact re-scoring, so every test must pass in either mode."""
    if request.param != "auto":
        monkeypatch.setenv("VB_GEMM_TERMS", request.param)
    return request.param


@pytest.mark.parametrize("metric", ["cosine", "inner_product", "negative_inner_product", "l2", "l2_squared"])
@pytest.mark.parametrize("n,d,nq,k", [(5000, 128, 40, 10), (20000, 768, 300, 10), (3000, 64, 16, 100), (9000, 96, 513, 1)])
def test_batched_search_matches_oracle(metric, n, d, nq, k, gemm_terms):
    rows = _rows(n, d, n + d)
    queries = _rows(nq, d, 17)
    queries[3] = rows[42]            # an exact duplicate: self-match must rank first
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    assert len(got) == nq
    for qi in list(range(0, nq, max(1, nq // 12))) + [3, nq - 1]:
        exp = ok(oracle.flat_search_dense(metric, rows, ids, queries[qi], k))
        assert_hits_match(got[qi], exp)


def test_batched_and_single_query_kernels_agree(monkeypatch, gemm_terms):
    n, d, nq, k = 12000, 256, 64, 10
    rows, queries = _rows(n, d, 1), _rows(nq, d, 2)
    ids = [f"{i:06d}" for i in range(n)]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    gemm = ok(nifs.flat_search_batch(idx, queries, k))
    monkeypatch.setenv("VB_FLAT_NO_GEMM", "1")
    scan = ok(nifs.flat_search_batch(idx, queries, k))
    for a, b in zip(gemm, scan):
        assert_hits_match(a, b)


def test_batched_search_ties_resolve_by_id(gemm_terms):
    """Small-integer rows: many exact score ties inside and across tiles."""
    rng = np.random.default_rng(5)
    n, d, nq, k = 4000, 32, 32, 25
    rows = rng.integers(-1, 2, size=(n, d)).astype(np.float32)
    queries = rng.integers(-1, 2, size=(nq, d)).astype(np.float32)
    ids = [f"{(i * 31) % n:05d}" for i in range(n)]
    idx = nifs.flat_new_inner_product()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in range(nq):
        assert got[qi] == ok(oracle.flat_search_dense("inner_product", rows, ids, queries[qi], k))


@pytest.mark.parametrize("k", [40, 100])
def test_batched_search_large_k_many_tiles_per_cta(k, gemm_terms):
    """Enough rows per CTA that the per-(CTA, query) candidate lists are cut back several times
    (value-bisection select), with a continuous score distribution."""
    n, d, nq = 250_000, 64, 16
    rows, queries = _rows(n, d, 77), _rows(nq, d, 78)
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in range(nq):
        assert_hits_match(got[qi], ok(oracle.flat_search_dense("cosine", rows, ids, queries[qi], k)))


@pytest.mark.parametrize("k", [25, 100])
def test_batched_search_large_k_ties_at_the_cut(k, gemm_terms):
    """Small-integer rows over many tiles: the cut of a candidate list falls inside large groups of equal
    scores, so the id word decides which of them survive (second bisection)."""
    rng = np.random.default_rng(6)
    n, d, nq = 120_000, 32, 16
    rows = rng.integers(-1, 2, size=(n, d)).astype(np.float32)
    queries = rng.integers(-1, 2, size=(nq, d)).astype(np.float32)
    ids = [f"{(i * 31) % n:06d}" for i in range(n)]
    idx = nifs.flat_new_inner_product()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in range(nq):
        assert got[qi] == ok(oracle.flat_search_dense("inner_product", rows, ids, queries[qi], k))


def test_single_pass_filter_is_complete_on_adversarial_near_ties(monkeypatch):
    """Rows packed within the single-pass error bound of each other (all-positive, nearly parallel vectors at d = 32
    and 64: ADVICE r1): the kept candidate set cannot be proven complete from TF32 scores alone, so the queries must
    come back exact anyway (flagged and redone by the single-query kernel)."""
    rng = np.random.default_rng(11)
    for d in (32, 64):
        n, nq, k = 6000, 20, 10
        base = np.abs(rng.standard_normal(d)).astype(np.float32) + 1.0
        rows = (base[None, :] * (1.0 + 2e-4 * rng.standard_normal((n, d)))).astype(np.float32)
        rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
        queries = rows[:nq] + 1e-4 * rng.standard_normal((nq, d)).astype(np.float32)
        ids = [f"{(i * 7919) % n:05d}" for i in range(n)]
        for terms in ("1", "3"):
            monkeypatch.setenv("VB_GEMM_TERMS", terms)
            idx = nifs.flat_new_cosine()
            ok(nifs.flat_insert_matrix(idx, ids, rows))
            got = ok(nifs.flat_search_batch(idx, queries, k))
            for qi in range(nq):
                assert_hits_match(got[qi], ok(oracle.flat_search_dense("cosine", rows, ids, queries[qi], k)))


@pytest.mark.parametrize("metric", ["l2", "l2_squared"])
def test_batched_l2_family_unnormalised_rows_and_mutations(metric, gemm_terms):
    """The L2 family goes through the tensor cores as |x|^2 - 2 q.x with a row-norm mirror: rows of very different
    norms, an exact duplicate of a query (distance 0), and mutations that must refresh the mirror."""
    rng = np.random.default_rng(21)
    n, d, nq, k = 70_000, 64, 48, 10
    rows = (rng.standard_normal((n, d)) * rng.uniform(0.2, 5.0, size=(n, 1))).astype(np.float32)
    queries = (rng.standard_normal((nq, d)) * 2.0).astype(np.float32)
    queries[5] = rows[1234]
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in range(nq):
        assert_hits_match(got[qi], ok(oracle.flat_search_dense(metric, rows, ids, queries[qi], k)))
    assert got[5][0] == (ids[1234], 0.0)
    # delete the duplicate, upsert another row onto a query: the mirror must follow
    ok(nifs.flat_delete(idx, ids[1234]))
    ok(nifs.flat_insert(idx, ids[77], queries[9]))
    rows2 = rows.copy()
    rows2[77] = queries[9]
    keep = [i for i in range(n) if i != 1234]
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in (5, 9, 20, nq - 1):
        assert_hits_match(got[qi], ok(oracle.flat_search_dense(metric, rows2[keep], [ids[i] for i in keep], queries[qi], k)))
    assert got[9][0] == (ids[77], 0.0)


@pytest.mark.parametrize("metric,n,d,nq,k", [("cosine", 70000, 768, 300, 10), ("l2", 9000, 96, 513, 100), ("inner_product", 3000, 64, 16, 1)])
def test_cta_pair_form_matches_oracle(monkeypatch, metric, n, d, nq, k):
    """VB_GEMM_PAIR=1: the single-pass kernel as CTA pairs (clusters of 2, tcgen05 cta_group::2, 128-query blocks, two
    accumulator sets). Opt-in (measured slower than the single-CTA form), but it must stay exact."""
    monkeypatch.setenv("VB_GEMM_PAIR", "1")
    rows = _rows(n, d, n + d)
    if metric == "l2":
        rows = (rows * np.linspace(0.5, 2.0, n, dtype=np.float32)[:, None]).astype(np.float32)
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    queries = _rows(nq, d, 5)
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in list(range(0, nq, max(1, nq // 24))) + [nq - 1]:
        assert_hits_match(got[qi], ok(oracle.flat_search_dense(metric, rows, ids, queries[qi], k)))


@pytest.mark.parametrize("metric", ["inner_product", "l2_squared"])
def test_batched_search_keeps_the_overflow_semantics(metric):
    """distances.rs:59-98 through the batched path: magnitudes at which |q| * |row| can overflow fp32 send the batch to the
    exact kernels (f64 recovery, "metric overflow"), so the answers equal the oracle's, error string included."""
    n, d, nq, k = 3000, 64, 24, 5
    rng = np.random.default_rng(2)
    rows = (rng.standard_normal((n, d)) * 1.0e18).astype(np.float32)
    ids = [f"{i:05d}" for i in range(n)]
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    small_q = rng.standard_normal((nq, d)).astype(np.float32)            # finite scores (~1e19): recovered, not an error
    got = nifs.flat_search_batch(idx, small_q, k)
    for qi in (0, nq - 1):
        exp = oracle.flat_search_dense(metric, rows, ids, small_q[qi], k)
        assert exp[0] == "ok"
        if metric == "inner_product":
            assert_hits_match(ok(got)[qi], exp[1])
    big_q = (rng.standard_normal((nq, d)) * 1.0e20).astype(np.float32)    # |q| |row| ~ 6e39: beyond fp32, and beyond f32 after f64 recovery
    exp = oracle.flat_search_dense(metric, rows, ids, big_q[0], k)
    got = nifs.flat_search_batch(idx, big_q, k)
    assert got[0] == exp[0]
    if exp[0] == "error":
        assert got[1] == exp[1]


@pytest.mark.parametrize("metric,k,dup", [("cosine", 100, False), ("l2", 40, False), ("inner_product", 100, True)])
def test_long_pre_pass_sample_two_phase_select(monkeypatch, metric, k, dup):
    """The dense pre-pass with a sample longer than the select kernel's shared-memory head (32 768 scores): the tail of
    the sample is streamed and filtered against the head's bound. `dup`: every sample score equal (the survivor list
    overflows, the head's bound stands) — either way the bound only filters, the answers stay exact."""
    monkeypatch.setenv("VB_GEMM_TERMS", "1")
    monkeypatch.setenv("VB_GEMM_SAMPLE", "60000")
    n, d, nq = 90000, 64, 40
    rows = _rows(n, d, 77)
    if dup:
        rows[:70000] = rows[0]
    if metric == "l2":
        rows = (rows * np.linspace(0.5, 2.0, n, dtype=np.float32)[:, None]).astype(np.float32)
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    queries = _rows(nq, d, 78)
    idx = getattr(nifs, f"flat_new_{metric}")()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_search_batch(idx, queries, k))
    for qi in (0, 7, nq - 1):
        assert_hits_match(got[qi], ok(oracle.flat_search_dense(metric, rows, ids, queries[qi], k)))
