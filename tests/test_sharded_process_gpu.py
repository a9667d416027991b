"""The multi-GPU flat index INSIDE ONE PROCESS (vb_flat_new_sharded): the reference surface is one BEAM process
holding one FlatResource (flat.rs:131-134, nifs.rs:297-309), so the G-GPU path must answer the ordinary
flat_insert / flat_search calls exactly like the single-GPU index and the oracle. Shards land on device
s % device_count: 8 shards share the one GPU of a single-GPU box and spread over the GPUs of a bigger one."""
import json
import os
import threading

import numpy as np
import pytest

import oracle
from helpers import METRICS, assert_hits_match
from test_golden_fixtures import KNOWN, _run_case
from vettore_b200 import _lib, nifs

pytestmark = pytest.mark.gpu


def _rows(n, d, seed, ints=False):
    rng = np.random.default_rng(seed)
    if ints:
        return rng.integers(-2, 3, size=(n, d)).astype(np.float32)
    return rng.standard_normal((n, d)).astype(np.float32)


class _ShardedFlat:
    shards = 8

    def __init__(self, metric):
        self.idx = nifs.flat_new_sharded(metric, self.shards)
    def insert(self, i, v): return nifs.flat_insert(self.idx, i, v)
    def insert_many(self, items): return nifs.flat_insert_many(self.idx, items)
    def delete(self, i): return nifs.flat_delete(self.idx, i)
    def search(self, q, k): return nifs.flat_search(self.idx, q, k)


@pytest.mark.parametrize("case", [c for c in KNOWN if c["fn"] == "flat_script"], ids=lambda c: c["source"].split(" ")[0])
def test_reference_flat_known_answers_through_a_sharded_handle(case):
    """flat.rs:165-180, 208-249, 252-281 and the hardening tests: same answers from 8 shards."""
    _run_case(case, nifs, _ShardedFlat)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("shards,n,d,k,ints", [(2, 5000, 96, 10, False), (8, 4000, 64, 100, True), (3, 700, 33, 1000, False)])
def test_sharded_search_equals_oracle_and_single_index(metric, shards, n, d, k, ints):
    rows = _rows(n, d, n + d, ints)
    if metric in ("cosine",) and not ints:
        rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    q = _rows(1, d, 3, ints)[0]
    sh = nifs.flat_new_sharded(metric, shards)
    assert nifs.flat_reserve(sh, n) == ("ok", ())
    assert nifs.flat_insert_matrix(sh, ids, rows) == ("ok", ())
    assert nifs.flat_info(sh) == (n, d)
    st, got = nifs.flat_search(sh, q, k)
    assert st == "ok", got
    st, exp = oracle.flat_search_dense(metric, rows, ids, q, k)
    assert st == "ok"
    assert_hits_match(got, exp, exact_ids=ints)
    one = getattr(nifs, f"flat_new_{metric}")()
    assert nifs.flat_insert_matrix(one, ids, rows) == ("ok", ())
    assert nifs.flat_search(one, q, k) == ("ok", got)     # bit-identical to the single-GPU index


def test_sharded_mutations_batches_and_errors():
    n, d, k = 3000, 128, 10
    rows = _rows(n, d, 11)
    ids = [f"id-{i:05d}" for i in range(n)]
    sh = nifs.flat_new_sharded("l2", 4)
    assert nifs.flat_search(sh, [1.0, 2.0], 5) == ("ok", [])                      # empty index: any finite query
    assert nifs.flat_insert_many(sh, [("a", [1.0, 2.0]), ("b", [1.0])]) == ("error", "dimension mismatch")
    assert nifs.flat_info(sh) == (0, None)                                         # all-or-nothing across shards
    assert nifs.flat_insert_many(sh, [("a", [1.0, float("inf")])]) == ("error", "vector contains a non-finite value")
    assert nifs.flat_insert_matrix(sh, ids, rows) == ("ok", ())
    assert nifs.flat_insert(sh, "x", [1.0]) == ("error", "dimension mismatch")
    assert nifs.flat_search(sh, np.zeros(d + 1), 3) == ("error", "dimension mismatch")
    assert nifs.flat_search(sh, [], 3) == ("error", "vector must not be empty")
    assert nifs.flat_search(sh, np.zeros(d + 1), 0) == ("ok", [])                  # flat.rs:97-99
    # upsert moves a row to the query; delete removes the runner-up
    q = _rows(1, d, 5)[0]
    assert nifs.flat_insert(sh, ids[77], q) == ("ok", ())
    rows2 = rows.copy()
    rows2[77] = q
    st, got = nifs.flat_search(sh, q, k)
    assert st == "ok" and got[0] == (ids[77], 0.0)
    assert nifs.flat_delete(sh, got[1][0]) == ("ok", ())
    gone = ids.index(got[1][0])
    keep = [i for i in range(n) if i != gone]
    st, got = nifs.flat_search(sh, q, k)
    assert_hits_match(got, oracle.flat_search_dense("l2", rows2[keep], [ids[i] for i in keep], q, k)[1])
    assert nifs.flat_info(sh) == (n - 1, d)
    # a batch of queries (K2 is not eligible for l2: per-query K1 on every shard) and an inner-product batch (K2)
    qs = _rows(40, d, 9)
    st, res = nifs.flat_search_batch(sh, qs, k)
    assert st == "ok"
    for i in range(40):
        assert_hits_match(res[i], oracle.flat_search_dense("l2", rows2[keep], [ids[j] for j in keep], qs[i], k)[1])
    ip = nifs.flat_new_sharded("inner_product", 4)
    assert nifs.flat_insert_matrix(ip, ids, rows) == ("ok", ())
    st, res = nifs.flat_search_batch(ip, qs, k)
    assert st == "ok"
    for i in range(0, 40, 5):
        assert_hits_match(res[i], oracle.flat_search_dense("inner_product", rows, ids, qs[i], k)[1])
    # delete everything: the dimension resets to None (flat.rs:90-92)
    small = nifs.flat_new_sharded("cosine", 3)
    assert nifs.flat_insert_many(small, [("a", [1.0, 0.0]), ("b", [0.0, 1.0])]) == ("ok", ())
    assert nifs.flat_delete(small, "a") == ("ok", ()) and nifs.flat_delete(small, "b") == ("ok", ())
    assert nifs.flat_info(small) == (0, None)
    assert nifs.flat_insert(small, "c", [1.0, 2.0, 3.0]) == ("ok", ())
    # the stream-ordered device-level entries need a single device and answer loudly
    rc = _lib.lib().vb_flat_device_status(sh.handle, None)
    assert rc != 0 and "sharded" in _lib.last_error()


def test_eight_shards_driven_by_concurrent_callers():
    n, d, k = 8000, 256, 10
    rows = _rows(n, d, 21)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ids = [f"{i:06d}" for i in range(n)]
    sh = nifs.flat_new_sharded("cosine", 8)
    assert nifs.flat_insert_matrix(sh, ids, rows) == ("ok", ())
    qs = _rows(8, d, 22)
    ref = [oracle.flat_search_dense("cosine", rows, ids, qs[i], k)[1] for i in range(8)]
    errors = []

    def caller(t):
        try:
            for _ in range(40):
                st, hits = nifs.flat_search(sh, qs[t], k)
                assert st == "ok"
                assert_hits_match(hits, ref[t])
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=caller, args=(t,)) for t in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]


# ------------------------------------------------------------------ resident pipelines over the sharded corpus
def _pipeline_pair(metric, shards, n, d, seed):
    rows = _rows(n, d, seed)
    rows = (rows / np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    single = getattr(nifs, f"flat_new_{metric}")()
    sharded = nifs.flat_new_sharded(metric, shards)
    for idx in (single, sharded):
        assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    return rows, ids, single, sharded


@pytest.mark.parametrize("metric_code", [2, 3, 0, 5])
@pytest.mark.parametrize("shards", [2, 8])
def test_sharded_funnel_and_prefix_equal_the_single_index_and_the_oracle(metric_code, shards):
    n, d = 6000, 96
    rows, ids, single, sharded = _pipeline_pair("cosine", shards, n, d, 11)
    q = _rows(1, d, 5)[0]
    for stages, cand, limit in (([16, 48], 200, 10), ([32], 1500, 25), ([], 50, 7), ([8, 24, 64], 64, 64)):
        a = nifs.flat_funnel_search(single, q, metric_code, stages, cand, limit)
        b = nifs.flat_funnel_search(sharded, q, metric_code, stages, cand, limit)
        assert a[0] == b[0] == "ok"
        assert a[1] == b[1], (stages, cand, limit)
    # vector_top_k over all rows / a listed subset (search.rs:38-73) against the oracle
    vectors = [(ids[i], rows[i]) for i in range(n)]
    for dims, limit in ((24, 50), (d, 10)):
        exp = oracle.vector_top_k(vectors, q, metric_code, dims, limit)
        got = nifs.flat_prefix_top_k(sharded, None, q, metric_code, dims, limit)
        assert got[0] == exp[0] == "ok"
        assert_hits_match(got[1], exp[1])
    some = ids[::7] + ["no-such-id"]
    exp = oracle.vector_top_k([(i, rows[ids.index(i)]) for i in ids[::7]], q, metric_code, 40, 30)
    assert_hits_match(nifs.flat_prefix_top_k(sharded, some, q, metric_code, 40, 30)[1], exp[1])
    # error strings and degenerate limits behave like the single index
    for args in ((q, metric_code, [200], 10, 5), (q, 9, [16], 10, 5), (q, metric_code, [16], 0, 5), (q, metric_code, [16], 10, 0)):
        assert nifs.flat_funnel_search(sharded, *args) == nifs.flat_funnel_search(single, *args)
    assert nifs.flat_prefix_top_k(sharded, None, q, metric_code, 0, 5) == nifs.flat_prefix_top_k(single, None, q, metric_code, 0, 5)


@pytest.mark.parametrize("shards", [2, 8])
def test_sharded_quantized_search_equals_the_single_index(shards):
    n, d = 20000, 128
    rows, ids, single, sharded = _pipeline_pair("cosine", shards, n, d, 23)
    q = _rows(1, d, 9)[0]
    for code in (2, 3, 0):
        for cand, limit in ((100, 10), (1500, 40), (5, 10), (n + 5, 3)):
            a = nifs.flat_quantized_search(single, q, code, cand, limit)
            b = nifs.flat_quantized_search(sharded, q, code, cand, limit)
            assert a[0] == b[0] == "ok"
            assert a[1] == b[1], (code, cand, limit)
    # and the oracle's composition: binary_top_k over the sign codes, then vector_top_k over the candidates
    codes = [(ids[i], oracle.compress_sign_bits(rows[i])) for i in range(n)]
    st, cands = oracle.binary_top_k(codes, oracle.compress_sign_bits(q), d, 300)
    assert st == "ok"
    pos = {i: r for r, i in enumerate(ids)}
    exp = oracle.vector_top_k([(i, rows[pos[i]]) for i, _ in cands], q, 2, d, 10)
    assert_hits_match(nifs.flat_quantized_search(sharded, q, 2, 300, 10)[1], exp[1])
    for args in ((q[:5], 2, 10, 5), (q, 9, 10, 5), (q, 2, 0, 5), (q, 2, 10, 0)):
        assert nifs.flat_quantized_search(sharded, *args) == nifs.flat_quantized_search(single, *args)


# ------------------------------------------------------------------ multi-vector collection over several GPUs
def _mv_docs(n, tmin, tmax, dim, seed):
    rng = np.random.default_rng(seed)
    docs = []
    for i in range(n):
        t = int(rng.integers(tmin, tmax + 1))
        docs.append((f"doc-{(i * 7919) % n:05d}", rng.standard_normal((t, dim)).astype(np.float32)))
    return docs


@pytest.mark.parametrize("metric", ["inner_product", "cosine", "l2", "manhattan"])
@pytest.mark.parametrize("shards", [2, 8])
def test_sharded_multi_vector_collection_equals_single_and_oracle(metric, shards):
    dim, tq = 64, 12
    docs = _mv_docs(900, 0, 90, dim, shards + len(metric))
    q = np.random.default_rng(4).standard_normal((tq, dim)).astype(np.float32)
    single, sharded = nifs.mv_new(metric), nifs.mv_new_sharded(metric, shards)
    for idx in (single, sharded):
        assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    assert nifs.mv_info(sharded) == nifs.mv_info(single)
    code = nifs.METRIC_CODE[metric]
    for limit in (1, 10, 200, 2000):
        a, b = nifs.mv_search(single, q, limit), nifs.mv_search(sharded, q, limit)
        assert a[0] == b[0] == "ok"
        assert a[1] == b[1], limit
        assert_hits_match(b[1], oracle.multi_vector_top_k(docs, q, code, limit)[1])
    # upsert, delete, empty query, validation errors: like the single collection
    best = nifs.mv_search(single, q, 1)[1][0][0]
    for idx in (single, sharded):
        assert nifs.mv_insert_many(idx, [(best, np.zeros((2, dim), dtype=np.float32))]) == ("ok", ())
        assert nifs.mv_delete(idx, docs[3][0]) == ("ok", ())
    assert nifs.mv_search(sharded, q, 25) == nifs.mv_search(single, q, 25)
    assert nifs.mv_search(sharded, [], 5) == nifs.mv_search(single, [], 5)
    assert nifs.mv_search(sharded, q[:, :5], 5) == nifs.mv_search(single, q[:, :5], 5) == ("error", "dimension mismatch")
    assert nifs.mv_insert_many(sharded, [("x", np.zeros((1, dim + 1), dtype=np.float32))]) == ("error", "dimension mismatch")
    assert nifs.mv_info(sharded) == nifs.mv_info(single)
