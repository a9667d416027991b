"""Concurrent use of one resident index, the way BEAM dirty schedulers drive the reference's FlatResource
(nifs.rs:266-309: flat_search takes the RwLock's read side from many threads at once, insert / delete take
the write side): 8 reader threads search through the C ABI while one writer thread inserts, upserts and
deletes. Every reader result must be the oracle's answer for a state the index actually was in — never a
torn one (a moved row under a stale id, a half-updated code mirror)."""
import threading

import numpy as np
import pytest

import oracle
from helpers import assert_hits_match
from vettore_b200 import nifs

pytestmark = pytest.mark.gpu


def _unit(x):
    return (x / np.linalg.norm(x.astype(np.float64), axis=-1, keepdims=True)).astype(np.float32)


def test_readers_and_one_writer_never_see_a_torn_state():
    rng = np.random.default_rng(42)
    n, d, k, readers, rounds = 6000, 256, 10, 8, 120
    rows = _unit(rng.standard_normal((n, d)).astype(np.float32))
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    queries = _unit(rows[:readers] + 0.1 * rng.standard_normal((readers, d)).astype(np.float32))
    far = _unit(-queries.sum(axis=0, keepdims=True))[0]          # negative dot with every query: never a hit,
    idx = nifs.flat_new_cosine()                                  # and the opposite sign code: never a candidate
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    base = [oracle.flat_search_dense("cosine", rows, ids, queries[r], k)[1] for r in range(readers)]
    hot = "hot-row"                                               # upserted to query 0 itself, then deleted, repeatedly
    rows_hot, ids_hot = np.vstack([rows, queries[:1]]), ids + [hot]
    with_hot = [oracle.flat_search_dense("cosine", rows_hot, ids_hot, queries[r], k)[1] for r in range(readers)]
    assert with_hot[0][0][0] == hot
    from test_pipelines_gpu import ref_quantized
    vectors = [(ids[i], rows[i]) for i in range(n)]
    qbase = [ref_quantized(vectors, queries[r], 2, 2000, k) for r in range(readers)]
    qwith = [ref_quantized(vectors + [(hot, queries[0])], queries[r], 2, 2000, k) for r in range(readers)]

    def either(hits, a, b):
        """The index was in one of two states while the search held the read lock."""
        present = any(h[0] == hot for h in hits)
        assert_hits_match(hits, b if present else a)
    assert nifs.flat_quantized_search(idx, queries[0], 2, 2000, k)[0] == "ok"   # build the code mirror up front

    stop = threading.Event()
    errors = []

    def writer():
        try:
            live = []
            i = 0
            while not stop.is_set():
                cid = f"zz-churn-{i:05d}"
                assert nifs.flat_insert(idx, cid, far) == ("ok", ())
                live.append(cid)
                if len(live) > 6:                                  # delete an OLD churn row: the last row moves into its hole
                    assert nifs.flat_delete(idx, live.pop(0)) == ("ok", ())
                if i % 3 == 0:
                    assert nifs.flat_insert(idx, hot, queries[0]) == ("ok", ())
                elif i % 3 == 1:
                    assert nifs.flat_delete(idx, hot) == ("ok", ())
                if i % 7 == 0:                                     # upsert in place of an existing churn row
                    assert nifs.flat_insert(idx, live[0], far) == ("ok", ())
                i += 1
        except Exception as e:   # noqa: BLE001
            errors.append(("writer", repr(e)))

    def reader(r):
        try:
            for it in range(rounds):
                st, hits = nifs.flat_search(idx, queries[r], k)
                assert st == "ok", hits
                either(hits, base[r], with_hot[r])
                if it % 4 == 0:
                    st, qh = nifs.flat_quantized_search(idx, queries[r], 2, 2000, k)
                    assert st == "ok", qh
                    either(qh, qbase[r], qwith[r])
        except Exception as e:   # noqa: BLE001
            errors.append((f"reader {r}", repr(e)))

    w = threading.Thread(target=writer)
    rs = [threading.Thread(target=reader, args=(r,)) for r in range(readers)]
    w.start()
    for t in rs:
        t.start()
    for t in rs:
        t.join()
    stop.set()
    w.join()
    assert not errors, errors[:3]
    # the index is still consistent afterwards
    nifs.flat_delete(idx, hot)
    st, hits = nifs.flat_search(idx, queries[1], k)
    assert st == "ok"
    assert_hits_match(hits, base[1])


def test_concurrent_batched_and_single_searches_share_no_scratch():
    """Readers mixing single queries (K1) and 32-query batches (K2) from the context pool at the same time."""
    rng = np.random.default_rng(7)
    n, d, k = 20000, 128, 10
    rows = _unit(rng.standard_normal((n, d)).astype(np.float32))
    ids = [f"{i:06d}" for i in range(n)]
    idx = nifs.flat_new_inner_product()
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    qs = _unit(rng.standard_normal((32, d)).astype(np.float32))
    ref = [oracle.flat_search_dense("inner_product", rows, ids, qs[i], k)[1] for i in range(32)]
    errors = []

    def single(t):
        try:
            for it in range(60):
                i = (t * 7 + it) % 32
                st, hits = nifs.flat_search(idx, qs[i], k)
                assert st == "ok"
                assert_hits_match(hits, ref[i])
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    def batched():
        try:
            for _ in range(15):
                st, res = nifs.flat_search_batch(idx, qs, k)
                assert st == "ok"
                for i in range(32):
                    assert_hits_match(res[i], ref[i])
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=single, args=(t,)) for t in range(6)] + [threading.Thread(target=batched) for _ in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
