"""GPU parity of the RAGGED tensor-core MaxSim kernel (csrc/maxsim_tcr.cu) against the oracle
(multi_vector.rs:65-132): documents of any length packed back to back, documents that cross chunk, tile and
CTA boundaries, empty documents, tombstones, up to 64 query tokens, dimensions that are not a multiple of 32."""
import numpy as np
import pytest

import oracle
from helpers import assert_hits_match
from vettore_b200 import _lib, nifs

pytestmark = pytest.mark.gpu

TC_METRICS = ["inner_product", "negative_inner_product", "cosine"]


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def path():
    """0 general kernel, 1 tensor-core uniform, 2 tensor-core ragged, 3 / 4 flagged and redone (maxsim.cu)."""
    return _lib.lib().vb_debug_maxsim_path()


def ragged_docs(lengths, dim, seed, normalise=True):
    rng = np.random.default_rng(seed)
    n = len(lengths)
    docs = []
    for i, t in enumerate(lengths):
        v = rng.standard_normal((int(t), dim)).astype(np.float32)
        if t and normalise:
            v = (v / np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
        docs.append((f"doc-{(i * 7919) % n:06d}" if np.gcd(7919, n) == 1 else f"doc-{i:06d}", v))
    return docs


def query(tq, dim, seed, normalise=True):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((tq, dim)).astype(np.float32)
    if normalise:
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    return q


def lengths(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "colbert":        # 40..180 tokens, the shape of a real late-interaction corpus
        return rng.integers(40, 181, n)
    if kind == "tiny":           # many documents per 32-token chunk
        return rng.integers(1, 4, n)
    if kind == "mixed":          # empty, tiny and multi-tile documents side by side
        return rng.choice([0, 1, 2, 31, 32, 33, 127, 128, 129, 300, 700], n)
    if kind == "long":           # every document spans several tiles (carry chain across both epilogue groups)
        return rng.integers(500, 2500, n)
    raise AssertionError(kind)


@pytest.mark.parametrize("metric", TC_METRICS)
@pytest.mark.parametrize("kind,ndocs,dim,tq", [("colbert", 3000, 128, 32), ("tiny", 5000, 64, 8), ("mixed", 1500, 128, 32),
                                               ("long", 120, 96, 20), ("colbert", 800, 70, 37), ("mixed", 900, 128, 64),
                                               ("colbert", 600, 20, 1)])
def test_ragged_tensor_core_maxsim_matches_oracle(metric, kind, ndocs, dim, tq):
    docs = ragged_docs(lengths(kind, ndocs, seed=ndocs + dim), dim, seed=dim + tq, normalise=(metric != "cosine"))
    q = query(tq, dim, seed=3, normalise=(metric != "cosine"))
    idx = nifs.mv_new(metric)
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    code = nifs.METRIC_CODE[metric]
    for limit in (1, 10, 100, 1000):
        got = ok(nifs.mv_search(idx, q, limit))
        assert path() == 2, "the ragged tensor-core kernel must answer this query"
        assert_hits_match(got, ok(oracle.multi_vector_top_k(docs, q, code, limit)))


def test_ragged_index_mutation_and_both_kernels_agree(monkeypatch):
    dim, tq = 128, 32
    docs = ragged_docs(lengths("colbert", 2000, seed=1), dim, seed=2)
    q = query(tq, dim, seed=9)
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    exp = ok(oracle.multi_vector_top_k(docs, q, 3, 20))
    assert_hits_match(ok(nifs.mv_search(idx, q, 20)), exp)
    # upsert the best document with a shorter useless one, delete the runner-up, add an empty document
    best, second = exp[0][0], exp[1][0]
    new_doc = (best, np.zeros((3, dim), dtype=np.float32))
    empty = ("aaa-empty", np.zeros((0, dim), dtype=np.float32))
    assert nifs.mv_insert_many(idx, [new_doc, empty]) == ("ok", ())
    assert nifs.mv_delete(idx, second) == ("ok", ())
    docs2 = [d for d in docs if d[0] not in (best, second)] + [new_doc, empty]
    tc = ok(nifs.mv_search(idx, q, 2000))
    assert path() == 2
    assert_hits_match(tc, ok(oracle.multi_vector_top_k(docs2, q, 3, 2000)))
    monkeypatch.setenv("VB_MAXSIM_NO_TCR", "1")
    general = ok(nifs.mv_search(idx, q, 2000))
    assert path() == 0
    assert_hits_match(tc[:1000], general[:1000])


def test_by_value_top_k_takes_the_ragged_kernel():
    docs = ragged_docs(lengths("mixed", 400, seed=5), 128, seed=6)
    q = query(32, 128, seed=7)
    for code in (3, 4, 2):
        got = ok(nifs.multi_vector_top_k(docs, q, code, 50))
        assert path() == 2
        assert_hits_match(got, ok(oracle.multi_vector_top_k(docs, q, code, 50)))


def test_one_document_longer_than_a_cta_share_and_single_token_corpus():
    dim = 64
    docs = ragged_docs([40000, 5, 17], dim, seed=8)          # the first document spans most CTAs' token shares
    q = query(16, dim, seed=1)
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    assert_hits_match(ok(nifs.mv_search(idx, q, 3)), ok(oracle.multi_vector_top_k(docs, q, 3, 3)))
    assert path() == 2
    one = ragged_docs([1], dim, seed=2)
    idx = nifs.mv_new("cosine")
    assert nifs.mv_insert_many(idx, one) == ("ok", ())
    assert_hits_match(ok(nifs.mv_search(idx, q, 5)), ok(oracle.multi_vector_top_k(one, q, 2, 5)))


def test_overflowing_pairs_fall_back_to_the_reference_semantics():
    # distances.rs:59-98: a pair whose f32 dot is not finite is recomputed in f64; what cannot be represented is
    # "metric overflow", a non-finite sum of finite maxima "score overflow" (multi_vector.rs:81-86).
    dim = 32
    big = np.full((5, dim), 3.0e37, dtype=np.float32)
    docs = [("a", big), ("b", np.ones((70, dim), dtype=np.float32))]
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    got = nifs.mv_search(idx, np.full((2, dim), 3.0e37, dtype=np.float32), 2)
    assert got == oracle.multi_vector_top_k(docs, np.full((2, dim), 3.0e37, dtype=np.float32), 3, 2)
    assert path() == 4
    q = np.full((4, dim), 1.0e18, dtype=np.float32)
    docs = [("a", np.full((3, dim), 1.0e19, dtype=np.float32)), ("b", np.ones((40, dim), dtype=np.float32))]
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    assert nifs.mv_search(idx, q, 2) == oracle.multi_vector_top_k(docs, q, 3, 2)


def test_ragged_device_ingest_equals_host_ingest():
    torch = pytest.importorskip("torch")
    dim, tq = 128, 32
    lens = lengths("mixed", 700, seed=4)
    docs = ragged_docs(lens, dim, seed=5)
    q = query(tq, dim, seed=6)
    host = nifs.mv_new("cosine")
    assert nifs.mv_insert_many(host, docs) == ("ok", ())
    flat = np.concatenate([d[1] for d in docs], axis=0)
    dev_tokens = torch.from_numpy(flat).cuda()
    doc_tok = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    dev = nifs.mv_new("cosine")
    assert nifs.mv_insert_ragged_device(dev, [d[0] for d in docs], dev_tokens.data_ptr(), doc_tok, dim) == ("ok", ())
    torch.cuda.synchronize()
    assert nifs.mv_info(dev) == nifs.mv_info(host)
    a, b = ok(nifs.mv_search(host, q, 300)), ok(nifs.mv_search(dev, q, 300))
    assert path() == 2
    assert a == b
    assert_hits_match(a, ok(oracle.multi_vector_top_k(docs, q, 2, 300)))


def test_compaction_rebuilds_the_token_owner_map():
    """Deleting most of a ragged index compacts the token matrix (tombstones outweigh live tokens): the per-token
    owner map and the document offsets must be rebuilt consistently, and later inserts must extend them."""
    dim, tq = 64, 16
    lens = lengths("colbert", 400, seed=21)
    docs = ragged_docs(lens, dim, seed=22)
    q = query(tq, dim, seed=23)
    idx = nifs.mv_new("inner_product")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    keep = docs[::5]
    for d in docs:
        if d[0] not in {k[0] for k in keep}:
            assert nifs.mv_delete(idx, d[0]) == ("ok", ())
    # an upsert after the deletes triggers the compaction check inside insert_many
    extra = ragged_docs([3, 77, 0, 140], dim, seed=24)
    extra = [(f"extra-{i}", v) for i, (_, v) in enumerate(extra)]
    assert nifs.mv_insert_many(idx, extra) == ("ok", ())
    live = keep + extra
    assert nifs.mv_info(idx)[0] == len(live)
    got = ok(nifs.mv_search(idx, q, 50))
    assert path() == 2
    assert_hits_match(got, ok(oracle.multi_vector_top_k(live, q, 3, 50)))


@pytest.mark.parametrize("tq", [32, 33, 64])
@pytest.mark.parametrize("dim", [4, 100, 128])
def test_query_token_and_dimension_boundaries(tq, dim):
    docs = ragged_docs(lengths("mixed", 300, seed=tq + dim), dim, seed=dim)
    q = query(tq, dim, seed=tq)
    idx = nifs.mv_new("cosine")
    assert nifs.mv_insert_many(idx, docs) == ("ok", ())
    got = ok(nifs.mv_search(idx, q, 40))
    assert path() == 2
    assert_hits_match(got, ok(oracle.multi_vector_top_k(docs, q, 2, 40)))


def test_by_value_limit_beyond_the_fused_collector_and_more_than_64_query_tokens():
    docs = ragged_docs(lengths("tiny", 3000, seed=31), 32, seed=32)
    q = query(8, 32, seed=33)
    got = ok(nifs.multi_vector_top_k(docs, q, 3, 2500))        # k > 1024: every score dumped and radix-sorted
    assert path() == 2 and len(got) == 2500
    assert_hits_match(got, ok(oracle.multi_vector_top_k(docs, q, 3, 2500)))
    q65 = query(65, 32, seed=34)                                # more than 64 query tokens: the general kernel
    got = ok(nifs.multi_vector_top_k(docs[:500], q65, 3, 10))
    assert path() == 0
    assert_hits_match(got, ok(oracle.multi_vector_top_k(docs[:500], q65, 3, 10)))
