"""Parity at the BENCHMARKED configurations themselves (BASELINE.json configs[1], and the per-GPU shapes of
configs[3]/[4]): the CUDA path through the C ABI against the oracle on the same seeded inputs, full size.
Reference behaviour: flat.rs:96-124 (K1/K2), search.rs:76-92 (K3), multi_vector.rs:90-132 (K5)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_hits_match
from vettore_b200 import _lib, nifs

pytestmark = pytest.mark.gpu

SEED = 20_260_721
THREADS = os.cpu_count() or 1


def _normal_rows(n, d, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn(n, d, generator=g, device=dev)
    for s in range(0, n, 131072):
        b = x[s:s + 131072].double()
        x[s:s + 131072] = (b / b.norm(dim=1, keepdim=True)).float()
    return x


@pytest.fixture(scope="module")
def corpus_1m():
    dev = torch.device("cuda", 0)
    n, d = 1_000_000, 768
    x = _normal_rows(n, d, SEED, dev)
    rows = x.cpu().numpy()
    idx = nifs.flat_new_cosine()
    assert nifs.flat_insert_device(idx, nifs.decimal_ids(0, n), x.data_ptr(), d) == ("ok", ())
    del x
    torch.cuda.empty_cache()
    return idx, rows


@pytest.mark.parametrize("k", [10, 100, 1000])
def test_k1_single_query_1m_x_768_cosine(corpus_1m, k):
    """configs[1], batch of 1: every hit of 16 queries equals the oracle's over the whole 1M-row corpus."""
    idx, rows = corpus_1m
    q = _normal_rows(16 if k <= 100 else 4, 768, SEED + 1, torch.device("cuda", 0)).cpu().numpy()   # k = 1000: merge tree + rank merge
    _, ref = oracle.flat_scan_timed("cosine", rows, q, k, THREADS)
    for qi in range(q.shape[0]):
        st, hits = nifs.flat_search(idx, q[qi], k)
        assert st == "ok", hits
        assert_hits_match([(int(h[0]), h[1]) for h in hits], ref[qi])


def test_k2_batch_1024_queries_1m_x_768_cosine(corpus_1m):
    """configs[1], batch of 1024 (tcgen05 3xTF32 filter + exact re-scoring): 64 of the batch's queries equal
    the oracle, and the whole batch equals the single-query kernel on a further sample."""
    idx, rows = corpus_1m
    k = 10
    q = _normal_rows(1024, 768, SEED + 2, torch.device("cuda", 0)).cpu().numpy()
    st, batch = nifs.flat_search_batch(idx, q, k)
    assert st == "ok", batch
    _, ref = oracle.flat_scan_timed("cosine", rows, q[:64], k, THREADS)
    for qi in range(64):
        assert_hits_match([(int(h[0]), h[1]) for h in batch[qi]], ref[qi])
    for qi in range(64, 1024, 97):
        st, one = nifs.flat_search(idx, q[qi], k)
        assert st == "ok" and one == batch[qi], qi


def test_k3_hamming_10m_x_1024_bits_1000_candidates():
    """configs[3] per-GPU shape (10M of the 100M codes): the 1000 candidates, ids and distances, are bit-exact."""
    n, dims, cand = 10_000_000, 1024, 1000
    rng = np.random.default_rng(SEED)
    codes = rng.integers(0, 2 ** 63, size=(n, dims // 64), dtype=np.uint64)
    codes ^= rng.integers(0, 2, size=(n, 1), dtype=np.uint64) << np.uint64(63)
    q = codes[12345] ^ rng.integers(0, 2 ** 40, size=dims // 64, dtype=np.uint64)
    _, ref = oracle.binary_scan_timed(codes, dims, q[None, :], cand, 1)
    ids = nifs.decimal_ids(0, n)
    blob, ioff = nifs._ids_blob(ids)
    import ctypes as C
    from vettore_b200 import _lib
    woff = np.arange(n + 1, dtype=np.uint64) * np.uint64(dims // 64)
    h = C.c_void_p()
    rc = _lib.lib().vb_binary_top_k(n, blob, ioff.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    codes.ctypes.data_as(C.POINTER(C.c_uint64)), woff.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    q.ctypes.data_as(C.POINTER(C.c_uint64)), dims // 64, dims, cand, C.byref(h))
    assert rc == 0, _lib.last_error()
    hits = nifs._take_hits(h)
    assert [(int(i), v) for i, v in hits] == ref[0]


def _oracle_maxsim_parallel(metric, tokens, q, k):
    """One query over many documents: the oracle runs whole queries per thread, so split the documents."""
    nd = tokens.shape[0]
    bounds = np.linspace(0, nd, THREADS + 1).astype(int)

    def part(i):
        lo, hi = bounds[i], bounds[i + 1]
        if lo == hi:
            return []
        _, r = oracle.maxsim_scan_timed(metric, tokens[lo:hi], q[None, :, :], k, 1)
        return [(lo + d, s) for d, s in r[0]]

    with ThreadPoolExecutor(THREADS) as ex:
        parts = list(ex.map(part, range(THREADS)))
    pool = [e for p in parts for e in p]
    pool.sort(key=lambda e: (-e[1], e[0]))
    return pool[:k]


@pytest.mark.parametrize("metric", ["inner_product", "cosine"])
def test_k5_maxsim_tc_50k_docs_128x128_32_query_tokens(metric):
    """configs[4] per-GPU shape on the tensor-core kernel: 50k docs x 128 tokens x 128 dims, 32-token query."""
    dev = torch.device("cuda", 0)
    nd, td, d, tq, k = 50_000, 128, 128, 32, 10
    idx = nifs.mv_new(metric)
    assert nifs.mv_reserve(idx, nd, nd * td, d) == ("ok", ())
    host = np.empty((nd, td, d), dtype=np.float32)
    for s in range(0, nd, 10_000):
        x = _normal_rows(10_000 * td, d, SEED + s, dev)
        assert nifs.mv_insert_device(idx, nifs.decimal_ids(s, 10_000), x.data_ptr(), td, d) == ("ok", ())
        host[s:s + 10_000] = x.cpu().numpy().reshape(10_000, td, d)
        del x
    q = _normal_rows(tq, d, SEED + 5, dev).cpu().numpy()
    st, hits = nifs.mv_search(idx, q, k)
    assert st == "ok", hits
    ref = _oracle_maxsim_parallel(metric, host, q, k)
    assert_hits_match([(int(i), s) for i, s in hits], ref)


@pytest.mark.parametrize("metric", ["inner_product", "cosine"])
def test_k5_ragged_maxsim_30k_docs_of_40_to_180_tokens(metric):
    """The C5 shape with ragged documents (the ragged tensor-core kernel, csrc/maxsim_tcr.cu): 30k documents of 40-180
    tokens x 128 dims (3.3M tokens, 1.7 GB), 32-token query, device ingest; the oracle scores every document."""
    dev = torch.device("cuda", 0)
    nd, d, tq, k = 30_000, 128, 32, 10
    lens = np.random.default_rng(3).integers(40, 181, nd)
    idx = nifs.mv_new(metric)
    assert nifs.mv_reserve(idx, nd, int(lens.sum()), d) == ("ok", ())
    docs = []
    for s in range(0, nd, 10_000):
        doc_tok = np.concatenate([[0], np.cumsum(lens[s:s + 10_000])]).astype(np.uint64)
        x = _normal_rows(int(doc_tok[-1]), d, SEED + s, dev)
        assert nifs.mv_insert_ragged_device(idx, nifs.decimal_ids(s, 10_000), x.data_ptr(), doc_tok, d) == ("ok", ())
        host = x.cpu().numpy()
        docs.extend(host[int(doc_tok[i]):int(doc_tok[i + 1])] for i in range(10_000))
        del x
    q = _normal_rows(tq, d, SEED + 5, dev).cpu().numpy()
    st, hits = nifs.mv_search(idx, q, k)
    assert st == "ok", hits
    assert _lib.lib().vb_debug_maxsim_path() == 2
    code = nifs.METRIC_CODE[metric]
    bounds = np.linspace(0, nd, THREADS + 1).astype(int)

    def part(i):
        lo, hi = bounds[i], bounds[i + 1]
        r = oracle.multi_vector_top_k([(f"{j:09d}", docs[j]) for j in range(lo, hi)], q, code, k)
        assert r[0] == "ok"
        return r[1]

    with ThreadPoolExecutor(THREADS) as ex:
        pool = [e for p in ex.map(part, range(THREADS)) for e in p]
    pool.sort(key=lambda e: (-e[1], e[0]))
    assert_hits_match(hits, pool[:k])
