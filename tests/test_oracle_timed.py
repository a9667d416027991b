"""The timed oracle legs bench.py uses as CPU baseline AND parity checker must agree with the oracle's own
top-k restatements (which tests/test_oracle_golden.py pins to the reference's known answers)."""
import numpy as np

import oracle
from vettore_b200 import nifs


def test_flat_scan_timed_returns_every_querys_hits():
    rng = np.random.default_rng(1)
    rows = rng.integers(-3, 4, size=(400, 12)).astype(np.float32)      # many exact ties: id order decides
    q = rng.integers(-3, 4, size=(5, 12)).astype(np.float32)
    for metric in ("cosine", "inner_product", "l2"):
        _, res = oracle.flat_scan_timed(metric, rows, q, 9, 3)
        for qi in range(5):
            st, ref = oracle.flat_search_dense(metric, rows, None, q[qi], 9)
            assert st == "ok" and [(f"{r:09d}", v) for r, v in res[qi]] == ref


def test_binary_scan_timed_equals_binary_top_k():
    rng = np.random.default_rng(2)
    n, dims, nw = 500, 130, 3
    codes = rng.integers(0, 2 ** 63, size=(n, nw), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, nw), dtype=np.uint64)
    codes[:, :2] &= np.uint64(0xF)                                      # narrow distances: big ties
    ids = [f"{i:09d}" for i in range(n)]
    _, res = oracle.binary_scan_timed(codes, dims, codes[:3], 25, 2)
    for qi in range(3):
        st, ref = oracle.binary_top_k([(ids[i], codes[i]) for i in range(n)], codes[qi], dims, 25)
        assert st == "ok" and [(ids[r], v) for r, v in res[qi]] == ref


def test_maxsim_scan_timed_equals_multi_vector_top_k():
    rng = np.random.default_rng(3)
    nd, td, d, tq = 60, 5, 16, 3
    tok = rng.standard_normal((nd, td, d)).astype(np.float32)
    q = rng.standard_normal((2, tq, d)).astype(np.float32)
    ids = [f"{i:09d}" for i in range(nd)]
    for metric in ("inner_product", "cosine", "l2", "manhattan"):
        _, res = oracle.maxsim_scan_timed(metric, tok, q, 7, 2)
        for qi in range(2):
            st, ref = oracle.multi_vector_top_k([(ids[i], tok[i]) for i in range(nd)], q[qi], oracle.METRIC_CODE[metric], 7)
            assert st == "ok" and [(ids[r], v) for r, v in res[qi]] == ref


def test_decimal_id_blob_matches_formatted_strings():
    b = nifs.decimal_ids(999_995, 12)
    assert [b[i] for i in range(12)] == [f"{999_995 + i:09d}" for i in range(12)]
    assert len(b) == 12 and int(b.off[-1]) == 12 * 9
