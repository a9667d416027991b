"""GPU parity of the lane-per-row true-cosine prefix kernel (csrc/prefix_lane.cu) against the oracle
(search.rs:38-73 with distances.rs:160-177): every row scored over its first `dims` <= 128 columns, through the
resident prefix scan, the funnel's first stage (main matrix and dense prefix mirror) and narrow whole rows."""
import numpy as np
import pytest

import oracle
from helpers import assert_hits_match
from vettore_b200 import _lib, nifs

pytestmark = pytest.mark.gpu


def ok(x):
    assert x[0] == "ok", x
    return x[1]


def scan_path():
    """0 kernel A, 1 kernel B whole rows, 2 kernel B prefix box, 3 kernel C one row per lane (flat_scan.cu)."""
    return _lib.lib().vb_debug_scan_path()


def _rows(n, d, seed, scale=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if scale:   # true cosine must not depend on the row norms
        x *= rng.uniform(0.1, 30.0, (n, 1)).astype(np.float32)
    return x


@pytest.mark.parametrize("n,d,dims", [(5000, 256, 128), (4097, 256, 100), (9000, 96, 33), (6000, 64, 5), (4100, 160, 32),
                                      (7001, 128, 128), (5003, 40, 40), (4999, 768, 64)])
def test_prefix_scan_of_every_row_matches_oracle(n, d, dims):
    rows = _rows(n, d, n + d)
    rows[17] = 0.0                                     # a zero row scores 0 (distances.rs:170-172)
    rows[18, :dims] = 0.0                              # zero prefix, data beyond it
    ids = [f"{(i * 7919) % n:06d}" if np.gcd(7919, n) == 1 else f"{i:06d}" for i in range(n)]
    q = _rows(1, d, 3)[0]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    vectors = [(ids[i], rows[i]) for i in range(n)]
    for limit in (1, 10, 100, 1000):
        got = ok(nifs.flat_prefix_top_k(idx, None, q, 2, dims, limit))
        assert scan_path() == 3, "the lane-per-row kernel must answer this scan"
        assert_hits_match(got, ok(oracle.vector_top_k(vectors, q, 2, dims, limit)))


def test_lane_kernel_agrees_with_the_warp_per_row_kernels(monkeypatch):
    n, d, dims = 30000, 256, 128
    rows = _rows(n, d, 5)
    rows[100:140] = rows[99]                           # equal scores: ties resolve by id in both kernels
    ids = [f"{(i * 7919) % n:06d}" for i in range(n)]
    q = _rows(1, d, 6)[0]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    lane = ok(nifs.flat_prefix_top_k(idx, None, q, 2, dims, 500))
    assert scan_path() == 3
    monkeypatch.setenv("VB_SCAN_NO_LANE", "1")
    warp = ok(nifs.flat_prefix_top_k(idx, None, q, 2, dims, 500))
    assert scan_path() != 3
    assert_hits_match(lane, warp)


@pytest.mark.parametrize("warps,depth", [("4", "0"), ("6", "0"), ("8", "1"), ("2", "3"), ("12", "0"), ("3", "2")])
def test_ring_geometries(monkeypatch, warps, depth):
    """Consumer warps and ring depth (tiles per warp) are launch parameters: a box slot always belongs to one warp."""
    monkeypatch.setenv("VB_LANE_WARPS", warps)
    if depth != "0":
        monkeypatch.setenv("VB_LANE_DEPTH", depth)
    n, d, dims = 20011, 128, 96
    rows = _rows(n, d, 9)
    ids = [f"{i:06d}" for i in range(n)]
    q = _rows(1, d, 10)[0]
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_prefix_top_k(idx, None, q, 2, dims, 100))
    assert scan_path() == 3
    assert_hits_match(got, ok(oracle.vector_top_k([(ids[i], rows[i]) for i in range(n)], q, 2, dims, 100)))


def test_overflowing_true_cosine_is_the_reference_error():
    # distances.rs:160-177: the f64 quotient of finite f32 inputs is always finite, so nothing overflows here; rows of
    # huge and tiny magnitude must still score exactly like the oracle (f64 accumulation, no f32 intermediate).
    n, d, dims = 5000, 64, 48
    rows = _rows(n, d, 11, scale=False)
    rows[::3] *= 1.0e18
    rows[1::3] *= 1.0e-18
    ids = [f"{i:05d}" for i in range(n)]
    q = (_rows(1, d, 12, scale=False)[0] * 1.0e15).astype(np.float32)
    idx = nifs.flat_new_cosine()
    ok(nifs.flat_insert_matrix(idx, ids, rows))
    got = ok(nifs.flat_prefix_top_k(idx, None, q, 2, dims, 50))
    assert scan_path() == 3
    assert_hits_match(got, ok(oracle.vector_top_k([(ids[i], rows[i]) for i in range(n)], q, 2, dims, 50)))
