"""GPU parity of the row-sharded path: per-shard fused scan (device-level C-ABI entry),
packed records laid out exactly as an all-gather would, then the K7 merge kernel. Shards
live on one device here; the NCCL exchange itself is exercised by bench.py --gpus N."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from helpers import assert_hits_match
from vettore_b200 import _lib, nifs
from vettore_b200.sharded import ShardedFlat, packed_layout, set_global_ranks

pytestmark = pytest.mark.gpu


def _rows(n, d, seed, ints=False):
    rng = np.random.default_rng(seed)
    if ints:
        return rng.integers(-2, 3, size=(n, d)).astype(np.float32)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("ints", [False, True])
@pytest.mark.parametrize("shards,n_per,d,k,nq", [(2, 3000, 384, 10, 1), (4, 1500, 96, 100, 3), (8, 700, 768, 10, 2)])
def test_sharded_search_equals_single_index_and_oracle(shards, n_per, d, k, nq, ints):
    dev = torch.device("cuda", 0)
    rows = _rows(shards * n_per, d, seed=shards * 1000 + d, ints=ints)
    queries = _rows(nq, d, seed=7, ints=ints)
    ids = [f"{i:09d}" for i in range(shards * n_per)]
    metric = "inner_product"
    lay = packed_layout(nq, k)
    gathered = torch.zeros(shards * lay["bytes"], dtype=torch.uint8, device=dev)
    dq = torch.from_numpy(queries).to(dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    keep = []
    for s in range(shards):
        idx = nifs.flat_new_inner_product()
        assert nifs.flat_insert_matrix(idx, ids[s * n_per:(s + 1) * n_per], rows[s * n_per:(s + 1) * n_per]) == ("ok", ())
        set_global_ranks(idx, s * n_per, n_per)
        keep.append(idx)
        base = gathered.data_ptr() + s * lay["bytes"]
        rc = _lib.lib().vb_flat_search_device(idx.handle, C.c_void_p(dq.data_ptr()), nq, d, k,
                                              C.c_void_p(base + lay["keys"]), C.c_void_p(base + lay["values"]),
                                              C.c_void_p(base + lay["rows"]), C.c_void_p(base + lay["counts"]), stream)
        assert rc == 0, _lib.last_error()
    out_keys = torch.zeros(nq * k, dtype=torch.int64, device=dev)
    out_vals = torch.zeros(nq * k, dtype=torch.float32, device=dev)
    out_rows = torch.zeros(nq * k, dtype=torch.int64, device=dev)
    out_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    g = gathered.data_ptr()
    rc = _lib.lib().vb_topk_merge_device(C.c_void_p(g + lay["keys"]), C.c_void_p(g + lay["values"]),
                                         C.c_void_p(g + lay["rows"]), C.c_void_p(g + lay["counts"]), lay["bytes"],
                                         nq, shards, k, k, C.c_void_p(out_keys.data_ptr()),
                                         C.c_void_p(out_vals.data_ptr()), C.c_void_p(out_rows.data_ptr()),
                                         C.c_void_p(out_cnt.data_ptr()), stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    vals = out_vals.cpu().numpy().reshape(nq, k)
    rws = out_rows.cpu().numpy().astype(np.uint64).reshape(nq, k)
    cnt = out_cnt.cpu().numpy()
    for q in range(nq):
        got = [(ids[int(rws[q, i] >> np.uint64(32)) * n_per + int(rws[q, i] & np.uint64(0xFFFFFFFF))], float(vals[q, i]))
               for i in range(int(cnt[q]))]
        st, exp = oracle.flat_search_dense(metric, rows, ids, queries[q], k)
        assert st == "ok"
        assert_hits_match(got, exp, exact_ids=ints)


def test_single_shard_wrapper_roundtrip():
    n, d, k = 5000, 128, 10
    rows = _rows(n, d, 3)
    ids = [f"{i:09d}" for i in range(n)]
    idx = nifs.flat_new_cosine()
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    sh = ShardedFlat(idx, k=k, nq=1)
    q = torch.from_numpy(_rows(1, d, 4)).pin_memory()
    hits = sh.search(q)[0]
    got = [(ids[h.row], h.value) for h in hits]
    assert_hits_match(got, oracle.flat_search_dense("cosine", rows, ids, q[0].numpy(), k)[1])
