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


# ---------------------------------------------------------------------------------------------
# Row-sharded quantized_search and document-sharded MaxSim:
# the shards live on one device here and the "all-gather" is the layout NCCL would produce.
from test_pipelines_gpu import ref_quantized  # noqa: E402  (oracle composition of collection.ex:699-713)
from vettore_b200.sharded import ShardedMv, ShardedQuantized, set_global_mv_ranks  # noqa: E402


def _merge(gathered, lay, lists, k_in, k_out, dev, stream):
    out_keys = torch.zeros(k_out, dtype=torch.int64, device=dev)
    out_vals = torch.zeros(k_out, dtype=torch.float32, device=dev)
    out_rows = torch.zeros(k_out, dtype=torch.int64, device=dev)
    out_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    g = gathered.data_ptr()
    rc = _lib.lib().vb_topk_merge_device(C.c_void_p(g + lay["keys"]), C.c_void_p(g + lay["values"]),
                                         C.c_void_p(g + lay["rows"]), C.c_void_p(g + lay["counts"]), lay["bytes"],
                                         1, lists, k_in, k_out, C.c_void_p(out_keys.data_ptr()),
                                         C.c_void_p(out_vals.data_ptr()), C.c_void_p(out_rows.data_ptr()),
                                         C.c_void_p(out_cnt.data_ptr()), stream)
    assert rc == 0, _lib.last_error()
    return out_vals, out_rows, out_cnt


@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product"])
@pytest.mark.parametrize("shards,n_per,d,cand,k", [(2, 2000, 256, 100, 10), (4, 900, 128, 300, 10), (3, 50, 96, 1000, 7)])
def test_sharded_quantized_equals_single_index_and_oracle(shards, n_per, d, cand, k, metric):
    dev = torch.device("cuda", 0)
    n = shards * n_per
    rows = _rows(n, d, seed=shards * 77 + d)
    q = _rows(1, d, seed=11)
    ids = [f"{i:09d}" for i in range(n)]
    code = nifs.METRIC_CODE[metric]
    cand_eff, k_eff = min(cand, 1024), k
    lay_c, lay_k = packed_layout(1, cand_eff), packed_layout(1, k_eff)
    gath_c = torch.zeros(shards * lay_c["bytes"], dtype=torch.uint8, device=dev)
    gath_k = torch.zeros(shards * lay_k["bytes"], dtype=torch.uint8, device=dev)
    dq = torch.from_numpy(q).to(dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shard_idx = []
    for s in range(shards):
        idx = getattr(nifs, f"flat_new_{metric}")()
        assert nifs.flat_insert_matrix(idx, ids[s * n_per:(s + 1) * n_per], rows[s * n_per:(s + 1) * n_per]) == ("ok", ())
        set_global_ranks(idx, s * n_per, n_per)
        shard_idx.append(idx)
        b = gath_c.data_ptr() + s * lay_c["bytes"]
        rc = _lib.lib().vb_flat_hamming_device(idx.handle, C.c_void_p(dq.data_ptr()), 1, d, cand_eff,
                                               C.c_void_p(b + lay_c["keys"]), C.c_void_p(b + lay_c["values"]),
                                               C.c_void_p(b + lay_c["rows"]), C.c_void_p(b + lay_c["counts"]), stream)
        assert rc == 0, _lib.last_error()
    _, c_rows, c_cnt = _merge(gath_c, lay_c, shards, cand_eff, cand_eff, dev, stream)
    torch.cuda.synchronize()
    assert int(c_cnt.item()) == min(cand_eff, n)
    for s, idx in enumerate(shard_idx):
        b = gath_k.data_ptr() + s * lay_k["bytes"]
        rc = _lib.lib().vb_flat_rerank_owned_device(idx.handle, C.c_void_p(dq.data_ptr()), d, code,
                                                    C.c_void_p(c_rows.data_ptr()), C.c_void_p(c_cnt.data_ptr()), cand_eff, s,
                                                    k_eff, C.c_void_p(b + lay_k["keys"]), C.c_void_p(b + lay_k["values"]),
                                                    C.c_void_p(b + lay_k["rows"]), C.c_void_p(b + lay_k["counts"]), stream)
        assert rc == 0, _lib.last_error()
    vals, rws, cnt = _merge(gath_k, lay_k, shards, k_eff, k_eff, dev, stream)
    torch.cuda.synchronize()
    vals, rws = vals.cpu().numpy(), rws.cpu().numpy().astype(np.uint64)
    got = [(ids[int(rws[i] >> np.uint64(32)) * n_per + int(rws[i] & np.uint64(0xFFFFFFFF))], float(vals[i]))
           for i in range(int(cnt.item()))]
    vectors = [(ids[i], rows[i]) for i in range(n)]
    assert_hits_match(got, ref_quantized(vectors, q[0], code, cand_eff, k_eff))
    whole = getattr(nifs, f"flat_new_{metric}")()
    assert nifs.flat_insert_matrix(whole, ids, rows) == ("ok", ())
    st, single = nifs.flat_quantized_search(whole, q[0], code, cand_eff, k_eff)
    assert st == "ok"
    assert_hits_match(got, single)


def test_sharded_quantized_wrapper_single_rank():
    n, d = 6000, 192
    rows, q = _rows(n, d, 21), _rows(1, d, 22)
    ids = [f"{i:09d}" for i in range(n)]
    idx = nifs.flat_new_cosine()
    assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
    sh = ShardedQuantized(idx, candidates=200, limit=10, metric_code=nifs.METRIC_CODE["cosine"])
    hits = sh.search(torch.from_numpy(q).pin_memory())
    got = [(ids[h.row], h.value) for h in hits]
    st, exp = nifs.flat_quantized_search(idx, q[0], nifs.METRIC_CODE["cosine"], 200, 10)
    assert st == "ok"
    assert_hits_match(got, exp)


def _docs(ndocs, td, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((ndocs, td, d)).astype(np.float32)
    return (x / np.linalg.norm(x.astype(np.float64), axis=2, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("metric", ["inner_product", "cosine", "l2"])
@pytest.mark.parametrize("shards,docs_per,td,d,tq,k", [(2, 300, 32, 64, 8, 10), (4, 100, 128, 128, 32, 10), (3, 40, 5, 48, 3, 50)])
def test_sharded_maxsim_equals_single_index_and_oracle(shards, docs_per, td, d, tq, k, metric):
    dev = torch.device("cuda", 0)
    ndocs = shards * docs_per
    docs = _docs(ndocs, td, d, seed=shards * 31 + td)
    query = _docs(1, tq, d, seed=5)[0]
    ids = [f"{i:09d}" for i in range(ndocs)]
    lay = packed_layout(1, k)
    gathered = torch.zeros(shards * lay["bytes"], dtype=torch.uint8, device=dev)
    keep = []
    for s in range(shards):
        mv = nifs.mv_new(metric)
        lo, hi = s * docs_per, (s + 1) * docs_per
        assert nifs.mv_insert_tensor(mv, ids[lo:hi], docs[lo:hi])[0] == "ok"
        set_global_mv_ranks(mv, lo, docs_per)
        keep.append(mv)
        b = gathered.data_ptr() + s * lay["bytes"]
        res = nifs.mv_search_packed_device(mv, query, k, C.c_void_p(b + lay["keys"]), C.c_void_p(b + lay["values"]),
                                           C.c_void_p(b + lay["rows"]), C.c_void_p(b + lay["counts"]))
        assert res[0] == "ok", res
    torch.cuda.synchronize()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vals, rws, cnt = _merge(gathered, lay, shards, k, k, dev, stream)
    torch.cuda.synchronize()
    vals, rws = vals.cpu().numpy(), rws.cpu().numpy().astype(np.uint64)
    got = [(ids[int(rws[i] >> np.uint64(32)) * docs_per + int(rws[i] & np.uint64(0xFFFFFFFF))], float(vals[i]))
           for i in range(int(cnt.item()))]
    whole = nifs.mv_new(metric)
    assert nifs.mv_insert_tensor(whole, ids, docs)[0] == "ok"
    st, single = nifs.mv_search(whole, query, k)
    assert st == "ok"
    assert_hits_match(got, single)
    st, exp = oracle.multi_vector_top_k([(ids[i], docs[i]) for i in range(ndocs)], query, nifs.METRIC_CODE[metric], k)
    assert st == "ok"
    assert_hits_match(got, exp)


def test_sharded_maxsim_wrapper_single_rank():
    docs, query = _docs(500, 32, 64, 9), _docs(1, 16, 64, 10)[0]
    ids = [f"{i:09d}" for i in range(500)]
    mv = nifs.mv_new("inner_product")
    assert nifs.mv_insert_tensor(mv, ids, docs)[0] == "ok"
    hits = ShardedMv(mv, k=10).search(query)
    st, exp = nifs.mv_search(mv, query, 10)
    assert st == "ok"
    assert_hits_match([(ids[h.row], h.value) for h in hits], exp)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL exchange between real ranks)")
def test_nccl_two_rank_sharded_paths_match_single_index():
    """tools/sharded_check.py under torchrun: flat (K1 and K2 batch), quantized and MaxSim over 2 ranks."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "sharded_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "sharded paths match the single-index results" in res.stdout


# ---------------------------------------------------------------------------------------------
# NVLink peer-memory exchange + fused select (csrc/peer_exchange.cu). `world` virtual ranks live on one
# device here (vb_peer_connect_local), each on its own stream; the result of every rank must equal the K7
# merge of the same records laid out as an all-gather would lay them out.
def _random_record(rng, lay, nq, k, shard):
    rec = np.zeros(lay["bytes"], dtype=np.uint8)
    keys = rec[lay["keys"]:lay["keys"] + nq * k * 8].view(np.uint64).reshape(nq, k)
    vals = rec[lay["values"]:lay["values"] + nq * k * 4].view(np.float32).reshape(nq, k)
    rows = rec[lay["rows"]:lay["rows"] + nq * k * 4].view(np.uint32).reshape(nq, k)
    cnts = rec[lay["counts"]:lay["counts"] + nq * 4].view(np.uint32)
    for q in range(nq):
        c = int(rng.integers(0, k + 1)) if (q + shard) % 3 == 0 else k
        # few distinct high words: plenty of cross-shard ties decided by the low (id rank) word
        hi = np.sort(rng.integers(0, 6, size=c).astype(np.uint64))
        lo = rng.permutation(1 << 20)[:c].astype(np.uint64) * np.uint64(8) + np.uint64(shard)
        kk = np.sort((hi << np.uint64(32)) | lo)
        keys[q, :c] = kk
        keys[q, c:] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vals[q, :c] = (kk >> np.uint64(32)).astype(np.float32) + 0.25 * shard
        rows[q, :c] = (kk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        cnts[q] = c
    return rec


@pytest.mark.parametrize("world,nq,k_in,k_out", [(2, 1, 10, 10), (8, 1, 10, 10), (4, 1, 1000, 1000), (8, 3, 100, 100),
                                                 (8, 64, 100, 100), (3, 300, 10, 10), (2, 1, 1000, 1000),
                                                 (8, 2, 1000, 1000), (8, 1, 1024, 100), (5, 1, 300, 300)])
def test_peer_exchange_merge_equals_gather_then_merge(world, nq, k_in, k_out):
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(world * 1000 + nq)
    lay = packed_layout(nq, k_in)
    peers = []
    for r in range(world):
        px = C.c_void_p()
        assert L.vb_peer_new(world, r, lay["bytes"], C.byref(px), None) == 0, _lib.last_error()
        peers.append(px)
    arr = (C.c_void_p * world)(*peers)
    assert L.vb_peer_connect_local(arr, world) == 0, _lib.last_error()
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    o_keys, o_vals = 0, nq * k_out * 8
    o_rows = (o_vals + nq * k_out * 4 + 7) // 8 * 8
    o_cnt = o_rows + nq * k_out * 8
    outs = [torch.zeros(o_cnt + nq * 4 + 16, dtype=torch.uint8, device=dev) for _ in range(world)]
    try:
        for step in range(3):          # several epochs: both buffer parities and the flag protocol
            recs = [_random_record(rng, lay, nq, k_in, r) for r in range(world)]
            d_recs = [torch.from_numpy(x).to(dev) for x in recs]
            gathered = torch.from_numpy(np.concatenate(recs)).to(dev)
            torch.cuda.synchronize()
            p = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
            if nq <= 4:
                for r in range(world):   # fused push + wait + select, one launch per rank, concurrent streams
                    rc = L.vb_peer_exchange_merge(peers[r], p(d_recs[r]), nq, k_in, k_out, lay["keys"], lay["values"],
                                                  lay["rows"], lay["counts"], p(outs[r], o_keys), p(outs[r], o_vals),
                                                  p(outs[r], o_rows), p(outs[r], o_cnt), C.c_void_p(streams[r].cuda_stream))
                    assert rc == 0, _lib.last_error()
            else:
                for r in range(world):   # one host thread drives every rank: all pushes first
                    rc = L.vb_peer_push(peers[r], p(d_recs[r]), nq, k_in, k_out, lay["keys"], lay["values"], lay["rows"],
                                        lay["counts"], C.c_void_p(streams[r].cuda_stream))
                    assert rc == 0, _lib.last_error()
                for r in range(world):
                    rc = L.vb_peer_wait_merge(peers[r], nq, k_in, k_out, lay["keys"], lay["values"], lay["rows"],
                                              lay["counts"], p(outs[r], o_keys), p(outs[r], o_vals), p(outs[r], o_rows),
                                              p(outs[r], o_cnt), C.c_void_p(streams[r].cuda_stream))
                    assert rc == 0, _lib.last_error()
            torch.cuda.synchronize()
            ref = torch.zeros_like(outs[0])
            g = gathered.data_ptr()
            rc = L.vb_topk_merge_device(C.c_void_p(g + lay["keys"]), C.c_void_p(g + lay["values"]), C.c_void_p(g + lay["rows"]),
                                        C.c_void_p(g + lay["counts"]), lay["bytes"], nq, world, k_in, k_out, p(ref, o_keys),
                                        p(ref, o_vals), p(ref, o_rows), p(ref, o_cnt),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, _lib.last_error()
            torch.cuda.synchronize()
            ref_h = ref.cpu().numpy()
            for r in range(world):
                err = C.c_uint32(9)
                assert L.vb_peer_error(peers[r], C.byref(err)) == 0 and err.value == 0
                assert np.array_equal(outs[r].cpu().numpy()[:o_cnt + nq * 4], ref_h[:o_cnt + nq * 4]), (step, r)
    finally:
        torch.cuda.synchronize()
        for px in peers:
            L.vb_peer_free(px)
