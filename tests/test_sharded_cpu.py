"""Host-side logic of the row-sharded path on CPU: world_size-2 gloo process group, one
packed top-k record per rank, ONE all-gather, then the final select. The CUDA kernels are
not involved here (no GPU in this container): per-shard lists come from the oracle and the
final select is a plain sort, which is exactly the property the K7 kernel must satisfy:
merge of per-shard top-k lists == top-k of the whole corpus, ties broken by global id rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import total_order_key
from vettore_b200.sharded import packed_layout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_per, d, k, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1234)
    rows = rng.integers(-2, 3, size=(world * n_per, d)).astype(np.float32)  # small ints: many exact ties
    q = rng.integers(-2, 3, size=d).astype(np.float32)
    base = rank * n_per
    ids = [f"{base + i:09d}" for i in range(n_per)]
    st, hits = oracle.flat_search_dense("inner_product", rows[base:base + n_per], ids, q, k)
    assert st == "ok"
    lay = packed_layout(1, k)
    rec = np.zeros(lay["bytes"], dtype=np.uint8)
    keys = rec[lay["keys"]:lay["keys"] + k * 8].view(np.uint64)
    vals = rec[lay["values"]:lay["values"] + k * 4].view(np.float32)
    rws = rec[lay["rows"]:lay["rows"] + k * 4].view(np.uint32)
    cnt = rec[lay["counts"]:lay["counts"] + 4].view(np.uint32)
    for i, (hid, raw) in enumerate(hits):
        rank_key = total_order_key(oracle.rank_value("inner_product", raw))
        keys[i] = (rank_key << 32) | int(hid)          # global id rank == global row number
        vals[i] = raw
        rws[i] = int(hid) - base
    cnt[0] = len(hits)
    local = torch.from_numpy(rec)
    gathered = torch.zeros(world * lay["bytes"], dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, local)
    g = gathered.numpy()
    cands = []
    for shard in range(world):
        o = shard * lay["bytes"]
        c = int(g[o + lay["counts"]:o + lay["counts"] + 4].view(np.uint32)[0])
        ks = g[o + lay["keys"]:o + lay["keys"] + k * 8].view(np.uint64)
        vs = g[o + lay["values"]:o + lay["values"] + k * 4].view(np.float32)
        rs = g[o + lay["rows"]:o + lay["rows"] + k * 4].view(np.uint32)
        cands += [(int(ks[i]), shard, int(rs[i]), float(vs[i])) for i in range(c)]
    cands.sort()
    merged = [(f"{s * n_per + r:09d}", v) for _, s, r, v in cands[:k]]
    all_ids = [f"{i:09d}" for i in range(world * n_per)]
    st, expect = oracle.flat_search_dense("inner_product", rows, all_ids, q, k)
    assert merged == expect, (merged, expect)
    if rank == 0:
        open(out_path, "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [1, 10, 64])
def test_two_rank_gloo_all_gather_merge_equals_global_top_k(tmp_path, k):
    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(2, _free_port(), 500, 8, k, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_packed_layout_is_aligned_and_disjoint():
    for nq, k in [(1, 10), (3, 7), (1024, 100)]:
        lay = packed_layout(nq, k)
        assert lay["keys"] % 8 == 0 and lay["values"] % 4 == 0 and lay["rows"] % 4 == 0 and lay["counts"] % 4 == 0
        assert lay["values"] - lay["keys"] == nq * k * 8 and lay["rows"] - lay["values"] == nq * k * 4
        assert lay["bytes"] % 16 == 0 and lay["bytes"] >= lay["counts"] + 4 * nq


def _quantized_worker(rank, world, port, n_per, d, cand, k, out_path):
    """Two-stage exchange of the row-sharded quantized_search (ShardedQuantized's host logic): local Hamming
    candidates -> all-gather -> global candidate set -> every owner reranks its survivors -> all-gather -> top-k.
    Must equal the reference composition (collection.ex:699-713) over the whole corpus."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    rows = rng.integers(-3, 4, size=(world * n_per, d)).astype(np.float32)   # coarse values: many Hamming ties
    rows[rows == 0] = 1.0
    q = rng.integers(-3, 4, size=d).astype(np.float32)
    q[q == 0] = -1.0
    code = 2  # cosine: exact rerank uses the true cosine (search.rs:56-60)
    ids = [f"{i:09d}" for i in range(world * n_per)]
    base = rank * n_per
    qcode = oracle.compress_sign_bits(q)
    local_codes = [(ids[base + i], oracle.compress_sign_bits(rows[base + i])) for i in range(n_per)]
    st, local_cands = oracle.binary_top_k(local_codes, qcode, d, cand)
    assert st == "ok"
    gathered = [None] * world
    dist.all_gather_object(gathered, local_cands)
    merged = sorted((h for lst in gathered for h in lst), key=lambda h: (h[1], h[0]))[:cand]
    mine = [(i, rows[int(i)]) for i, _ in merged if base <= int(i) < base + n_per]
    st, local_top = oracle.vector_top_k(mine, q, code, d, k) if mine else ("ok", [])
    assert st == "ok"
    gathered2 = [None] * world
    dist.all_gather_object(gathered2, local_top)
    final = sorted((h for lst in gathered2 for h in lst),
                   key=lambda h: (total_order_key(oracle.rank_value("cosine", h[1])), h[0]))[:k]
    # reference composition over the whole corpus
    codes = [(ids[i], oracle.compress_sign_bits(rows[i])) for i in range(world * n_per)]
    keep = {h[0] for h in oracle.binary_top_k(codes, qcode, d, cand)[1]}
    st, expect = oracle.vector_top_k([(ids[i], rows[i]) for i in range(world * n_per) if ids[i] in keep], q, code, d, k)
    assert st == "ok" and final == expect, (final, expect)
    if rank == 0:
        open(out_path, "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("cand,k", [(20, 5), (150, 10)])
def test_two_rank_gloo_quantized_two_stage_exchange_equals_reference_composition(tmp_path, cand, k):
    out = tmp_path / "ok.txt"
    mp.spawn(_quantized_worker, args=(2, _free_port(), 120, 96, cand, k, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def _maxsim_worker(rank, world, port, docs_per, k, out_path):
    """Document-sharded MaxSim (ShardedMv's host logic): per-shard top-k -> all-gather -> merge by
    (score descending, id ascending) must equal multi_vector_top_k over the whole corpus (multi_vector.rs:90-132)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    docs = rng.integers(-2, 3, size=(world * docs_per, 4, 8)).astype(np.float32)   # coarse values: many score ties
    query = rng.integers(-2, 3, size=(3, 8)).astype(np.float32)
    ids = [f"{i:06d}" for i in range(world * docs_per)]
    lo = rank * docs_per
    st, local = oracle.multi_vector_top_k([(ids[i], docs[i]) for i in range(lo, lo + docs_per)], query, 3, k)
    assert st == "ok"
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = sorted((h for lst in gathered for h in lst), key=lambda h: (-total_order_key(h[1]), h[0]))[:k]
    st, expect = oracle.multi_vector_top_k([(ids[i], docs[i]) for i in range(world * docs_per)], query, 3, k)
    assert st == "ok" and merged == expect, (merged, expect)
    if rank == 0:
        open(out_path, "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_maxsim_merge_equals_global_top_k(tmp_path):
    out = tmp_path / "ok.txt"
    mp.spawn(_maxsim_worker, args=(2, _free_port(), 150, 10, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
