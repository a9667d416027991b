#!/usr/bin/env python
"""Multi-process check of the sharded paths over NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/sharded_check.py
Every rank owns one shard on its own GPU; rank 0 also builds the whole corpus in a single index on its GPU and
compares: flat search (one query: K1; a batch of 32: K2), quantized search (K3 -> all-gather -> K7 -> owner
rerank K4 -> all-gather -> K7) and MaxSim (K5 -> all-gather -> K7). Prints one JSON line with timings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from helpers import assert_hits_match
from vettore_b200 import nifs
from vettore_b200.sharded import (ShardedFlat, ShardedMv, ShardedQuantized, set_global_mv_ranks, set_global_ranks)


def unit(x):
    return (x / np.linalg.norm(x.astype(np.float64), axis=-1, keepdims=True)).astype(np.float32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_per, d, k = 40_000, 256, 10
    n = world * n_per
    rng = np.random.default_rng(2026)
    rows = unit(rng.standard_normal((n, d)).astype(np.float32))
    queries = unit(rng.standard_normal((32, d)).astype(np.float32))
    ids = [f"{i:09d}" for i in range(n)]
    lo, hi = rank * n_per, (rank + 1) * n_per
    report = {"world": world}

    # ---- flat (cosine): one query (K1) and a batch of 32 (K2)
    idx = nifs.flat_new_cosine()
    assert nifs.flat_insert_matrix(idx, ids[lo:hi], rows[lo:hi]) == ("ok", ())
    set_global_ranks(idx, lo, n_per)
    whole = None
    if rank == 0:
        whole = nifs.flat_new_cosine()
        assert nifs.flat_insert_matrix(whole, ids, rows) == ("ok", ())
    for nq in (1, 32):
        sh = ShardedFlat(idx, k=k, nq=nq)
        qh = torch.from_numpy(queries[:nq]).pin_memory()
        hits = sh.search(qh)
        t0 = time.perf_counter()
        for _ in range(5):
            hits = sh.search(qh)
        report[f"flat_nq{nq}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        if rank == 0:
            for qi in range(nq):
                got = [(ids[h.shard * n_per + h.row], h.value) for h in hits[qi]]
                st, exp = nifs.flat_search(whole, queries[qi], k)
                assert st == "ok"
                assert_hits_match(got, exp)

    # ---- quantized (cosine), 1000 candidates
    for cand in (100, 1000):
        sq = ShardedQuantized(idx, candidates=cand, limit=k, metric_code=nifs.METRIC_CODE["cosine"])
        qh = torch.from_numpy(queries[:1]).pin_memory()
        hits = sq.search(qh)
        t0 = time.perf_counter()
        for _ in range(5):
            hits = sq.search(qh)
        report[f"quantized_c{cand}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        if rank == 0:
            got = [(ids[h.shard * n_per + h.row], h.value) for h in hits]
            st, exp = nifs.flat_quantized_search(whole, queries[0], nifs.METRIC_CODE["cosine"], cand, k)
            assert st == "ok"
            assert_hits_match(got, exp)

    # ---- MaxSim (inner product, tensor-core path shape; cosine through the same kernel)
    docs_per, td, dd, tq = 2000, 32, 64, 16
    docs = unit(rng.standard_normal((world * docs_per, td, dd)).astype(np.float32))
    query = unit(rng.standard_normal((tq, dd)).astype(np.float32))
    dids = [f"{i:09d}" for i in range(world * docs_per)]
    for metric in ("inner_product", "cosine"):
        mv = nifs.mv_new(metric)
        assert nifs.mv_insert_tensor(mv, dids[rank * docs_per:(rank + 1) * docs_per],
                                     docs[rank * docs_per:(rank + 1) * docs_per])[0] == "ok"
        set_global_mv_ranks(mv, rank * docs_per, docs_per)
        sm = ShardedMv(mv, k=k)
        hits = sm.search(query)
        t0 = time.perf_counter()
        for _ in range(5):
            hits = sm.search(query)
        report[f"maxsim_{metric}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        if rank == 0:
            allmv = nifs.mv_new(metric)
            assert nifs.mv_insert_tensor(allmv, dids, docs)[0] == "ok"
            st, exp = nifs.mv_search(allmv, query, k)
            assert st == "ok"
            assert_hits_match([(dids[h.shard * docs_per + h.row], h.value) for h in hits], exp)
    dist.barrier()
    if rank == 0:
        report["status"] = "sharded paths match the single-index results"
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
