#!/usr/bin/env python
"""Weak-scaling check of the sharded batch (C3) and quantized (C4) paths over NCCL, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        tools/bench_sharded_scale.py [--rows-per-gpu 4000000]
Every rank ingests its shard on the device (vb_flat_insert_many_device), then
  batch:     ShardedFlat(nq=1024, k=100).search_device  (K2 per shard -> all-gather -> K7), device-timed, max over ranks
  quantized: ShardedQuantized(1000 candidates -> k=10).search (K6+K3 -> all-gather -> K7 -> owner rerank K4 -> all-gather -> K7), wall clock
Rank 0 prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
from vettore_b200.sharded import ShardedFlat, ShardedQuantized, set_global_ranks

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=4_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = a.rows_per_gpu
idx = nifs.flat_new_inner_product()
assert nifs.flat_reserve(idx, n) == ("ok", ())
base = rank * n
for s in range(0, n, 1_000_000):
    m = min(1_000_000, n - s)
    blk = make_rows_torch(m, a.dim, SEED + 1000 * rank + s // 1_000_000, dev)
    assert nifs.flat_insert_device(idx, [f"{base + i:010d}" for i in range(s, s + m)], blk.data_ptr(), a.dim) == ("ok", ())
    del blk
set_global_ranks(idx, base, n)
queries = make_rows_torch(a.nq, a.dim, SEED + 1, dev)   # same on every rank (same seed)

def max_over_ranks(ms):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

out = {"world": world, "rows_per_gpu": n, "rows_total": n * world, "dim": a.dim}
sh = ShardedFlat(idx, k=a.k, nq=a.nq)
for _ in range(2):
    sh.search_device(queries)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    res = sh.search_device(queries)
e1.record()
torch.cuda.synchronize()
ms = max_over_ranks(e0.elapsed_time(e1) / a.steps)
out["batch"] = {"nq": a.nq, "k": a.k, "ms_per_batch": ms, "queries_per_s": a.nq / ms * 1e3,
                "corpus_rows_per_s": a.nq * n * world / ms * 1e3}

sq = ShardedQuantized(idx, candidates=1000, limit=10, metric_code=nifs.METRIC_CODE["inner_product"])
qh = queries[:1].cpu().pin_memory()
for _ in range(3):
    hits = sq.search(qh)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    hits = sq.search(qh)
qms = max_over_ranks((time.perf_counter() - t0) / 20 * 1e3)
out["quantized"] = {"candidates": 1000, "k": 10, "ms_per_query": qms, "queries_per_s": 1e3 / qms,
                    "code_gb_total": n * world * ((a.dim + 63) // 64) * 8 / 1e9,
                    "top": [[h.shard, h.row, h.value] for h in hits[:2]]}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
