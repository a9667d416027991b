#!/usr/bin/env python
"""Prefix scan of every row (1M x 768, first 128 columns, true cosine) over the result size and the stage-elimination
masks (VB_SCAN_DEBUG: 1 skip emit, 2 skip the FMAs, 4 skip the collector checkpoints; results under a mask are garbage
by design): lane-per-row kernel against the warp-per-row kernels (VB_SCAN_NO_LANE=1). Wall clock through the resident
call (about 35 us of host work per call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
n, d, dims = 1_000_000, 768, int(os.environ.get("PREFIX", "128"))
dev = torch.device("cuda", 0)
idx = nifs.flat_new_cosine()
assert nifs.flat_reserve(idx, n) == ("ok", ())
blk = make_rows_torch(n, d, SEED, dev)
assert nifs.flat_insert_device(idx, [f"{i:09d}" for i in range(n)], blk.data_ptr(), d) == ("ok", ())
del blk
q = make_rows_torch(1, d, SEED + 1, dev)[0].cpu().numpy()
def timed(limit, iters=200):
    for _ in range(5): nifs.flat_prefix_top_k(idx, None, q, 2, dims, limit)
    t0 = time.perf_counter()
    for _ in range(iters): r = nifs.flat_prefix_top_k(idx, None, q, 2, dims, limit)
    return (time.perf_counter() - t0) / iters * 1e3
for lane in (0, 1):
    os.environ.pop("VB_SCAN_NO_LANE", None)
    if not lane: os.environ["VB_SCAN_NO_LANE"] = "1"
    for dbg in ("0", "1", "5", "7"):
        os.environ["VB_SCAN_DEBUG"] = dbg
        print("lane" if lane else "warp", "debug", dbg, {k: round(timed(k), 4) for k in ((1, 10, 100, 1000) if dbg == "0" else (100,))}, flush=True)
