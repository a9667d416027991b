#!/usr/bin/env python
"""Secondary benchmark: Matryoshka funnel_search (collection.ex:742-760) on the resident flat index:
stage prefixes over ALL rows (K4), survivors re-scored at the next width, exact final rerank.
1M x 768 cosine, stages and candidates from the command line. One JSON line (wall clock, host query in / ids out)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--stages", default="64,256")
ap.add_argument("--candidates", type=int, default=1000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--metric", default="cosine")
a = ap.parse_args()
dev = torch.device("cuda", 0)
stages = [int(x) for x in a.stages.split(",")]
idx = getattr(nifs, f"flat_new_{a.metric}")()
assert nifs.flat_reserve(idx, a.rows) == ("ok", ())
for s in range(0, a.rows, 1_000_000):
    m = min(1_000_000, a.rows - s)
    blk = make_rows_torch(m, a.dim, SEED + s, dev)
    assert nifs.flat_insert_device(idx, [f"{i:09d}" for i in range(s, s + m)], blk.data_ptr(), a.dim) == ("ok", ())
q = make_rows_torch(1, a.dim, SEED + 1, dev)[0].cpu().numpy()
code = nifs.METRIC_CODE[a.metric]
out = {"config": {"rows": a.rows, "dim": a.dim, "stages": stages, "candidates": a.candidates, "k": a.k, "metric": a.metric,
                  "prefix_mirror": not os.environ.get("VB_NO_PREFIX_MIRROR")}}
def timed(fn):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        r = fn()
    return (time.perf_counter() - t0) / a.iters * 1e3, r
ms, res = timed(lambda: nifs.flat_funnel_search(idx, q, code, stages, a.candidates, a.k))
assert res[0] == "ok", res
out["funnel_ms"] = ms
out["funnel_queries_per_s"] = 1e3 / ms
out["stage1_algorithmic_gbs"] = a.rows * stages[0] * 4 / ms / 1e6
for d in stages[:1] + [a.dim]:
    ms1, r1 = timed(lambda: nifs.flat_prefix_top_k(idx, None, q, code, d, a.candidates))
    assert r1[0] == "ok"
    out[f"prefix{d}_all_rows_ms"] = ms1
    out[f"prefix{d}_algorithmic_gbs"] = a.rows * d * 4 / ms1 / 1e6
ms2, r2 = timed(lambda: nifs.flat_search(idx, q, a.k))
out["exact_flat_ms"] = ms2
out["funnel_top"] = res[1][:3]
out["exact_top"] = r2[1][:3]
print(json.dumps(out))
