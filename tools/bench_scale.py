#!/usr/bin/env python
"""Per-GPU shard of the large BASELINE.json configs, built with the device bulk ingest:
  --mode batch      C3 shard: rows x 768 inner product, 1024-query batch, k = 100 (K2), plus one query (K1)
  --mode quantized  C4 on one GPU: rows x 1024 cosine, quantized_search with 1000 candidates -> k = 10 (K6+K3+K4)
Rows are generated on the device chunk by chunk (same generator as bench.py). One JSON line."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
from vettore_b200._lib import lib

ap = argparse.ArgumentParser()
ap.add_argument("--mode", choices=["batch", "quantized"], required=True)
ap.add_argument("--rows", type=int, default=12_500_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--candidates", type=int, default=1000)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
metric = "inner_product" if a.mode == "batch" else "cosine"
idx = getattr(nifs, f"flat_new_{metric}")()
assert nifs.flat_reserve(idx, a.rows) == ("ok", ())
t0 = time.perf_counter()
chunk = 1_000_000
for s in range(0, a.rows, chunk):
    m = min(chunk, a.rows - s)
    blk = make_rows_torch(m, a.dim, SEED + 7 * (s // chunk), dev)
    assert nifs.flat_insert_device(idx, [f"{i:09d}" for i in range(s, s + m)], blk.data_ptr(), a.dim) == ("ok", ())
    del blk
torch.cuda.synchronize()
ingest_s = time.perf_counter() - t0
out = {"mode": a.mode, "config": {"rows": a.rows, "dim": a.dim, "metric": metric}, "ingest_s": ingest_s,
       "ingest_rows_per_s": a.rows / ingest_s}
queries = make_rows_torch(a.nq, a.dim, SEED + 1, dev)

def dev_timed(fn, steps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

if a.mode == "batch":
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for nq, k in ((a.nq, a.k), (1, 10)):
        keys = torch.zeros(nq * k, dtype=torch.int64, device=dev)
        vals = torch.zeros(nq * k, dtype=torch.float32, device=dev)
        rws = torch.zeros(nq * k, dtype=torch.int32, device=dev)
        cnts = torch.zeros(nq, dtype=torch.int32, device=dev)
        dq = queries[:nq].contiguous()
        def step():
            rc = lib().vb_flat_search_device(idx.handle, C.c_void_p(dq.data_ptr()), nq, a.dim, k, C.c_void_p(keys.data_ptr()),
                                             C.c_void_p(vals.data_ptr()), C.c_void_p(rws.data_ptr()), C.c_void_p(cnts.data_ptr()), stream)
            assert rc == 0
        ms = dev_timed(step, a.steps if nq > 1 else 20)
        flops = 2.0 * nq * a.rows * a.dim
        res[f"nq{nq}_k{k}"] = {"device_ms": ms, "queries_per_s": nq / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9,
                              "tf32_tflops_issued": 3 * flops / ms / 1e9 if nq > 1 else None,
                              "hbm_gbs_if_single_pass": a.rows * a.dim * 4 / ms / 1e6}
        # spot check: the best hit of query 0 must carry the largest dot product of the whole shard
        if nq > 1:
            best_row, best_val = int(rws[0].item()), float(vals[0].item())
            res[f"nq{nq}_k{k}"]["best_of_query0"] = [best_row, best_val]
    # e2e through the host API for the batch
    qh = queries.cpu().numpy()
    nifs.flat_search_batch(idx, qh, a.k)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        st, hits = nifs.flat_search_batch(idx, qh, a.k)
    res["e2e_batch_ms"] = (time.perf_counter() - t0) / a.steps * 1e3
    assert st == "ok" and abs(hits[0][0][1] - res[f"nq{a.nq}_k{a.k}"]["best_of_query0"][1]) < 1e-5
    out["batch"] = res
else:
    q = queries[0].cpu().numpy()
    code = nifs.METRIC_CODE["cosine"]
    st, hits = nifs.flat_quantized_search(idx, q, code, a.candidates, 10)
    assert st == "ok", hits
    t0 = time.perf_counter()
    n_it = 20
    for _ in range(n_it):
        st, hits = nifs.flat_quantized_search(idx, q, code, a.candidates, 10)
    ms = (time.perf_counter() - t0) / n_it * 1e3
    nw = (a.dim + 63) // 64
    out["quantized"] = {"e2e_ms": ms, "queries_per_s": 1e3 / ms, "candidates": a.candidates,
                        "code_bytes": a.rows * nw * 8, "hamming_gbs_e2e": a.rows * nw * 8 / ms / 1e6, "top": hits[:3]}
    # exact search for comparison (recall of the sign-code candidate pass on random data is low by nature; not a parity metric)
    st, exact = nifs.flat_search(idx, q, 10)
    out["quantized"]["exact_top"] = exact[:3]
print(json.dumps(out))
