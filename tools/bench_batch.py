#!/usr/bin/env python
"""Secondary benchmark: batched flat search (BASELINE.json configs[1], batch of 1024 queries, k=10,
1M x 768 fp32 cosine) through the C ABI (host queries in, host hits out). One JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-gemm", action="store_true")
ap.add_argument("--no-check", action="store_true", help="timing experiments with VB_GEMM_DEBUG: results are garbage by design")
a = ap.parse_args()
if a.no_gemm:
    os.environ["VB_FLAT_NO_GEMM"] = "1"
import numpy as np
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
import oracle

dev = torch.device("cuda", 0)
rows = make_rows_torch(a.rows, a.dim, SEED, dev).cpu().numpy()
ids = [f"{i:09d}" for i in range(a.rows)]
idx = nifs.flat_new_cosine()
assert nifs.flat_insert_matrix(idx, ids, rows) == ("ok", ())
queries = make_rows_torch(a.nq, a.dim, SEED + 1, dev).cpu().numpy()
st, hits = nifs.flat_search_batch(idx, queries, a.k)
assert st == "ok" or a.no_check, hits
# parity spot check of 3 queries against the oracle on a 200k-row prefix
sub = min(a.rows, 200_000)
for qi in (() if a.no_check else (0, a.nq // 2, a.nq - 1)):
    ref = dict(oracle.flat_search_dense("cosine", rows[:sub], ids[:sub], queries[qi], a.k)[1])
    for hid, v in hits[qi]:
        if int(hid) < sub:
            assert hid in ref and abs(ref[hid] - v) <= 1e-5, (qi, hid, v, ref.get(hid))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    nifs.flat_search_batch(idx, queries, a.k)
dt = (time.perf_counter() - t0) / a.steps
# device-timed: queries resident in HBM, results stay on the device (vb_flat_search_device)
import ctypes as C
from vettore_b200._lib import lib
dq = torch.from_numpy(queries).to(dev)
keys = torch.zeros(a.nq * a.k, dtype=torch.int64, device=dev)
vals = torch.zeros(a.nq * a.k, dtype=torch.float32, device=dev)
rws = torch.zeros(a.nq * a.k, dtype=torch.int32, device=dev)
cnts = torch.zeros(a.nq, dtype=torch.int32, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def dev_step():
    rc = lib().vb_flat_search_device(idx.handle, C.c_void_p(dq.data_ptr()), a.nq, a.dim, a.k, C.c_void_p(keys.data_ptr()),
                                     C.c_void_p(vals.data_ptr()), C.c_void_p(rws.data_ptr()), C.c_void_p(cnts.data_ptr()), stream)
    assert rc == 0
for _ in range(2):
    dev_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    dev_step()
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / a.steps
# the device path must agree with the host path
hv = vals.cpu().numpy().reshape(a.nq, a.k)
for qi in (() if a.no_check else (0, a.nq - 1)):
    assert all(abs(hv[qi, i] - hits[qi][i][1]) <= 1e-6 for i in range(a.k)), (qi, hv[qi], hits[qi])
flops = 2.0 * a.nq * a.rows * a.dim
print(json.dumps({"metric": "batched flat queries/s (e2e through vb_flat_search_batch)", "value": a.nq / dt,
                  "ms_per_batch": dt * 1e3, "config": {"rows": a.rows, "dim": a.dim, "nq": a.nq, "k": a.k,
                                                       "kernel": "K1 per query" if a.no_gemm else "K2 tcgen05 3xTF32 + rescoring"},
                  "device_ms_per_batch": dev_ms, "device_queries_per_s": a.nq / (dev_ms * 1e-3),
                  "algorithmic_tflops_device": flops / (dev_ms * 1e-3) / 1e12,
                  "tf32_tflops_issued_device": 3 * flops / (dev_ms * 1e-3) / 1e12}))
