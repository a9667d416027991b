#!/usr/bin/env python
"""K1 (single-query flat scan, 1M x 768 cosine) over a sweep of k: device-timed through vb_flat_search_device."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
from vettore_b200._lib import lib

dev = torch.device("cuda", 0)
n, d = 1_000_000, 768
idx = nifs.flat_new_cosine()
assert nifs.flat_reserve(idx, n) == ("ok", ())
blk = make_rows_torch(n, d, SEED, dev)
assert nifs.flat_insert_device(idx, [f"{i:09d}" for i in range(n)], blk.data_ptr(), d) == ("ok", ())
del blk
q = make_rows_torch(8, d, SEED + 1, dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {}
for k in (1, 10, 50, 100, 200, 500, 1000):
    keys = torch.zeros(k, dtype=torch.int64, device=dev); vals = torch.zeros(k, dtype=torch.float32, device=dev)
    rws = torch.zeros(k, dtype=torch.int32, device=dev); cnts = torch.zeros(1, dtype=torch.int32, device=dev)
    def step(i):
        rc = lib().vb_flat_search_device(idx.handle, C.c_void_p(q[i % 8].data_ptr()), 1, d, k, C.c_void_p(keys.data_ptr()),
                                         C.c_void_p(vals.data_ptr()), C.c_void_p(rws.data_ptr()), C.c_void_p(cnts.data_ptr()), stream)
        assert rc == 0
    for i in range(5): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    out[f"k{k}"] = {"ms": round(ms, 4), "gbs": round(n * d * 4 / ms / 1e6, 1)}
print(json.dumps(out))
