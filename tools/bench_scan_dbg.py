#!/usr/bin/env python
"""K1 stage elimination (VB_SCAN_DEBUG bit mask: 1 skip the group reduce + emit, 2 skip the FMAs, 4 skip the
collector checkpoints): 1M x 768 cosine, k = 10, device-timed. Results under a mask are garbage by design."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
from vettore_b200._lib import lib
n, d, k = 1_000_000, 768, 10
dev = torch.device("cuda", 0)
idx = nifs.flat_new_cosine()
assert nifs.flat_reserve(idx, n) == ("ok", ())
blk = make_rows_torch(n, d, SEED, dev)
assert nifs.flat_insert_device(idx, [f"{i:09d}" for i in range(n)], blk.data_ptr(), d) == ("ok", ())
del blk
q = make_rows_torch(8, d, SEED + 1, dev)
keys = torch.zeros(k, dtype=torch.int64, device=dev); vals = torch.zeros(k, dtype=torch.float32, device=dev)
rws = torch.zeros(k, dtype=torch.int32, device=dev); cnts = torch.zeros(1, dtype=torch.int32, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def step(i):
    assert lib().vb_flat_search_device(idx.handle, C.c_void_p(q[i % 8].data_ptr()), 1, d, k, C.c_void_p(keys.data_ptr()),
                                       C.c_void_p(vals.data_ptr()), C.c_void_p(rws.data_ptr()), C.c_void_p(cnts.data_ptr()), stream) == 0
for dbg in (sys.argv[1:] or ["0"]):
    os.environ["VB_SCAN_DEBUG"] = dbg
    for i in range(5): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print("debug", dbg, "ms/query %.4f  (%.0f GB/s)" % (ms, n * d * 4 / ms / 1e6))
