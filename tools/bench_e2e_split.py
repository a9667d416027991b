#!/usr/bin/env python
"""Where the end-to-end overhead of a single-query flat search goes: the C call alone (vb_flat_search + vb_hits_free)
against the Python mirror (nifs.flat_search: argument marshalling + hit list construction). 1M x 768 cosine, k = 10."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import SEED, make_rows_torch
from vettore_b200 import nifs
from vettore_b200._lib import lib

dev = torch.device("cuda", 0)
n, d, k = 1_000_000, 768, 10
idx = nifs.flat_new_cosine()
assert nifs.flat_reserve(idx, n) == ("ok", ())
x = make_rows_torch(n, d, SEED, dev)
assert nifs.flat_insert_device(idx, nifs.decimal_ids(0, n), x.data_ptr(), d) == ("ok", ())
del x
qs = make_rows_torch(64, d, SEED + 1, dev).cpu().numpy()
f32p = C.POINTER(C.c_float)
ptrs = [q.ctypes.data_as(f32p) for q in qs]
L = lib()


def c_only(i):
    h = C.c_void_p()
    rc = L.vb_flat_search(idx.handle, ptrs[i % 64], d, k, C.byref(h))
    assert rc == 0
    L.vb_hits_free(h)


def mirror(i):
    st, hits = nifs.flat_search(idx, qs[i % 64], k)
    assert st == "ok"


out = {}
for name, fn in (("c_call_only", c_only), ("python_mirror", mirror)):
    for i in range(50):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(1000):
        fn(i)
    out[name + "_ms"] = (time.perf_counter() - t0)
print(json.dumps(out))
