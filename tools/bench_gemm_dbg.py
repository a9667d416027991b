import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_rows_torch, SEED
from vettore_b200 import nifs
from vettore_b200._lib import lib
nq, n, d, k = int(sys.argv[1]), 1_000_000, 768, 10
dev = torch.device("cuda", 0)
rows = make_rows_torch(n, d, SEED, dev).cpu().numpy()
idx = nifs.flat_new_cosine()
assert nifs.flat_insert_matrix(idx, [f"{i:09d}" for i in range(n)], rows) == ("ok", ())
dq = make_rows_torch(nq, d, SEED + 1, dev)
keys = torch.zeros(nq * k, dtype=torch.int64, device=dev); vals = torch.zeros(nq * k, dtype=torch.float32, device=dev)
rws = torch.zeros(nq * k, dtype=torch.int32, device=dev); cnts = torch.zeros(nq, dtype=torch.int32, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def step():
    assert lib().vb_flat_search_device(idx.handle, C.c_void_p(dq.data_ptr()), nq, d, k, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_void_p(rws.data_ptr()), C.c_void_p(cnts.data_ptr()), stream) == 0
for dbg in (sys.argv[2:] or ["0"]):
    os.environ["VB_GEMM_DEBUG"] = dbg
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    print("debug", dbg, "ms/batch %.3f" % (e0.elapsed_time(e1) / 3))
